"""`import softpool as sp` (reference model.py:12) -> the B200 implementation."""
from softpool_b200.softpool import Periodics, SoftPool, SoftPoolFeat, Sorter, train2cabins  # noqa: F401
