"""`import dist_chamfer as cd` (reference train.py:18-19, val.py:16-17) -> the B200 implementation."""
from softpool_b200.dist_chamfer import chamferDist, chamferFunction  # noqa: F401
