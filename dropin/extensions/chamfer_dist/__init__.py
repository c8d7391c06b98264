"""`from extensions.chamfer_dist import ChamferFunction, ChamferDistance` (reference GRNet) -> B200."""
from softpool_b200.chamfer_dist import ChamferDistance, ChamferFunction  # noqa: F401
