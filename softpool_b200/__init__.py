"""softpool_b200 -- B200-native (sm_100a) SoftPool sort/top-k + gather and Chamfer distance,
behind the Python surface of wangyida/softpool.  See DESIGN.md / INTEGRATION.md."""
from . import _lib, ops  # noqa: F401
from .softpool import Periodics, SoftPool, SoftPoolFeat, Sorter, train2cabins  # noqa: F401
from .dist_chamfer import chamferDist, chamferFunction  # noqa: F401
from .chamfer_dist import ChamferDistance, ChamferFunction  # noqa: F401

__version__ = "0.1.0"
