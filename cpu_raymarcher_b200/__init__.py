"""cpu_raymarcher_b200 — B200-native implementation of vxlerian/cpu-raymarcher's per-pixel raymarch
hot path, behind the reference's worker Job/Result contract.  The compute lives in librm_b200.so
(hand-written sm_100a CUDA behind the C ABI of include/rm.h); this package is the thin host side."""
from ._lib import ACCELS, ALGORITHMS, LIB_PATH, SHADERS, RmError  # noqa: F401
from .camera import Camera  # noqa: F401
from .renderer import Context, Frame, build_bvh, build_bvh_scene, build_octree, build_octree_scene  # noqa: F401
from .scene import Scene  # noqa: F401
from .pool import RaymarchPool  # noqa: F401
from .worker import RaymarchWorker  # noqa: F401
