"""cylindrical_epoch_b200 -- B200-native per-timestep PIC hot path of cylindrical EPOCH.

The product is the C-ABI library `libcylgpu.so` (hand-written CUDA for sm_100a, see
`csrc/` and `include/cylgpu.h`).  This package is the thin host-side mirror of the
reference's module procedures (`push_particles`, `update_eb_fields_half`, ...) over that
ABI, used by the tests and the benchmark in place of the Fortran driver, which cannot be
built in this environment.  There is no CPU fallback: without the CUDA library and a GPU
every compute call raises.
"""
from .constants import *  # noqa: F401,F403
from .decomp import slab_bounds, SlabGrid  # noqa: F401
from .hotpath import Slab, Species, Laser, CylGpuError  # noqa: F401
