"""The other particle shapes of the reference (-DPARTICLE_SHAPE_TOPHAT / -DPARTICLE_SHAPE_BSPLINE3).
As in the reference the shape is a compile-time choice -- it sets ng = png + 2 and with it the layout of every array
(constants.F90:524-545) -- so each shape is its own build of the oracle (oracle/Makefile `shapes`) and of the
library, selected for a whole process by CYL_SHAPE.  Here: the oracle builds of both shapes pass the physics pins that
do not assume the triangle -- exact discrete charge continuity of the deposit WITH THE SHAPE'S OWN WEIGHTS, the Gauss
law residual, the axis identities, vacuum propagation, the gaussian_pulse deck's focus, Boris rotation, the Langmuir
period, the two-stream deck, the moments' known answers."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.mark.parametrize("shape", ["tophat", "bspline3"])
def test_oracle_pins_hold_for_the_shape(shape):
    env = dict(os.environ, CYL_SHAPE=shape)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "test_oracle.py"),
                        os.path.join(HERE, "test_oracle_moments.py"), "-q", "-x", "-p", "no:cacheprovider"],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=1500)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:]
    assert r.returncode == 0 and " passed" in tail and "failed" not in tail, r.stdout[-3000:]
    assert int(tail.split(" passed")[0].split()[-1]) >= 13, tail


@pytest.mark.parametrize("shape", ["tophat", "bspline3"])
def test_product_kernels_of_the_shape_track_its_oracle(shape):
    """tests/test_kernel_emulation.py under CYL_SHAPE: the product's kernel sources compiled for the shape
    (-DCYL_SHAPE, csrc/shape.cuh, csrc/push_shapes.cuh) and run on the CPU against the oracle build of the same shape --
    whole steps from the generic push kernel (laser-plasma deck to 1e-10), the nine particle moments, number and
    charge density, particle_bcs, the field kernels with ng = 4 / 6, the communication-avoiding field phases."""
    env = dict(os.environ, CYL_SHAPE=shape)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "test_kernel_emulation.py"), "-q", "-x",
                        "-p", "no:cacheprovider"], capture_output=True, text=True, cwd=ROOT, env=env, timeout=1500)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:]
    assert r.returncode == 0 and " passed" in tail and "failed" not in tail, r.stdout[-3000:]
    assert int(tail.split(" passed")[0].split()[-1]) >= 60, tail


@pytest.mark.parametrize("shape,code,ng", [("triangle", 0, 5), ("tophat", 1, 4), ("bspline3", 2, 6)])
def test_every_shape_build_exports_the_whole_abi(shape, code, ng):
    """libcylgpu.so / libcylgpu_tophat.so / libcylgpu_bspline3.so: the same entry points (include/cylgpu.h), each
    telling its shape and its ng -- what the Fortran shim checks against its own -DPARTICLE_SHAPE_* """
    import ctypes as C
    from cylindrical_epoch_b200 import _lib, build
    path = build.lib_path(shape)
    if not os.path.exists(path):
        if not os.path.exists(build.NVCC):
            pytest.skip("no nvcc here and the library of this shape is not built")
        build.build(shape=shape)
    lib = C.CDLL(path)     # (RTLD_LOCAL: three builds of the same symbols side by side)
    for name in _lib.SYMBOLS:
        assert hasattr(lib, name), (shape, name)
    lib.cylgpu_shape.restype = C.c_int
    lib.cylgpu_ghost_cells.restype = C.c_int
    assert (lib.cylgpu_shape(), lib.cylgpu_ghost_cells()) == (code, ng)


def test_balancer_keeps_slabs_two_halos_wide():
    """calculate_breaks may cut down to ncell_min = (png + 1) / 2 + 1 columns (constants.F90:548); a handle needs
    2 ng: balance.widen_narrow_slabs moves such breaks with the reference's own backwards / forwards passes"""
    from cylindrical_epoch_b200.balance import widen_narrow_slabs
    from cylindrical_epoch_b200.constants import NG
    w = 2 * NG
    assert widen_narrow_slabs([(1, 32), (33, 64)], 64) == [(1, 32), (33, 64)]                 # nothing to do
    assert widen_narrow_slabs([(1, 3), (4, 64)], 64) == [(1, w), (w + 1, 64)]                 # first slab too narrow
    assert widen_narrow_slabs([(1, 61), (62, 64)], 64) == [(1, 64 - w), (64 - w + 1, 64)]     # last slab too narrow
    out = widen_narrow_slabs([(1, 40), (41, 43), (44, 64)], 64)                               # a narrow one in the middle
    assert out[0][0] == 1 and out[-1][1] == 64
    assert all(hi - lo + 1 >= w for lo, hi in out) and all(out[k + 1][0] == out[k][1] + 1 for k in range(2))
    with pytest.raises(RuntimeError):
        widen_narrow_slabs([(1, 5), (6, 10), (11, 15)], 15)                                   # 15 columns cannot hold three


@pytest.mark.parametrize("shape", ["tophat", "bspline3"])
def test_host_side_files_under_the_shape(shape):
    """the rest of the CPU suite with ng = 4 / 6: the SDF writer read back by the reference's own reader, the host
    mirror, the ABI, the counter-based column, the property checkers (Gauss law with the shape's weights), the slab
    re-balancer over gloo"""
    env = dict(os.environ, CYL_SHAPE=shape)
    files = ["test_sdf.py", "test_host_logic.py", "test_abi.py", "test_counter_insert.py", "test_properties_cpu.py",
             "test_balance_redistribute.py"]
    r = subprocess.run([sys.executable, "-m", "pytest"] + [os.path.join(HERE, f) for f in files] +
                       ["-q", "-x", "-m", "not gpu", "-p", "no:cacheprovider"], capture_output=True, text=True, cwd=ROOT,
                       env=env, timeout=1500)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:]
    assert r.returncode == 0 and " passed" in tail and "failed" not in tail, r.stdout[-3000:]
