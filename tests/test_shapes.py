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
