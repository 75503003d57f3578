"""GPU parity of the calc_df.F90 particle moments (cylgpu_particle_moment, csrc/moments.cuh)
against the oracle (oracle/cyl_moments.cpp), through the C-ABI.

The test_zz<n>_gpu_* modules sort after the other test modules on purpose: they were written after the
round's GPU budget was spent, so their first run on a B200 is the driver's; under `pytest -x` a failure
here cannot hide the results of the parity tests that were run on the GPU during development.  They are
numbered by confidence: first those whose kernels were checked on the CPU by the emulation tests.
"""
import numpy as np
import pytest

import decks
from parity import Pair

pytestmark = pytest.mark.gpu

# (kind, direction) as in tests/test_oracle_moments.py
CASES = [("mass_density", 0), ("number_density", 0), ("ekbar", 0), ("ekflux", 1), ("ekflux", -2), ("ekflux", 3),
         ("ppc", 0), ("average_weight", 0), ("temperature", 0), ("temperature", 2), ("species_current", 1),
         ("species_current", 3), ("average_momentum", 2)]
# deposit summation order differs (REDs); same bound as the density diagnostics
TOL_MOMENT = 1.0e-12
# moments that weigh with the momenta inherit the parity of the momenta themselves after the 5 steps
# (tests/parity.py TOL: 1e-10 between the device's and the oracle's particles)
TOL_MOMENTUM_MOMENT = 1.0e-10
MOMENTUM_MOMENTS = ("ekbar", "ekflux", "species_current", "average_momentum")
# the temperature subtracts cell means from particle momenta: the difference between the two particle sets
# enters sigma relative to the spread inside a cell, not to |p| (cold beams, coherent quiver motion)
TOL_TEMPERATURE = 1.0e-8


def _deck(name):
    return {"lwfa": lambda: decks.lwfa(nx=64, ny=24, n_mode=2, ppc_e=4, ppc_p=1),
            "thermal": lambda: decks.thermal(nx=64, ny=32, n_mode=2, ppc=8),
            "drift": lambda: decks.drift()}[name]()


@pytest.mark.parametrize("deckname,nranks", [("lwfa", 1), ("thermal", 1), ("drift", 1), ("lwfa", 2), ("thermal", 2)])
def test_particle_moments(deckname, nranks):
    d = _deck(deckname)
    p = Pair(d, nranks=nranks)
    try:
        p.step(5)
        # the diagnostics are compared on IDENTICAL particle sets: the oracle's lists go to the device (the
        # parity of the stepping itself is the business of tests/test_gpu_parity.py)
        for k, s in enumerate(p.slabs):
            for isp in range(len(d.species)):
                s.upload_particles(isp, p.oracle.particles(k, isp))
        for kind, direction in CASES:
            for isp in [-1] + list(range(len(d.species))):
                ref = p.oracle.moment(kind, isp, direction)
                got = [None] * nranks
                p.each(lambda s: got.__setitem__(p.slabs.index(s), s.moment(kind, isp, direction)))
                den = max(np.abs(r).max() for r in ref)
                assert den > 0, (kind, direction)
                tol = TOL_TEMPERATURE if kind == "temperature" else (
                    TOL_MOMENTUM_MOMENT if kind in MOMENTUM_MOMENTS else TOL_MOMENT)
                for k in range(nranks):
                    if kind in ("ppc", "average_weight"):
                        # integer cell assignment: bit-exact counts, weights to rounding of the sum order
                        if kind == "ppc":
                            assert np.array_equal(got[k], ref[k]), (deckname, k)
                            continue
                    err = np.abs(got[k] - ref[k]).max() / den
                    assert err < tol, (deckname, kind, direction, isp, k, err)
    finally:
        p.close()


def test_moment_argument_errors():
    import cylindrical_epoch_b200 as ce
    d = decks.lwfa(nx=32, ny=12, n_mode=1, ppc_e=1)
    p = Pair(d, init_half_step=False)
    try:
        s = p.slabs[0]
        for kind in ("species_current", "average_momentum"):   # calc_df.F90:1053-1059,1155-1161
            with pytest.raises(ce.CylGpuError, match="direction"):
                s.moment(kind, 0, 0)
        with pytest.raises(ce.CylGpuError, match="unknown moment"):
            s.moment(99, 0, 0)
        with pytest.raises(ce.CylGpuError, match="bad argument"):
            s.moment("ppc", 7, 0)
    finally:
        p.close()
