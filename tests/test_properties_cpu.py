"""The vectorised property checkers of tests/properties.py against the oracle at small size:
what the full-size GPU tests (tests/test_zz5_gpu_fullsize.py) rely on."""
import numpy as np

import decks
import properties as pr
from cylindrical_epoch_b200.constants import M0, Q0


def _residual(w, d, sc):
    info = w.rank_info(0)
    return pr.gauss_residual(w.field(0, "exm")[0].real, w.field(0, "erm")[0].real, w.particles(0, 0).reshape(-1, 7),
                             -Q0, M0, sc["dt"], info["x_grid_min_local"], sc["y_grid_min_local"], sc["dx"], sc["dy"],
                             d.nx, d.ny)


def test_gauss_residual_is_frozen_on_a_thermal_two_mode_plasma():
    """the deck shape of BASELINE.json configs[1] in small: periodic x, reflecting r_max, two modes, 1 keV"""
    d = decks.thermal(nx=32, ny=16, n_mode=2, ppc=8)
    w = decks.make_oracle(d)
    w.call("init_half_step")
    sc = w.scalars()
    w.step(3)
    r0, f0 = _residual(w, d, sc)
    w.step(20)
    r1, f1 = _residual(w, d, sc)
    moved = np.abs(f1 - f0).max()
    assert moved > 0
    assert np.abs(r1 - r0).max() < 1e-9 * moved, np.abs(r1 - r0).max() / moved
    # and the checker is not vacuous: a charge that is off by one part in 1e6 shows
    p = w.particles(0, 0).reshape(-1, 7).copy()
    p[:, 6] *= 1.0 + 1e-6
    w.set_particles(0, 0, p)
    r2, _ = _residual(w, d, sc)
    assert np.abs(r2 - r0).max() > 1e-7 * moved


def test_same_multiset():
    a = np.random.default_rng(1).random(1000)
    assert pr.same_multiset(a, a[::-1].copy())
    b = a.copy()
    b[3] = np.nextafter(b[3], 2.0)
    assert not pr.same_multiset(a, b)
