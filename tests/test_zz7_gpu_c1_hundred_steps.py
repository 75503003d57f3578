"""BASELINE.json configs[0] at its own size -- the scaled Wakefield_Lifschitz09 deck, 512 x 64 cells, m = 0..1,
16 + 4 particles per cell, laser from x_min -- for the N = 10 and N = 100 steps of SURVEY.md section 8(d)'s
parity report: integer outputs exact, fields / currents / phase space within the stated tolerance.

Sorts after the other test modules on purpose (see tests/test_zz1_gpu_moments.py): added after the round's
GPU budget was spent; it only exercises paths the earlier parity tests verified on B200s, at a size and
step count they did not reach (the oracle needs about a minute for it).
"""
import pytest

import decks
from parity import Pair

pytestmark = pytest.mark.gpu

# rounding differences between two correct FP64 implementations grow along the trajectories: 1e-10 after 10
# steps, 1e-9 after 50 (tests/test_gpu_parity.py); for 100 steps with the pulse inside the plasma
TOL_100 = 1.0e-7


def test_c1_deck_10_and_100_steps():
    d = decks.lwfa(nx=512, ny=64, n_mode=2, ppc_e=16, ppc_p=4, t_centre=12e-15)
    p = Pair(d)
    try:
        p.step(10)
        p.check_counts()
        p.check_cells()
        e10 = max(p.check_fields().values())
        q10 = p.check_particles()
        p.step(90)
        p.check_counts()
        # (per-particle cells are compared bit for bit at 10 steps only: after 100, two correct FP64 runs may
        # put a particle that sits within 1e-9 of a cell edge on either side)
        e100 = max(p.check_fields(TOL_100).values())
        q100 = p.check_particles(TOL_100)
        print(f"C1 parity: fields {e10:.2e} / particles {q10:.2e} after 10 steps, {e100:.2e} / {q100:.2e} after 100")
    finally:
        p.close()
