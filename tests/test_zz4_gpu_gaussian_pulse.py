"""example_decks/gaussian_pulse.deck on the GPU against the oracle (which tests/test_oracle.py pins to the
deck's own design target): the laser source with a phase function of r evaluated by the host mirror, the
per-mode FDTD and the outflow boundaries over 300 steps of a 500 x 100 vacuum grid.

Sorts after the other test modules on purpose (see tests/test_zz1_gpu_moments.py): written after
the round's GPU budget was spent, first run on a B200 is the driver's.
"""
import pytest

import decks
from parity import Pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nranks", [1, 2])
def test_gaussian_pulse_deck(nranks):
    d = decks.gaussian_pulse()
    p = Pair(d, nranks=nranks)
    try:
        p.step(300)
        errs = p.check_fields(1e-10)
        assert max(errs.values()) < 1e-10
    finally:
        p.close()
