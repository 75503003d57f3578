"""cylgpu_set_deferred_bcs: cylgpu_push returns before the leaver counts are known and the rest of particle_bcs
(count sync, compaction, neighbour exchange, arrivals) runs at the next call that touches particle state --
after current_finish and the field phases have been enqueued.  Results must be those of the ordinary order.

Sorts after the other test modules on purpose (see tests/test_zz1_gpu_moments.py): written after the round's
GPU budget was spent, first run on a B200 is the driver's.  The mode is opt-in and off by default.
"""
import numpy as np
import pytest

import decks
from parity import Pair, TOL_HOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("deckname,nranks", [("lwfa", 1), ("thermal", 1), ("thermal", 2), ("lwfa", 2), ("window", 2)])
def test_deferred_particle_bcs_gives_the_same_run(deckname, nranks):
    d = {"lwfa": lambda: decks.lwfa(nx=96, ny=32, n_mode=2, ppc_e=4, ppc_p=1),
         "thermal": lambda: decks.thermal(nx=64, ny=32, n_mode=2, ppc=8),
         "window": lambda: decks.lwfa(nx=64, ny=24, n_mode=2, ppc_e=4, ppc_p=1, window=True, t_centre=30e-15)}[deckname]()
    p = Pair(d, nranks=nranks)
    try:
        p.each(lambda s: s.set_deferred_bcs(True))
        tol = TOL_HOT if deckname == "thermal" else 1e-9
        for _ in range(3):
            p.step(10)
            # the outstanding particle_bcs contains the neighbour exchange: all slabs complete it together (the
            # comparisons below query one slab after the other from this thread)
            p.each(lambda s: s.synchronize())
            p.check_counts()
            p.check_fields(tol)
            p.check_particles(tol)
        p.check_cells()
        # and switching it off again mid-run is harmless
        p.each(lambda s: s.set_deferred_bcs(False))     # (completes what is outstanding, on all slabs together)
        p.step(5)
        p.check_counts()
        p.check_particles(tol)
    finally:
        p.close()
