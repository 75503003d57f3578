"""The slab re-balancer on the GPU (balance.F90 for nprocy = 1): a lopsided plasma runs on evenly
split slabs, balance_workload re-splits it (cylgpu_load_x on the device, cylgpu_calculate_breaks, whole columns and
their particles moved to new handles), and the run continues -- against the oracle, which keeps its even slabs: the
GLOBAL arrays and the global particle set must stay the reference's."""
import numpy as np
import pytest

import decks
from cylindrical_epoch_b200 import balance
from cylindrical_epoch_b200.constants import FIELD_NAMES, NG
from parity import Pair, TOL_HOT, J_FLOOR

pytestmark = pytest.mark.gpu


def _lopsided(oracle, nranks, keep_fraction, x_max):
    """drop the plasma to the right of keep_fraction * x_max from the oracle's initial state"""
    for k in range(nranks):
        for isp in range(oracle.n_species):
            p = oracle.particles(k, isp).reshape(-1, 7)
            oracle.set_particles(k, isp, np.ascontiguousarray(p[p[:, 0] < keep_fraction * x_max]))


def _compare_global(p, tol):
    nxg = p.deck.nx
    periodic = p.slabs[0].periodic_x
    ob = [(p.oracle.rank_info(k)["cell_x_min"], p.oracle.rank_info(k)["cell_x_max"]) for k in range(p.nranks)]
    sb = [(s.grid.cell_x_min, s.grid.cell_x_max) for s in p.slabs]
    qnc = sum(abs(sp.charge) * sp.density for sp in p.deck.species) * 2.99792458e8
    worst = 0.0
    for name in FIELD_NAMES:
        ref = balance.assemble_global([p.oracle.field(k, name) for k in range(p.nranks)], ob, nxg, periodic)
        got = balance.assemble_global([s.download_field(name) for s in p.slabs], sb, nxg, periodic)
        den = np.abs(ref).max()
        if name.startswith("j"):
            den = max(den, J_FLOOR * qnc)
        if den > 0:
            worst = max(worst, np.abs(got - ref).max() / den)
        if not np.abs(got - ref).max() <= tol * den:
            where = np.unravel_index(np.abs(got - ref).argmax(), ref.shape)
            bad_cols = np.unique(np.nonzero(np.abs(got - ref) > tol * den)[2])
            raise AssertionError((name, np.abs(got - ref).max() / max(den, 1e-300), "worst at (mode, row, col)", where,
                                  "columns off", bad_cols.tolist(), "oracle bounds", ob, "slab bounds", sb))
    for isp in range(len(p.deck.species)):
        ref = np.concatenate([p.oracle.particles(k, isp).reshape(-1, 7) for k in range(p.nranks)])
        got = np.concatenate([s.download_particles(isp) for s in p.slabs])
        assert ref.shape == got.shape
        ref, got = ref[np.lexsort((ref[:, 0], ref[:, 6]))], got[np.lexsort((got[:, 0], got[:, 6]))]
        assert np.array_equal(ref[:, 6], got[:, 6])
        for cols in (slice(0, 3), slice(3, 6)):
            den = np.abs(ref[:, cols]).max()
            if den > 0:
                assert np.abs(ref[:, cols] - got[:, cols]).max() <= tol * den
    return worst


@pytest.mark.parametrize("deckname,nranks", [("thermal", 3), ("window", 2), ("lwfa", 3)])
def test_rebalance_mid_run(deckname, nranks):
    d = {"thermal": lambda: decks.thermal(nx=96, ny=16, n_mode=2, ppc=6),
         "lwfa": lambda: decks.lwfa(nx=96, ny=16, n_mode=2, ppc_e=4, ppc_p=1),
         "window": lambda: decks.lwfa(nx=64, ny=16, n_mode=2, ppc_e=4, ppc_p=1, window=True, t_centre=30e-15)}[deckname]()
    tol = TOL_HOT if deckname == "thermal" else 1e-9
    p = Pair(d, nranks=nranks, prepare=lambda o: _lopsided(o, nranks, 0.3, d.x_max))
    try:
        p.step(6)
        _compare_global(p, tol)
        even = [(s.grid.cell_x_min, s.grid.cell_x_max) for s in p.slabs]
        n_before = [sum(s.particle_count(i) for i in range(len(d.species))) for s in p.slabs]
        p.slabs, report = balance.rebalance_slabs(p.slabs, transport_kw=lambda k: dict(fabric=p.fabric), over_ride=True)
        assert report["redistributed"], report
        new = [(s.grid.cell_x_min, s.grid.cell_x_max) for s in p.slabs]
        assert new != even and new[0][0] == 1 and new[-1][1] == d.nx
        n_after = [sum(s.particle_count(i) for i in range(len(d.species))) for s in p.slabs]
        assert sum(n_after) == sum(n_before) and max(n_after) < max(n_before)      # the load did even out
        assert report["after"] > 1.05 * report["balance"]
        _compare_global(p, tol)        # moving the columns changed nothing
        p.step(8)
        worst = _compare_global(p, tol)
        if deckname == "window":
            assert p.slabs[0].window_shifts_total == int(p.oracle.scalars()["window_shifts_total"]) >= 6
            assert p.slabs[-1].rng_get_state() == p.oracle.rng_state(nranks - 1)
        print(deckname, "bounds", even, "->", new, "balance %.3f -> %.3f" % (report["balance"], report["after"]),
              "worst field error after 8 more steps %.2e" % worst)
    finally:
        p.close()


@pytest.mark.parametrize("bounds", [[(1, 32), (33, 64)], [(1, 20), (21, 64)], [(1, 12), (13, 64)], [(1, 2 * NG), (2 * NG + 1, 64)]])
def test_window_deck_prescribed_splits(bounds):
    """the moving-window deck handed over to new handles with a prescribed split (the first = the split it has: a
    pure hand-over of the state), then 8 more steps"""
    d = decks.lwfa(nx=64, ny=16, n_mode=2, ppc_e=4, ppc_p=1, window=True, t_centre=30e-15)
    p = Pair(d, nranks=2, prepare=lambda o: _lopsided(o, 2, 0.3, d.x_max))
    try:
        p.step(6)
        p.slabs, report = balance.rebalance_slabs(p.slabs, transport_kw=lambda k: dict(fabric=p.fabric), over_ride=True,
                                                  force_bounds=bounds)
        _compare_global(p, 1e-9)
        for k in range(8):
            p.step(1)
            try:
                _compare_global(p, 1e-9)
            except AssertionError as e:
                raise AssertionError(("step after hand-over", k + 1, str(e)[:600]))
    finally:
        p.close()
