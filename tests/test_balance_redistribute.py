"""CPU tests of the slab re-balancer's data movement (cylindrical_epoch_b200/balance.py): the in-process
redistribution against a global picture, and the one-slab-per-process version over torch.distributed (gloo,
world_size 2 and 3) against the in-process one, on states cut from a 3-rank oracle world (so that ghost columns do
equal the neighbours' interiors, as after the exchanges that end every phase)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import decks
from cylindrical_epoch_b200 import balance
from cylindrical_epoch_b200.constants import FIELD_NAMES, NG, SNAP_NAMES


def _states(deck_name, nranks, seed=3):
    d = decks.thermal(nx=60, ny=10, n_mode=2, ppc=3) if deck_name == "thermal" else \
        decks.lwfa(nx=60, ny=10, n_mode=2, ppc_e=3, ppc_p=0)
    w = decks.make_oracle(d, nranks=nranks)
    w.call("init_half_step")
    w.step(4)
    states, bounds = [], []
    for k in range(nranks):
        info = w.rank_info(k)
        bounds.append((info["cell_x_min"], info["cell_x_max"]))
        states.append(dict(fields={n: w.field(k, n).copy() for n in FIELD_NAMES},
                           snaps={n: w.field(k, n).copy() for n in SNAP_NAMES},
                           particles=[w.particles(k, 0).reshape(-1, 7).copy()], rng=w.rng_state(k),
                           bounds=bounds[-1]))
    sc = w.scalars()
    return d, w, states, bounds, sc


def _edges(new_bounds, sc):
    x_grid_min = sc["x_min"] + sc["dx"] / 2.0
    return [(x_grid_min + (lo - 1) * sc["dx"] - 0.5 * sc["dx"], x_grid_min + (hi - 1) * sc["dx"] + 0.5 * sc["dx"])
            for lo, hi in new_bounds]


@pytest.mark.parametrize("deck_name", ["thermal", "lwfa"])
def test_in_process_redistribution_keeps_every_column_and_particle(deck_name):
    d, w, states, bounds, sc = _states(deck_name, 3)
    periodic = deck_name == "thermal"
    new_bounds = [(1, 11), (12, 47), (48, 60)]
    new = balance.redistribute_local(states, new_bounds, 60, _edges(new_bounds, sc), periodic)
    for n in FIELD_NAMES:
        G = balance.assemble_global([s["fields"][n] for s in states], bounds, 60, periodic)
        G2 = balance.assemble_global([s["fields"][n] for s in new], new_bounds, 60, periodic)
        assert np.array_equal(G, G2)
        # ... and the new slabs' ghost columns are their new neighbours' interiors
        for k in range(2):
            a, b = new[k]["fields"][n], new[k + 1]["fields"][n]
            nxa = new_bounds[k][1] - new_bounds[k][0] + 1
            assert np.array_equal(a[..., NG + nxa:], b[..., NG:2 * NG])
            assert np.array_equal(b[..., :NG], a[..., nxa:nxa + NG])
    before = np.concatenate([s["particles"][0] for s in states])
    after = np.concatenate([s["particles"][0] for s in new])
    assert np.array_equal(decks.sort_particles(before), decks.sort_particles(after))
    for k, (xl, xr) in enumerate(_edges(new_bounds, sc)):
        x = new[k]["particles"][0][:, 0]
        inside = (x >= xl) & (x < xr)
        assert inside.all() or k in (0, 2)      # only the end slabs may hold particles that left the box


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, deck_name, new_bounds, q):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d, w, states, bounds, sc = _states(deck_name, world)
    periodic = deck_name == "thermal"
    edges = _edges(new_bounds, sc)
    mine = balance.redistribute_dist(states[rank], bounds, new_bounds, 60, edges, rank, periodic)
    ref = balance.redistribute_local(states, new_bounds, 60, edges, periodic)[rank]
    ok = all(np.array_equal(mine["fields"][n], ref["fields"][n]) for n in FIELD_NAMES)
    ok = ok and np.array_equal(decks.sort_particles(mine["particles"][0]), decks.sort_particles(ref["particles"][0]))
    ok = ok and mine["bounds"] == ref["bounds"]
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("deck_name,world,new_bounds", [("thermal", 2, [(1, 17), (18, 60)]),
                                                        ("lwfa", 3, [(1, 30), (31, 36), (37, 60)]),
                                                        ("thermal", 3, [(1, 8), (9, 52), (53, 60)])])
def test_distributed_redistribution_over_gloo(deck_name, world, new_bounds):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, deck_name, new_bounds, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
    assert res == [(r, True) for r in range(world)], res


def test_plan_accepts_only_a_real_improvement(cylgpu_lib):
    """balance_workload's decision (balance.F90:143-225): a lopsided load is re-split, an even one is left alone"""
    nxg, ny = 120, 16
    bounds = [(1, 40), (41, 80), (81, 120)]
    even = [np.full(40 + 2 * NG, 0, dtype=np.int64) for _ in bounds]
    for c in even:
        c[NG:-NG] = 64
    new, frac, after = balance.plan(cylgpu_lib, even, bounds, nxg, ny, over_ride=True)
    assert new is None and frac > 0.99
    lop = [c.copy() for c in even]
    lop[0][NG:NG + 20] = 4000            # a dense slab in the first 20 columns
    new, frac, after = balance.plan(cylgpu_lib, lop, bounds, nxg, ny, over_ride=True)
    assert new is not None and after > 1.05 * frac
    assert new[0][1] < 40 and new[0][0] == 1 and new[-1][1] == nxg
    assert all(new[k + 1][0] == new[k][1] + 1 for k in range(2))
