"""CPU tests of the host-side logic: slab decomposition / grid / dt against the oracle's
restatement of mpi_routines.F90, setup.F90 and utilities.f90; BC normalisation; and the
N > 1 exchange protocol over gloo with world_size 2."""
import os
import socket
import sys

import numpy as np
from cylindrical_epoch_b200.constants import NG
import pytest
import torch
import torch.multiprocessing as mp

import decks
from cylindrical_epoch_b200 import decomp, hotpath
from cylindrical_epoch_b200.constants import *  # noqa: F401,F403


@pytest.mark.parametrize("nx,nranks", [(96, 1), (96, 2), (100, 3), (1000, 7), (8192, 8), (37, 4)])
def test_slab_bounds_match_reference_split(nx, nranks):
    d = decks.lwfa(nx=nx, ny=16, ppc_e=0)
    d.species = []
    w = decks.make_oracle(d, nranks=nranks, load=False)
    b = decomp.slab_bounds(nx, nranks)
    assert b[0][0] == 1 and b[-1][1] == nx
    for k in range(nranks):
        info = w.rank_info(k)
        assert b[k] == (info["cell_x_min"], info["cell_x_max"])
        g = decomp.SlabGrid(nx, 16, nranks, k, d.x_min, d.x_max, d.y_max)
        assert g.nx == info["nx"]
        # bit-exact grid scalars (they decide which slab owns a particle)
        assert g.x_grid_min_local == info["x_grid_min_local"]
        assert g.x_min_local == info["x_min_local"] and g.x_max_local == info["x_max_local"]
        sc = w.scalars()
        assert g.dt == sc["dt"] and g.dx == sc["dx"] and g.dy == sc["dy"]
        assert g.y_grid_min_local == sc["y_grid_min_local"]


def test_window_grid_shift_matches_oracle():
    d = decks.lwfa(nx=64, ny=16, ppc_e=0, window=True)
    d.species = []
    w = decks.make_oracle(d, nranks=2, load=False)
    grids = [decomp.SlabGrid(64, 16, 2, k, d.x_min, d.x_max, d.y_max) for k in range(2)]
    for _ in range(7):
        w.call("step")
    n = int(w.scalars()["window_shifts_total"])
    assert n >= 3
    for g in grids:
        for _ in range(n):
            g.shift()
    for k, g in enumerate(grids):
        info = w.rank_info(k)
        assert g.x_grid_min_local == info["x_grid_min_local"]
        assert g.x_min_local == info["x_min_local"] and g.x_max_local == info["x_max_local"]
    sc = w.scalars()
    assert grids[0].x_min == sc["x_min"] and grids[0].x_max == sc["x_max"]


def test_bc_normalisation_matches_setup_boundaries():
    raw = [BC_SIMPLE_LASER, BC_OPEN, 0, BC_REFLECT]
    d = decks.lwfa(nx=32, ny=16, ppc_e=1)
    d.bc_field = tuple(raw)
    d.species[0].bc_particle = (BC_SIMPLE_LASER, BC_OTHER, BC_OPEN, BC_CONDUCT)
    w = decks.make_oracle(d, load=False)
    bc, add_laser = hotpath.normalise_bc_field(raw)
    ref = w.bc_field()
    assert [bc[i] for i in (0, 1, 3)] == [ref[i] for i in (0, 1, 3)]
    assert add_laser[0] and not add_laser[1]
    assert hotpath.normalise_bc_particle(d.species[0].bc_particle) == w.bc_particle(0)


def test_laser_sources_match_oracle():
    d = decks.lwfa(nx=32, ny=24, ppc_e=0)
    d.species = []
    w = decks.make_oracle(d, load=False)
    # the Slab constructor needs a GPU; exercise the host function on a stand-in object
    class Stub:
        pass
    s = Stub()
    s.grid = decomp.SlabGrid(32, 24, 1, 0, d.x_min, d.x_max, d.y_max)
    # second laser: the curved phase front of example_decks/gaussian_pulse.deck (phase function of y)
    d.lasers.append(dict(decks.gaussian_pulse().lasers[0]))
    w.add_laser(**d.lasers[-1])
    s.lasers = [hotpath.Laser(**L) for L in d.lasers]
    s.add_laser = [True, False, False, False]
    for t in (0.0, 1.3e-14, 3.0e-14, 4.1e-14):
        w.set_time(t)
        s.time = t
        r1, r2 = w.laser_sources(BD_X_MIN)
        g1, g2 = hotpath.Slab.laser_sources(s, BD_X_MIN)
        np.testing.assert_allclose(g1, r1, rtol=1e-14, atol=1e-15 * np.abs(r1).max())
        np.testing.assert_allclose(g2, r2, rtol=1e-14, atol=1e-15 * max(np.abs(r2).max(), 1e-300))


# ------------------------------------------------------------------ world_size 2 over gloo
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ring_worker(rank, world, port, periodic, q):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cylindrical_epoch_b200.transport import TorchRing, neighbours
    import decks as dk
    ring = TorchRing()
    left, right = neighbours(rank, world, periodic)
    # --- field halo of a 2-rank oracle world, driven through the ring ---
    d = dk.thermal(nx=32, ny=12, ppc=2) if periodic else dk.lwfa(nx=32, ny=12, ppc_e=2)
    w = dk.make_oracle(d, nranks=world)
    rng = np.random.default_rng(5)
    for k in range(world):   # same random field on both processes
        f = w.field(k, "erm")
        f[...] = rng.standard_normal(f.shape) + 1j * rng.standard_normal(f.shape)
    mine = w.field(rank, "erm").copy()
    nx = w.rank_info(rank)["nx"]
    from cylindrical_epoch_b200.constants import NG as NGH
    send_l = torch.from_numpy(np.ascontiguousarray(mine[:, :, NGH:2 * NGH]))              # columns 1..ng
    send_r = torch.from_numpy(np.ascontiguousarray(mine[:, :, nx:nx + NGH]))              # nx+1-ng..nx
    recv_l = torch.zeros_like(send_l)
    recv_r = torch.zeros_like(send_r)
    ring.sendrecv(left, right, send_l if left >= 0 else None, recv_l if left >= 0 else None,
                  send_r if right >= 0 else None, recv_r if right >= 0 else None)
    if left >= 0:
        mine[:, :, 0:NGH] = recv_l.numpy()
    if right >= 0:
        mine[:, :, nx + NGH:nx + 2 * NGH] = recv_r.numpy()
    w.call("efield_bcs")    # the oracle's own halo + edge BCs on the same data
    ref = w.field(rank, "erm")
    ok_halo = True
    rows = slice(0, 12 + NGH - 1)   # rows above ny are rewritten by the r_max edge condition afterwards
    if left >= 0:
        ok_halo &= np.array_equal(mine[:, rows, 0:NGH], ref[:, rows, 0:NGH])
    if right >= 0:
        ok_halo &= np.array_equal(mine[:, rows, nx + NGH:], ref[:, rows, nx + NGH:])
    # --- variable-size particle exchange (counts, then payload) ---
    pl = torch.full((3 + rank, 7), float(10 * rank + 1), dtype=torch.float64)
    pr = torch.full((5 - rank, 7), float(10 * rank + 2), dtype=torch.float64)
    fl, fr = ring.exchange_counts_then_payload(left, right, pl, pr)
    other = 1 - rank
    exp_fl = (5 - other, 10 * other + 2) if left >= 0 else (0, None)    # left neighbour's right-going
    exp_fr = (3 + other, 10 * other + 1) if right >= 0 else (0, None)   # right neighbour's left-going
    ok_p = fl.shape[0] == exp_fl[0] and fr.shape[0] == exp_fr[0]
    if exp_fl[0]:
        ok_p &= bool((fl == exp_fl[1]).all())
    if exp_fr[0]:
        ok_p &= bool((fr == exp_fr[1]).all())
    q.put((rank, bool(ok_halo), bool(ok_p)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("periodic", [False, True])
def test_two_rank_exchange_protocol_gloo(periodic):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_ring_worker, args=(r, 2, port, periodic, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_halo, ok_p in res:
        assert ok_halo, f"rank {rank}: halo columns differ from the oracle's field_mode_bc"
        assert ok_p, f"rank {rank}: particle payload exchange wrong"


# ------------------------------------------------------------------ host mirror of the round-1e entry points
class _FakeLib:
    """stands in for libcylgpu.so: records the calls of the host mirror and returns success"""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def fn(*args):
            self.calls.append((name, args))
            return 0
        return fn


def _stub_slab(move_window=False):
    d = decks.lwfa(nx=32, ny=24, ppc_e=4, ppc_p=1, window=move_window)
    s = hotpath.Slab.__new__(hotpath.Slab)
    s.L = _FakeLib()
    s.h = None
    s.grid = decomp.SlabGrid(32, 24, 1, 0, d.x_min, d.x_max, d.y_max)
    s.n_mode = 2
    s.species = decks.product_species(d)
    s.lasers = []
    s.dt, s.time, s.step = s.grid.dt, 1.5e-15, 12
    s.window_shift_fraction, s.window_started, s.window_shifts_total = 0.25, True, 7
    s.device_insert_seed, s.insert_fn = 99, None
    return s


def test_host_mirror_builds_the_sdf_descriptor_and_moment_calls():
    import ctypes as C
    s = _stub_slab()
    d = s.sdf_dump("/tmp/x.sdf", ["electron", "proton"], npart_global=[10, 4], npart_offset=[3, 1], restart=True,
                   derived=("number_density", "temperature_x", "ekflux/y_min", "number_density_mode"))
    name, args = s.L.calls[-1]
    assert name == "cylgpu_sdf_dump" and args[1] == b"/tmp/x.sdf"
    assert (d.nx_global, d.ny_global, d.n_mode, d.n_species, d.nx_local, d.cell_x_min) == (32, 24, 2, 2, 32, 1)
    assert (d.step, d.restart) == (12, 1) and d.time == 1.5e-15
    assert d.derived_mask == (1 << 3) | (1 << 10) | (1 << 20) | (1 << 22) and d.derived_sum == 1 and d.derived_species == 1
    assert [d.species_name[i] for i in range(2)] == [b"electron", b"proton"]
    assert list(d.npart_global)[:2] == [10, 4] and list(d.npart_offset)[:2] == [3, 1]
    assert d.n_constants == 5 and d.constant_id[1] == b"window_shift_fraction" and d.constant_value[1] == 0.25
    assert d.constant_value[0] == s.dt and d.constant_value[2] == s.grid.x_grid_min
    assert (d.x_min, d.dx, d.dy) == (s.grid.xb_min, s.grid.dx, s.grid.dy)
    # restart: the loader's descriptor lists the same constants; the mirror re-derives the ghosts afterwards
    s.L.calls.clear()
    s.sdf_load("/tmp/x.sdf", ["electron", "proton"])
    names = [c[0] for c in s.L.calls]
    assert names[0] == "cylgpu_sdf_load"
    assert names[1:4] == ["cylgpu_efield_bcs", "cylgpu_bfield_bcs", "cylgpu_current_finish"]
    # moments: kind / species / direction reach the C-ABI as the enum of include/cylgpu.h
    s.L.calls.clear()
    a = s.moment("species_current", 1, 3)
    name, args = s.L.calls[-1]
    assert name == "cylgpu_particle_moment" and args[1:4] == (7, 1, 3) and a.shape == (24 + 2 * NG, 32 + 2 * NG)


def test_host_mirror_device_insertion_uses_the_shift_counter_as_column():
    s = _stub_slab(move_window=True)
    s.move_window = True
    s._shift_window_once()
    ins = [c for c in s.L.calls if c[0] == "cylgpu_insert_particles_device"]
    assert len(ins) == 2                       # both species of the deck, in deck order
    for isp, (_, args) in enumerate(ins):
        assert args[1] == isp and args[9] == 99 and args[10] == 7      # seed, column = shifts made so far
    assert s.L.calls[-1][0] == "cylgpu_window_shift" and s.window_shifts_total == 8


def test_calculate_breaks_matches_the_literal_restatement(cylgpu_lib):
    """cylgpu_calculate_breaks (csrc/balance.cu) against oracle/balance_ref.py, the literal restatement of
    balance.F90:2510-2653, on uniform, ramped, lopsided and random load profiles; plus the properties the reference
    relies on: contiguous cover of 1..sz, at least ncell_min cells per slab, and a spread of the slab loads that is no
    worse than the equal-width split's on the lopsided profiles."""
    import ctypes as C
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import balance_ref as br
    L = cylgpu_lib
    rng = np.random.default_rng(11)

    def product(load, nproc):
        a = np.ascontiguousarray(load, dtype=np.int64)
        mins = (C.c_int32 * nproc)()
        maxs = (C.c_int32 * nproc)()
        assert L.cylgpu_calculate_breaks(a.ctypes.data, len(load) - 2 * br.NG, nproc, mins, maxs) == 0
        return list(mins), list(maxs)

    cases = []
    for sz in (40, 97, 256):
        x = np.arange(sz + 2 * br.NG)
        cases += [np.full(sz + 2 * br.NG, 7), 3 + x, 1 + (x > sz // 4) * 50, 1 + 400 * (np.abs(x - sz * 0.7) < 4),
                  rng.integers(0, 1000, sz + 2 * br.NG), 5 * rng.poisson(3.0, sz + 2 * br.NG) + 12]
    for load in cases:
        sz = len(load) - 2 * br.NG
        for nproc in (1, 2, 3, 8):
            if nproc * br.NCELL_MIN > sz:
                continue
            got = product(load, nproc)
            ref = br.calculate_breaks(list(load), nproc)
            assert got == (ref[0], ref[1]), (sz, nproc, got, ref)
            mins, maxs = got
            assert mins[0] == 1 and maxs[-1] == sz
            assert all(mins[p] == maxs[p - 1] + 1 for p in range(1, nproc))
            assert all(maxs[p] - mins[p] + 1 >= br.NCELL_MIN for p in range(nproc))
    # a plasma slab in the left fifth of the box: the equal split leaves most slabs empty, the breaks do not
    sz, nproc = 200, 4
    load = np.full(sz + 2 * br.NG, 50, dtype=np.int64)
    load[br.NG:br.NG + 40] += 5 * 3000
    mins, maxs = product(load, nproc)
    per = [int(load[br.NG + mins[p] - 1: br.NG + maxs[p]].sum()) for p in range(nproc)]
    equal = [int(load[br.NG + p * 50: br.NG + (p + 1) * 50].sum()) for p in range(nproc)]
    assert max(per) < 0.5 * max(equal)
