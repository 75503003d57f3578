"""The two particle-exchange protocols of the library give the same run (include/cylgpu.h
cylgpu_set_exchange_capacity): device-resident counts with one fixed-size migration message per neighbour (the
default of hotpath.Slab, what every other GPU test runs) and the exact count-then-data protocol of
partlist.F90:842,869 with its two host syncs per species per step."""
import numpy as np
import pytest

import decks
from parity import Pair, TOL_HOT

pytestmark = pytest.mark.gpu

DECKS = {"lwfa": lambda: decks.lwfa(nx=96, ny=32, n_mode=2, ppc_e=4, ppc_p=1),
         "thermal": lambda: decks.thermal(nx=64, ny=32, n_mode=2, ppc=8),
         "window": lambda: decks.lwfa(nx=64, ny=24, n_mode=2, ppc_e=4, ppc_p=1, window=True, t_centre=30e-15)}


@pytest.mark.parametrize("deckname,nranks", [("lwfa", 1), ("thermal", 1), ("thermal", 2), ("lwfa", 2), ("window", 1),
                                             ("window", 2)])
def test_exact_protocol_gives_the_same_run(deckname, nranks):
    d = DECKS[deckname]()
    p = Pair(d, nranks=nranks, slab_kw=dict(exchange_capacity=0))
    try:
        tol = TOL_HOT if deckname == "thermal" else 1e-9
        for _ in range(3):
            p.step(10)
            p.check_counts()
            p.check_fields(tol)
            p.check_particles(tol)
        p.check_cells()
    finally:
        p.close()


@pytest.mark.parametrize("deckname,nranks", [("thermal", 2), ("window", 2)])
def test_switching_protocols_mid_run(deckname, nranks):
    d = DECKS[deckname]()
    p = Pair(d, nranks=nranks)
    try:
        tol = TOL_HOT if deckname == "thermal" else 1e-9
        cap = p.slabs[0].exchange_capacity
        assert cap > 0
        for k in range(4):
            p.each(lambda s: s.set_exchange_capacity(0 if k % 2 else cap))
            p.step(6)
            p.check_counts()
            p.check_particles(tol)
        p.check_fields(tol)
    finally:
        p.close()


def test_overflow_of_the_fixed_size_message_is_reported():
    """more leavers towards one neighbour than the message holds: a sticky error, not silent loss (one periodic slab:
    its own neighbour on both sides, so no second rank can be left waiting in an exchange)"""
    import cylindrical_epoch_b200 as ce
    d = decks.thermal(nx=64, ny=32, n_mode=1, ppc=8, temp_k=5.0e9)
    p = Pair(d, nranks=1, slab_kw=dict(exchange_capacity=4))
    try:
        with pytest.raises(ce.hotpath.CylGpuError, match="exchange overflow"):
            for _ in range(3):
                p.slabs[0].step_once()
            p.slabs[0].particle_count(0)
    finally:
        p.close()


@pytest.mark.parametrize("deckname,nranks", [("window", 1), ("window", 2), ("thermal", 1)])
def test_loop_body_one_entry_point_at_a_time(deckname, nranks):
    """every other GPU test runs whole steps inside the library (csrc/driver.cu, cylgpu_driver_step); this one drives
    the same steps one entry point at a time from the host mirror, as the Fortran driver of INTEGRATION.md would
    (fields_half, push, current_finish, fields_final with host-evaluated laser sources, the window logic, insertion,
    window_shift, particle_bcs): same run"""
    d = DECKS[deckname]()
    p = Pair(d, nranks=nranks, slab_kw=dict(native_driver=False))
    try:
        assert not p.slabs[0].native
        tol = TOL_HOT if deckname == "thermal" else 1e-9
        for _ in range(3):
            p.step(10)
            p.check_counts()
            p.check_fields(tol)
            p.check_particles(tol)
        if deckname == "window":
            assert p.slabs[0].window_shifts_total >= 15
            assert p.slabs[-1].rng_get_state() == p.oracle.rng_state(nranks - 1)
    finally:
        p.close()
