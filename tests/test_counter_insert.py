"""CPU pins of the counter-based plasma column: Philox4x32-10 known answers for
the oracle's and the product's generator, and the properties of the oracle's column that make it
a usable reference for the device kernel (decomposition invariance, statistics)."""
import ctypes as C
import math

import numpy as np

import decks
import pyoracle as po
from pyoracle import KB, M0

# Random123 1.09, examples/kat_vectors: philox4x32 10 <ctr x4> <key x2> <expected x4>
KAT = [
    ([0x00000000] * 4, [0x00000000] * 2, [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


def product_philox(lib, ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    assert lib.cylgpu_philox4x32(c, k, o) == 0
    return list(o)


def test_philox_known_answers_oracle_and_product(cylgpu_lib):
    for ctr, key, want in KAT:
        assert po.philox4x32(ctr, key) == want
        assert product_philox(cylgpu_lib, ctr, key) == want
    rng = np.random.default_rng(5)
    for _ in range(200):   # two independent restatements agree everywhere
        ctr = [int(v) for v in rng.integers(0, 2 ** 32, 4)]
        key = [int(v) for v in rng.integers(0, 2 ** 32, 2)]
        assert po.philox4x32(ctr, key) == product_philox(cylgpu_lib, ctr, key)


def _window_world(nranks, seed=1234, ppc_e=4.5):
    d = decks.lwfa(nx=32, ny=16, n_mode=1, ppc_e=ppc_e, ppc_p=0, window=True, t_centre=30e-15)
    d.species[0].temp = (2.0e5, 1.0e5, 3.0e5)
    w = decks.make_oracle(d, nranks=nranks, load=False)
    w.set_counter_insert(True, seed)
    return d, w


def _run(w, steps=12):
    w.call("init_half_step")
    w.step(steps)
    assert w.scalars()["window_shifts_total"] >= 5
    return np.concatenate([w.particles(k, 0).reshape(-1, 7) for k in range(w.nranks)])


def test_column_is_independent_of_the_decomposition():
    """no particles are loaded, so everything in the box came from insert_particles_counter: one
    slab or two must hold the same set -- weights bit for bit, phase space to the rounding of the
    field solve (the slabs sum their currents in a different order)"""
    _, w1 = _window_world(1)
    _, w2 = _window_world(2)
    a, b = decks.sort_particles(_run(w1)), decks.sort_particles(_run(w2))
    assert a.shape == b.shape and a.shape[0] > 0
    assert np.array_equal(a[:, 6], b[:, 6])
    for cols in (slice(0, 3), slice(3, 6)):
        assert np.abs(a[:, cols] - b[:, cols]).max() <= 1e-10 * np.abs(a[:, cols]).max()
    _, w3 = _window_world(1, seed=99)
    c = decks.sort_particles(_run(w3))
    assert c.shape != a.shape or not np.array_equal(a[:, 6], c[:, 6])   # the seed matters


def test_column_statistics():
    d, w = _window_world(1, ppc_e=64.25)
    p = _run(w, steps=12)
    sc = w.scalars()
    shifts = int(sc["window_shifts_total"])
    sp = d.species[0]
    # fractional ppc: every cell gets 64 or 65 particles, on average 64.25
    per_cell = p.shape[0] / (shifts * 16)
    assert 64.0 <= per_cell <= 65.0 and abs(per_cell - 64.25) < 0.2
    # the weights add up to density x the volume of the inserted columns (triangle-weighted profile = uniform)
    vol = math.pi * (16 * sc["dy"]) ** 2 * sc["dx"] * shifts
    assert abs(p[:, 6].sum() / (sp.density * vol) - 1.0) < 1e-2
    # thermal momenta: <p_i^2> = k_B T_i m per direction, zero mean, uncorrelated
    for i, T in enumerate(sp.temp):
        var = (p[:, 3 + i] ** 2).mean()
        assert abs(var / (KB * T * M0) - 1.0) < 0.05, i
        assert abs(p[:, 3 + i].mean()) < 0.05 * math.sqrt(KB * T * M0)
    cc = np.corrcoef(p[:, 3:6].T)
    assert np.abs(cc - np.eye(3)).max() < 0.05
    # positions: theta uniform, r uniform within each cell (helper.F90:571 convention)
    th = np.arctan2(p[:, 2], p[:, 1])
    assert abs(np.cos(th).mean()) < 0.03 and abs(np.sin(th).mean()) < 0.03
    r = np.hypot(p[:, 1], p[:, 2]) / sc["dy"]
    assert abs((r - np.floor(r)).mean() - 0.5) < 0.02
