"""Pins for the oracle's calc_df.F90 moments (oracle/cyl_moments.cpp), CPU only.

The reference ships no fixtures for these diagnostics, so the restatement is pinned by
 (i)  an independent vectorised numpy restatement (different summation order, 1e-12),
 (ii) physical known answers of a uniform thermal / drifting load (density, temperature,
      mean kinetic energy, current, mean momentum), and
 (iii) decomposition invariance: two x-slabs give the one-slab answer on every interior cell.
"""
import math

import numpy as np
import pytest

import decks
import pyoracle as po
import pyoracle
from pyoracle import NG, BC_PERIODIC, BC_REFLECT, BD_X_MIN, BD_X_MAX, BD_Y_MAX, C_LIGHT, KB, M0, Q0

C_TINY = np.finfo(np.float64).tiny


def _to_grid(parts, xg, yg, dx, dy):
    """include/particle_to_grid.inc + triangle/gxfac.inc, vectorised"""
    r = np.sqrt(parts[:, 1] ** 2 + parts[:, 2] ** 2)
    cxr = (parts[:, 0] - xg) / dx
    cyr = (r - yg) / dy
    cx = np.floor(cxr + 0.5).astype(np.int64)
    cy = np.floor(cyr + 0.5).astype(np.int64)
    fx = cx - cxr
    fy = cy - cyr
    gx = np.stack([0.5 * (0.25 + fx * fx + fx), 0.75 - fx * fx, 0.5 * (0.25 + fx * fx - fx)], axis=1)
    gy = np.stack([0.5 * (0.25 + fy * fy + fy), 0.75 - fy * fy, 0.5 * (0.25 + fy * fy - fy)], axis=1)
    fold = r < dy
    gy[fold, 1] += gy[fold, 0]
    gy[fold, 0] = 0.0
    return cx + 1, cy + 1, gx, gy, r


def _scatter(shape, cx, cy, gx, gy, val):
    a = np.zeros(shape)
    for iy in range(3):
        for ix in range(3):
            np.add.at(a, (cy + iy - 1 + NG - 1, cx + ix - 1 + NG - 1), gx[:, ix] * gy[:, iy] * val)
    return a


def _sum_bcs(a, nx, ny, bcp):
    """processor_summation_bcs on one rank (boundary.F90:833-914 real variant, :1019-1129)"""
    o = NG - 1   # array index of Fortran index 0
    if bcp[BD_X_MIN] == BC_REFLECT:
        for i in range(1, NG):
            a[:, o + i] += a[:, o + 1 - i]
            a[:, o + 1 - i] = 0.0
    if bcp[BD_X_MAX] == BC_REFLECT:
        for i in range(1, NG + 1):
            a[:, o + nx + 1 - i] += a[:, o + nx + i]
            a[:, o + nx + i] = 0.0
    if bcp[BD_Y_MAX] == BC_REFLECT:
        for i in range(1, NG + 1):
            a[o + ny + 1 - i, :] += a[o + ny + i, :]
            a[o + ny + i, :] = 0.0
    if bcp[BD_X_MIN] == BC_PERIODIC:
        lo = a[:, o + 1 - NG:o + 1].copy()          # ghosts 1-ng..0
        hi = a[:, o + nx + 1:o + nx + NG + 1].copy()  # ghosts nx+1..nx+ng
        a[:, o + 1:o + NG + 1] += hi
        a[:, o + nx + 1 - NG:o + nx + 1] += lo
    return a


def _zero_gradient(a, nx, ny, bcf):
    o = NG - 1
    if bcf[BD_X_MIN] != BC_PERIODIC:
        for i in range(1, NG + 1):
            a[:, o + i - NG] = a[:, o + NG + 1 - i]
    if bcf[BD_X_MAX] != BC_PERIODIC:
        for i in range(1, NG + 1):
            a[:, o + nx + i] = a[:, o + nx + 1 - i]
    for i in range(1, NG + 1):   # r_min is never periodic
        a[o + i - NG, :] = a[o + NG + 1 - i, :]
    if bcf[BD_Y_MAX] != BC_PERIODIC:
        for i in range(1, NG + 1):
            a[o + ny + i, :] = a[o + ny + 1 - i, :]
    return a


def _halo_periodic(a, nx):
    o = NG - 1
    a[:, o + nx + 1:o + nx + NG + 1] = a[:, o + 1:o + NG + 1]
    a[:, o + 1 - NG:o + 1] = a[:, o + nx + 1 - NG:o + nx + 1]
    return a


def numpy_moment(w, deck, kind, isp, direction):
    """independent restatement for one rank, one species"""
    sc = w.scalars()
    dx, dy = sc["dx"], sc["dy"]
    info = w.rank_info(0)
    nx, ny = info["nx"], info["ny"]
    xg, yg = info["x_grid_min_local"], sc["y_grid_min_local"]
    parts = w.particles(0, isp).reshape(-1, 7)
    spec = deck.species[isp]
    bcp, bcf = w.bc_particle(isp), w.bc_field()
    shape = (ny + 2 * NG, nx + 2 * NG)
    cx, cy, gx, gy, r = _to_grid(parts, xg, yg, dx, dy)
    wgt = parts[:, 6]
    p = parts[:, 3:6]
    vol = 2.0 * math.pi * dx * dy * r
    if kind in ("ppc", "average_weight"):
        ccx = np.floor((parts[:, 0] - xg) / dx + 0.5).astype(np.int64) + 1
        ccy = np.floor((r - yg) / dy + 0.5).astype(np.int64) + 1
        cnt = np.zeros(shape)
        np.add.at(cnt, (ccy + NG - 1, ccx + NG - 1), 1.0)
        if kind == "ppc":
            return cnt
        a = np.zeros(shape)
        np.add.at(a, (ccy + NG - 1, ccx + NG - 1), wgt)
        return a / np.maximum(cnt, C_TINY)
    if kind == "temperature":
        pm = p / math.sqrt(spec.mass)
        dirs = [direction - 1] if direction > 0 else [0, 1, 2]
        cnt = _sum_bcs(_scatter(shape, cx, cy, gx, gy, wgt), nx, ny, bcp)
        cnt = np.maximum(cnt, 1e-6)
        means = {}
        for d in dirs:
            m = _sum_bcs(_scatter(shape, cx, cy, gx, gy, wgt * pm[:, d]), nx, ny, bcp) / cnt
            if bcf[BD_X_MIN] == BC_PERIODIC:
                m = _halo_periodic(m, nx)
            means[d] = m
        sig = np.zeros(shape)
        cnt2 = np.zeros(shape)
        for iy in range(3):
            for ix in range(3):
                jj, ii = cy + iy - 1 + NG - 1, cx + ix - 1 + NG - 1
                gf = gx[:, ix] * gy[:, iy]
                wd = sum((pm[:, d] - means[d][jj, ii]) ** 2 for d in dirs)
                np.add.at(sig, (jj, ii), gf * wd)
                np.add.at(cnt2, (jj, ii), gf)
        sig = _sum_bcs(sig, nx, ny, bcp)
        cnt2 = _sum_bcs(cnt2, nx, ny, bcp)
        return sig / np.maximum(cnt2, 1e-6) / KB / (1.0 if direction > 0 else 3.0)
    averaged = False
    if kind == "mass_density":
        val = spec.mass * wgt / vol
    elif kind == "number_density":
        val = wgt / vol
    elif kind == "species_current":
        mc = C_LIGHT * spec.mass
        val = spec.charge * wgt * p[:, direction - 1] / np.sqrt(mc * mc + (p ** 2).sum(axis=1)) * C_LIGHT / vol
    elif kind in ("ekbar", "ekflux"):
        averaged = True
        mc = C_LIGHT * spec.mass
        u = p / mc
        u2 = (u ** 2).sum(axis=1)
        gam = np.sqrt(u2 + 1.0)
        val = u2 / (gam + 1.0) * (mc * wgt * C_LIGHT)
        if kind == "ekflux":
            d = abs(direction) - 1
            fac = [C_LIGHT * dy, C_LIGHT * dx, C_LIGHT * dx * dy][d]
            flux = fac * u[:, d] / gam
            val = val * np.maximum(flux, 0.0) if direction > 0 else -val * np.minimum(flux, 0.0)
    elif kind == "average_momentum":
        averaged = True
        val = wgt * p[:, direction - 1]
    else:
        raise ValueError(kind)
    a = _sum_bcs(_scatter(shape, cx, cy, gx, gy, val), nx, ny, bcp)
    if averaged:
        wt = _sum_bcs(_scatter(shape, cx, cy, gx, gy, wgt), nx, ny, bcp)
        a = a / np.maximum(wt, C_TINY)
    return _zero_gradient(a, nx, ny, bcf)


CASES = [("mass_density", 0), ("number_density", 0), ("ekbar", 0), ("ekflux", 1), ("ekflux", -2), ("ekflux", 3),
         ("ppc", 0), ("average_weight", 0), ("temperature", 0), ("temperature", 2), ("species_current", 1),
         ("species_current", 3), ("average_momentum", 2)]


def _deck(name):
    if name == "thermal":
        return decks.thermal(nx=24, ny=12, n_mode=2, ppc=6)
    if name == "drift":
        return decks.drift(nx=20, ny=10, n_mode=2)
    return decks.lwfa(nx=24, ny=10, n_mode=2, ppc_e=4, ppc_p=2)


TRIANGLE_ONLY = pytest.mark.skipif(pyoracle.SHAPE != "triangle", reason="the numpy restatement in this file is the triangle's")


@TRIANGLE_ONLY
@pytest.mark.parametrize("deck_name", ["thermal", "drift", "lwfa"])
def test_moments_match_independent_numpy_restatement(deck_name):
    d = _deck(deck_name)
    w = decks.make_oracle(d)
    w.call("init_half_step")
    w.step(3)   # particles off their load positions, some in the ghost cells / reflected
    for kind, direction in CASES:
        for isp in range(len(d.species)):
            got = w.moment(kind, isp, direction)[0]
            want = numpy_moment(w, d, kind, isp, direction)
            scale = np.abs(want).max()
            assert scale > 0.0, (kind, direction)
            assert np.abs(got - want).max() <= 1e-12 * scale, (deck_name, kind, direction, isp)


def test_species_sum_skips_tracers_and_adds_the_rest():
    d = decks.lwfa(nx=24, ny=10, n_mode=2, ppc_e=4, ppc_p=2)
    w = decks.make_oracle(d)
    tot = w.moment("mass_density", -1)[0]
    parts = w.moment("mass_density", 0)[0] + w.moment("mass_density", 1)[0]
    assert np.abs(tot - parts).max() <= 1e-13 * np.abs(tot).max()
    # the real-valued number density is the m = 0 real part of calc_number_density_modes
    nd = w.moment("number_density", 0)[0]
    ndm = w.number_density_modes(0)[0][0].real
    assert np.array_equal(nd, ndm)
    # charge density = charge * number density (same deposit, calc_df.F90:479 vs :566)
    rho = w.charge_density(0)[0]
    assert np.abs(rho - d.species[0].charge * nd).max() <= 1e-13 * np.abs(rho).max()


@TRIANGLE_ONLY
def test_known_answers_of_a_uniform_thermal_load():
    T, n0 = 1.16e7, 1.0e24
    d = decks.thermal(nx=48, ny=24, n_mode=1, ppc=64, temp_k=T, density=n0)
    w = decks.make_oracle(d)
    inner = (slice(NG + 2, NG + 24 - 2), slice(NG, NG + 48))
    nd = w.moment("number_density", 0)[0][inner]
    assert abs(nd.mean() / n0 - 1.0) < 0.02
    # example_decks/uniform_density_load.deck: a uniform load must show a uniform number_density down to the
    # axis -- the r < dy fold of the weights (triangle/gxfac.inc) and the 2 pi r macro-particle volume
    # (partlist.F90:999-1013) have to cancel; row means over 48 columns, axis row included
    rows = w.moment("number_density", 0)[0][NG:NG + 22, NG:NG + 48].mean(axis=1)
    assert np.abs(rows / n0 - 1.0).max() < 0.03, rows / n0
    rho_m = w.moment("mass_density", 0)[0][inner]
    assert np.abs(rho_m - M0 * nd).max() <= 1e-13 * rho_m.max()
    # every cell holds ppc particles at load time (helper.F90:552-583), weights sum to n0 * cell volume
    ppc = w.moment("ppc", 0)[0]
    assert ppc.sum() == 48 * 24 * 64
    assert np.all(ppc[NG:NG + 24, NG:NG + 48] == 64.0)
    aw = w.moment("average_weight", 0)[0]
    sc = w.scalars()
    rr = (np.arange(24) + 0.5) * sc["dy"]
    # weights follow n * macro-particle volume / ppc with the loader's own deposit normalisation: within a few %
    cellvol = 2.0 * math.pi * rr * sc["dx"] * sc["dy"]
    assert np.allclose(aw[NG + 2:NG + 22, NG + 5] * 64 / (n0 * cellvol[2:22]), 1.0, rtol=0.05)
    # temperature: unbiased towards T(1 - 1/N_eff); 64 ppc * 9 cells -> a few % statistical
    for direction in (0, 1, 2, 3):
        t = w.moment("temperature", 0, direction)[0][inner]
        assert abs(t.mean() / T - 1.0) < 0.05, direction
    # <(gamma - 1) m c^2> = 3/2 k T (1 + O(kT / mc^2)), kT/mc^2 = 2e-3
    ek = w.moment("ekbar", 0)[0][inner]
    assert abs(ek.mean() / (1.5 * KB * T) - 1.0) < 0.05
    # forward + backward energy flux are equal in a drift-free load
    fp = w.moment("ekflux", 0, 1)[0][inner].mean()
    fm = w.moment("ekflux", 0, -1)[0][inner].mean()
    assert fp > 0.0 and abs(fp / fm - 1.0) < 0.1


def test_known_answers_of_a_cold_drifting_beam():
    d = decks.drift(nx=32, ny=16, n_mode=1)
    w = decks.make_oracle(d)
    sp = d.species[0]
    inner = (slice(NG + 2, NG + 14), slice(NG + 2, NG + 30))
    g = 1.0 / math.sqrt(1 - 0.09)
    v = 0.3 * C_LIGHT
    nd = w.moment("number_density", 0)[0][inner]
    jx = w.moment("species_current", 0, 1)[0][inner]
    jy = w.moment("species_current", 0, 2)[0][inner]
    # p = (g m v, -g m v / 2, 0): gamma from |p|, v_x = p_x / (gamma m)
    px, py = sp.drift[0], sp.drift[1]
    gam = math.sqrt(1.0 + (px * px + py * py) / (M0 * C_LIGHT) ** 2)
    assert np.allclose(jx / (sp.charge * nd * px / (gam * M0)), 1.0, rtol=1e-3)
    assert np.allclose(jy / (sp.charge * nd * py / (gam * M0)), 1.0, rtol=1e-3)
    pm = w.moment("average_momentum", 0, 1)[0][inner]
    assert np.allclose(pm / px, 1.0, rtol=1e-3)
    assert abs(px / (M0 * g * v) - 1.0) < 1e-12


@pytest.mark.skipif(pyoracle.SHAPE == "tophat", reason="top-hat: calc_ppc counts into ghost cells that nothing sums "
                    "back (calc_df.F90:696-706 has no calc_boundary), so it depends on the decomposition in the reference too")
@pytest.mark.parametrize("deck_name", ["thermal", "lwfa"])
def test_two_slabs_give_the_one_slab_answer(deck_name):
    d = _deck(deck_name)
    w1 = decks.make_oracle(d, nranks=1)
    w2 = decks.make_oracle(d, nranks=2, load=False)
    # same particles: split rank 0's list of the one-slab world by x
    n0 = w2.rank_info(0)
    for isp in range(len(d.species)):
        p = w1.particles(0, isp).reshape(-1, 7)
        left = p[:, 0] < n0["x_max_local"]
        w2.set_particles(0, isp, p[left])
        w2.set_particles(1, isp, p[~left])
    nx0 = n0["nx"]
    for kind, direction in CASES:
        a1 = w1.moment(kind, 0, direction)[0]
        a2 = w2.moment(kind, 0, direction)
        joined = np.concatenate([a2[0][NG:-NG, NG:NG + nx0], a2[1][NG:-NG, NG:-NG]], axis=1)
        ref = a1[NG:-NG, NG:-NG]
        assert np.abs(joined - ref).max() <= 1e-12 * np.abs(ref).max(), (kind, direction)
