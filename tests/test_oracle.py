"""CPU tests that pin the oracle (the reference ships no golden vectors for the cylindrical
path -- SURVEY.md section 4/8c -- so the anchors are analytic and structural):
  * current_density_test.deck known answer (DOCUMENTATION.pdf section 7.2),
  * exact discrete charge continuity of the mode-0 deposit,
  * axis-condition identities of update_e/b_field,
  * vacuum propagation of an injected pulse at c,
  * KISS / Box-Muller sanity, loader statistics,
  * committed golden vectors (tests/golden) guarding the oracle against drift."""
import math
import os

import numpy as np
import pytest

import decks
from cylindrical_epoch_b200.constants import *  # noqa: F401,F403

NGH = 5


def fidx(ix, ir):
    """(ix, ir) Fortran indices -> numpy [ir, ix] offsets of one mode plane"""
    return ir + NGH - 1, ix + NGH - 1


def test_current_density_known_answer():
    """uniform n = 1e5 m^-3 electrons with p = (-1.23e20, +1.23e20, 0): ultra-relativistic, v =
    c(-1,1,0)/sqrt(2), so J_x = +3.40e-6, J_y = -3.40e-6 A/m^2; Jx only in m=0, J_perp only in
    m=1 with Jr_1 = J_y and Jtheta_1 = -i J_y (doc section 7.2)."""
    open4 = (BC_OPEN,) * 4
    sp = [decks.SpeciesSpec(-Q0, M0, open4, 36, 1.0e5, drift=(-1.23e20, 1.23e20, 0.0))]
    d = decks.Deck("cdt", 40, 20, 2, 0.0, 20e-6 * 40 / 500, 5e-6 * 20 / 100,
                   (BC_SIMPLE_LASER, BC_OPEN, 0, BC_OPEN), sp)
    w = decks.make_oracle(d)
    w.call("push_no_bcs")
    J = Q0 * 1.0e5 * C_LIGHT / math.sqrt(2.0)
    assert abs(J - 3.40e-6) < 0.01e-6
    inner = (slice(NGH + 4, NGH + 16), slice(NGH + 5, NGH + 35))
    jx0 = w.field(0, "jxm")[0][inner]
    jr1 = w.field(0, "jrm")[1][inner]
    jt1 = w.field(0, "jtm")[1][inner]
    assert abs(jx0.real.mean() / J - 1.0) < 0.03
    assert abs(jr1.real.mean() / (-J) - 1.0) < 0.05
    assert abs(jt1.imag.mean() / J - 1.0) < 0.05          # -i * J_y = +i J
    # parity selection: Jx has no m=1 content, J_perp no m=0 content (statistical noise only)
    assert abs(w.field(0, "jxm")[1][inner].mean()) < 0.05 * J
    assert abs(w.field(0, "jrm")[0][inner].mean()) < 0.05 * J
    assert abs(jx0.imag).max() == 0.0


def _tables(ny, dx, dy):
    """particles.F90:190-217 restated in numpy for the continuity check"""
    idx = np.arange(-NGH, ny + NGH + 1)
    r_low = dy / 2 - NGH * dy + (idx - (1 - NGH)) * dy      # r_low used at index iy
    area_rt = np.pi * np.abs((r_low + dy) ** 2 - r_low ** 2)
    on_axis = np.rint(2 * r_low / dy) == -1
    area_rt[on_axis] = np.pi * (0.5 * dy) ** 2
    area_xt = 2 * np.pi * np.abs(r_low + dy) * dx
    return {int(i): (a, b) for i, a, b in zip(idx, area_rt, area_xt)}


def test_mode0_deposit_is_exactly_charge_conserving():
    """A_rt(cy) [Jx(cx+1,cy) - Jx(cx,cy)] + A_xt(cy) Jr(cx,cy+1) - A_xt(cy-1) Jr(cx,cy)
       = -(Q_new - Q_old)/dt per node, with Q = q w fac gx gy evaluated at t+1/2 and t+3/2."""
    open4 = (BC_OPEN,) * 4
    sp = [decks.SpeciesSpec(-Q0, M0, open4, 0, 1.0)]
    nx, ny = 24, 24
    d = decks.Deck("cont", nx, ny, 2, 0.0, nx * 1e-6, ny * 1e-6, (BC_CLAMP, BC_CLAMP, 0, BC_CLAMP), sp)
    w = decks.make_oracle(d, load=False)
    sc = w.scalars()
    dx, dy, dt = sc["dx"], sc["dy"], sc["dt"]
    rng = np.random.default_rng(3)
    npart = 40
    p = np.zeros((npart, 7))
    p[:, 0] = rng.uniform(8e-6, 16e-6, npart)
    r = rng.uniform(8e-6, 16e-6, npart)
    th = rng.uniform(0, 2 * np.pi, npart)
    p[:, 1], p[:, 2] = r * np.cos(th), r * np.sin(th)
    p[:, 3:6] = rng.normal(0, 2.0, (npart, 3)) * M0 * C_LIGHT      # relativistic: crosses cells
    p[:, 6] = rng.uniform(1, 2, npart) * 1e3
    w.set_particles(0, 0, p)
    w.call("push_no_bcs")
    after = w.particles(0, 0)
    u = after[:, 3:6] / (M0 * C_LIGHT)
    delta = u * (C_LIGHT * dt / 2.0) / np.sqrt(1 + (u * u).sum(1))[:, None]
    tabs = _tables(ny, dx, dy)

    def charge(pos):
        Qg = np.zeros((ny + 2 * NGH, nx + 2 * NGH))
        xr = (pos[:, 0] - sc["x_grid_min"]) / dx
        rr = (np.hypot(pos[:, 1], pos[:, 2]) - sc["y_grid_min_local"]) / dy
        for k in range(npart):
            out = []
            for c_r in (xr[k], rr[k]):
                c2 = math.floor(c_r)
                f = c2 - c_r + 0.5
                out.append((c2 + 1, [0.25 + f * f + f, 1.5 - 2 * f * f, 0.25 + f * f - f]))
            (cx2, wx), (cy2, wy) = out
            for a in range(3):
                for b in range(3):
                    Qg[fidx(cx2 - 1 + b, cy2 - 1 + a)] += -Q0 * after[k, 6] * 0.25 * wx[b] * wy[a]
        return Qg

    dQ = (charge(after[:, 0:3] + delta) - charge(after[:, 0:3] - delta)) / dt
    jx = w.field(0, "jxm")[0].real
    jr = w.field(0, "jrm")[0].real
    div = np.zeros_like(dQ)
    for cy in range(3, ny - 2):
        a_rt, a_xt = tabs[cy]
        a_xt_m = tabs[cy - 1][1]
        for cx in range(3, nx - 2):
            div[fidx(cx, cy)] = (a_rt * (jx[fidx(cx + 1, cy)] - jx[fidx(cx, cy)])
                                 + a_xt * jr[fidx(cx, cy + 1)] - a_xt_m * jr[fidx(cx, cy)])
    scale = np.abs(dQ).max()
    assert scale > 0
    assert np.abs(div + dQ).max() < 1e-12 * scale
    assert np.abs(w.field(0, "jxm")[0].imag).max() == 0.0


def test_axis_identities_after_field_updates():
    d = decks.lwfa(nx=40, ny=20, n_mode=4, ppc_e=2)
    w = decks.make_oracle(d)
    w.call("init_half_step")
    w.step(15)
    w.call("update_e")
    ex, er, et = (w.field(0, n) for n in ("exm", "erm", "etm"))
    r0, r1, r2 = NGH - 1, NGH, NGH + 1
    assert np.all(et[0, r0] == 0) and np.array_equal(er[0, r0], -er[0, r1])
    assert np.all(ex[1, r0] == 0)
    np.testing.assert_allclose(et[1, r0], -1j / 8 * (9 * er[1, r1] - er[1, r2]), rtol=1e-15)
    for m in (2, 3):
        assert np.all(ex[m, r0] == 0) and np.all(et[m, r0] == 0)
        np.testing.assert_allclose(er[m, r1], er[m, r2] / 9.0, rtol=1e-15)
    # mirror parity below the axis: Ex even for even m, odd for odd m
    for m in range(4):
        sgn = 1.0 if m % 2 == 0 else -1.0
        for k in range(1, NGH):
            assert np.array_equal(ex[m, r0 - k], sgn * ex[m, r0 + k])
            assert np.array_equal(er[m, r0 - k], -sgn * er[m, r0 + k + 1])
    w.call("update_b")
    bx, br, bt = (w.field(0, n) for n in ("bxm", "brm", "btm"))
    assert np.all(br[0, r0] == 0) and np.array_equal(bx[0, r0], bx[0, r1]) and np.array_equal(bt[0, r0], -bt[0, r1])
    np.testing.assert_allclose(bt[1, r0], -2j * br[1, r0] - bt[1, r1], rtol=1e-15)


def test_vacuum_pulse_propagates_at_c():
    d = decks.lwfa(nx=400, ny=24, n_mode=2, ppc_e=0, t_centre=12e-15)
    d.species = []
    w = decks.make_oracle(d)
    w.call("init_half_step")
    pos = []
    for n in (420, 520):
        w.step(n - int(w.scalars()["step"]))
        env = np.abs(w.field(0, "etm")[1, NGH + 2, NGH:NGH + 400])
        xs = (np.arange(400) + 0.5) * w.scalars()["dx"]
        pos.append(((env ** 2 * xs).sum() / (env ** 2).sum(), w.scalars()["time"]))
    v = (pos[1][0] - pos[0][0]) / (pos[1][1] - pos[0][1])
    assert abs(v / C_LIGHT - 1.0) < 0.02, v / C_LIGHT


def test_rng_and_loader_statistics():
    import pyoracle
    d = decks.thermal(nx=32, ny=16, ppc=16, temp_k=1.0e7)
    w = decks.make_oracle(d)
    u = np.array([pyoracle.lib().cylo_rng_uniform(w.h, 0) for _ in range(20000)])
    assert 0.0 <= u.min() and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.01 and abs(u.var() - 1 / 12) < 0.003
    p = w.particles(0, 0)
    assert p.shape == (32 * 16 * 16, 7)
    sd = math.sqrt(1.0e7 * KB * M0)
    for k in (3, 4, 5):
        assert abs(p[:, k].std() / sd - 1.0) < 0.03 and abs(p[:, k].mean()) < 0.05 * sd
    # weights sum to n * volume of the cylinder (helper.F90:771-778)
    vol = math.pi * d.y_max ** 2 * (d.x_max - d.x_min)
    assert abs(p[:, 6].sum() / (1.0e24 * vol) - 1.0) < 0.02
    # two ranks reseed with 7842432 + rank: different streams
    w2 = decks.make_oracle(d, nranks=2)
    assert not np.array_equal(w2.particles(0, 0)[:50, 1], w2.particles(1, 0)[:50, 1])


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lwfa_48x16_m2_20steps.npz")


def test_oracle_reproduces_golden_vectors():
    """fixture made by tests/golden/make_golden.py from the oracle itself: guards the restatement
    (and its compiler flags) against drift; the GPU suite checks the CUDA path against the same file"""
    from golden.make_golden import run_case
    ref = np.load(GOLDEN)
    got = run_case()
    for k in ref.files:
        a, b = ref[k], got[k]
        assert a.shape == b.shape, k
        den = np.abs(a).max()
        assert np.abs(a - b).max() <= 1e-12 * max(den, 1e-300), k
