"""Synthetic decks shared by the tests and bench.py (shapes follow BASELINE.json configs and
the reference's example_decks; sizes here are the small parity variants).

`make_oracle` builds the CPU oracle world (TEST INFRASTRUCTURE); `make_slabs` builds the
product `Slab`s (one per rank) and copies the oracle's initial state into them so that both
start from identical particles and fields.
"""
import math
from dataclasses import dataclass, field

import numpy as np

import cylindrical_epoch_b200 as ce
from cylindrical_epoch_b200.constants import *  # noqa: F401,F403

LAMBDA0 = 0.8e-6
OMEGA0 = 2.0 * math.pi * C_LIGHT / LAMBDA0


@dataclass
class SpeciesSpec:
    charge: float
    mass: float
    bc_particle: tuple
    ppc: float
    density: float
    temp: tuple = (0.0, 0.0, 0.0)
    drift: tuple = (0.0, 0.0, 0.0)
    immobile: bool = False
    zero_current: bool = False


@dataclass
class Deck:
    name: str
    nx: int
    ny: int
    n_mode: int
    x_min: float
    x_max: float
    y_max: float
    bc_field: tuple
    species: list
    lasers: list = field(default_factory=list)
    move_window: bool = False
    window_v_x: float = 0.0
    window_start_time: float = 0.0
    window_stop_time: float = 1e300
    bc_x_min_after_move: int = BC_SIMPLE_OUTFLOW
    bc_x_max_after_move: int = BC_SIMPLE_OUTFLOW
    dt_multiplier: float = 0.95


def laser_amp(intensity_w_cm2):
    """deck_laser_block.f90:135-139"""
    return 100.0 * math.sqrt(intensity_w_cm2 / (C_LIGHT * EPSILON0 / 2.0))


def lwfa(nx=128, ny=32, n_mode=2, ppc_e=4, ppc_p=0, window=False, t_centre=None):
    """Scaled Wakefield_Lifschitz09_JCompPhys228.deck: dx = lambda/25, dy = lambda/3-ish,
    n = 7.5e24 m^-3, a0 ~ 1.26 Gaussian pulse from x_min, open x_max / r_max."""
    dx, dy = LAMBDA0 / 25.0, LAMBDA0 / 3.0
    open4 = (BC_OPEN, BC_OPEN, BC_OPEN, BC_OPEN)
    sp = [SpeciesSpec(-Q0, M0, open4, ppc_e, 7.5e24)]
    if ppc_p:
        sp.append(SpeciesSpec(Q0, M0 * 1836.2, open4, ppc_p, 7.5e24))
    wt = 10.0e-15 if t_centre is None else t_centre / 3.0
    tc = 3.0 * wt
    las = [dict(boundary=BD_X_MIN, amp=laser_amp(3.4e18), omega=OMEGA0, t_centre=tc, t_width=wt,
                r_width=min(5.0e-6, 0.4 * ny * dy), phase=0.0, pol_angle=0.0)]
    return Deck("lwfa", nx, ny, n_mode, 0.0, nx * dx, ny * dy,
                (BC_SIMPLE_LASER, BC_OPEN, 0, BC_OPEN), sp, las,
                move_window=window, window_v_x=3.0e8, window_start_time=0.0)


def thermal(nx=64, ny=32, n_mode=2, ppc=8, temp_k=1.16e7, density=1.0e24):
    """Periodic-x thermal plasma; r_max field zero_b + particle reflect (two_stream deck style)."""
    dx = dy = 0.5e-6
    bcp = (BC_PERIODIC, BC_PERIODIC, BC_OPEN, BC_REFLECT)
    sp = [SpeciesSpec(-Q0, M0, bcp, ppc, density, temp=(temp_k, temp_k, temp_k))]
    return Deck("thermal", nx, ny, n_mode, 0.0, nx * dx, ny * dy,
                (BC_PERIODIC, BC_PERIODIC, 0, BC_ZERO_B), sp)


def drift(nx=48, ny=24, n_mode=3):
    """current_density_test.deck analogue: cold drifting beam, reflecting walls everywhere."""
    dx = dy = 1.0e-3
    bcp = (BC_REFLECT, BC_REFLECT, BC_OPEN, BC_REFLECT)
    v = 0.3 * C_LIGHT
    g = 1.0 / math.sqrt(1 - 0.09)
    sp = [SpeciesSpec(-Q0, M0, bcp, 6, 1.0e12, temp=(50.0, 50.0, 50.0), drift=(M0 * g * v, -0.5 * M0 * g * v, 0.0))]
    return Deck("drift", nx, ny, n_mode, 0.0, nx * dx, ny * dy, (BC_CLAMP, BC_CLAMP, 0, BC_CLAMP), sp)


def gaussian_pulse(nx=500, ny=100):
    """example_decks/gaussian_pulse.deck of the reference: a CW beam of 1 um light launched from x_min with the
    curved phase front and width of a Gaussian beam that focuses 10 um into the box to a 1.5 um FWHM spot of
    1e15 W/cm^2; vacuum, m = 0..1, open x_max / r_max.  The deck's own constants block is restated here."""
    lam, i_fwhm, i_peak, foc = 1.0e-6, 1.5e-6, 1.0e15, 10.0e-6
    k = 2.0 * math.pi / lam
    w0 = i_fwhm / math.sqrt(2.0 * math.log(2.0))
    zr = math.pi * w0 ** 2 / lam
    w_b = w0 * math.sqrt(1.0 + (foc / zr) ** 2)
    i_b = i_peak * (w0 / w_b) ** 2
    rc = foc * (1.0 + (zr / foc) ** 2)
    gouy = math.atan(-foc / rc)
    las = [dict(boundary=BD_X_MIN, amp=laser_amp(i_b), omega=2.0 * math.pi * C_LIGHT / lam, r_width=w_b,
                phase=-gouy, phase_curv=k / (2.0 * rc), pol_angle=0.0)]
    d = Deck("gaussian_pulse", nx, ny, 2, 0.0, 20.0e-6, 5.0e-6, (BC_SIMPLE_LASER, BC_OPEN, 0, BC_OPEN), [], las)
    d.expect = dict(w0=w0, e_peak=math.sqrt(2.0 * i_peak * 1.0e4 / (EPSILON0 * C_LIGHT)), focus=foc, rayleigh=zr)
    return d


def make_oracle(deck, nranks=1, load=True):
    import pyoracle as po
    w = po.OracleWorld(deck.nx, deck.ny, deck.n_mode, deck.x_min, deck.x_max, deck.y_max, list(deck.bc_field),
                       nranks=nranks, dt_multiplier=deck.dt_multiplier, move_window=deck.move_window,
                       window_v_x=deck.window_v_x, window_start_time=deck.window_start_time,
                       window_stop_time=deck.window_stop_time, bc_x_min_after_move=deck.bc_x_min_after_move,
                       bc_x_max_after_move=deck.bc_x_max_after_move)
    for s in deck.species:
        w.add_species(s.charge, s.mass, list(s.bc_particle), ppc=s.ppc, density=s.density, temp=s.temp,
                      drift=s.drift, immobile=s.immobile, zero_current=s.zero_current)
    for L in deck.lasers:
        w.add_laser(**L)
    if load:
        for i in range(len(deck.species)):
            w.load_uniform(i)
    return w


def product_species(deck):
    return [ce.Species(s.charge, s.mass, tuple(s.bc_particle), s.immobile, s.zero_current, s.ppc, s.density,
                       tuple(s.temp), tuple(s.drift)) for s in deck.species]


def make_slab(deck, rank=0, nranks=1, **kw):
    lasers = [ce.Laser(**L) for L in deck.lasers]
    return ce.Slab(deck.nx, deck.ny, deck.n_mode, deck.x_min, deck.x_max, deck.y_max, list(deck.bc_field),
                   product_species(deck), rank=rank, nranks=nranks, dt_multiplier=deck.dt_multiplier,
                   lasers=lasers, move_window=deck.move_window, window_v_x=deck.window_v_x,
                   window_start_time=deck.window_start_time, window_stop_time=deck.window_stop_time,
                   bc_x_min_after_move=deck.bc_x_min_after_move, bc_x_max_after_move=deck.bc_x_max_after_move,
                   **kw)


def copy_state(oracle, slab, k=0):
    """oracle rank k -> product slab: all 15 mode arrays, 12 snapshots, particles, time."""
    for name in FIELD_NAMES:
        slab.upload_field(name, oracle.field(k, name))
    for name in SNAP_NAMES:
        slab.upload_snapshot(name, oracle.field(k, name))
    for isp in range(oracle.n_species):
        slab.upload_particles(isp, oracle.particles(k, isp))
    sc = oracle.scalars()
    slab.time = sc["time"]
    slab.step = int(sc["step"])


def rel_err(a, b):
    """max |a-b| / max |b| (0 when both vanish)."""
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.abs(b).max() if b.size else 0.0
    num = np.abs(a - b).max() if b.size else 0.0
    if den == 0.0:
        return 0.0 if num == 0.0 else float("inf")
    return float(num / den)


def sort_particles(aos):
    """canonical order for comparing particle sets whose storage order differs."""
    a = np.asarray(aos).reshape(-1, 7)
    idx = np.lexsort((a[:, 6], a[:, 2], a[:, 1], a[:, 0]))
    return a[idx]
