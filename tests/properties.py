"""Size-independent properties of the hot path, vectorised so that they can be evaluated at
BASELINE.json's full sizes (tens of millions of particles) in seconds of numpy time.  The same
functions are checked against the oracle at small sizes in tests/test_properties_cpu.py, so the
GPU tests at full size rest on validated checkers."""
import numpy as np

from cylindrical_epoch_b200.constants import C_LIGHT, EPSILON0

from cylindrical_epoch_b200.constants import NG, SHAPE  # ng = png + 2 and the particle shape of the build in use (CYL_SHAPE)


def area_tables(ny, dx, dy):
    """particles.F90:190-217: face areas of the staggered control volumes, index iy = -NG .. ny+NG
    -> arrays offset by NG (area_rt[iy + NG]); the volume centred on the axis is a disc (:199-201)"""
    idx = np.arange(-NG, ny + NG + 1)
    r_low = dy / 2 - NG * dy + (idx - (1 - NG)) * dy
    area_rt = np.pi * np.abs((r_low + dy) ** 2 - r_low ** 2)
    area_rt[np.rint(2 * r_low / dy) == -1] = np.pi * (0.5 * dy) ** 2
    area_xt = 2 * np.pi * np.abs(r_low + dy) * dx
    return area_rt, area_xt


def stag_weights(c_r):
    """staggered shape weights of one direction with their factor (<shape>/hx_dcell.inc; particles.F90:145-153) for
    the particle shape of the build in use: (first node index, tuple of weight arrays)"""
    if SHAPE == "tophat":
        c_r = c_r - 0.5
    c2 = np.floor(c_r)
    f = c2 - c_r + 0.5
    first = c2.astype(np.int64) + 1
    if SHAPE == "tophat":
        return first, (0.5 + f, 0.5 - f)
    f2 = f * f
    if SHAPE == "bspline3":
        w = ((0.5 + f) ** 4, 4.75 + 11.0 * f + 4.0 * f2 * (1.5 - f - f2), 14.375 + 6.0 * f2 * (f2 - 2.5),
             4.75 - 11.0 * f + 4.0 * f2 * (1.5 + f - f2), (0.5 - f) ** 4)
        return first - 2, tuple(v / 24.0 for v in w)
    return first - 1, (0.5 * (0.25 + f2 + f), 0.5 * (1.5 - 2 * f2), 0.5 * (0.25 + f2 - f))


def node_charge(pos, weight, q, x_grid_min, y_grid_min_local, dx, dy, nx, ny):
    """charge on the staggered nodes with the deposit's own weights (particles.F90:369-388, DOCUMENTATION eq. 96):
    Q(cx, cy) = q w fac hx hy, as a [ir + NG - 1, ix + NG - 1] array.  np.bincount: one pass per node of the shape."""
    SX, SY = nx + 2 * NG, ny + 2 * NG
    xr = (pos[:, 0] - x_grid_min) / dx
    rr = (np.hypot(pos[:, 1], pos[:, 2]) - y_grid_min_local) / dy
    (cx0, wx), (cy0, wy) = stag_weights(xr), stag_weights(rr)
    qw = q * weight
    Q = np.zeros(SX * SY)
    for a in range(len(wy)):
        for b in range(len(wx)):
            ix = cx0 + b + NG - 1
            iy = cy0 + a + NG - 1
            ok = (ix >= 0) & (ix < SX) & (iy >= 0) & (iy < SY)
            idx, val = iy * SX + ix, qw * wx[b] * wy[a]
            if not ok.all():          # (full-size runs: everything is inside, no masked copies)
                idx, val = idx[ok], val[ok]
            Q += np.bincount(idx, weights=val, minlength=SX * SY)
    return Q.reshape(SY, SX)


def gauss_residual(ex0, er0, parts, q, mass, dt, x_grid_min, y_grid_min_local, dx, dy, nx, ny, margin=3):
    """Integral Gauss law of mode 0 on the deposit's control volumes,
         A_rt(cy) [Ex(cx+1,cy) - Ex(cx,cy)] + A_xt(cy) Er(cx,cy+1) - A_xt(cy-1) Er(cx,cy) - Q(cx,cy)/eps0,
    Q = mean of the node charges half a step before and after the stored positions (E at the end of a
    step has seen half of each of the two currents).  Returns (residual, flux) on the interior
    [margin : n - margin).  ex0 / er0: real parts of mode 0 of exm / erm, [ir + NG - 1, ix + NG - 1]."""
    u = parts[:, 3:6] / (mass * C_LIGHT)
    delta = u * (C_LIGHT * dt / 2.0) / np.sqrt(1 + (u * u).sum(1))[:, None]
    args = (q, x_grid_min, y_grid_min_local, dx, dy, nx, ny)
    Q = 0.5 * (node_charge(parts[:, 0:3] + delta, parts[:, 6], *args) +
               node_charge(parts[:, 0:3] - delta, parts[:, 6], *args))
    a_rt, a_xt = area_tables(ny, dx, dy)
    o = NG - 1
    cy = np.arange(margin, ny - margin)
    cx = np.arange(margin, nx - margin)
    J, I = np.meshgrid(cy + o, cx + o, indexing="ij")
    flux = (a_rt[cy + NG][:, None] * (ex0[J, I + 1] - ex0[J, I])
            + a_xt[cy + NG][:, None] * er0[J + 1, I] - a_xt[cy - 1 + NG][:, None] * er0[J, I])
    return flux - Q[J, I] / EPSILON0, flux


def same_multiset(a, b):
    """two float arrays hold the same values (order-free, bit-exact): what a sort or a migration
    must preserve for the carried quantities (weights)"""
    return a.shape == b.shape and np.array_equal(np.sort(a, axis=None), np.sort(b, axis=None))
