"""TEST INFRASTRUCTURE: a literal Python restatement of calculate_breaks (balance.F90:2510-2653) and of the load
profile of get_load (balance.F90:2322-2365, part_load_func :2453-2478), against which the product's
cylgpu_calculate_breaks / cylgpu_load_x are checked.  Only tests/ may import this module."""
import math

import numpy as np

from pyoracle import NG   # ng = png + 2 of the particle shape in use (CYL_SHAPE)

PNG = NG - 2
NCELL_MIN = (PNG + 1) // 2 + 1       # constants.F90:548
PUSH_PER_FIELD = 5                   # shared_data.F90:761


def calculate_breaks(load, nproc):
    """load: sequence holding load(1-ng : sz+ng); returns (mins, maxs), 1-based inclusive"""
    sz = len(load) - 2 * NG
    L = lambda i: int(load[i - (1 - NG)])      # noqa: E731   Fortran index
    mins = [1] * nproc
    maxs = [sz] * (nproc + 2)                  # 1-based, maxs[nproc] = sz
    if nproc < 2:
        return [1], [sz]
    load_per_proc_ideal = math.floor(sum(L(i) for i in range(1, sz + 1)) / nproc + 0.5)
    proc, old, total = 0, 1, 0
    for idim in range(1, sz + 1):
        total_old = total
        total = total + L(idim)
        if total >= load_per_proc_ideal:
            proc += 1
            if load_per_proc_ideal - total_old < total - load_per_proc_ideal:
                maxs[proc] = idim - 1
            else:
                maxs[proc] = idim
            nextra = old - maxs[proc] + NCELL_MIN
            if nextra > 0:
                maxs[proc] = maxs[proc] + nextra
            if proc == nproc - 1:
                break
            old = maxs[proc]
            total = total - load_per_proc_ideal

    def backwards():
        o = sz
        for p in range(nproc - 1, 0, -1):
            if o - maxs[p] < NCELL_MIN:
                maxs[p] = o - NCELL_MIN
            o = maxs[p]

    def spread():
        lmax, lmin, i0 = -1, None, 1
        for p in range(1, nproc + 1):
            i1 = maxs[p]
            l = sum(L(i) for i in range(i0, i1 + 1))
            lmax = max(lmax, l)
            lmin = l if lmin is None else min(lmin, l)
            i0 = i1 + 1
        return lmax, lmin

    backwards()
    best = None
    lmax = lmin = 0
    for _ in range(1000):
        done = False
        for i in range(1, nproc):
            old_maxs = maxs[i]
            o = 0 if i == 1 else maxs[i - 1]
            new_maxs = old_maxs - 1 if old_maxs - o - 1 >= NG else old_maxs
            if new_maxs != old_maxs:
                maxs[i] = new_maxs
                lmax, lmin = spread()
                if best is None or lmax - lmin < best:
                    done = True
                    break
                maxs[i] = old_maxs
            old_maxs = maxs[i]
            o = maxs[i + 1]
            new_maxs = old_maxs + 1 if o - old_maxs - 1 >= NG else old_maxs
            if new_maxs != old_maxs:
                maxs[i] = new_maxs
                lmax, lmin = spread()
                if best is None or lmax - lmin < best:
                    done = True
                    break
                maxs[i] = old_maxs
        _ = done
        if best is None or lmax - lmin < best:
            best = lmax - lmin
        else:
            break
    backwards()
    o = 0
    for p in range(1, nproc):
        if maxs[p] - o < NCELL_MIN:
            maxs[p] = o + NCELL_MIN
        o = maxs[p]
    out_max = [maxs[p] for p in range(1, nproc)] + [sz]
    for p in range(2, nproc + 1):
        mins[p - 1] = maxs[p - 1] + 1
    return mins, out_max


def load_x_profile(particle_x, x_grid_min, dx, nx_global, ny_global):
    """get_load's load_x from global particle positions: push_per_field * count per column + ny_global on 1..nx"""
    load = np.zeros(nx_global + 2 * NG, dtype=np.int64)
    cell = np.floor((np.asarray(particle_x) - x_grid_min) / dx + 1.5).astype(np.int64)
    cell = np.clip(cell, 1 - NG, nx_global + NG)
    np.add.at(load, cell - (1 - NG), 1)
    load *= PUSH_PER_FIELD
    load[NG:NG + nx_global] += ny_global
    return load
