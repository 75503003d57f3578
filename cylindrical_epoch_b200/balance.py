"""Slab re-balancer for the z-only (x-slab) decomposition: balance_workload of the reference
(housekeeping/balance.F90:93-299 with nprocy = 1) over the cylgpu handles.

What the reference does in balance_workload -- measure the imbalance, build the load profile (get_load with
part_load_func), place the new slab boundaries (calculate_breaks), accept them if the imbalance improves by more
than 5 %, then redistribute_domain + distribute_particles -- is host orchestration there too.  The arithmetic is in
the library (cylgpu_load_x on the device, cylgpu_calculate_breaks: csrc/balance.cu); this module does the
orchestration and moves whole columns and their particles to the new owners:

  * `plan(...)`              the decision of balance_workload from the slabs' column loads (pure host logic),
  * `redistribute_local(...)` states of all slabs held in one process (the in-process fabric of the tests),
  * `redistribute_dist(...)`  one slab per process over torch.distributed (gloo / nccl): every old slab sends each new
                             owner the columns and particles that fall into its range, point to point,
  * `rebalance_slabs(...)`   plan + redistribution + re-creation of the handles for slabs in one process.

A slab's state = its 15 mode arrays with ghosts, 12 boundary snapshots, the particle lists, the random stream of the
rank and the loop scalars.  Because the ghost columns of an x-slab equal the neighbour's interior columns after the
exchanges that end every phase, the global arrays are assembled from the interiors (plus the outer ghosts of the two
end slabs) and cut again with ghosts: nothing is recomputed, the step after a re-balance continues bit for bit
(up to the order of the deposit sums) -- tests/test_zz8_gpu_rebalance.py.
"""
import ctypes as C

import numpy as np

from .constants import FIELD_NAMES, NG, SNAP_NAMES

PUSH_PER_FIELD = 5          # shared_data.F90:761
DLB_THRESHOLD = 1.0         # deck default: always consider (deck_control_block.F90); callers pass their own


# ------------------------------------------------------------------------------------------------ decision
def calculate_breaks(lib, load, nproc):
    """cylgpu_calculate_breaks on load(1-ng : sz+ng); returns [(cell_min, cell_max)] per slab, 1-based inclusive"""
    a = np.ascontiguousarray(load, dtype=np.int64)
    sz = a.shape[0] - 2 * NG
    mins = (C.c_int32 * nproc)()
    maxs = (C.c_int32 * nproc)()
    if lib.cylgpu_calculate_breaks(a.ctypes.data, sz, nproc, mins, maxs) != 0:
        raise RuntimeError(lib.cylgpu_last_error().decode())
    return widen_narrow_slabs([(int(mins[p]), int(maxs[p])) for p in range(nproc)], sz)


def widen_narrow_slabs(bounds, sz, width=2 * NG):
    """The reference lets a slab shrink to ncell_min = (png + 1) / 2 + 1 cells (constants.F90:548); a handle needs
    two halos' worth of columns (cylgpu_create: nx >= 2 ng).  Breaks that cut narrower are moved with the
    reference's own backwards / forwards passes (balance.F90:2572-2577, 2646-2651) at that width -- the one place
    where the slabs may differ from the reference's, and only where it would cut below 2 ng columns."""
    n = len(bounds)
    if n * width > sz:
        raise RuntimeError(f"{sz} columns cannot hold {n} slabs of at least {width}")
    mx = [hi for _, hi in bounds]
    o = sz
    for p in range(n - 2, -1, -1):
        if o - mx[p] < width:
            mx[p] = o - width
        o = mx[p]
    o = 0
    for p in range(n - 1):
        if mx[p] - o < width:
            mx[p] = o + width
        o = mx[p]
    mx[n - 1] = sz
    return [(1 if p == 0 else mx[p - 1] + 1, mx[p]) for p in range(n)]


def global_load_x(column_counts, bounds, nx_global, ny_global):
    """get_load (balance.F90:2322-2365): the slabs' particle counts per local column (cylgpu_load_x, index
    1-ng..nx+ng) summed into the global profile load_x(1-ng : nx_global+ng), times push_per_field, plus ny_global
    per interior column"""
    load = np.zeros(nx_global + 2 * NG, dtype=np.int64)
    for (lo, hi), cnt in zip(bounds, column_counts):
        cnt = np.asarray(cnt, dtype=np.int64)
        assert cnt.shape[0] == hi - lo + 1 + 2 * NG
        # local column ix <-> global column lo - 1 + ix; ghosts beyond the global ghosts are clamped into them
        g0 = lo - 1 + (1 - NG)                      # global index of the first local entry
        for k, v in enumerate(cnt):
            gi = min(max(g0 + k, 1 - NG), nx_global + NG)
            load[gi - (1 - NG)] += v
    load *= PUSH_PER_FIELD
    load[NG:NG + nx_global] += ny_global
    return load


def balance_fraction(loads):
    """(load_av + sqrt(load_av)) / (load_max + sqrt(load_max)), balance.F90:152-155"""
    loads = np.asarray(loads, dtype=np.float64)
    av, mx = loads.mean(), loads.max()
    return float((av + np.sqrt(av)) / (mx + np.sqrt(mx)))


def plan(lib, column_counts, bounds, nx_global, ny_global, over_ride=False, dlb_threshold=DLB_THRESHOLD):
    """balance_workload's decision (balance.F90:143-225): returns (new_bounds or None, balance_frac,
    balance_frac_final).  column_counts[k]: cylgpu_load_x of slab k; bounds[k] = (cell_x_min, cell_x_max)."""
    nproc = len(bounds)
    if nproc == 1:
        return None, 1.0, 1.0
    npart = [int(np.asarray(c).sum()) for c in column_counts]
    load_local = [PUSH_PER_FIELD * n + (hi - lo + 1) * ny_global for n, (lo, hi) in zip(npart, bounds)]
    frac = balance_fraction(load_local)
    if not over_ride and frac > dlb_threshold:
        return None, frac, frac
    load = global_load_x(column_counts, bounds, nx_global, ny_global)
    new_bounds = calculate_breaks(lib, load, nproc)
    # calculate_new_load_imbalance (balance.F90:2820-2929): push_per_field * particles + 1 per cell of the new slabs
    per_col = (load[NG:NG + nx_global] - ny_global) // PUSH_PER_FIELD
    new_load = [PUSH_PER_FIELD * int(per_col[lo - 1:hi].sum()) + (hi - lo + 1) * ny_global for lo, hi in new_bounds]
    frac_final = balance_fraction(new_load)
    improvement = (frac_final - frac) / frac
    if improvement > 0.05 and new_bounds != list(bounds):      # balance.F90:186-190
        return new_bounds, frac, frac_final
    return None, frac, frac_final


# ------------------------------------------------------------------------------------------------ state
def slab_state(slab):
    """everything a slab owns, on the host"""
    st = dict(fields={n: slab.download_field(n) for n in FIELD_NAMES},
              snaps={n: slab.download_snapshot(n) for n in SNAP_NAMES},
              particles=[slab.download_particles(i) for i in range(len(slab.species))],
              rng=slab.rng_get_state(), bounds=(slab.grid.cell_x_min, slab.grid.cell_x_max))
    return st


def _columns_for(lo, hi, nx_global, periodic):
    """global columns (1-based, may lie in the ghosts) a slab (lo, hi) holds, ghosts included"""
    cols = np.arange(lo - NG, hi + NG + 1)
    if periodic:
        cols = (cols - 1) % nx_global + 1
    return cols


def assemble_global(field_pieces, bounds, nx_global, periodic):
    """global array (M, SY, nx_global + 2 ng) from the slabs' arrays: interiors, plus the outer ghost columns of the
    two end slabs (a periodic x wraps instead)"""
    first = field_pieces[0]
    G = np.zeros(first.shape[:-1] + (nx_global + 2 * NG,), dtype=first.dtype)
    for (lo, hi), a in zip(bounds, field_pieces):
        G[..., NG + lo - 1:NG + hi] = a[..., NG:NG + (hi - lo + 1)]
    G[..., :NG] = field_pieces[0][..., :NG]
    G[..., NG + nx_global:] = field_pieces[-1][..., -NG:]
    if periodic:
        G[..., :NG] = G[..., nx_global:nx_global + NG]
        G[..., NG + nx_global:] = G[..., NG:2 * NG]
    return G


def cut_slab(G, lo, hi):
    return np.ascontiguousarray(G[..., lo - 1:hi + 2 * NG])


def redistribute_local(states, new_bounds, nx_global, x_edges, periodic=False):
    """states[k] of all old slabs -> states of the new slabs.  x_edges[k] = (x_min_local, x_max_local) of NEW slab k
    (boundary.F90:1607,1685: a particle belongs to x_min_local <= x < x_max_local)."""
    old_bounds = [s["bounds"] for s in states]
    nsp = len(states[0]["particles"])
    out = []
    globals_ = {n: assemble_global([s["fields"][n] for s in states], old_bounds, nx_global, periodic)
                for n in FIELD_NAMES}
    allp = [np.concatenate([s["particles"][i] for s in states]) for i in range(nsp)]
    for k, (lo, hi) in enumerate(new_bounds):
        xl, xr = x_edges[k]
        parts = []
        for i in range(nsp):
            x = allp[i][:, 0]
            sel = (x >= xl) & (x < xr)
            if k == 0:
                sel |= x < xl            # what has left the domain but not yet the list stays with the end slabs
            if k == len(new_bounds) - 1:
                sel |= x >= xr
            parts.append(np.ascontiguousarray(allp[i][sel]))
        out.append(dict(fields={n: cut_slab(globals_[n], lo, hi) for n in FIELD_NAMES},
                        # the x_min / x_max snapshots are only used by the slabs that own those walls
                        snaps={n: (states[0] if n.endswith("_x_min") else states[-1])["snaps"][n] for n in SNAP_NAMES},
                        particles=parts, rng=states[k]["rng"], bounds=(lo, hi)))
    return out


def redistribute_dist(state, old_bounds, new_bounds, nx_global, x_edges, rank, periodic=False, group=None):
    """the same with one slab per process: point-to-point over torch.distributed.  Every old slab sends each new owner
    the interior columns of its 15 arrays that fall into the owner's range WITH GHOSTS, and the particles inside the
    owner's x range; sizes first (one all_gather of small lists), then the payloads in a fixed order."""
    import torch
    import torch.distributed as dist
    world = len(old_bounds)
    lo_old, hi_old = old_bounds[rank]
    nsp = len(state["particles"])
    shape = state["fields"][FIELD_NAMES[0]].shape[:-1]

    def send_cols(dest):
        """(global columns this rank contributes to new slab `dest`, positions in dest's array, in my array)"""
        lo, hi = new_bounds[dest]
        cols = _columns_for(lo, hi, nx_global, periodic)
        mine = (cols >= lo_old) & (cols <= hi_old)
        if not periodic:
            if rank == 0:
                mine |= cols < 1                    # outer ghosts live with the end slabs
            if rank == world - 1:
                mine |= cols > nx_global
        pos = np.nonzero(mine)[0]
        return pos, cols[pos] - lo_old + NG          # index in my local array (column c <-> c - lo_old + ng)

    outgoing = []
    for d in range(world):
        pos, loc = send_cols(d)
        xl, xr = x_edges[d]
        plist = []
        for i in range(nsp):
            x = state["particles"][i][:, 0]
            sel = (x >= xl) & (x < xr)
            if d == 0:
                sel |= x < xl
            if d == world - 1:
                sel |= x >= xr
            plist.append(np.ascontiguousarray(state["particles"][i][sel]))
        outgoing.append((pos, loc, plist))
    meta = [(len(o[0]), [p.shape[0] for p in o[2]]) for o in outgoing]
    gathered = [None] * world
    dist.all_gather_object(gathered, meta, group=group)      # gathered[src][dest] = (ncols, [n per species])

    def pack(d):
        pos, loc, plist = outgoing[d]
        chunks = [pos.astype(np.float64)]
        for n in FIELD_NAMES:
            chunks.append(np.ascontiguousarray(state["fields"][n][..., loc]).view(np.float64).ravel())
        chunks += [p.ravel() for p in plist]
        return np.concatenate(chunks) if chunks else np.zeros(0)

    reqs, recv_bufs = [], {}
    for src in range(world):
        ncols, nps = gathered[src][rank]
        size = ncols + len(FIELD_NAMES) * int(np.prod(shape)) * ncols * 2 + 7 * sum(nps)
        if src == rank or size == 0:
            continue
        recv_bufs[src] = torch.empty(size, dtype=torch.float64)
        reqs.append(dist.irecv(recv_bufs[src], src=src, group=group))
    send_keep = []
    for d in range(world):
        if d == rank:
            continue
        buf = torch.from_numpy(pack(d))
        if buf.numel():
            send_keep.append(buf)
            reqs.append(dist.isend(buf, dst=d, group=group))
    for r in reqs:
        r.wait()
    lo, hi = new_bounds[rank]
    ncol_new = hi - lo + 1 + 2 * NG
    new_fields = {n: np.zeros(shape + (ncol_new,), dtype=np.complex128) for n in FIELD_NAMES}
    new_parts = [[] for _ in range(nsp)]
    for src in range(world):
        ncols, nps = gathered[src][rank]
        if src == rank:
            flat = pack(rank)
        elif src in recv_bufs:
            flat = recv_bufs[src].numpy()
        else:
            continue
        pos = flat[:ncols].astype(np.int64)
        off = ncols
        per = int(np.prod(shape)) * ncols * 2
        for n in FIELD_NAMES:
            if ncols:
                new_fields[n][..., pos] = flat[off:off + per].view(np.complex128).reshape(shape + (ncols,))
            off += per
        for i, npart in enumerate(nps):
            new_parts[i].append(flat[off:off + 7 * npart].reshape(npart, 7))
            off += 7 * npart
    # snapshots travel from the wall owners to the (same-numbered) wall owners: ranks 0 and world - 1 keep theirs
    return dict(fields=new_fields, snaps=state["snaps"],
                particles=[np.concatenate(p) if p else np.zeros((0, 7)) for p in new_parts], rng=state["rng"],
                bounds=(lo, hi))


# ------------------------------------------------------------------------------------------------ handles
def rebalance_slabs(slabs, transport_kw=lambda k: {}, over_ride=False, dlb_threshold=DLB_THRESHOLD, force_bounds=None):
    """balance_workload for slabs held in ONE process (tests, single-process drivers).  transport_kw(rank): the
    transport arguments of the new handle of `rank` (e.g. the in-process fabric); returns (slabs, report)."""
    lib = slabs[0].L
    g0 = slabs[0].grid
    bounds = [(s.grid.cell_x_min, s.grid.cell_x_max) for s in slabs]
    counts = [s.load_x() for s in slabs]
    new_bounds, frac, frac_final = plan(lib, counts, bounds, g0.nx_global, g0.ny_global, over_ride, dlb_threshold)
    if force_bounds is not None:     # (tests: a prescribed split, e.g. the present one = a pure hand-over of the state)
        new_bounds = [tuple(b) for b in force_bounds]
    report = dict(balance=frac, after=frac_final, redistributed=new_bounds is not None, bounds=new_bounds or bounds)
    if new_bounds is None:
        return slabs, report
    periodic = slabs[0].periodic_x
    states = [slab_state(s) for s in slabs]
    new_slabs = [s.respawn(new_bounds, **transport_kw(k)) for k, s in enumerate(slabs)]
    for s in slabs:
        s.close()
    x_edges = [(s.grid.x_min_local, s.grid.x_max_local) for s in new_slabs]
    new_states = redistribute_local(states, new_bounds, g0.nx_global, x_edges, periodic)
    for s, st in zip(new_slabs, new_states):
        s.load_state(st)
    return new_slabs, report
