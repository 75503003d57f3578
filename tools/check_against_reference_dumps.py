#!/usr/bin/env python
"""Pin the oracle (and, with --gpu, the CUDA path) against dumps of the REAL reference.

Neither this container nor the GPU box can build the reference (Fortran 2003 + MPI), so its
physics is not pinned by reference output here.  Anyone who can build it
elsewhere closes that gap with two restart dumps of one short run:

    make COMPILER=gfortran                      # in epoch_axial/
    mpirun -n 1 ./bin/epoch2d                   # deck with  restart dumps at step A and step B > A
    python tools/check_against_reference_dumps.py deck.json A.sdf B.sdf [--gpu]

deck.json describes what the dumps do not hold (it mirrors the deck's control / boundaries /
species / laser blocks):

    {"nx": 512, "ny": 64, "n_mode": 2, "x_min": 0.0, "x_max": 1.6e-5, "y_max": 2.1e-5,
     "bc_field": [3, 5, 0, 5], "dt_multiplier": 0.95, "nranks": 1,
     "species": [{"name": "electron", "charge": -1.602176565e-19, "mass": 9.10938291e-31,
                  "bc_particle": [5, 5, 5, 5]}],
     "lasers": [{"boundary": 0, "amp": ..., "omega": ..., "t_centre": ..., "t_width": ..., "r_width": ...}]}

The tool loads dump A (the 15 mode arrays and the particle lists; ghosts are re-derived by the
boundary routines exactly as the reference's own restart does, housekeeping/setup.F90:1196-1260),
advances B.step - A.step steps with the oracle (and the CUDA path), and compares with dump B:
particle counts exactly, per-array and per-particle maximum errors relative to the array maximum.
What a restart cannot carry (absorbing-boundary ghost columns, the E row at r_max under zero_b)
limits the agreement to ~1e-5 on laser decks and makes conducting-box decks exact; see DESIGN.md 5a.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from cylindrical_epoch_b200.constants import NG  # noqa: E402  (ng of the build in use, CYL_SHAPE)


def read_dump(path, deck, k, info):
    """this slab's view of a dump through the product's host-level reader (cylgpu_sdf_read_host)"""
    from cylindrical_epoch_b200 import _lib
    from cylindrical_epoch_b200.constants import FIELD_NAMES
    lib = _lib.load()
    nsp = len(deck["species"])
    d = _lib.SdfDesc()
    d.nx_global, d.ny_global, d.n_mode, d.n_species = deck["nx"], deck["ny"], deck["n_mode"], nsp
    d.nx_local, d.cell_x_min = info["nx"], info["cell_x_min"]
    names = [s["name"].encode() for s in deck["species"]]
    for i, n in enumerate(names):
        d.species_name[i] = n
    shape = (deck["n_mode"], deck["ny"] + 2 * NG, info["nx"] + 2 * NG)
    fields = [np.zeros(shape, dtype=np.complex128) for _ in FIELD_NAMES]
    fp = (C.c_void_p * 15)(*[f.ctypes.data for f in fields])
    rc = lib.cylgpu_sdf_read_host(path.encode(), C.byref(d), fp, info["x_min_local"], info["x_max_local"], None, None)
    if rc:
        raise SystemExit(lib.cylgpu_last_error().decode())
    counts = [int(d.npart_local[i]) for i in range(nsp)]
    bufs = [np.zeros((max(c, 1), 7)) for c in counts]
    pp = (C.c_void_p * 8)(*([b.ctypes.data for b in bufs] + [None] * (8 - nsp)))
    cap = (C.c_int64 * 8)(*([b.shape[0] for b in bufs] + [0] * (8 - nsp)))
    rc = lib.cylgpu_sdf_read_host(path.encode(), C.byref(d), fp, info["x_min_local"], info["x_max_local"], pp, cap)
    if rc:
        raise SystemExit(lib.cylgpu_last_error().decode())
    return int(d.step), float(d.time), dict(zip(FIELD_NAMES, fields)), [b[:c] for b, c in zip(bufs, counts)]


def make_world(deck):
    import pyoracle as po
    w = po.OracleWorld(deck["nx"], deck["ny"], deck["n_mode"], deck["x_min"], deck["x_max"], deck["y_max"],
                       list(deck["bc_field"]), nranks=deck.get("nranks", 1),
                       dt_multiplier=deck.get("dt_multiplier", 0.95))
    for s in deck["species"]:
        w.add_species(s["charge"], s["mass"], list(s["bc_particle"]), immobile=s.get("immobile", False),
                      zero_current=s.get("zero_current", False))
    for L in deck.get("lasers", []):
        w.add_laser(**L)
    return w


def sorted_particles(a):
    a = np.asarray(a).reshape(-1, 7)
    return a[np.lexsort((a[:, 0], a[:, 6]))]


def compare(tag, got_fields, got_parts, ref_fields, ref_parts, deck, report):
    from cylindrical_epoch_b200.constants import FIELD_NAMES
    for name in FIELD_NAMES[:9]:
        st = name in ("exm", "etm", "brm", "jxm", "jtm")
        rows = slice(NG - 1, NG - 1 + deck["ny"]) if st else slice(NG, NG + deck["ny"])
        den = max(np.abs(r[name][:, rows, NG:-NG]).max() for r in ref_fields)
        err = max(np.abs(g[name][:, rows, NG:-NG] - r[name][:, rows, NG:-NG]).max() for g, r in zip(got_fields, ref_fields))
        report[f"{tag}:{name}"] = float(err / den) if den > 0 else float(err)
    for i, s in enumerate(deck["species"]):
        g = sorted_particles(np.concatenate([p[i] for p in got_parts]))
        r = sorted_particles(np.concatenate([p[i] for p in ref_parts]))
        report[f"{tag}:count:{s['name']}"] = [int(g.shape[0]), int(r.shape[0])]
        if g.shape == r.shape and g.shape[0]:
            report[f"{tag}:weights_bit_exact:{s['name']}"] = bool(np.array_equal(g[:, 6], r[:, 6]))
            for cols, nm in ((slice(0, 3), "pos"), (slice(3, 6), "p")):
                den = np.abs(r[:, cols]).max()
                report[f"{tag}:{nm}:{s['name']}"] = float(np.abs(g[:, cols] - r[:, cols]).max() / den) if den > 0 else 0.0


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("deck")
    ap.add_argument("dump_a")
    ap.add_argument("dump_b")
    ap.add_argument("--gpu", action="store_true", help="also advance the CUDA path (needs a B200)")
    ap.add_argument("--tol", type=float, default=None, help="fail (exit 1) if any relative error exceeds this")
    args = ap.parse_args(argv)
    deck = json.load(open(args.deck))
    from cylindrical_epoch_b200.constants import FIELD_NAMES
    w = make_world(deck)
    nr = w.nranks
    infos = [w.rank_info(k) for k in range(nr)]
    step_a = time_a = None
    for k in range(nr):
        step_a, time_a, fields, parts = read_dump(args.dump_a, deck, k, infos[k])
        for name in FIELD_NAMES:
            w.field(k, name)[...] = fields[name]
        for i in range(len(deck["species"])):
            w.set_particles(k, i, parts[i])
    w.set_time(time_a)
    ref = [read_dump(args.dump_b, deck, k, infos[k]) for k in range(nr)]
    step_b = ref[0][0]
    nsteps = step_b - step_a
    if nsteps <= 0:
        raise SystemExit(f"dump B (step {step_b}) must be later than dump A (step {step_a})")
    report = {"steps": nsteps, "from_step": step_a, "to_step": step_b}
    slabs = []
    if args.gpu:
        import cylindrical_epoch_b200 as ce
        sp = [ce.Species(s["charge"], s["mass"], tuple(s["bc_particle"]), s.get("immobile", False),
                         s.get("zero_current", False)) for s in deck["species"]]
        if nr != 1:
            raise SystemExit("--gpu: one slab per process; run with nranks = 1")
        sl = ce.Slab(deck["nx"], deck["ny"], deck["n_mode"], deck["x_min"], deck["x_max"], deck["y_max"],
                     list(deck["bc_field"]), sp, dt_multiplier=deck.get("dt_multiplier", 0.95),
                     lasers=[ce.Laser(**L) for L in deck.get("lasers", [])])
        sl.sdf_load(args.dump_a, [s["name"] for s in deck["species"]])
        sl.snapshot_field_boundaries() if hasattr(sl, "snapshot_field_boundaries") else None
        slabs.append(sl)
    # ghosts the file does not hold: the boundary routines rebuild them (setup.F90 restart does the same)
    w.call("snapshot_boundaries")
    w.call("efield_bcs")
    w.call("bfield_bcs")
    w.call("current_finish")
    w.step(nsteps)
    got_fields = [{n: w.field(k, n) for n in FIELD_NAMES} for k in range(nr)]
    got_parts = [[w.particles(k, i).reshape(-1, 7) for i in range(len(deck["species"]))] for k in range(nr)]
    compare("oracle", got_fields, got_parts, [r[2] for r in ref], [r[3] for r in ref], deck, report)
    for sl in slabs:
        for _ in range(nsteps):
            sl.step_once()
        gf = [{n: sl.download_field(n) for n in FIELD_NAMES}]
        gp = [[sl.download_particles(i) for i in range(len(deck["species"]))]]
        compare("cuda", gf, gp, [r[2] for r in ref], [r[3] for r in ref], deck, report)
        sl.close()
    print(json.dumps(report, indent=1))
    if args.tol is not None:
        bad = {k: v for k, v in report.items() if isinstance(v, float) and v > args.tol}
        bad.update({k: v for k, v in report.items() if ":count:" in k and v[0] != v[1]})
        if bad:
            print("FAIL", json.dumps(bad), file=sys.stderr)
            return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
