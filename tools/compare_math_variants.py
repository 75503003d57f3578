#!/usr/bin/env python
"""How far the branch-free division / square root / arcsine of push.cuh (rcp.approx / rsqrt.approx seeds + Newton
steps, <= 1 ulp) move a run: the same decks stepped by the default library and by the -DCYL_REFERENCE_MATH build
(IEEE division, sqrt, atan2), and both against the oracle.  Usage (GPU box): python tools/compare_math_variants.py
Each library runs in its own process (CYLGPU_LIB); the states travel through .npz files."""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
WORKER = r'''
import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, "oracle")
import decks
from parity import Pair
name, steps, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
d = {"lwfa": lambda: decks.lwfa(nx=128, ny=32, n_mode=2, ppc_e=8, ppc_p=2, t_centre=12e-15),
     "thermal": lambda: decks.thermal(nx=64, ny=32, n_mode=2, ppc=16),
     "window": lambda: decks.lwfa(nx=96, ny=24, n_mode=2, ppc_e=4, ppc_p=1, window=True, t_centre=30e-15)}[name]()
p = Pair(d, nranks=1)
p.step(steps)
res = {}
from cylindrical_epoch_b200.constants import FIELD_NAMES
for n in FIELD_NAMES[:9]:
    res["f_" + n] = p.slabs[0].download_field(n)
    res["o_" + n] = p.oracle.field(0, n).copy()
for i in range(len(d.species)):
    a = p.slabs[0].download_particles(i); b = p.oracle.particles(0, i).reshape(-1, 7)
    res["p%d" % i] = a[np.lexsort((a[:, 0], a[:, 6]))]
    res["q%d" % i] = b[np.lexsort((b[:, 0], b[:, 6]))]
np.savez(out, **res)
p.close()
'''


def run(lib, name, steps, out):
    env = dict(os.environ)
    if lib:
        env["CYLGPU_LIB"] = lib
    subprocess.run([sys.executable, "-c", WORKER, name, str(steps), out], check=True, cwd=ROOT, env=env)
    return np.load(out)


def rel(a, b):
    den = np.abs(b).max()
    return 0.0 if den == 0 else float(np.abs(a - b).max() / den)


def main():
    ref_lib = os.path.join(ROOT, "cylindrical_epoch_b200", "libcylgpu_refmath.so")
    assert os.path.exists(ref_lib), "tools/build_variant.sh refmath -DCYL_REFERENCE_MATH first"
    print("deck steps | worst field: default vs IEEE build, default vs oracle, IEEE build vs oracle | same for particle positions / momenta")
    with tempfile.TemporaryDirectory() as tmp:
        for name, steps in (("lwfa", 40), ("window", 40), ("thermal", 20)):
            a = run(None, name, steps, os.path.join(tmp, "a.npz"))
            b = run(ref_lib, name, steps, os.path.join(tmp, "b.npz"))
            fk = [k[2:] for k in a.files if k.startswith("f_")]
            f_ab = max(rel(a["f_" + k], b["f_" + k]) for k in fk)
            f_ao = max(rel(a["f_" + k], a["o_" + k]) for k in fk if not k.startswith("j"))
            f_bo = max(rel(b["f_" + k], b["o_" + k]) for k in fk if not k.startswith("j"))
            pk = [k for k in a.files if k.startswith("p")]
            assert all(a[k].shape == b[k].shape for k in pk)
            p_ab = max(max(rel(a[k][:, :3], b[k][:, :3]), rel(a[k][:, 3:6], b[k][:, 3:6])) for k in pk)
            p_ao = max(max(rel(a[k][:, :3], a["q" + k[1:]][:, :3]), rel(a[k][:, 3:6], a["q" + k[1:]][:, 3:6])) for k in pk)
            p_bo = max(max(rel(b[k][:, :3], b["q" + k[1:]][:, :3]), rel(b[k][:, 3:6], b["q" + k[1:]][:, 3:6])) for k in pk)
            print("%-8s %3d | fields %.2e %.2e %.2e | particles %.2e %.2e %.2e" % (name, steps, f_ab, f_ao, f_bo, p_ab, p_ao, p_bo))


if __name__ == "__main__":
    main()
