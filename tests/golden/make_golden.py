"""Generates tests/golden/lwfa_48x16_m2_20steps.npz from the CPU oracle (and, run under CYL_SHAPE=tophat / bspline3,
lwfa_48x16_m2_20steps_<shape>.npz from the oracle build of that particle shape).

The reference ships no fixtures for the cylindrical path and cannot be built here (Fortran +
MPI), so these vectors come from the oracle restatement; they pin it against drift and give
the GPU suite a file-based target that does not need the oracle at run time.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(HERE), os.path.dirname(os.path.dirname(HERE)),
          os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

NSTEPS = 20


def deck():
    import decks
    return decks.lwfa(nx=48, ny=16, n_mode=2, ppc_e=3, ppc_p=1, t_centre=6e-15)


def run_case():
    import decks
    d = deck()
    w = decks.make_oracle(d)
    out = {}
    for isp in range(2):
        out[f"init_particles_{isp}"] = w.particles(0, isp).copy()
    w.call("init_half_step")
    w.step(NSTEPS)
    for name in ("exm", "erm", "etm", "bxm", "brm", "btm", "jxm", "jrm", "jtm"):
        out[name] = w.field(0, name).copy()
    for isp in range(2):
        p = w.particles(0, isp)
        out[f"particles_{isp}"] = p[np.lexsort((p[:, 0], p[:, 6]))].copy()
    st = w.stats(0)
    out["counts"] = np.array([w.nparticles(0, 0), w.nparticles(0, 1), st["removed"]], dtype=np.int64)
    return out


def golden_path():
    import pyoracle
    tag = "" if pyoracle.SHAPE == "triangle" else "_" + pyoracle.SHAPE
    return os.path.join(HERE, f"lwfa_48x16_m2_20steps{tag}.npz")


if __name__ == "__main__":
    np.savez_compressed(golden_path(), **run_case())
    print("written", golden_path())
