"""Diagnostic for tests/test_zz2_gpu_counter_insert.py: per-block error growth of the
window deck with (a) the device-generated hot column, (b) the host KISS column at the same temperature,
(c) the device-generated column cold.  If (a) and (b) grow alike the drift is the deck's conditioning
(hot particles near the Taylor switch of particles.F90:593), not CUDA's log / sin / cos in the column kernel.

    python tools/diag_zz2.py > gpurun_out/diag_zz2.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import decks  # noqa: E402
from parity import Pair  # noqa: E402

SEED = 0x1234_5678_9ABC


def run(label, temp, device, nranks=1, blocks=8, per=5):
    d = decks.lwfa(nx=64, ny=24, n_mode=2, ppc_e=4, ppc_p=1, window=True, t_centre=30e-15)
    d.species[0].temp = temp
    if device:
        p = Pair(d, nranks=nranks, prepare=lambda o: o.set_counter_insert(True, SEED),
                 slab_kw=dict(device_insert_seed=SEED))
    else:
        p = Pair(d, nranks=nranks)
    try:
        for b in range(blocks):
            p.step(per)
            e = p.field_errors()
            try:
                w = p.check_particles(1.0)
            except AssertionError as ex:   # noqa: BLE001
                w = str(ex)
            worst = max(e, key=e.get)
            print(f"{label:28s} step {(b + 1) * per:3d} shifts {p.slabs[0].window_shifts_total:3d} "
                  f"worst field {worst}={e[worst]:.2e} jtm={e['jtm']:.2e} jxm={e['jxm']:.2e} "
                  f"exm={e['exm']:.2e} particles={w if isinstance(w, str) else format(w, '.2e')}", flush=True)
    finally:
        p.close()


if __name__ == "__main__":
    hot = (2.0e5, 1.0e5, 3.0e5)
    run("device column, hot", hot, True)
    run("host KISS column, hot", hot, False)
    run("device column, cold", (0.0, 0.0, 0.0), True)
    run("device column, hot, 2 ranks", hot, True, nranks=2)
