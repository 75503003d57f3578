"""GPU side of the SDF dump / restart row (cylgpu_sdf_dump / cylgpu_sdf_load): the file written from
the device mirrors equals the file the host-level writer produces from the oracle's arrays (which
tests/test_sdf.py checks through the reference's own reader), and a run restarted from it continues
like the uninterrupted one.

Sorts after the other test modules on purpose (see tests/test_zz1_gpu_moments.py): written after
the round's GPU budget was spent, first run on a B200 is the driver's.
"""
import numpy as np
import pytest

import decks
from parity import Pair, TOL, TOL_HOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nranks", [1, 2])
def test_dump_from_device_then_restart(tmp_path, nranks):
    # Conducting box (clamp on every wall, reflecting particles): every value the file does not hold --
    # ghosts, the x_min face column 0 and the r_max row of the r-staggered arrays, which the reference's
    # writer drops as well (io/diagnostics.F90:2085-2097 writes rows 0..ny-1 of those) -- is re-derived by
    # efield_bcs / bfield_bcs, so the restarted run must track the uninterrupted one.
    d = decks.drift(nx=48, ny=16, n_mode=2)
    p = Pair(d, nranks=nranks)
    q = None
    try:
        p.step(4)
        path = str(tmp_path / "0004.sdf")
        counts = [s.particle_count(0) for s in p.slabs]
        total = sum(counts)
        ds = [None] * nranks

        def dump(s):
            k = p.slabs.index(s)
            ds[k] = s.sdf_dump(path, ["electron"], npart_global=[total], npart_offset=[sum(counts[:k])], restart=True,
                               derived=("number_density", "temperature", "jx"))
        p.each(dump)
        # the derived blocks were computed on the device (cylgpu_particle_moment): compare with the oracle's
        # moments through the reference's reader when oracle/_ref travelled with the snapshot
        import os
        import subprocess
        exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "sdf_ref_dump")
        if os.path.exists(exe):
            out = str(tmp_path / "out")
            os.makedirs(out)
            r = subprocess.run([exe, path, out], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            ids = [ln.split(" id=")[1].split(" ")[0] for ln in r.stdout.splitlines()[1:]]
            for bid, (kind, direction) in (("number_density", ("number_density", 0)), ("temperature", ("temperature", 0)),
                                           ("jx", ("species_current", 1))):
                for suffix, isp in (("", -1), ("/electron", 0)):
                    n = ids.index(bid + suffix)
                    got = np.fromfile(os.path.join(out, f"{n}.bin")).reshape(d.ny, d.nx)
                    ref = np.concatenate([m[5:-5, 5:-5] for m in p.oracle.moment(kind, isp, direction)], axis=1)
                    # the device computed them from ITS particles, the oracle from its own (equal to ~1e-12 after
                    # 4 steps); the temperature of this cold beam (drift / spread = 3400) amplifies that difference
                    tol = 1e-5 if bid == "temperature" else 1e-9
                    assert np.abs(got - ref).max() <= tol * np.abs(ref).max(), (bid, suffix)
        # a second pair of slabs restarts from the file and both continue
        q = Pair(d, nranks=nranks)
        q.oracle = p.oracle            # one oracle: it is the uninterrupted run
        q.each(lambda s: s.sdf_load(path, ["electron"]))
        for s, t in zip(p.slabs, q.slabs):
            assert t.step == s.step and t.time == s.time
            assert t.particle_count(0) == s.particle_count(0)
            for name in ("exm", "erm", "etm", "bxm", "brm", "btm", "jxm", "jrm", "jtm"):
                a, b = s.download_field(name), t.download_field(name)
                rows = slice(4, -6) if name in ("exm", "etm", "brm", "jxm", "jtm") else slice(5, -5)
                assert np.array_equal(a[:, rows, 5:-5], b[:, rows, 5:-5]), name
        p.step(3)
        for _ in range(3):
            q.each(lambda s: s.step_once())
        q.check_counts()
        q.check_fields(TOL_HOT)
        q.check_particles(TOL_HOT)
    finally:
        p.close()
        if q is not None:
            q.oracle = None
            q.close()
