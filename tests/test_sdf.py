"""SDF dump / restart of the hot-path state (csrc/sdf_io.cu), CPU only.

The writer is checked THROUGH THE REFERENCE'S OWN SDF READER: oracle/sdf_ref/Makefile compiles the
reference's SDF C library and its sdf2ascii utility from /root/reference where they lie (into
oracle/_ref/, test infrastructure), and oracle/sdf_ref/sdf_ref_dump.c exports what that reader
parsed.  So this row of SURVEY.md section 8 is pinned by reference code run here, not only by
our own restatement.  (The host-level entry points work on host arrays: no GPU involved.)
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import decks
from cylindrical_epoch_b200 import _lib
from cylindrical_epoch_b200.constants import FIELD_NAMES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
from cylindrical_epoch_b200.constants import NG  # ng = png + 2 of the build in use (CYL_SHAPE)

# io/diagnostics.F90:497-575: block order, names, units; stagger codes constants.F90:291-299 / sdf_common.f90:184-195
GROUPS = [("Electric Field Modes", "V/m", [("exm", 2), ("erm", 1), ("etm", 3)]),
          ("Magnetic Field Modes", "T", [("bxm", 1), ("brm", 2), ("btm", 4)]),
          ("Magnetic Field Modes", "T", [("bxm_old", 1), ("brm_old", 2), ("btm_old", 4)]),
          ("Current Modes", "A/m^2", [("jxm", 2), ("jrm", 1), ("jtm", 3)]),
          ("Current Modes", "A/m^2", [("jxm_old", 2), ("jrm_old", 1), ("jtm_old", 3)])]
NAMES = [b"electron", b"proton"]


def ref_tools():
    """oracle/_ref/{sdf_ref_dump,sdf2ascii}: built from the reference tree when it is present"""
    exe = os.path.join(REF_DIR, "sdf_ref_dump")
    if not os.path.exists(exe) and os.path.isdir("/root/reference/SDF/C/src"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "sdf_ref"), "-s"])
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return exe, os.path.join(REF_DIR, "sdf2ascii")


def make_desc(w, d, k, nranks, counts, offsets, totals, step=7):
    sc, info = w.scalars(), w.rank_info(k)
    desc = _lib.SdfDesc()
    desc.nx_global, desc.ny_global, desc.n_mode, desc.n_species = d.nx, d.ny, d.n_mode, len(d.species)
    desc.nx_local, desc.cell_x_min = info["nx"], info["cell_x_min"]
    desc.step, desc.time, desc.restart, desc.jobid1, desc.jobid2 = step, sc["time"], 1, 1234, 56
    desc.x_min, desc.dx, desc.dy = sc["xb_min"], sc["dx"], sc["dy"]
    for i in range(len(d.species)):
        desc.species_name[i] = NAMES[i]
        desc.npart_local[i], desc.npart_offset[i], desc.npart_global[i] = counts[i], offsets[i], totals[i]
    return desc


def write_world(w, d, path, order=None):
    nr = w.nranks
    nsp = len(d.species)
    lib = _lib.load()
    counts = [[w.nparticles(k, i) for i in range(nsp)] for k in range(nr)]
    totals = [sum(counts[k][i] for k in range(nr)) for i in range(nsp)]
    state = []
    for k in (order or range(nr)):
        offsets = [sum(counts[q][i] for q in range(k)) for i in range(nsp)]
        desc = make_desc(w, d, k, nr, counts[k], offsets, totals)
        fields = [np.ascontiguousarray(w.field(k, n)) for n in FIELD_NAMES]
        parts = [np.ascontiguousarray(w.particles(k, i).reshape(-1, 7)) for i in range(nsp)]
        fp = (C.c_void_p * 15)(*[f.ctypes.data for f in fields])
        pp = (C.c_void_p * 8)(*([p.ctypes.data for p in parts] + [None] * (8 - nsp)))
        rc = lib.cylgpu_sdf_write_host(path.encode(), C.byref(desc), fp, pp, None)
        assert rc == 0, lib.cylgpu_last_error()
        state.append((k, fields, parts))
    return sorted(state, key=lambda t: t[0])


def parse_ref(exe, path, outdir):
    os.makedirs(outdir, exist_ok=True)
    r = subprocess.run([exe, path, outdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    hdr = dict(t.split("=", 1) for t in lines[0].split()[1:])
    blocks = []
    for ln in lines[1:]:
        # name= may contain blanks: split on the known keys
        keys = ["n", "id", "name", "blocktype", "datatype", "ndims", "dims", "units", "mesh", "stagger", "species",
                "geometry", "extents", "labels", "const"]
        pos = [(ln.find(" " + k + "="), k) for k in keys if (" " + k + "=") in ln]
        pos.sort()
        b = {}
        for (p, k), nxt in zip(pos, pos[1:] + [(len(ln), None)]):
            b[k] = ln[p + len(k) + 2:nxt[0]]
        blocks.append(b)
    return hdr, blocks


def global_mode_array(state, d, name, part, shift):
    """what the file must hold for one block: (n_mode, ny, nx_global), row j <- array row j - shift"""
    cols = []
    for _, fields, _ in state:
        a = fields[FIELD_NAMES.index(name)]
        a = a.real if part == 0 else a.imag
        cols.append(a[:, NG - shift:NG - shift + d.ny, NG:-NG])
    return np.concatenate(cols, axis=2)


@pytest.mark.parametrize("nranks,order", [(1, None), (2, [1, 0]), (3, [2, 0, 1])])
def test_dump_is_read_back_by_the_reference_reader(tmp_path, nranks, order, cylgpu_lib):
    exe, ascii_exe = ref_tools()
    d = decks.lwfa(nx=24, ny=10, n_mode=2, ppc_e=3, ppc_p=1)
    w = decks.make_oracle(d, nranks=nranks)
    w.call("init_half_step")
    w.step(5)
    path = str(tmp_path / "0007.sdf")
    state = write_world(w, d, path, order)
    hdr, blocks = parse_ref(exe, path, str(tmp_path / "out"))
    assert hdr["step"] == "7" and hdr["code"] == "Epoch2d" and hdr["version"] == "1.4" and hdr["restart"] == "1"
    assert float(hdr["time"]) == w.scalars()["time"]
    assert int(hdr["nblocks"]) == len(blocks) == 1 + 30 + 2 + 8

    def data(n, sub=None):
        f = os.path.join(str(tmp_path / "out"), f"{n}.bin" if sub is None else f"{n}.{sub}.bin")
        return np.fromfile(f, dtype=np.float64)

    sc = w.scalars()
    g = blocks[0]
    assert (g["id"], g["name"], g["blocktype"], g["ndims"], g["labels"]) == ("grid", "Grid/Grid", "1", "2", "X,Y")
    assert np.array_equal(data(0, 0), sc["xb_min"] + np.arange(d.nx + 1) * sc["dx"])
    assert np.array_equal(data(0, 1), np.arange(d.ny + 1) * sc["dy"])
    n = 1
    for group, units, comps in GROUPS:
        for part, tag in enumerate(("real", "imag")):
            for stem, stagger in comps:
                b = blocks[n]
                assert b["id"] == f"{stem}_{tag}"
                assert b["name"] == f"{group}/{stem.capitalize()}/{tag}"
                assert (b["blocktype"], b["datatype"], b["ndims"]) == ("3", "4", "3")
                assert b["dims"] == f"{d.nx},{d.ny},{d.n_mode}" and b["units"] == units and b["mesh"] == "mode_grid"
                assert int(b["stagger"]) == stagger
                shift = 1 if stagger in (2, 3) else 0      # io/diagnostics.F90:2085-2097
                want = global_mode_array(state, d, stem, part, shift)
                got = data(n).reshape(d.n_mode, d.ny, d.nx)
                assert np.array_equal(got, want), b["id"]
                n += 1
    allp = [np.concatenate([st[2][i] for st in state]) for i in range(2)]
    for i, nm in enumerate(("electron", "proton")):
        b = blocks[n]
        assert (b["id"], b["name"], b["blocktype"], b["species"]) == (f"grid/{nm}", f"Grid/Particles/{nm}", "2", nm)
        for c in range(3):
            assert np.array_equal(data(n, c), allp[i][:, c])
        n += 1
    for var, units, col in (("Weight", "", 6), ("Px", "kg.m/s", 3), ("Py", "kg.m/s", 4), ("Pz", "kg.m/s", 5)):
        for i, nm in enumerate(("electron", "proton")):
            b = blocks[n]
            assert b["id"] == f"{var.lower()}/{nm}" and b["name"] == f"Particles/{var}/{nm}"
            assert (b["blocktype"], b["units"], b["mesh"], b["species"]) == ("4", units, f"grid/{nm}", nm)
            assert np.array_equal(data(n), allp[i][:, col])
            n += 1
    # the reference's own utility walks the file without complaint and lists every block
    r = subprocess.run([ascii_exe, "-c", "-v", "exm_real", path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "exm_real" in r.stdout


@pytest.mark.parametrize("nranks", [1, 2])
def test_read_host_round_trip(tmp_path, nranks, cylgpu_lib):
    d = decks.thermal(nx=24, ny=12, n_mode=2, ppc=4)
    w = decks.make_oracle(d, nranks=nranks)
    w.call("init_half_step")
    w.step(3)
    path = str(tmp_path / "r.sdf")
    lib = _lib.load()
    state = write_world(w, d, path)
    for k, fields, parts in state:
        info = w.rank_info(k)
        desc = make_desc(w, d, k, nranks, [0], [0], [0], step=0)
        desc.time = 0.0
        out = [np.full_like(f, 7.0 + 1.0j) for f in fields]          # ghosts must stay untouched
        fp = (C.c_void_p * 15)(*[a.ctypes.data for a in out])
        cap = (C.c_int64 * 8)(*([parts[0].shape[0] + 10] + [0] * 7))
        pbuf = np.zeros((parts[0].shape[0] + 10, 7))
        pp = (C.c_void_p * 8)(*([pbuf.ctypes.data] + [None] * 7))
        rc = lib.cylgpu_sdf_read_host(path.encode(), C.byref(desc), fp, info["x_min_local"], info["x_max_local"], pp, cap)
        assert rc == 0, lib.cylgpu_last_error()
        assert desc.step == 7 and desc.time == w.scalars()["time"]
        assert desc.npart_local[0] == parts[0].shape[0]
        assert desc.npart_global[0] == sum(s[2][0].shape[0] for s in state)
        # same particles (file order within a slab is the list order)
        assert np.array_equal(pbuf[:parts[0].shape[0]], parts[0])
        for name, f, o in zip(FIELD_NAMES, fields, out):
            stag = dict(exm=1, etm=1, brm=1, jxm=1, jtm=1).get(name.replace("_old", ""), 0)
            rows = slice(NG - stag, NG - stag + d.ny)
            assert np.array_equal(o[:, rows, NG:-NG], f[:, rows, NG:-NG]), name
            mask = np.ones(o.shape, dtype=bool)
            mask[:, rows, NG:-NG] = False
            assert np.all(o[mask] == 7.0 + 1.0j), name


def test_bad_descriptors_fail_loudly(tmp_path, cylgpu_lib):
    lib = _lib.load()
    desc = _lib.SdfDesc()
    fp = (C.c_void_p * 15)()
    pp = (C.c_void_p * 8)()
    assert lib.cylgpu_sdf_write_host(str(tmp_path / "x.sdf").encode(), C.byref(desc), fp, pp, None) != 0
    assert b"descriptor" in lib.cylgpu_last_error()
    assert lib.cylgpu_sdf_read_host(b"/nonexistent/file.sdf", C.byref(desc), fp, 0.0, 1.0, None, None) != 0
    assert b"cannot open" in lib.cylgpu_last_error()


@pytest.mark.parametrize("nranks", [1, 2])
def test_restart_through_the_file_continues_exactly_in_a_conducting_box(tmp_path, nranks, cylgpu_lib):
    """dump -> load -> continue on the oracle, through the real file: the format keeps rows 0..ny-1 of the
    r-staggered arrays and columns 1..nx (io/diagnostics.F90:2085-2105), everything else is re-derived by
    efield_bcs / bfield_bcs / current_finish after the load (as the reference's restart does).  With clamp on
    every wall that is lossless: the restarted world stays bit-identical to the uninterrupted one."""
    lib = _lib.load()
    d = decks.drift(nx=36, ny=12, n_mode=2)
    a = decks.make_oracle(d, nranks=nranks)
    b = decks.make_oracle(d, nranks=nranks)
    for w in (a, b):
        w.call("init_half_step")
        w.step(4)
    path = str(tmp_path / "restart.sdf")
    write_world(b, d, path)
    for k in range(nranks):
        info = b.rank_info(k)
        n0 = b.nparticles(k, 0)
        desc = make_desc(b, d, k, nranks, [0], [0], [0])
        loaded = [np.zeros_like(b.field(k, n)) for n in FIELD_NAMES]
        fp = (C.c_void_p * 15)(*[x.ctypes.data for x in loaded])
        pbuf = np.zeros((n0 + 8, 7))
        pp = (C.c_void_p * 8)(*([pbuf.ctypes.data] + [None] * 7))
        cap = (C.c_int64 * 8)(*([n0 + 8] + [0] * 7))
        rc = lib.cylgpu_sdf_read_host(path.encode(), C.byref(desc), fp, info["x_min_local"], info["x_max_local"], pp, cap)
        assert rc == 0, lib.cylgpu_last_error()
        for name, arr in zip(FIELD_NAMES, loaded):
            b.field(k, name)[...] = arr
        b.set_particles(k, 0, pbuf[:desc.npart_local[0]])
    b.call("efield_bcs")
    b.call("bfield_bcs")
    b.call("current_finish")
    a.step(3)
    b.step(3)
    for k in range(nranks):
        for name in FIELD_NAMES[:9]:
            assert np.array_equal(a.field(k, name), b.field(k, name)), name
        assert np.array_equal(a.particles(k, 0), b.particles(k, 0))


def test_reference_dump_bridge_self_check(tmp_path, cylgpu_lib):
    """tools/check_against_reference_dumps.py is the one-command way to pin the oracle against dumps of the
    real reference (which cannot be built here).  Self-check with two dumps of the oracle itself: the tool
    must load A, advance to B's step and report agreement to rounding in a conducting box."""
    import json
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import check_against_reference_dumps as bridge
    d = decks.drift(nx=36, ny=12, n_mode=2)
    w = decks.make_oracle(d)
    w.call("init_half_step")
    w.step(2)
    a, b = str(tmp_path / "a.sdf"), str(tmp_path / "b.sdf")

    def dump(path, step):
        st = write_world(w, d, path)
        # write_world stamps step 7; restamp through the descriptor of a second write
        lib = _lib.load()
        k, fields, parts = st[0]
        desc = make_desc(w, d, 0, 1, [parts[0].shape[0]], [0], [parts[0].shape[0]], step=step)
        fp = (C.c_void_p * 15)(*[f.ctypes.data for f in fields])
        pp = (C.c_void_p * 8)(*([parts[0].ctypes.data] + [None] * 7))
        assert lib.cylgpu_sdf_write_host(path.encode(), C.byref(desc), fp, pp, None) == 0

    dump(a, 2)
    w.step(5)
    dump(b, 7)
    sp = d.species[0]
    deck = dict(nx=d.nx, ny=d.ny, n_mode=d.n_mode, x_min=d.x_min, x_max=d.x_max, y_max=d.y_max,
                bc_field=list(d.bc_field), dt_multiplier=d.dt_multiplier, nranks=1,
                species=[dict(name="electron", charge=sp.charge, mass=sp.mass, bc_particle=list(sp.bc_particle))])
    deck_path = str(tmp_path / "deck.json")
    json.dump(deck, open(deck_path, "w"))
    assert bridge.main([deck_path, a, b, "--tol", "1e-12"]) == 0


def test_derived_variable_blocks(tmp_path, cylgpu_lib):
    """write_nspecies_field (io/diagnostics.F90:765-835,2222-2232,2396-2431): 'Derived/<Name>[/<species>]' blocks
    on the 'grid' mesh, cell centred, interior of the calc_df arrays -- here fed with the oracle's moments and
    read back through the reference's reader"""
    exe, _ = ref_tools()
    lib = _lib.load()
    d = decks.lwfa(nx=24, ny=10, n_mode=2, ppc_e=3, ppc_p=1)
    w = decks.make_oracle(d)
    w.call("init_half_step")
    w.step(3)
    counts = [w.nparticles(0, i) for i in range(2)]
    desc = make_desc(w, d, 0, 1, counts, [0, 0], counts)
    sel = {3: ("number_density", "Number_Density", "1/m^3", ("number_density", 0)),
           9: ("temperature", "Temperature", "K", ("temperature", 0)),
           14: ("jy", "Jy", "A/m^2", ("species_current", 2)),
           19: ("ekflux/x_min", "Particle_Energy_Flux/x_min", "W/m^2", ("ekflux", -1))}
    for v in sel:
        desc.derived_mask |= 1 << v
    desc.derived_sum, desc.derived_species = 1, 1
    assert lib.cylgpu_sdf_derived_count(C.byref(desc)) == 4 * 3
    arrays, expect = [], []
    for v in sorted(sel):
        bid, name, units, (kind, direction) = sel[v]
        for s in (-1, 0, 1):
            a = np.ascontiguousarray(w.moment(kind, s, direction)[0])
            arrays.append(a)
            suffix = "" if s < 0 else "/" + NAMES[s].decode()
            expect.append((bid + suffix, "Derived/" + name + suffix, units, a[NG:-NG, NG:-NG]))
    fields = [np.ascontiguousarray(w.field(0, n)) for n in FIELD_NAMES]
    parts = [np.ascontiguousarray(w.particles(0, i).reshape(-1, 7)) for i in range(2)]
    fp = (C.c_void_p * 15)(*[f.ctypes.data for f in fields])
    pp = (C.c_void_p * 8)(*([p.ctypes.data for p in parts] + [None] * 6))
    dp = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
    path = str(tmp_path / "derived.sdf")
    assert lib.cylgpu_sdf_write_host(path.encode(), C.byref(desc), fp, pp, dp) == 0, lib.cylgpu_last_error()
    hdr, blocks = parse_ref(exe, path, str(tmp_path / "out"))
    assert int(hdr["nblocks"]) == 41 + 12
    for n, (bid, name, units, want) in enumerate(expect, start=41):
        b = blocks[n]
        assert (b["id"], b["name"], b["units"], b["mesh"], b["stagger"], b["blocktype"], b["ndims"]) == \
            (bid, name, units, "grid", "0", "3", "2"), b
        assert b["dims"].startswith(f"{d.nx},{d.ny}")
        got = np.fromfile(os.path.join(str(tmp_path / "out"), f"{n}.bin")).reshape(d.ny, d.nx)
        assert np.array_equal(got, want), bid
    # number_density_mode: per species, complex, on the mode grid (io/diagnostics.F90:2596-2676)
    desc.derived_mask |= 1 << 22
    assert lib.cylgpu_sdf_derived_count(C.byref(desc)) == 4 * 3 + 2
    modes = [np.ascontiguousarray(w.number_density_modes(s)[0]) for s in range(2)]
    dp2 = (C.c_void_p * (len(arrays) + 2))(*([a.ctypes.data for a in arrays] + [m.ctypes.data for m in modes]))
    assert lib.cylgpu_sdf_write_host(path.encode(), C.byref(desc), fp, pp, dp2) == 0, lib.cylgpu_last_error()
    hdr, blocks = parse_ref(exe, path, str(tmp_path / "out2"))
    assert int(hdr["nblocks"]) == 41 + 12 + 4
    n = 53
    for s in range(2):
        for tag, arr in (("Real", modes[s].real), ("Imaginary", modes[s].imag)):
            b = blocks[n]
            nm = NAMES[s].decode()
            assert (b["id"], b["name"], b["units"], b["mesh"], b["stagger"], b["ndims"]) == \
                (f"number_density_mode/{nm}/{tag}"[:32], f"Number_Density_Mode/{nm}/{tag}", "1/m^3", "mode_grid", "0", "3")
            got = np.fromfile(os.path.join(str(tmp_path / "out2"), f"{n}.bin")).reshape(d.n_mode, d.ny, d.nx)
            assert np.array_equal(got, arr[:, NG:-NG, NG:-NG])
            n += 1
    # selected but not supplied: refused
    assert lib.cylgpu_sdf_write_host(path.encode(), C.byref(desc), fp, pp, None) != 0


def test_constant_blocks_round_trip(tmp_path, cylgpu_lib):
    """sdf_write_srl constants (io/diagnostics.F90:403-416: dt, window_shift_fraction, x_grid_min ...): read by the
    reference's reader with their values, and handed back by the product's reader for a restart"""
    exe, _ = ref_tools()
    lib = _lib.load()
    d = decks.drift(nx=20, ny=10, n_mode=1)
    w = decks.make_oracle(d)
    n0 = w.nparticles(0, 0)
    desc = make_desc(w, d, 0, 1, [n0], [0], [n0])
    consts = [(b"dt", b"Time increment", 1.25e-16), (b"window_shift_fraction", b"Window Shift Fraction", 0.375),
              (b"x_grid_min", b"Minimum grid position", -3.5e-6)]
    desc.n_constants = len(consts)
    for k, (bid, name, val) in enumerate(consts):
        desc.constant_id[k], desc.constant_name[k], desc.constant_value[k] = bid, name, val
    fields = [np.ascontiguousarray(w.field(0, n)) for n in FIELD_NAMES]
    parts = [np.ascontiguousarray(w.particles(0, 0).reshape(-1, 7))]
    fp = (C.c_void_p * 15)(*[f.ctypes.data for f in fields])
    pp = (C.c_void_p * 8)(*([parts[0].ctypes.data] + [None] * 7))
    path = str(tmp_path / "c.sdf")
    assert lib.cylgpu_sdf_write_host(path.encode(), C.byref(desc), fp, pp, None) == 0, lib.cylgpu_last_error()
    hdr, blocks = parse_ref(exe, path, str(tmp_path / "out"))
    for k, (bid, name, val) in enumerate(consts):
        b = blocks[k]
        assert (b["id"], b["name"], b["blocktype"], b["datatype"]) == (bid.decode(), name.decode(), "5", "4")
        assert float(b["const"]) == val
    assert blocks[len(consts)]["id"] == "grid"
    # reading back: listed ids are filled and flagged, an id the file does not hold is not
    rd = make_desc(w, d, 0, 1, [0], [0], [0])
    want = [b"x_grid_min", b"no_such_constant", b"dt"]
    rd.n_constants = 3
    for k, bid in enumerate(want):
        rd.constant_id[k] = bid
    out = [np.zeros_like(f) for f in fields]
    op = (C.c_void_p * 15)(*[a.ctypes.data for a in out])
    assert lib.cylgpu_sdf_read_host(path.encode(), C.byref(rd), op, -1.0, 1.0, None, None) == 0, lib.cylgpu_last_error()
    assert rd.constants_found == 0b101
    assert rd.constant_value[0] == -3.5e-6 and rd.constant_value[2] == 1.25e-16
