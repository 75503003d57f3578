"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports
exactly the symbols include/cylgpu.h declares; compute calls fail loudly without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "cylgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cylgpu_[a-z0-9_]+)\s*\(", src)) - {"cylgpu_sendrecv_fn"})


def test_library_exports_every_header_symbol(cylgpu_lib):
    from cylindrical_epoch_b200 import _lib
    names = header_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(cylgpu_lib, n), f"libcylgpu.so does not export {n}"
    # and the ctypes table binds every one of them (no stale entries either way)
    assert sorted(_lib.SYMBOLS) == names


def test_struct_layouts_match_header(cylgpu_lib):
    from cylindrical_epoch_b200 import _lib
    # sizes follow from the field lists in the header (LP64)
    assert ctypes.sizeof(_lib.SpeciesC) == 16 + 16 + 8
    assert ctypes.sizeof(_lib.Stats) == 8 * (8 + 7) + 8 * 5 + 16
    assert ctypes.sizeof(_lib.Config) == 4 * 16 + 8 * 10 + 8 * 5


def test_struct_layouts_match_the_c_compiler(cylgpu_lib, tmp_path):
    """sizeof and the offset of every member of the structs that cross the boundary, as gcc lays out include/cylgpu.h,
    against the ctypes mirrors of _lib.py (what a BIND(C) type of the Fortran shim has to reproduce as well)"""
    import subprocess
    from cylindrical_epoch_b200 import _lib
    structs = {"cylgpu_config": _lib.Config, "cylgpu_species": _lib.SpeciesC, "cylgpu_stats_t": _lib.Stats,
               "cylgpu_laser": _lib.LaserC, "cylgpu_insert_profile": _lib.InsertProfileC,
               "cylgpu_driver_config": _lib.DriverConfig, "cylgpu_driver_state": _lib.DriverState,
               "cylgpu_sdf_desc": _lib.SdfDesc}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "cylgpu.h"', "int main(void) {"]
    for cname, ct in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(ln.split() for ln in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, ct in structs.items():
        assert int(got[cname]) == ctypes.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, (cname, fname)


def test_no_cpu_fallback(cylgpu_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cylindrical_epoch_b200 as ce
    with pytest.raises(ce.CylGpuError, match="no CPU fallback"):
        ce.Slab(64, 32, 2, 0.0, 1e-5, 1e-5, [3, 5, 0, 5], [ce.Species(-1.6e-19, 9.1e-31)])


def test_product_never_touches_the_oracle():
    """the product package and its CUDA sources must not import/link/include oracle/"""
    pkg = os.path.join(ROOT, "cylindrical_epoch_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "cyl_oracle" not in txt, f
                assert not re.search(r"(import|from|include)\s+.*\boracle\b", txt), f


def test_integration_shim_binds_only_what_the_library_exports():
    """every BIND(C, NAME='cylgpu_...') of the Fortran shim in INTEGRATION.md names an entry point of include/cylgpu.h
    (the shim cannot be compiled in this image, so at least its names are held against the header)"""
    import re
    from cylindrical_epoch_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    bound = set(re.findall(r"NAME='(cylgpu_[a-z_0-9]+)'", text))
    assert len(bound) >= 30
    header = open(os.path.join(root, "include", "cylgpu.h")).read()
    for name in sorted(bound):
        assert name in _lib.SYMBOLS, name
        assert re.search(r"\b%s\s*\(" % name, header), name
    # ... with as many dummy arguments as the C prototype has parameters
    flat = re.sub(r"&\s*\n\s*", " ", text)
    n_checked = 0
    for m in re.finditer(r"FUNCTION\s+\w+\s*\(([^)]*)\)\s*(?:RESULT\(\w+\)\s*)?BIND\(C,\s*NAME='(cylgpu_\w+)'\)", flat):
        fargs = [a for a in m.group(1).split(",") if a.strip()]
        proto = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % m.group(2), header, re.S)
        cargs = [a for a in proto.group(1).split(",") if a.strip() and a.strip() != "void"]
        assert len(fargs) == len(cargs), (m.group(2), fargs, cargs)
        n_checked += 1
    assert n_checked >= 30
