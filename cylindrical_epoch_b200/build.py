"""Builds cylindrical_epoch_b200/libcylgpu.so in-tree with nvcc for sm_100a.

No CMake, no JIT cache: the .so sits next to the package so that it travels to the GPU box
with the repo snapshot.  `python -m cylindrical_epoch_b200.build [--force]`.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["api.cu", "fields.cu", "bcs.cu", "particles.cu", "transport.cu", "window_insert.cu", "sdf_io.cu", "driver.cu", "balance.cu"]
HEADERS = ["ctx.cuh", "push.cuh", "deposit_mma.cuh", "moments.cuh", "philox.cuh", "geom.cuh", "bc_kernels.cuh", "moments_kernels.cuh",
           "insert_kernel.cuh", "push_v0.cuh", "pbcs_kernels.cuh", "field_kernels.cuh", "compact_kernels.cuh", "field_ranges.cuh", "shape.cuh", "push_shapes.cuh", os.path.join("..", "..", "include", "cylgpu.h")]
LIB = os.path.join(HERE, "libcylgpu.so")
# The particle shape is a compile-time choice (csrc/shape.cuh, as -DPARTICLE_SHAPE_* is in the reference): one
# library per shape, libcylgpu.so (triangle), libcylgpu_tophat.so, libcylgpu_bspline3.so; CYL_SHAPE selects at run time
SHAPES = {"triangle": 0, "tophat": 1, "bspline3": 2}
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall"] + os.environ.get("CYLGPU_DEFS", "").split()


# window_insert.cu: the device-side column must give the oracle's weights and x positions bit for bit,
# so its one kernel is compiled without FMA contraction (it runs once per window shift)
EXTRA = {"window_insert.cu": ["-fmad=false"]}


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def lib_path(shape="triangle"):
    return LIB if shape == "triangle" else os.path.join(HERE, f"libcylgpu_{shape}.so")


def build(force=False, verbose=False, shape="triangle"):
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    procs = []
    bdir = os.path.join(HERE, "build") if shape == "triangle" else os.path.join(HERE, "build", shape)
    lib = lib_path(shape)
    defs = [] if shape == "triangle" else [f"-DCYL_SHAPE={SHAPES[shape]}"]
    os.makedirs(bdir, exist_ok=True)
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(bdir, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC] + FLAGS + defs + EXTRA.get(s, []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if force or procs or _stale(lib, objs):
        cmd = [NVCC, "-shared", "-o", lib] + objs + ["-ldl", "-lpthread"]
        subprocess.check_call(cmd)
    return lib


def build_all(force=False, verbose=False):
    return [build(force, verbose, shape) for shape in SHAPES]


if __name__ == "__main__":
    if "--all" in sys.argv:
        print("\n".join(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)))
    else:
        shape = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--shape=")), "triangle")
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, shape=shape))
