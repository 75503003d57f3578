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
           "insert_kernel.cuh", "push_v0.cuh", "pbcs_kernels.cuh", "field_kernels.cuh", "compact_kernels.cuh", "field_ranges.cuh", os.path.join("..", "..", "include", "cylgpu.h")]
LIB = os.path.join(HERE, "libcylgpu.so")
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


def build(force=False, verbose=False):
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC] + FLAGS + EXTRA.get(s, []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if force or procs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-ldl", "-lpthread"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
