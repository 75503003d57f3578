#!/usr/bin/env python
"""bench.py -- throughput of the cylindrical-EPOCH per-timestep PIC hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one pass of the hot path in driver order (epoch2d.F90:189-266):
update_eb_fields_half -> push_particles (gather + Boris + deposit + particle_bcs) ->
current_finish -> update_eb_fields_final -> moving_window, on synthetic plasma of a BASELINE.json shape.

Default workload = BASELINE.json configs[2], the configuration the north-star target is quoted on: laser
wakefield acceleration, m = 0..1, 8192 x 512 grid, 32 ppc (134 M macro-particles), window moving at c.
N > 1: the SAME grid z-decomposed into N x-slabs (strong scaling, 8192 / N x 512 cells per GPU, the
reference's nprocx = N, nprocy = 1 of mpi_routines.F90:312-337), NCCL send/recv for field halos, additive
J ghosts and migrating particles.  `--workload lwfa_4096x1024_m2_ppc60` is the weak-scaling C5 line
(4096 x 1024 cells and 2.5e8 particles per GPU, 2e9 on 8); the thermal / five-mode workloads keep one slab
per GPU in a periodic ring (weak).

Prints ONE JSON line (rank 0).  `value` is whole-job particle-steps/s with all state resident in HBM; `e2e`
is the same step driven through the C-ABI with HOST buffers: the particle list and the nine E/B/J mode arrays
live in pinned host memory and make the round trip every step (cylgpu_push_host streams the list through the
GPU in chunks, upload | push | download overlapped; a moving window over several slabs: upload -> step ->
download).  `cpu_baseline` / `--impl reference`: the CPU restatement of the reference (oracle port; the
Fortran + MPI reference cannot be built in this image) on the same deck shape, all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # name: nx, ny, n_mode, ppc, kind, scaling.  scaling "strong": nx is the GLOBAL grid, split over the GPUs;
    # "weak": nx is the slab of ONE GPU.
    # BASELINE.json configs[2]: laser wakefield, moving window at c, open boundaries -- the headline
    "lwfa_8192x512_m2_ppc32": dict(nx=8192, ny=512, n_mode=2, ppc=32, kind="lwfa", scaling="strong"),
    # BASELINE.json configs[4], weak scaling: 4096 x 1024 per GPU at 60 ppc = 2.5e8 particles per GPU (2e9 on 8)
    "lwfa_4096x1024_m2_ppc60": dict(nx=4096, ny=1024, n_mode=2, ppc=60, kind="lwfa", scaling="weak"),
    "lwfa_1024x128_m2_ppc16": dict(nx=1024, ny=128, n_mode=2, ppc=16, kind="lwfa", scaling="strong"),
    # BASELINE.json configs[1]: thermal periodic plasma (energy-conservation / deposit-correctness check)
    "thermal_2048x256_m2_ppc64": dict(nx=2048, ny=256, n_mode=2, ppc=64, kind="thermal", scaling="weak"),
    "thermal_1024x256_m2_ppc64": dict(nx=1024, ny=256, n_mode=2, ppc=64, kind="thermal", scaling="weak"),
    "thermal_512x128_m2_ppc16": dict(nx=512, ny=128, n_mode=2, ppc=16, kind="thermal", scaling="weak"),
    # BASELINE.json configs[3]: m = 0..4.  dt_multiplier 0.5: the reference's per-mode FDTD is unstable on the
    # axis rows for m = 4 at the default 0.95 (tests/test_oracle.py::test_high_modes_need_a_smaller_dt_multiplier)
    "modes5_4096x512_m5_ppc16": dict(nx=4096, ny=512, n_mode=5, ppc=16, kind="thermal", scaling="weak",
                                     dt_multiplier=0.5),
}
DEFAULT_WORKLOAD = "lwfa_8192x512_m2_ppc32"

TEMP_K = 1.16e7        # ~1 keV
DENSITY = 1.0e24       # m^-3
DXY = 0.5e-6
LWFA_LAMBDA = 0.8e-6   # scaled Wakefield_Lifschitz09 deck: dx = lambda/25, dy = lambda/3, n = 7.5e24 m^-3
LWFA_DENSITY = 7.5e24
LWFA_INTENSITY = 3.4e18


def algorithmic_bytes_per_particle_step(n_mode, ppc):
    """7 doubles read + 6 written per particle, plus 6M complex field reads and
    3M complex J writes per cell amortised over ppc."""
    return 104.0 + 144.0 * n_mode / ppc


def ce_shape():
    """particle shape of the library in use (CYL_SHAPE: triangle unless a top-hat / B-spline build is selected)"""
    from cylindrical_epoch_b200.constants import SHAPE
    return SHAPE


def committed_ncu(workload):
    """ncu --set full capture of the dominant kernel on this workload, committed under profiles/
    (bench.py cannot run under a profiler): DRAM traffic per launch and the pipe utilisations
    that explain the roofline fraction."""
    if ce_shape() != "triangle":
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "push_kernel_ncu.json")) as f:
            d = json.load(f)
        return d if d.get("workload") == workload else None
    except Exception:   # noqa: BLE001
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"], "measured"
    except Exception:   # noqa: BLE001
        return 6650.0, "fallback"


def thermal_particles(rng, nx, ny, ppc, dx, dy, x_grid_min_local, mass, temp_k=None, density=None):
    """Uniform thermal load in the shape of helper.F90:552-583 / particle_temperature.F90:388-398
    (positions uniform in x and r per cell, theta uniform, weight = n 2 pi r dx dy / ppc).
    numpy RNG -- the bench does not need the reference's KISS stream."""
    from cylindrical_epoch_b200.constants import KB
    n = nx * ny * ppc
    out = np.empty((n, 7), dtype=np.float64)
    ix = np.repeat(np.tile(np.arange(nx, dtype=np.float64), ny), ppc)
    iy = np.repeat(np.arange(ny, dtype=np.float64), nx * ppc)
    out[:, 0] = x_grid_min_local + ix * dx + (rng.random(n) - 0.5) * dx
    r = 0.5 * dy + iy * dy + (rng.random(n) - 0.5) * dy
    th = 2.0 * np.pi * rng.random(n)
    out[:, 1] = r * np.cos(th)
    out[:, 2] = r * np.sin(th)
    sd = np.sqrt((TEMP_K if temp_k is None else temp_k) * KB * mass)
    out[:, 3:6] = rng.normal(0.0, sd, size=(n, 3)) if sd > 0 else 0.0
    out[:, 6] = (DENSITY if density is None else density) * 2.0 * np.pi * dx * dy * r / ppc
    return out


class ClockSampler:
    """SM clock + throttle reasons of this rank's GPU sampled DURING the timed region (the numbers
    `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints, read through NVML in-process:
    one nvidia-smi per rank enumerating all GPUs under the driver lock stalled 4- and 8-rank runs)."""

    def __init__(self, index=0):
        self.index = index
        self.samples = []      # (sm_mhz, sm_max_mhz, reasons bitmask)
        self.on = False
        self.nvml = None
        self.handle = None
        self.thread = None
        try:   # in-process NVML: no nvidia-smi start-up (which enumerates every GPU under the driver lock)
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:   # noqa: BLE001
            self.nvml = None

    @staticmethod
    def _physical_index(local):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if local < len(ids) and ids[local].isdigit():
                return int(ids[local])
        return local

    def _loop(self):
        nv = self.nvml
        while self.on:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((sm, rs))
            except Exception:   # noqa: BLE001
                pass
            time.sleep(0.02)

    def start(self):
        if self.nvml is None:
            return
        self.on = True
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        self.on = False
        if self.thread:
            self.thread.join(timeout=1.0)
        nv = self.nvml
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        reasons = sorted(nm for nm, bit in names.items() if any(rs & bit for _, rs in self.samples))
        sm = [v for v, _ in self.samples]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(sm), "source": "NVML, sampled every 20 ms during the timed region"}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------
# reference arm: the CPU restatement of the reference (oracle port) on the host cores
# ------------------------------------------------------------------------------------------
def host_threads():
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:   # noqa: BLE001
        pass
    return max(1, min(cores, 64))


def oracle_deck(wl, nx_sample):
    """the workload's deck shape (ny, n_mode, ppc, species, boundaries, laser, window) on an x extent of
    nx_sample cells: what the CPU arm runs"""
    import decks
    if wl["kind"] == "lwfa":
        d = decks.lwfa(nx=nx_sample, ny=wl["ny"], n_mode=wl["n_mode"], ppc_e=wl["ppc"], ppc_p=0, window=True,
                       t_centre=30e-15)
        d.window_v_x = 2.99792458e8
    else:
        d = decks.thermal(nx=nx_sample, ny=wl["ny"], n_mode=wl["n_mode"], ppc=wl["ppc"], temp_k=TEMP_K,
                          density=DENSITY)
    d.dt_multiplier = wl.get("dt_multiplier", 0.95)
    return d


def cpu_port_rate(wl, steps=3, warmup=1, cells_per_thread=64, max_seconds=60.0):
    """particle-steps/s and cell-mode-updates/s of the oracle (kind "port": the reference is Fortran + MPI
    and cannot be compiled in this image) on a BOUNDED sample of the workload: same deck (ny, n_mode, ppc,
    species, boundaries, laser and moving window), x extent cut to `cells_per_thread` cells per host thread,
    one x-slab per thread -- the reference's own MPI decomposition (nprocx = threads, nprocy = 1), its ranks
    played by the threads of an OpenMP loop."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import decks
    import pyoracle
    nranks = host_threads()
    # torchrun exports OMP_NUM_THREADS=1 to its children: the thread count is set explicitly
    threads = pyoracle.set_threads(nranks)
    nx_sample = cells_per_thread * nranks
    d = oracle_deck(wl, nx_sample)
    t0 = time.time()
    w = decks.make_oracle(d, nranks=nranks)
    w.call("init_half_step")
    t_load = time.time() - t0
    for _ in range(warmup):
        w.call("step")
    t_push = t_fields = 0.0
    done = 0
    psteps = 0
    t_start = time.time()
    for _ in range(steps):
        psteps += sum(w.nparticles(k, 0) for k in range(nranks))
        tf = w.call("fields_half")
        tp = w.call("push")
        w.call("current_finish")
        w.call("advance_half_time")
        w.call("flush_rng")
        w.call("advance_half_time")
        tf2 = w.call("fields_final")
        w.call("moving_window")
        t_push += tp
        t_fields += tf + tf2
        done += 1
        if time.time() - t_start > max_seconds:
            break
    total = time.time() - t_start
    npart = psteps // max(done, 1)
    return dict(value=psteps / total, unit="particle-steps/s", cores=threads, kind="port",
                sample=f"{wl['kind']} deck, {nx_sample}x{wl['ny']} cells, m=0..{wl['n_mode'] - 1}, {wl['ppc']} ppc = "
                       f"{npart} particles, {done} whole steps in {total:.2f} s after {warmup} warm-up, {nranks} x-slabs "
                       f"on {threads} OpenMP threads (load {t_load:.1f}s untimed)",
                steps_timed=done, seconds=total, ms_per_step=1e3 * total / max(done, 1),
                field_cell_mode_updates_per_s=nx_sample * wl["ny"] * wl["n_mode"] * done / max(t_fields, 1e-9),
                push_only_particle_steps_per_s=psteps / max(t_push, 1e-9))


def run_reference(args, wl_name, wl):
    """--impl reference: the CPU arm alone.  Under torchrun rank 0 runs it, the other ranks exit at once."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    # each "step" of this arm is one whole step of the bounded sample; a few of them bound the run to ~a minute
    steps = max(1, min(args.steps, 5))
    warmup = max(0, min(args.warmup, 1))
    res = cpu_port_rate(wl, steps=steps, warmup=warmup, max_seconds=120.0)
    line = {
        "impl": "reference", "metric": "particle-steps/sec (push+gather+deposit)", "value": res["value"],
        "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": res["steps_timed"], "warmup": warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": wl["scaling"],
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl_name, "grid": [wl["nx"], wl["ny"]], "n_mode": wl["n_mode"], "ppc": wl["ppc"],
                   "kind": wl["kind"], "steps_requested": args.steps,
                   "note": "CPU restatement of the reference (oracle port; the Fortran + MPI reference cannot be built "
                           "in this image), same deck shape, x extent cut to a bounded sample: rates are per "
                           "particle-step, so the sample stands for the full grid"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "field_cell_mode_updates_per_s": res["field_cell_mode_updates_per_s"],
        "e2e": {"value": res["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def make_slab(wl, rank, world, local, uid):
    """the product slab of this rank for the workload (hotpath.Slab mirrors the reference's driver)"""
    import cylindrical_epoch_b200 as ce
    from cylindrical_epoch_b200.constants import (BC_OPEN, BC_PERIODIC, BC_REFLECT, BC_SIMPLE_LASER, BC_ZERO_B, BD_X_MIN,
                                                  C_LIGHT, EPSILON0, M0, Q0, TRANSPORT_NCCL, TRANSPORT_NONE)
    ny, M, ppc = wl["ny"], wl["n_mode"], wl["ppc"]
    nxg = wl["nx"] if wl["scaling"] == "strong" else wl["nx"] * world
    transport = TRANSPORT_NCCL if world > 1 else TRANSPORT_NONE
    dtm = wl.get("dt_multiplier", 0.95)
    if wl["kind"] == "lwfa":
        # scaled Wakefield_Lifschitz09 deck (example_decks): dx = lambda/25, dy = lambda/3, n = 7.5e24 m^-3,
        # a0 ~ 1.26 pulse from x_min, open boundaries, window moving at c from t = 0, cold electrons
        dx_, dy_ = LWFA_LAMBDA / 25.0, LWFA_LAMBDA / 3.0
        open4 = (BC_OPEN, BC_OPEN, BC_OPEN, BC_OPEN)
        species = [ce.Species(-Q0, M0, open4, False, False, ppc, LWFA_DENSITY, (0.0,) * 3)]
        amp = 100.0 * np.sqrt(LWFA_INTENSITY / (C_LIGHT * EPSILON0 / 2.0))
        lasers = [ce.Laser(boundary=BD_X_MIN, amp=amp, omega=2.0 * np.pi * C_LIGHT / LWFA_LAMBDA, t_centre=30e-15,
                           t_width=10e-15, r_width=min(5.0e-6, 0.4 * ny * dy_))]
        slab = ce.Slab(nxg, ny, M, 0.0, nxg * dx_, ny * dy_, [BC_SIMPLE_LASER, BC_OPEN, 0, BC_OPEN], species,
                       rank=rank, nranks=world, transport=transport, device=local, nccl_unique_id=uid,
                       lasers=lasers, move_window=True, window_v_x=C_LIGHT, window_start_time=0.0,
                       dt_multiplier=dtm)
        return slab, LWFA_DENSITY, 0.0
    bcp = (BC_PERIODIC, BC_PERIODIC, BC_OPEN, BC_REFLECT)
    species = [ce.Species(-Q0, M0, bcp, False, bool(int(os.environ.get('BENCH_ZERO_CURRENT', '0'))), ppc, DENSITY,
                          (TEMP_K,) * 3)]
    slab = ce.Slab(nxg, ny, M, 0.0, nxg * DXY, ny * DXY, [BC_PERIODIC, BC_PERIODIC, 0, BC_ZERO_B], species,
                   rank=rank, nranks=world, transport=transport, device=local, nccl_unique_id=uid, dt_multiplier=dtm)
    return slab, DENSITY, TEMP_K


def run_ours(args, wl_name, wl):
    import ctypes as C
    import torch
    from cylindrical_epoch_b200 import build as cbuild
    from cylindrical_epoch_b200.constants import FIELD_NAMES, M0
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    if not os.path.exists(os.path.join(ROOT, "cylindrical_epoch_b200", "libcylgpu.so")):
        cbuild.build()
    torch.cuda.set_device(local)
    dist = None
    uid = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from cylindrical_epoch_b200 import _lib
        buf = C.create_string_buffer(128)
        if rank == 0:
            rc = _lib.load().cylgpu_nccl_unique_id(buf)
            assert rc == 0, _lib.load().cylgpu_last_error()
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        uid = bytes(t.cpu().numpy().tobytes())

    ny, M, ppc = wl["ny"], wl["n_mode"], wl["ppc"]
    lwfa = wl["kind"] == "lwfa"
    slab, dens, temp = make_slab(wl, rank, world, local, uid)
    g = slab.grid
    nx, nxg = g.nx, g.nx_global
    rng = np.random.default_rng(7842432 + rank)
    # pinned host image of the particle list (also the e2e upload source)
    n0 = nx * ny * ppc
    host_p = torch.empty((n0 + n0 // 8 + 4 * ny * ppc, 7), dtype=torch.float64, pin_memory=True)
    hp = host_p.numpy()
    hp[:n0] = thermal_particles(rng, nx, ny, ppc, g.dx, g.dy, g.x_grid_min_local, M0, temp, dens)
    slab.upload_particles(0, hp[:n0])
    # run the library on a torch-owned stream so that torch.cuda.Event brackets its work
    tstream = torch.cuda.Stream(device=local)
    slab.set_stream(tstream.cuda_stream)
    if dist is not None:
        dist.barrier()   # the ranks' host-side set-up took different times; the first neighbour exchange is next
    slab.init_half_step()
    slab.L.cylgpu_set_timing(slab.h, 1)
    if "BENCH_VARIANT" in os.environ:
        slab.set_push_variant(int(os.environ["BENCH_VARIANT"]))
    if "BENCH_EXCHANGE_CAPACITY" in os.environ:   # 0 = the exact count-then-data protocol (two host syncs per step)
        slab.set_exchange_capacity(int(os.environ["BENCH_EXCHANGE_CAPACITY"]))
    if "BENCH_SORT_INTERVAL" in os.environ:
        slab.set_sort_interval(int(os.environ["BENCH_SORT_INTERVAL"]))

    def barrier():
        slab.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        slab.synchronize()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        slab.step_once()
    e_f0, e_k0 = slab.energy()
    shifts0 = slab.window_shifts_total
    slab.reset_stats()
    clocks = ClockSampler(local) if rank == 0 else None   # rank 0's GPU stands for the box
    barrier()
    n_start = slab.particle_count(0)
    if clocks:
        clocks.start()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record(tstream)
    for _ in range(args.steps):
        slab.step_once()
    ev1.record(tstream)
    barrier()
    wall_host = time.perf_counter() - t0
    wall = ev0.elapsed_time(ev1) * 1e-3   # device time, CUDA events on the launching stream
    clk = clocks.stop() if clocks else None
    st = slab.stats()
    n_end = slab.particle_count(0)
    # particle-steps of the timed region: the list changes by a few columns per step at most (window, migration);
    # the mean of the counts at both ends times the steps (no per-step count read: that would be a host sync)
    n_steps_particles = 0.5 * (n_start + n_end) * args.steps
    if os.environ.get("BENCH_RANK_PHASES"):
        sys.stderr.write("rank %d: wall %.3f ms/step fields %.3f push %.3f (kernel %.3f sort %.3f) bcs %.3f exchange %.3f\n" % (
            rank, 1e3 * ev0.elapsed_time(ev1) * 1e-3 / args.steps, st.ms_fields / args.steps, st.ms_push / args.steps,
            st.ms_push_kernel / args.steps, st.ms_sort / args.steps, st.ms_bcs / args.steps, st.ms_exchange / args.steps))
    e_f1, e_k1 = slab.energy()
    # max over ranks of the device-event time; particle-steps summed over ranks
    tt = torch.tensor([wall, float(n_steps_particles)], dtype=torch.float64, device="cuda")
    if dist is not None:
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        wall = float(tmax[0])
        total_psteps = float(tsum[1])
    else:
        total_psteps = float(n_steps_particles)
    value = total_psteps / wall
    field_rate = nxg * ny * M * args.steps / (st.ms_fields * 1e-3) if st.ms_fields > 0 else None

    # roofline of the dominant kernel (fused push + gather + deposit), per launch
    peak, peak_kind = measured_peaks()
    bp = algorithmic_bytes_per_particle_step(M, ppc)
    roof = None
    if st.n_push_kernel > 0:
        dur = st.ms_push_kernel * 1e-3 / st.n_push_kernel
        per_launch = n_steps_particles / args.steps
        achieved = bp * per_launch / dur / 1e9
        ncu = committed_ncu(wl_name)
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu["dram_bytes_per_launch"] if ncu else None, "peak_source": peak_kind,
                "kernel": ("k_push_v2 (strip CTAs: gather + Boris + DMMA deposit)" if ce_shape() == "triangle" else
                           "k_push_generic (one thread per particle, %s shape: gather + Boris + deposit by L2 reductions)" % ce_shape()),
                "algorithmic_bytes_per_launch": bp * per_launch,
                "binding_pipes_from_ncu": ({k: ncu[k] for k in ncu if k not in ("workload", "dram_bytes_per_launch")}
                                           if ncu else None),
                "note": "FP64 arithmetic as the reference prescribes it puts the FP64-pipe floor of this kernel several "
                        "times above its HBM floor; it is co-limited by the shared-memory data pipe, the FP64 pipe and "
                        "issue, see DESIGN.md 3.1",
                "kernel_ms_per_launch": dur * 1e3, "algorithmic_bytes_per_particle_step": bp,
                "kernel_particle_steps_per_s": per_launch / dur,
                "kernel_share_of_step": (st.ms_push_kernel / args.steps) / (1e3 * ev0.elapsed_time(ev1) * 1e-3 / args.steps)}

    line = {
        "metric": "particle-steps/sec (push+gather+deposit)", "value": value, "unit": "particle-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl_name, "particle_shape": ce_shape(), "grid_global": [nxg, ny], "grid_per_gpu": [nx, ny], "n_mode": M, "ppc": ppc,
                   "particles_per_gpu": n0, "dt_multiplier": wl.get("dt_multiplier", 0.95),
                   "decomposition": f"{world} x-slabs, " + ("open ends, moving window (%d shifts in the timed steps)"
                                                            % (slab.window_shifts_total - shifts0) if lwfa
                                                            else "periodic ring"),
                   "l2": "inputs (%.1f GB of particles per GPU) exceed the 126 MB L2; no flush needed" % (n0 * 56 / 1e9),
                   "timing": "CUDA events on the library's stream, max over ranks", "host_wall_s": wall_host,
                   "exchange_capacity": slab.exchange_capacity,
                   "neighbour_links": dict(zip(("transport", "left_mailbox", "right_mailbox", "mailbox_slot_KiB"),
                                               slab.transport_info()))},
        "field_cell_mode_updates_per_s": field_rate,
        "phase_ms_per_step": {"fields": st.ms_fields / args.steps, "push_total": st.ms_push / args.steps,
                              "push_kernel": st.ms_push_kernel / args.steps, "sort": st.ms_sort / args.steps,
                              "current_finish": st.ms_bcs / args.steps,
                              "neighbour_exchanges_incl_waiting": st.ms_exchange / args.steps},
        "energy": {"field_J": e_f1, "kinetic_J": e_k1,
                   "relative_drift_over_timed_steps": (abs((e_f1 + e_k1) - (e_f0 + e_k0)) / (e_f0 + e_k0)
                                                       if not lwfa else None),
                   "note": ("open system: the laser enters and the window inserts / drops plasma, no conserved total"
                            if lwfa else "closed periodic box")},
        "roofline": roof, "clocks": clk, "gpu_launches": int(st.kernel_launches),
    }

    # ---- e2e: host-authoritative round trip through the C-ABI every step ----
    if not args.no_e2e:
        names = FIELD_NAMES[:9]
        host_f = {nm: torch.empty(slab.field_shape, dtype=torch.complex128, pin_memory=True) for nm in names}
        for nm in names:
            host_f[nm].numpy()[...] = slab.download_field(nm)
        nn = C.c_int64()
        slab._ck(slab.L.cylgpu_download_particles(slab.h, 0, hp.shape[0], hp.ctypes.data, C.byref(nn)))
        e2e_steps = max(1, min(args.steps, 3 if lwfa else 5))
        h2d = d2h = 0
        nfield_bytes = len(names) * host_f[names[0]].numel() * 16
        if lwfa and world > 1:
            # moving window over several slabs: the list is uploaded, stepped on the device (the window drops the
            # plasma behind it, appends the new column, particles migrate) and downloaded again, every step
            count = [int(nn.value)]

            def e2e_step():
                n_before = count[0]
                for nm in names:
                    slab._ck(slab.L.cylgpu_upload_field(slab.h, FIELD_NAMES.index(nm), host_f[nm].numpy().ctypes.data))
                slab._ck(slab.L.cylgpu_upload_particles(slab.h, 0, n_before, hp.ctypes.data))
                slab.step_once()
                slab._ck(slab.L.cylgpu_download_particles(slab.h, 0, hp.shape[0], hp.ctypes.data, C.byref(nn)))
                for nm in names:
                    slab._ck(slab.L.cylgpu_download_field(slab.h, FIELD_NAMES.index(nm), host_f[nm].numpy().ctypes.data))
                count[0] = int(nn.value)
                return n_before, count[0]
            what = ("particle list and the 9 E/B/J mode arrays live in pinned host memory: every step uploads them "
                    "(cylgpu_upload_particles / _field), runs the full step incl. the moving window on the device and "
                    "downloads them")
        else:
            # from here on the particle list lives in the pinned host array and is streamed through
            # the GPU by cylgpu_push_host (upload | push + deposit + particle_bcs | download overlap)
            slab.attach_host_lists([hp], [int(nn.value)])
            slab.L.cylgpu_upload_particles(slab.h, 0, 0, hp.ctypes.data)   # nothing stays on the device

            def e2e_step():
                n_before = slab.host_counts[0]
                for nm in names:
                    slab._ck(slab.L.cylgpu_upload_field(slab.h, FIELD_NAMES.index(nm), host_f[nm].numpy().ctypes.data))
                slab.step_once()
                for nm in names:
                    slab._ck(slab.L.cylgpu_download_field(slab.h, FIELD_NAMES.index(nm), host_f[nm].numpy().ctypes.data))
                return n_before, slab.host_counts[0]
            what = ("particle list and the 9 E/B/J mode arrays live in pinned host memory: every step uploads them, runs "
                    "the full step and downloads them (cylgpu_push_host streams the list in chunks, both PCIe "
                    "directions busy" + ("; the window's new column joins the host list, the plasma behind the window "
                                         "is dropped as the list streams through" if lwfa else "") + ")")
        e2e_step()   # warm-up: staging buffers, streams
        barrier()
        t0 = time.perf_counter()
        psteps = 0
        for _ in range(e2e_steps):
            n_before, n_after = e2e_step()
            psteps += n_before
            h2d = n_before * 56 + nfield_bytes
            d2h = n_after * 56 + nfield_bytes
        barrier()
        w2 = time.perf_counter() - t0
        tt = torch.tensor([w2, float(psteps)], dtype=torch.float64, device="cuda")
        if dist is not None:
            a = tt.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
            b = tt.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
            w2, psteps = float(a[0]), float(b[1])
        line["e2e"] = {"value": psteps / w2, "unit": "particle-steps/s", "h2d_bytes_per_step": int(h2d),
                       "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": 1e3 * w2 / e2e_steps,
                       "what": what}

    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cb = cpu_port_rate(wl, steps=3, warmup=1, max_seconds=30.0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"]["field_cell_mode_updates_per_s"] = cb["field_cell_mode_updates_per_s"]
        except Exception as e:   # noqa: BLE001
            line["cpu_baseline"] = {"error": str(e)}
    slab.close()
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    wl = WORKLOADS[args.workload]
    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner,
    # torchrun notices) goes to stderr while the run is in progress
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args, args.workload, wl)
    else:
        run_ours(args, args.workload, wl)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
