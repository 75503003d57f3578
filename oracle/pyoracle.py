"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (cylindrical_epoch_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# The particle shape is a compile-time choice (cyl_oracle.hpp CYLO_SHAPE, as -DPARTICLE_SHAPE_* is in the reference):
# CYL_SHAPE = triangle (default) | tophat | bspline3 selects the library of that shape for the whole process, on the
# oracle's side here and on the product's side in cylindrical_epoch_b200/_lib.py.  ng = png + 2 follows the shape.
SHAPE = os.environ.get("CYL_SHAPE", "triangle") or "triangle"
NG = {"triangle": 5, "tophat": 4, "bspline3": 6}[SHAPE]
# boundary-condition codes (constants.F90:55-72)
BC_PERIODIC, BC_OTHER, BC_SIMPLE_LASER, BC_SIMPLE_OUTFLOW, BC_OPEN = 1, 2, 3, 4, 5
BC_ZERO_GRADIENT, BC_CLAMP, BC_REFLECT, BC_CONDUCT, BC_THERMAL = 7, 8, 9, 10, 11
BC_ZERO_B = 16
BD_X_MIN, BD_X_MAX, BD_Y_MIN, BD_Y_MAX = 0, 1, 2, 3

FIELD_NAMES = ["exm", "erm", "etm", "bxm", "brm", "btm", "jxm", "jrm", "jtm",
               "bxm_old", "brm_old", "btm_old", "jxm_old", "jrm_old", "jtm_old"]
SNAP_NAMES = [f"{n}_x_min" for n in FIELD_NAMES[:6]] + [f"{n}_x_max" for n in FIELD_NAMES[:6]]

OPS = dict(step=0, fields_half=1, push=2, current_finish=3, fields_final=4, moving_window=5,
           init_half_step=6, particle_bcs=7, efield_bcs=8, bfield_bcs_mpi=9, bfield_final_bcs=10,
           update_e=11, update_b=12, snapshot_boundaries=13, advance_half_time=14, push_no_bcs=15,
           current_bcs=16, flush_rng=17, bfield_bcs=18)

# calc_df.F90 moments (cyl_moments.cpp / include/cylgpu.h CYLGPU_MOM_*)
MOMENTS = dict(mass_density=0, number_density=1, ekbar=2, ekflux=3, ppc=4, average_weight=5, temperature=6,
               species_current=7, average_momentum=8)

Q0 = 1.602176565e-19
M0 = 9.10938291e-31
C_LIGHT = 2.99792458e8
KB = 1.3806488e-23
EPSILON0 = 8.854187817620389850536563031710750e-12


class CyloConfig(C.Structure):
    _fields_ = [("nx_global", C.c_int32), ("ny_global", C.c_int32), ("n_mode", C.c_int32),
                ("nranks", C.c_int32), ("x_min", C.c_double), ("x_max", C.c_double),
                ("y_max", C.c_double), ("dt_multiplier", C.c_double), ("bc_field", C.c_int32 * 4),
                ("move_window", C.c_int32), ("window_v_x", C.c_double),
                ("window_start_time", C.c_double), ("window_stop_time", C.c_double),
                ("bc_x_min_after_move", C.c_int32), ("bc_x_max_after_move", C.c_int32)]


def build(force=False):
    """Compile oracle/libcyl_oracle.so with the committed Makefile (g++, no FMA contraction)."""
    so = os.path.join(_HERE, "libcyl_oracle.so" if SHAPE == "triangle" else f"libcyl_oracle_{SHAPE}.so")
    srcs = [os.path.join(_HERE, f) for f in ("cyl_oracle.cpp", "cyl_moments.cpp", "cyl_philox.cpp", "cyl_oracle_capi.cpp", "cyl_oracle.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", os.path.basename(so)])
    return so


def philox4x32(ctr, key):
    """Philox4x32-10 of the oracle (cyl_philox.cpp): 4 counter words, 2 key words -> 4 words"""
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().cylo_philox4x32(c, k, o)
    return list(o)


def lib():
    global _LIB
    if _LIB is None:
        # idle OpenMP workers must sleep, not spin: the oracle's parallel regions are short and
        # the test box is usually oversubscribed
        os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
        L = C.CDLL(build())
        L.cylo_ng.restype = C.c_int
        assert L.cylo_ng() == NG, (SHAPE, NG, L.cylo_ng())
        L.cylo_create.restype = C.c_void_p
        L.cylo_create.argtypes = [C.POINTER(CyloConfig)]
        L.cylo_destroy.argtypes = [C.c_void_p]
        L.cylo_add_species.restype = C.c_int
        L.cylo_add_species.argtypes = [C.c_void_p, C.c_double, C.c_double, C.POINTER(C.c_int32), C.c_int,
                                       C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double),
                                       C.POINTER(C.c_double)]
        L.cylo_add_laser.argtypes = [C.c_void_p, C.c_int] + [C.c_double] * 10
        L.cylo_load_uniform.argtypes = [C.c_void_p, C.c_int]
        L.cylo_get_scalars.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.cylo_set_dt.argtypes = [C.c_void_p, C.c_double]
        L.cylo_set_time.argtypes = [C.c_void_p, C.c_double]
        L.cylo_number_density_modes.restype = None
        L.cylo_number_density_modes.argtypes = [C.c_void_p, C.c_int]
        L.cylo_set_counter_insert.restype = None
        L.cylo_set_counter_insert.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.cylo_philox4x32.restype = None
        L.cylo_philox4x32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.cylo_moment.restype = None
        L.cylo_moment.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.cylo_moment_ptr.restype = C.c_void_p
        L.cylo_moment_ptr.argtypes = [C.c_void_p, C.c_int]
        L.cylo_charge_density.restype = None
        L.cylo_charge_density.argtypes = [C.c_void_p, C.c_int]
        L.cylo_wk_ptr.restype = C.c_void_p
        L.cylo_wk_ptr.argtypes = [C.c_void_p, C.c_int]
        L.cylo_set_threads.restype = None
        L.cylo_set_threads.argtypes = [C.c_int]
        L.cylo_get_max_threads.restype = C.c_int
        L.cylo_set_taylor_switch.restype = None
        L.cylo_set_taylor_switch.argtypes = [C.c_void_p, C.c_double]
        L.cylo_set_reference_quirks.restype = None
        L.cylo_set_reference_quirks.argtypes = [C.c_void_p, C.c_int]
        L.cylo_set_hc_push.restype = None
        L.cylo_set_hc_push.argtypes = [C.c_void_p, C.c_int]
        L.cylo_set_smoothing.restype = None
        L.cylo_set_smoothing.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]
        L.cylo_get_bc_field.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        L.cylo_get_bc_particle.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32)]
        L.cylo_rank_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        L.cylo_field_ptr.restype = C.c_void_p
        L.cylo_field_ptr.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.cylo_nparticles.restype = C.c_int64
        L.cylo_nparticles.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.cylo_get_particles.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.cylo_set_particles.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p]
        L.cylo_stats.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        L.cylo_laser_sources.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.cylo_call.restype = C.c_double
        L.cylo_call.argtypes = [C.c_void_p, C.c_int]
        L.cylo_rng_uniform.restype = C.c_double
        L.cylo_rng_uniform.argtypes = [C.c_void_p, C.c_int]
        L.cylo_rng_get_state.restype = None
        L.cylo_rng_get_state.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int),
                                         C.POINTER(C.c_double)]
        _LIB = L
    return _LIB


def set_threads(n):
    """OpenMP threads of the rank loop (one x-slab per thread): explicit, whatever OMP_NUM_THREADS the launcher set"""
    lib().cylo_set_threads(int(n))
    return lib().cylo_get_max_threads()


class OracleWorld:
    """All ranks (x-slabs) of one simulation, stepped in-process by the CPU oracle."""

    def __init__(self, nx, ny, n_mode, x_min, x_max, y_max, bc_field, nranks=1, dt_multiplier=0.95,
                 move_window=False, window_v_x=0.0, window_start_time=0.0, window_stop_time=1e300,
                 bc_x_min_after_move=BC_SIMPLE_OUTFLOW, bc_x_max_after_move=BC_SIMPLE_OUTFLOW):
        self.L = lib()
        cfg = CyloConfig(nx, ny, n_mode, nranks, x_min, x_max, y_max, dt_multiplier,
                         (C.c_int32 * 4)(*bc_field), int(move_window), window_v_x, window_start_time,
                         window_stop_time, bc_x_min_after_move, bc_x_max_after_move)
        self.h = C.c_void_p(self.L.cylo_create(C.byref(cfg)))
        self.nranks = nranks
        self.n_mode = n_mode
        self.nx_global, self.ny_global = nx, ny
        self.n_species = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.cylo_destroy(self.h)
            self.h = None

    # --- configuration -------------------------------------------------------------------
    def add_species(self, charge, mass, bc_particle, ppc=0.0, density=0.0, temp=(0, 0, 0),
                    drift=(0, 0, 0), immobile=False, zero_current=False):
        t = (C.c_double * 3)(*temp)
        d = (C.c_double * 3)(*drift)
        i = self.L.cylo_add_species(self.h, charge, mass, (C.c_int32 * 4)(*bc_particle), int(immobile),
                                    int(zero_current), float(ppc), float(density), t, d)
        self.n_species = i + 1
        return i

    def add_laser(self, boundary, amp, omega, pol_angle=0.0, t_start=0.0, t_end=1e300, t_centre=0.0,
                  t_width=0.0, r_width=0.0, phase=0.0, phase_curv=0.0):
        self.L.cylo_add_laser(self.h, boundary, amp, omega, pol_angle, t_start, t_end, t_centre, t_width,
                              r_width, phase, phase_curv)

    def load_uniform(self, isp):
        self.L.cylo_load_uniform(self.h, isp)

    # --- scalars ---------------------------------------------------------------------------
    def scalars(self):
        out = (C.c_double * 16)()
        self.L.cylo_get_scalars(self.h, out)
        keys = ["dx", "dy", "dt", "time", "x_min", "x_max", "y_max", "x_grid_min", "xb_min",
                "y_grid_min_local", "length_x", "window_shift_fraction", "step", "window_started",
                "window_shifts_total"]
        return dict(zip(keys, list(out)[:15]))

    def set_dt(self, dt):
        self.L.cylo_set_dt(self.h, dt)

    def set_smoothing(self, enable, its=1, comp_its=0, strides=()):
        """smooth_currents, smooth_its, smooth_compensation, smooth_strides of the control block"""
        arr = (C.c_int32 * max(len(strides), 1))(*strides)
        self.L.cylo_set_smoothing(self.h, int(enable), int(its), int(comp_its), len(strides), arr)

    def set_counter_insert(self, on, seed=0):
        """moving window: generate the new column from Philox counters (cyl_philox.cpp) instead of KISS"""
        self.L.cylo_set_counter_insert(self.h, int(on), int(seed))

    def moment(self, kind, species=-1, direction=0):
        """one of the real-valued calc_df.F90 moments (MOMENTS keys; direction: +-1/2/3 = c_dir_x/y/z,
        0 = absent): list of per-rank real arrays [ir+NG-1, ix+NG-1] (copies)"""
        self.L.cylo_moment(self.h, MOMENTS[kind] if isinstance(kind, str) else int(kind), int(species),
                           int(direction))
        out = []
        for k in range(self.nranks):
            info = self.rank_info(k)
            shape = (self.n_mode, info["ny"] + 2 * NG, info["nx"] + 2 * NG)
            buf = (C.c_double * (2 * shape[0] * shape[1] * shape[2])).from_address(
                self.L.cylo_moment_ptr(self.h, k))
            out.append(np.frombuffer(buf, dtype=np.complex128).reshape(shape)[0].real.copy())
        return out

    def charge_density(self, species=-1):
        """calc_charge_density (calc_df.F90:442-519): list of per-rank real arrays [ir+NG-1, ix+NG-1]"""
        return [a[0].real.copy() for a in self.number_density_modes(species, _charge=True)]

    def number_density_modes(self, species=-1, _charge=False):
        """calc_number_density_modes (calc_df.F90:588-661) for one species (or all, -1): list of
        per-rank complex arrays [im, ir+NG-1, ix+NG-1] (copies)"""
        if _charge:
            self.L.cylo_charge_density(self.h, int(species))
        else:
            self.L.cylo_number_density_modes(self.h, int(species))
        out = []
        for k in range(self.nranks):
            info = self.rank_info(k)
            shape = (self.n_mode, info["ny"] + 2 * NG, info["nx"] + 2 * NG)
            buf = (C.c_double * (2 * shape[0] * shape[1] * shape[2])).from_address(self.L.cylo_wk_ptr(self.h, k))
            out.append(np.frombuffer(buf, dtype=np.complex128).reshape(shape).copy())
        return out

    def set_reference_quirks(self, on):
        """laser.f90's array-section and REAL-for-imaginary quirks (cyl_oracle.hpp reference_quirks); default on"""
        self.L.cylo_set_reference_quirks(self.h, int(bool(on)))

    def set_taylor_switch(self, v):
        """|m dtheta| below which the deposit uses the small-angle series (particles.F90:593: 1.0e-4); moved only by
        the test that isolates the conditioning of the closed forms just above the switch"""
        self.L.cylo_set_taylor_switch(self.h, float(v))

    def set_hc_push(self, on):
        """the reference's -DHC_PUSH build: Higuera-Cary gamma instead of Boris'"""
        self.L.cylo_set_hc_push(self.h, int(on))

    def set_time(self, t):
        self.L.cylo_set_time(self.h, t)

    def bc_field(self):
        out = (C.c_int32 * 4)()
        self.L.cylo_get_bc_field(self.h, out)
        return list(out)

    def bc_particle(self, isp):
        out = (C.c_int32 * 4)()
        self.L.cylo_get_bc_particle(self.h, isp, out)
        return list(out)

    def rank_info(self, k):
        io = (C.c_int32 * 6)()
        do = (C.c_double * 4)()
        self.L.cylo_rank_info(self.h, k, io, do)
        return dict(nx=io[0], ny=io[1], cell_x_min=io[2], cell_x_max=io[3], x_min_boundary=bool(io[4]),
                    x_max_boundary=bool(io[5]), x_grid_min_local=do[0], x_grid_max_local=do[1],
                    x_min_local=do[2], x_max_local=do[3])

    # --- arrays (numpy views onto the oracle's memory; index [im, ir+NG-1, ix+NG-1]) -------
    def field(self, k, name):
        info = self.rank_info(k)
        nx, ny = info["nx"], info["ny"]
        if name in FIELD_NAMES:
            fid = FIELD_NAMES.index(name)
            shape = (self.n_mode, ny + 2 * NG, nx + 2 * NG)
        else:
            fid = 15 + SNAP_NAMES.index(name)
            shape = (self.n_mode, ny + 2 * NG)
        ptr = self.L.cylo_field_ptr(self.h, k, fid)
        n = int(np.prod(shape))
        buf = (C.c_double * (2 * n)).from_address(ptr)
        return np.frombuffer(buf, dtype=np.complex128).reshape(shape)

    def nparticles(self, k, isp):
        return int(self.L.cylo_nparticles(self.h, k, isp))

    def particles(self, k, isp):
        """(n, 7) array: x, y, z, px, py, pz, weight (the reference's 7-double wire format)."""
        n = self.nparticles(k, isp)
        out = np.empty((n, 7), dtype=np.float64)
        if n:
            self.L.cylo_get_particles(self.h, k, isp, out.ctypes.data)
        return out

    def set_particles(self, k, isp, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1, 7)
        self.L.cylo_set_particles(self.h, k, isp, arr.shape[0], arr.ctypes.data)

    def stats(self, k):
        out = (C.c_int64 * 4)()
        self.L.cylo_stats(self.h, k, out)
        return dict(sent_left=out[0], sent_right=out[1], removed=out[2], received=out[3])

    def laser_sources(self, bd, k=0):
        ny = self.rank_info(k)["ny"]
        s1 = np.zeros(ny + 1)
        s2 = np.zeros(ny + 1)
        self.L.cylo_laser_sources(self.h, bd, k, s1.ctypes.data, s2.ctypes.data)
        return s1, s2

    def rng_state(self, k):
        """(x, y, z, w), box_muller_cached, cached_random_value of rank k's KISS stream"""
        xyzw = (C.c_int32 * 4)()
        cached = C.c_int()
        cv = C.c_double()
        self.L.cylo_rng_get_state(self.h, k, xyzw, C.byref(cached), C.byref(cv))
        return list(xyzw), int(cached.value), float(cv.value)

    # --- operators -------------------------------------------------------------------------
    def call(self, op):
        t = self.L.cylo_call(self.h, OPS[op])
        if t < 0:
            raise ValueError(op)
        return t

    def step(self, n=1):
        t = 0.0
        for _ in range(n):
            t += self.call("step")
        return t
