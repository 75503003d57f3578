"""Host-side mirror of the reference's hot-path module procedures over the cylgpu C-ABI.

One `Slab` is one MPI rank of the reference (one x-slab, one GPU).  Method names and call
order are the reference's (epoch2d.F90:189-266): `update_eb_fields_half`, `push_particles`,
`current_finish`, `update_eb_fields_final`, `moving_window`, plus the pieces the reference
also calls on their own (`particle_bcs`, `efield_bcs`, `bfield_bcs`, `bfield_final_bcs`).
What stays on the host in the reference stays on the host here: time/step bookkeeping, the
laser source evaluation (laser.f90:276-328,442-461), the window trigger logic
(window.F90:330-376) and the generation of the freshly inserted plasma column.
"""
import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .constants import *  # noqa: F401,F403
from .constants import (BC_CLAMP, BC_CONDUCT, BC_CPML_LASER, BC_CPML_OUTFLOW, BC_OPEN, BC_OTHER, BC_PERIODIC,
                        BC_REFLECT, BC_SIMPLE_LASER, BC_SIMPLE_OUTFLOW, BD_X_MAX, BD_X_MIN, BD_Y_MIN,
                        FIELD_NAMES, NG, SNAP_NAMES, TRANSPORT_NONE)
from .decomp import SlabGrid


class CylGpuError(RuntimeError):
    pass


@dataclass
class Species:   # shared_data.F90:190-280, hot-path members
    charge: float
    mass: float
    bc_particle: tuple = (BC_OPEN, BC_OPEN, BC_OPEN, BC_OPEN)
    immobile: bool = False
    zero_current: bool = False
    # used only for moving-window insertion (window.F90:157-300): uniform plasma
    npart_per_cell: float = 0.0
    density: float = 0.0
    temp: tuple = (0.0, 0.0, 0.0)
    drift: tuple = (0.0, 0.0, 0.0)
    density_min: float = 0.0          # initial_conditions%density_min / density_max
    density_max: float = 1.0e300


@dataclass
class Laser:     # laser.f90 laser_block restricted to what the decks in scope use
    boundary: int
    amp: float
    omega: float
    pol_angle: float = 0.0
    t_start: float = 0.0
    t_end: float = 1e300
    t_centre: float = 0.0
    t_width: float = 0.0      # <= 0 -> constant temporal envelope
    r_width: float = 0.0      # <= 0 -> flat radial profile
    phase: float = 0.0
    phase_curv: float = 0.0   # phase(y) = phase + phase_curv y^2: the deck's phase function (laser.f90:203-226,454)


def normalise_bc_field(bc):
    """setup_boundaries, boundary.F90:30-75: returns (bc_field, add_laser)."""
    bc = list(bc)
    add_laser = [False] * 4
    for i in range(4):
        if bc[i] == BC_OTHER:
            bc[i] = BC_CLAMP
        if bc[i] == BC_SIMPLE_LASER:
            add_laser[i] = True
        if bc[i] == BC_REFLECT:
            bc[i] = BC_CLAMP
        if bc[i] == BC_OPEN:
            bc[i] = BC_SIMPLE_OUTFLOW
    return bc, add_laser


def normalise_bc_particle(bc):
    """boundary.F90:109-123."""
    bc = list(bc)
    for i in range(4):
        if i == BD_Y_MIN:
            continue
        if bc[i] in (BC_OTHER, BC_CONDUCT):
            bc[i] = BC_REFLECT
        if bc[i] in (BC_SIMPLE_LASER, BC_SIMPLE_OUTFLOW, BC_CPML_LASER, BC_CPML_OUTFLOW):
            bc[i] = BC_OPEN
    return bc


class Slab:
    def __init__(self, nx, ny, n_mode, x_min, x_max, y_max, bc_field, species, rank=0, nranks=1,
                 dt_multiplier=0.95, lasers=(), transport=TRANSPORT_NONE, device=-1, fabric=None,
                 nccl_unique_id=None, sendrecv=None, move_window=False, window_v_x=0.0,
                 window_start_time=0.0, window_stop_time=1e300, bc_x_min_after_move=BC_SIMPLE_OUTFLOW,
                 bc_x_max_after_move=BC_SIMPLE_OUTFLOW, insert_fn=None, device_insert_seed=None,
                 exchange_capacity=None, native_driver=None, cell_bounds=None, grid_like=None):
        self.L = _lib.load()
        # what respawn() (the slab re-balancer) needs to build this slab again with other bounds
        self._ctor = dict(nx=nx, ny=ny, n_mode=n_mode, y_max=y_max, nranks=nranks, dt_multiplier=dt_multiplier,
                          transport=transport, device=device, move_window=move_window, window_v_x=window_v_x,
                          window_start_time=window_start_time, window_stop_time=window_stop_time,
                          bc_x_min_after_move=bc_x_min_after_move, bc_x_max_after_move=bc_x_max_after_move,
                          insert_fn=insert_fn, device_insert_seed=device_insert_seed)
        if grid_like is not None:      # a re-balanced slab: the (possibly shifted) grid of its predecessor, new bounds
            self.grid = SlabGrid.like(grid_like, cell_bounds, rank)
        else:
            self.grid = SlabGrid(nx, ny, nranks, rank, x_min, x_max, y_max, dt_multiplier, bounds=cell_bounds)
        g = self.grid
        self.n_mode = n_mode
        self.raw_bc_field = list(bc_field)
        self.bc_field, self.add_laser = normalise_bc_field(bc_field)
        self.species = list(species)
        self.lasers = list(lasers)
        self.dt = g.dt
        self._time = 0.0
        self._step = 0
        self.move_window = move_window
        self.window_v_x = window_v_x
        self.window_start_time = window_start_time
        self.window_stop_time = window_stop_time
        self.bc_after_move = (bc_x_min_after_move, bc_x_max_after_move)
        self.window_started = False
        self.window_shift_fraction = 0.0
        self.window_shifts_total = 0
        self.insert_fn = insert_fn
        # not None: the plasma column of the moving window is generated on the device from the
        # counter-based stream of cylgpu_insert_particles_device (seeded with this number)
        self.device_insert_seed = device_insert_seed
        self.host_lists = None
        self.host_counts = None
        self._keep = []   # ctypes objects that must outlive the handle

        cfg = _lib.Config()
        cfg.nx, cfg.ny, cfg.nx_global, cfg.ny_global = g.nx, g.ny, g.nx_global, g.ny_global
        cfg.n_mode, cfg.rank, cfg.nranks = n_mode, rank, nranks
        cfg.x_min_boundary, cfg.x_max_boundary = int(g.x_min_boundary), int(g.x_max_boundary)
        cfg.bc_field = (C.c_int32 * 4)(*self.bc_field)
        cfg.n_species = len(self.species)
        cfg.device = device
        cfg.transport = transport
        cfg.dx, cfg.dy, cfg.dt = g.dx, g.dy, g.dt
        cfg.x_grid_min_local, cfg.y_grid_min_local = g.x_grid_min_local, g.y_grid_min_local
        cfg.x_min, cfg.x_max, cfg.y_max = g.x_min, g.x_max, g.y_max
        cfg.x_min_local, cfg.x_max_local = g.x_min_local, g.x_max_local
        if nccl_unique_id is not None:
            buf = C.create_string_buffer(bytes(nccl_unique_id), 128)
            self._keep.append(buf)
            cfg.nccl_unique_id = C.cast(buf, C.c_void_p)
        if sendrecv is not None:
            cb = _lib.SENDRECV_FN(sendrecv)
            self._keep.append(cb)
            cfg.sendrecv = cb
        if fabric is not None:
            cfg.fabric = fabric
        self.h = C.c_void_p()
        self._ck(self.L.cylgpu_create(C.byref(cfg), C.byref(self.h)))
        self.rng_init(7842432 + rank)         # setup.F90:563-567
        for i, sp in enumerate(self.species):
            sc = _lib.SpeciesC(sp.charge, sp.mass, (C.c_int32 * 4)(*normalise_bc_particle(sp.bc_particle)),
                               int(sp.immobile), int(sp.zero_current))
            self._ck(self.L.cylgpu_set_species(self.h, i, C.byref(sc)))
        # device-resident particle counts by default: particles move less than a cell per step, so the leavers
        # towards one neighbour are bounded by the population of its two boundary columns (plus the window's)
        if exchange_capacity is None:
            ppc = sum(int(math.ceil(sp.npart_per_cell)) for sp in self.species)
            exchange_capacity = max(4096, 4 * (g.ny + 2) * max(ppc, 1))
        self.set_exchange_capacity(exchange_capacity)
        # whole steps run inside the library (csrc/driver.cu: the same loop body, natively) unless the plasma
        # column of the moving window comes from a Python callback; native_driver=False keeps the step in this
        # mirror, one entry point at a time, as the Fortran driver would call them
        self.native = (insert_fn is None) if native_driver is None else bool(native_driver)
        if self.native:
            self._configure_driver()

    # ------------------------------------------------------------------ native main-loop body (csrc/driver.cu)
    def _configure_driver(self):
        g = self.grid
        cfg = _lib.DriverConfig()
        cfg.cell_x_min = g.cell_x_min
        cfg.move_window = int(bool(self.move_window))
        cfg.raw_bc_field = (C.c_int32 * 4)(*self.raw_bc_field)
        cfg.bc_x_min_after_move, cfg.bc_x_max_after_move = self.bc_after_move
        cfg.n_lasers = len(self.lasers)
        cfg.insert_mode = 0 if self.device_insert_seed is None else 1
        cfg.insert_seed = 0 if self.device_insert_seed is None else int(self.device_insert_seed)
        cfg.x_grid_min = g.x_grid_min
        cfg.xb_min = g.xb_min
        cfg.window_v_x, cfg.window_start_time = self.window_v_x, self.window_start_time
        cfg.window_stop_time = self.window_stop_time
        arr = (_lib.LaserC * max(len(self.lasers), 1))()
        for k, L in enumerate(self.lasers):
            arr[k] = _lib.LaserC(L.boundary, 0, L.amp, L.omega, L.pol_angle, L.t_start, L.t_end, L.t_centre, L.t_width,
                                 L.r_width, L.phase, L.phase_curv)
        self._keep.append(arr)
        cfg.lasers = C.cast(arr, C.POINTER(_lib.LaserC))
        for i, sp in enumerate(self.species):
            cfg.insert[i] = _lib.InsertProfileC(float(sp.npart_per_cell), float(sp.density), (C.c_double * 3)(*sp.temp),
                                                (C.c_double * 3)(*sp.drift), float(sp.density_min), float(sp.density_max))
        cfg.time, cfg.step = self.time, self.step
        cfg.window_shift_fraction = self.window_shift_fraction
        cfg.window_shifts_total = self.window_shifts_total
        cfg.window_started = int(self.window_started)
        self._ck(self.L.cylgpu_driver_configure(self.h, C.byref(cfg)))

    def _refresh_from_driver(self):
        """host-side state the native loop advanced: no device work, no sync"""
        st = _lib.DriverState()
        self._ck(self.L.cylgpu_driver_get_state(self.h, C.byref(st)))
        self._time, self._step = st.time, int(st.step)
        self.window_started = bool(st.window_started)
        self.window_shift_fraction = st.window_shift_fraction
        g = self.grid
        shifts = int(st.window_shifts_total) - self.window_shifts_total
        for _ in range(shifts):
            g.shift()
        self.window_shifts_total = int(st.window_shifts_total)
        self.bc_field = list(st.bc_field)
        self.raw_bc_field = list(st.raw_bc_field)
        self.add_laser = normalise_bc_field(self.raw_bc_field)[1]

    @property
    def time(self):
        return self._time

    @time.setter
    def time(self, v):
        self._time = float(v)
        if getattr(self, "native", False) and getattr(self, "h", None):
            self._ck(self.L.cylgpu_driver_set_time(self.h, self._time, int(self._step)))

    @property
    def step(self):
        return self._step

    @step.setter
    def step(self, v):
        self._step = int(v)
        if getattr(self, "native", False) and getattr(self, "h", None):
            self._ck(self.L.cylgpu_driver_set_time(self.h, float(self._time), self._step))

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc):
        if rc != 0:
            raise CylGpuError(self.L.cylgpu_last_error().decode())

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.L.cylgpu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def field_shape(self):
        return (self.n_mode, self.grid.ny + 2 * NG, self.grid.nx + 2 * NG)

    def upload_field(self, name, arr):
        a = np.ascontiguousarray(arr, dtype=np.complex128)
        assert a.shape == self.field_shape, (a.shape, self.field_shape)
        self._ck(self.L.cylgpu_upload_field(self.h, FIELD_NAMES.index(name), a.ctypes.data))

    def download_field(self, name):
        a = np.empty(self.field_shape, dtype=np.complex128)
        self._ck(self.L.cylgpu_download_field(self.h, FIELD_NAMES.index(name), a.ctypes.data))
        return a

    def upload_snapshot(self, name, arr):
        a = np.ascontiguousarray(arr, dtype=np.complex128)
        assert a.shape == (self.n_mode, self.grid.ny + 2 * NG)
        self._ck(self.L.cylgpu_upload_snapshot(self.h, SNAP_NAMES.index(name), a.ctypes.data))

    def download_snapshot(self, name):
        a = np.empty((self.n_mode, self.grid.ny + 2 * NG), dtype=np.complex128)
        self._ck(self.L.cylgpu_download_snapshot(self.h, SNAP_NAMES.index(name), a.ctypes.data))
        return a

    def upload_particles(self, isp, aos):
        a = np.ascontiguousarray(aos, dtype=np.float64).reshape(-1, 7)
        self._ck(self.L.cylgpu_upload_particles(self.h, isp, a.shape[0], a.ctypes.data))

    def append_particles(self, isp, aos):
        a = np.ascontiguousarray(aos, dtype=np.float64).reshape(-1, 7)
        self._ck(self.L.cylgpu_append_particles(self.h, isp, a.shape[0], a.ctypes.data))

    def particle_count(self, isp):
        n = C.c_int64()
        self._ck(self.L.cylgpu_particle_count(self.h, isp, C.byref(n)))
        return n.value

    def download_particles(self, isp):
        n = self.particle_count(isp)
        a = np.empty((n, 7), dtype=np.float64)
        nn = C.c_int64()
        self._ck(self.L.cylgpu_download_particles(self.h, isp, n, a.ctypes.data, C.byref(nn)))
        return a

    def particle_cells(self, isp):
        n = self.particle_count(isp)
        a = np.empty((n, 2), dtype=np.int32)
        self._ck(self.L.cylgpu_particle_cells(self.h, isp, n, a.ctypes.data))
        return a

    def stats(self):
        s = _lib.Stats()
        self._ck(self.L.cylgpu_stats(self.h, C.byref(s)))
        return s

    def reset_stats(self):
        self._ck(self.L.cylgpu_reset_stats(self.h))

    def number_density_modes(self, isp=-1):   # calc_df.F90:588-661
        a = np.empty(self.field_shape, dtype=np.complex128)
        self._ck(self.L.cylgpu_number_density_modes(self.h, int(isp), a.ctypes.data))
        return a

    def charge_density(self, isp=-1):         # calc_df.F90:442-519
        a = np.empty(self.field_shape[1:], dtype=np.float64)
        self._ck(self.L.cylgpu_charge_density(self.h, int(isp), a.ctypes.data))
        return a

    # calc_df.F90 moments (include/cylgpu.h CYLGPU_MOM_*): calc_mass_density :59, calc_number_density :523,
    # calc_ekbar :140, calc_ekflux :249, calc_ppc :665, calc_average_weight :716, calc_temperature :782,
    # calc_per_species_current :1037, calc_average_momentum :1143
    MOMENTS = dict(mass_density=0, number_density=1, ekbar=2, ekflux=3, ppc=4, average_weight=5, temperature=6,
                   species_current=7, average_momentum=8)

    def moment(self, kind, isp=-1, direction=0):
        """real array [ir+ng-1, ix+ng-1] of one calc_df.F90 moment of species isp (< 0: all that carry
        current); direction 1/2/3 = c_dir_x/y/z, negative for the backward ekflux, 0 = absent"""
        a = np.empty(self.field_shape[1:], dtype=np.float64)
        k = self.MOMENTS[kind] if isinstance(kind, str) else int(kind)
        self._ck(self.L.cylgpu_particle_moment(self.h, k, int(isp), int(direction), a.ctypes.data))
        return a

    # ------------------------------------------------------------------ SDF dump / restart
    def _sdf_desc(self, species_names, step=None, time=None, restart=False):
        g = self.grid
        d = _lib.SdfDesc()
        d.nx_global, d.ny_global, d.n_mode, d.n_species = g.nx_global, g.ny_global, self.n_mode, len(self.species)
        d.nx_local, d.cell_x_min = g.nx, g.cell_x_min
        d.step = self.step if step is None else int(step)
        d.time = self.time if time is None else float(time)
        d.restart = int(bool(restart))
        d.x_min, d.dx, d.dy = g.xb_min, g.dx, g.dy
        self._sdf_names = [n.encode() if isinstance(n, str) else n for n in species_names]
        for i, n in enumerate(self._sdf_names):
            d.species_name[i] = n
        return d

    # derived variables of write_nspecies_field in the reference's order (io/diagnostics.F90:765-835)
    SDF_DERIVED = ["ekbar", "mass_density", "charge_density", "number_density", "ppc", "average_weight",
                   "average_px", "average_py", "average_pz", "temperature", "temperature_x", "temperature_y",
                   "temperature_z", "jx", "jy", "jz", "ekflux/x_max", "ekflux/y_max", "ekflux/z_max",
                   "ekflux/x_min", "ekflux/y_min", "ekflux/z_min", "number_density_mode"]

    def _sdf_constants(self, d, ids, values=None):
        self._sdf_const_keep = [(i.encode(), i.encode()) for i in ids]
        d.n_constants = len(ids)
        for k, (bid, name) in enumerate(self._sdf_const_keep):
            d.constant_id[k], d.constant_name[k] = bid, name
            if values is not None:
                d.constant_value[k] = float(values[k])

    # what the driver owns and a restart needs back (sdf_write_srl, io/diagnostics.F90:408-416)
    SDF_RESTART_CONSTANTS = ("dt", "window_shift_fraction", "x_grid_min", "window_started", "window_shifts_total")

    def sdf_dump(self, path, species_names, npart_global=None, npart_offset=None, restart=False, derived=(),
                 derived_sum=True, derived_species=True):
        """one SDF file in the reference's layout (io/diagnostics.F90:497-575,2033-2110,3040-3160) written
        from the device mirrors; with several ranks pass the global particle counts and this rank's offsets
        (the reference's species_offset) -- every rank calls with the same path"""
        d = self._sdf_desc(species_names, restart=restart)
        for name in derived:      # computed on the device from the resident lists (cylgpu_particle_moment)
            d.derived_mask |= 1 << self.SDF_DERIVED.index(name)
        d.derived_sum, d.derived_species = int(derived_sum), int(derived_species)
        self._sdf_constants(d, self.SDF_RESTART_CONSTANTS,
                            (self.dt, self.window_shift_fraction, self.grid.x_grid_min, float(self.window_started),
                             float(self.window_shifts_total)))
        if npart_global is not None:
            for i in range(len(self.species)):
                d.npart_global[i] = int(npart_global[i])
                d.npart_offset[i] = int(npart_offset[i])
        self._ck(self.L.cylgpu_sdf_dump(self.h, str(path).encode(), C.byref(d)))
        return d

    def sdf_load(self, path, species_names):
        """restart of the hot-path state from such a file (housekeeping/setup.F90:1196-1260,1424-1466):
        interior of the 15 mode arrays, this slab's particles, step and time; the ghosts are then
        re-derived by the boundary routines"""
        d = self._sdf_desc(species_names)
        self._sdf_constants(d, self.SDF_RESTART_CONSTANTS)
        self._ck(self.L.cylgpu_sdf_load(self.h, str(path).encode(), C.byref(d)))
        self.step, self.time = int(d.step), float(d.time)
        if d.constants_found & 2:       # the window state travels in the file (io/diagnostics.F90:411-414)
            self.window_shift_fraction = float(d.constant_value[1])
        self.sdf_constants = {i: float(d.constant_value[k]) for k, i in enumerate(self.SDF_RESTART_CONSTANTS)
                              if d.constants_found & (1 << k)}
        if getattr(self, "native", False):
            self._configure_driver()
        self.efield_bcs()
        self.bfield_bcs(False)
        # J ghosts: the halo of current_finish without smoothing the stored (already smoothed) currents again
        sm = getattr(self, "_smoothing", (False, 1, 0, ()))
        if sm[0]:
            self.set_current_smoothing(False)
        self._ck(self.L.cylgpu_current_finish(self.h))
        if sm[0]:
            self.set_current_smoothing(*sm)
        return d

    def energy(self):
        out = (C.c_double * 2)()
        self._ck(self.L.cylgpu_energy(self.h, out))
        return out[0], out[1]

    def synchronize(self):
        self._ck(self.L.cylgpu_synchronize(self.h))

    def set_stream(self, stream_ptr):
        self._ck(self.L.cylgpu_set_stream(self.h, stream_ptr))

    def set_dt(self, dt):
        self.dt = dt
        self._ck(self.L.cylgpu_set_dt(self.h, dt))

    # ------------------------------------------------------------------ random stream (window insertion)
    def rng_init(self, seed):                 # random_init, random_generator.f90:81-108
        self._ck(self.L.cylgpu_rng_init(self.h, int(seed)))

    def rng_set_state(self, xyzw, cached, cached_value):
        self._ck(self.L.cylgpu_rng_set_state(self.h, (C.c_int32 * 4)(*[int(v) for v in xyzw]), int(cached),
                                             float(cached_value)))

    def rng_get_state(self):
        xyzw = (C.c_int32 * 4)()
        cached = C.c_int()
        cv = C.c_double()
        self._ck(self.L.cylgpu_rng_get_state(self.h, xyzw, C.byref(cached), C.byref(cv)))
        return list(xyzw), int(cached.value), float(cv.value)

    def rng_flush_cache(self):                # random_flush_cache, diagnostics.F90:235
        self._ck(self.L.cylgpu_rng_flush_cache(self.h))

    def rng_uniform(self):
        v = C.c_double()
        self._ck(self.L.cylgpu_rng_uniform(self.h, C.byref(v)))
        return v.value

    def insert_particles(self, isp):          # window.F90:157-300, uniform profiles of `Species`
        sp = self.species[isp]
        nrow = self.grid.ny + 2
        dens = np.full(nrow, float(sp.density))
        temp = np.repeat(np.asarray(sp.temp, dtype=np.float64), nrow)     # (3, ny+2), radial index fastest
        drift = np.repeat(np.asarray(sp.drift, dtype=np.float64), nrow)
        n = C.c_int64()
        self._ck(self.L.cylgpu_insert_particles(self.h, isp, self.grid.x_grid_max, float(sp.npart_per_cell),
                                                dens.ctypes.data, temp.ctypes.data, drift.ctypes.data,
                                                float(sp.density_min), float(sp.density_max), C.byref(n)))
        return n.value

    def insert_particles_host(self, isp, host_aos, count):
        """insert_particles into a list that lives in host memory; returns the new count"""
        sp = self.species[isp]
        nrow = self.grid.ny + 2
        dens = np.full(nrow, float(sp.density))
        temp = np.repeat(np.asarray(sp.temp, dtype=np.float64), nrow)
        drift = np.repeat(np.asarray(sp.drift, dtype=np.float64), nrow)
        n = C.c_int64(int(count))
        self._ck(self.L.cylgpu_insert_particles_host(
            self.h, isp, self.grid.x_grid_max, float(sp.npart_per_cell), dens.ctypes.data, temp.ctypes.data,
            drift.ctypes.data, float(sp.density_min), float(sp.density_max), host_aos.ctypes.data, host_aos.shape[0],
            C.byref(n)))
        return n.value

    def insert_particles_device(self, isp, seed, column):
        """insert_particles (window.F90:157-300) generated by one kernel from the Philox stream
        (seed, species, column): no host loop, no upload, independent of the number of ranks"""
        sp = self.species[isp]
        nrow = self.grid.ny + 2
        dens = np.full(nrow, float(sp.density))
        temp = np.repeat(np.asarray(sp.temp, dtype=np.float64), nrow)
        drift = np.repeat(np.asarray(sp.drift, dtype=np.float64), nrow)
        n = C.c_int64()
        self._ck(self.L.cylgpu_insert_particles_device(
            self.h, isp, self.grid.x_grid_max, float(sp.npart_per_cell), dens.ctypes.data, temp.ctypes.data,
            drift.ctypes.data, float(sp.density_min), float(sp.density_max), int(seed), int(column), C.byref(n)))
        return n.value

    def set_pusher(self, higuera_cary):       # -DHC_PUSH, particles.F90:409-421
        self._ck(self.L.cylgpu_set_pusher(self.h, int(bool(higuera_cary))))

    def set_taylor_switch(self, v):
        """test knob: |m dtheta| below which the deposit uses the small-angle series (particles.F90:593, 1.0e-4)"""
        self._ck(self.L.cylgpu_set_taylor_switch(self.h, float(v)))

    # ------------------------------------------------------------------ slab re-balancer (balance.py)
    @property
    def periodic_x(self):
        from .constants import BC_PERIODIC
        return self.bc_field[BD_X_MIN] == BC_PERIODIC

    def load_x(self):
        """part_load_func (balance.F90:2453-2478) on the device: particles of all species per local column
        1-ng .. nx+ng"""
        a = np.zeros(self.grid.nx + 2 * NG, dtype=np.int64)
        self._ck(self.L.cylgpu_load_x(self.h, a.ctypes.data))
        return a

    def respawn(self, bounds, **transport_kw):
        """this rank's slab again for the slab bounds `bounds` of all ranks (balance.F90 redistribute_domain: new
        extents, same run): same species, lasers, window and loop state, the random stream of the rank, empty arrays
        -- balance.redistribute_* then brings the columns and particles.  transport_kw: fabric / nccl_unique_id /
        sendrecv of the new handle (a communicator is good for one handle)."""
        k = dict(self._ctor)
        g = self.grid
        s = Slab(g.nx_global, k["ny"], k["n_mode"], g.x_min, g.x_max, g.y_max, list(self.raw_bc_field), self.species,
                 rank=g.rank, nranks=k["nranks"], dt_multiplier=k["dt_multiplier"], lasers=self.lasers,
                 transport=k["transport"], device=k["device"], move_window=k["move_window"],
                 window_v_x=k["window_v_x"], window_start_time=k["window_start_time"],
                 window_stop_time=k["window_stop_time"], bc_x_min_after_move=k["bc_x_min_after_move"],
                 bc_x_max_after_move=k["bc_x_max_after_move"], insert_fn=k["insert_fn"],
                 device_insert_seed=k["device_insert_seed"], exchange_capacity=self.exchange_capacity,
                 native_driver=self.native, cell_bounds=bounds, grid_like=g, **transport_kw)
        s.window_started = self.window_started
        s.window_shift_fraction = self.window_shift_fraction
        s.window_shifts_total = self.window_shifts_total
        s._time, s._step = self._time, self._step
        sm = getattr(self, "_smoothing", None)
        if sm is not None:
            s.set_current_smoothing(*sm)
        s.rng_set_state(*self.rng_get_state())
        if s.native:
            s._configure_driver()
        return s

    def load_state(self, st):
        """arrays and particle lists of balance.slab_state / redistribute_*"""
        for n in FIELD_NAMES:
            self.upload_field(n, st["fields"][n])
        for n in SNAP_NAMES:
            self.upload_snapshot(n, st["snaps"][n])
        for i, p in enumerate(st["particles"]):
            self.upload_particles(i, p)

    def transport_info(self):
        """(transport kind, left link through peer-memory mailboxes, right link, mailbox slot KiB)"""
        out = (C.c_int32 * 4)()
        self._ck(self.L.cylgpu_transport_info(self.h, out))
        return tuple(out)

    def set_exchange_capacity(self, particles):
        """> 0: device-resident particle counts, one fixed-size migration message per neighbour and no host sync
        inside a step (include/cylgpu.h); 0: the exact count-then-data protocol of partlist.F90:842,869"""
        self.exchange_capacity = int(particles)
        self._ck(self.L.cylgpu_set_exchange_capacity(self.h, int(particles)))

    def set_reference_quirks(self, on):
        """laser.f90's array-section and REAL-for-imaginary quirks (include/cylgpu.h); reproduced by default"""
        self._ck(self.L.cylgpu_set_reference_quirks(self.h, int(bool(on))))

    def set_sort_interval(self, n):
        self._ck(self.L.cylgpu_set_sort_interval(self.h, n))

    def set_push_variant(self, v):
        self._ck(self.L.cylgpu_set_push_variant(self.h, v))

    def sort_particles(self):
        self._ck(self.L.cylgpu_sort_particles(self.h))

    # ------------------------------------------------------------------ host-side laser
    def laser_sources(self, bd):
        """source1/source2 on ir = 0..ny (laser.f90:442-461); host work in the reference too."""
        ny = self.grid.ny
        s1 = np.zeros(ny + 1)
        s2 = np.zeros(ny + 1)
        if not self.add_laser[bd]:
            return s1, s2
        yv = self.grid.y_grid_min_local + (np.arange(0, ny + 1, dtype=np.float64) - 1.0) * self.grid.dy
        for L in self.lasers:
            if L.boundary != bd or not (L.t_start <= self.time <= L.t_end):
                continue
            tprof = 1.0
            if L.t_width > 0.0:
                a = (self.time - L.t_centre) / L.t_width
                tprof = math.exp(-(a * a))
            t_env = tprof * L.amp
            prof = np.ones(ny + 1)
            if L.r_width > 0.0:
                a = (yv - 0.0) / L.r_width
                prof = np.exp(-(a * a))
            # libm's sin, element by element: the sources then equal the reference's (and the oracle's) bit for
            # bit, which numpy's vectorised sin does not guarantee
            if L.phase_curv == 0.0:
                base = t_env * prof * math.sin(L.omega * self.time + (L.phase + 0.0))
            else:
                arg = L.omega * self.time + (L.phase + L.phase_curv * (yv * yv))
                base = t_env * prof * np.array([math.sin(v) for v in arg])
            s1 = s1 + base * math.cos(L.pol_angle)
            s2 = s2 + base * math.sin(L.pol_angle)
        return s1, s2

    def _src_ptrs(self):
        s1a, s2a = self.laser_sources(BD_X_MIN)
        s1b, s2b = self.laser_sources(BD_X_MAX)
        self._src_keep = (s1a, s2a, s1b, s2b)
        return [a.ctypes.data for a in self._src_keep]

    # ------------------------------------------------------------------ the hot path
    def update_eb_fields_half(self):          # fields.f90:316-337
        self._ck(self.L.cylgpu_fields_half(self.h))

    def push_particles(self):                 # particles.F90:28-734
        if self.host_lists is not None:       # particle lists stay on the host (streamed push)
            self.host_counts = self.push_particles_host(self.host_lists, self.host_counts)
            return
        self._ck(self.L.cylgpu_push(self.h))

    def attach_host_lists(self, lists, counts):
        """From now on the species whose entry in `lists` is an (capacity, 7) float64 array live
        in host memory: push_particles streams them through the GPU (cylgpu_push_host)."""
        self.host_lists = list(lists)
        self.host_counts = [int(c) for c in counts]

    def push_particles_host(self, lists, counts):
        """push_particles + particle_bcs for particle lists that stay in host memory.
        lists[isp]: writable C-contiguous float64 array (capacity, 7) in pack_particle order
        (or None: species isp is device-resident); counts[isp]: particles in it.  Survivors
        and arrivals are written back in place; returns the new counts."""
        nsp = len(self.species)
        n_in = (C.c_int64 * nsp)(*[int(c) for c in counts])
        cap = (C.c_int64 * nsp)()
        ptr = (C.c_void_p * nsp)()
        for i, a in enumerate(lists):
            if a is None:
                continue
            assert a.dtype == np.float64 and a.flags.c_contiguous and a.flags.writeable and a.shape[1] == 7
            cap[i] = a.shape[0]
            ptr[i] = a.ctypes.data
        n_out = (C.c_int64 * nsp)(*[int(c) for c in counts])
        self._ck(self.L.cylgpu_push_host(self.h, n_in, ptr, cap, n_out))
        return [int(v) for v in n_out]

    def set_host_chunk(self, particles):
        self._ck(self.L.cylgpu_set_host_chunk(self.h, int(particles)))

    def push_particles_no_bcs(self):
        self._ck(self.L.cylgpu_push_no_bcs(self.h))

    def set_current_smoothing(self, enable, its=1, comp_its=0, strides=()):
        """smooth_currents, smooth_its, smooth_compensation, smooth_strides (deck_control_block.F90:447-466)"""
        arr = (C.c_int32 * max(len(strides), 1))(*strides)
        self._smoothing = (bool(enable), int(its), int(comp_its), tuple(strides))
        self._ck(self.L.cylgpu_set_current_smoothing(self.h, int(enable), int(its), int(comp_its), len(strides), arr))

    def current_finish(self):                 # current_smooth.F90:29-45
        self._ck(self.L.cylgpu_current_finish(self.h))

    def update_eb_fields_final(self):         # fields.f90:341-353
        self._ck(self.L.cylgpu_fields_final(self.h, *self._src_ptrs()))

    def update_e_field(self):
        self._ck(self.L.cylgpu_update_e_field(self.h))

    def update_b_field(self):
        self._ck(self.L.cylgpu_update_b_field(self.h))

    def efield_bcs(self):
        self._ck(self.L.cylgpu_efield_bcs(self.h))

    def bfield_bcs(self, mpi_only):
        self._ck(self.L.cylgpu_bfield_bcs(self.h, int(mpi_only)))

    def bfield_final_bcs(self):
        self._ck(self.L.cylgpu_bfield_final_bcs(self.h, *self._src_ptrs()))

    def particle_bcs(self):
        self._ck(self.L.cylgpu_particle_bcs(self.h))

    def current_bcs(self):
        self._ck(self.L.cylgpu_current_bcs(self.h))

    def snapshot_field_boundaries(self):      # setup.F90:393-423
        self._ck(self.L.cylgpu_snapshot_field_boundaries(self.h))

    def init_half_step(self):                 # epoch2d.F90:143-161
        if getattr(self, "native", False):
            self._ck(self.L.cylgpu_driver_init_half_step(self.h))
            self._refresh_from_driver()
            return
        self.particle_bcs()
        self.efield_bcs()
        dt_store = self.dt
        self.set_dt(self.dt / 2.0)
        self.time = self.time + self.dt
        self.bfield_final_bcs()
        self.set_dt(dt_store)

    def moving_window(self):                  # window.F90:330-376
        if not self.move_window:
            return
        if not self.window_started:
            if self.window_start_time <= self.time < self.window_stop_time:
                raw = list(self.raw_bc_field)
                raw[BD_X_MIN], raw[BD_X_MAX] = self.bc_after_move
                self.raw_bc_field = raw
                self.bc_field, self.add_laser = normalise_bc_field(raw)
                self._ck(self.L.cylgpu_set_bc_field(self.h, (C.c_int32 * 4)(*self.bc_field)))
                self.window_shift_fraction = 0.0
                self.window_started = True
        if not self.window_started or self.time >= self.window_stop_time or self.window_v_x <= 0.0:
            return
        self.window_shift_fraction = self.window_shift_fraction + self.dt * self.window_v_x / self.grid.dx
        cells = int(math.floor(self.window_shift_fraction))
        if cells > 0:
            for _ in range(cells):
                self._shift_window_once()
            self.particle_bcs()
            self.window_shift_fraction = self.window_shift_fraction - float(cells)

    def _shift_window_once(self):             # window.F90:62-94, one cell
        nsp = len(self.species)
        n_new = (C.c_int64 * max(nsp, 1))()
        ptrs = (C.c_void_p * max(nsp, 1))()
        keep = []
        if getattr(self, "host_lists", None) is not None:
            # the lists live in host memory (cylgpu_push_host): the column joins them there, and the plasma behind
            # the window is dropped by the next streamed push
            if self.grid.nranks > 1:
                raise CylGpuError("host-resident lists with a moving window: one slab only")
            for isp in range(nsp):
                if self.species[isp].npart_per_cell > 0 and self.species[isp].density > 0 and \
                        self.host_lists[isp] is not None:
                    self.host_counts[isp] = self.insert_particles_host(isp, self.host_lists[isp], self.host_counts[isp])
        elif self.device_insert_seed is not None:
            for isp in range(nsp):
                if self.species[isp].npart_per_cell > 0 and self.species[isp].density > 0:
                    self.insert_particles_device(isp, self.device_insert_seed, self.window_shifts_total)
        elif self.insert_fn is None:
            # insert_particles with the rank's KISS stream, species in deck order (window.F90:191)
            for isp in range(nsp):
                if self.species[isp].npart_per_cell > 0 and self.species[isp].density > 0:
                    self.insert_particles(isp)
        elif self.grid.x_max_boundary:
            for isp in range(nsp):
                aos = self.insert_fn(self, isp)      # insert_particles stays on the host
                if aos is not None and len(aos):
                    a = np.ascontiguousarray(aos, dtype=np.float64).reshape(-1, 7)
                    keep.append(a)
                    n_new[isp] = a.shape[0]
                    ptrs[isp] = a.ctypes.data
        self.grid.shift()
        g = self.grid
        grid5 = (C.c_double * 5)(g.x_grid_min_local, g.x_min, g.x_max, g.x_min_local, g.x_max_local)
        self._ck(self.L.cylgpu_window_shift(self.h, n_new, ptrs, grid5))
        self.window_shifts_total += 1

    def run_steps(self, n):
        """n whole steps inside the library (native driver only)"""
        assert self.native and self.host_lists is None
        self._ck(self.L.cylgpu_driver_step(self.h, int(n)))
        self._refresh_from_driver()

    def step_once(self, flush_rng=None):      # epoch2d.F90:189-266 loop body, optional physics off
        if getattr(self, "native", False) and self.host_lists is None and flush_rng is None:
            self._ck(self.L.cylgpu_driver_step(self.h, 1))
            self._refresh_from_driver()
            return
        self.update_eb_fields_half()
        self.push_particles()
        self.current_finish()
        self.step += 1
        self.time = self.time + self.dt / 2.0
        self.rng_flush_cache()                # output_routines -> random_flush_cache, diagnostics.F90:235
        if flush_rng is not None:
            flush_rng()
        self.time = self.time + self.dt / 2.0
        self.update_eb_fields_final()
        self.moving_window()
        if getattr(self, "native", False):       # this step ran in the mirror: the library's copy of the loop state follows
            self._configure_driver()
