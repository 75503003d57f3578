"""ctypes binding of libcylgpu.so (include/cylgpu.h).  Fails loudly when the library is missing."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
from .constants import NG, SHAPE

# CYL_SHAPE: the library of that particle shape (constants.py); CYLGPU_LIB: an alternative build of the same library
# (kernel-tuning experiments)
LIB_PATH = os.environ.get("CYLGPU_LIB") or os.path.join(
    _HERE, "libcylgpu.so" if SHAPE == "triangle" else f"libcylgpu_{SHAPE}.so")
MAX_SPECIES = 8

SENDRECV_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int,
                          C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                          C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p)


class Config(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nx_global", C.c_int32), ("ny_global", C.c_int32),
                ("n_mode", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32),
                ("x_min_boundary", C.c_int32), ("x_max_boundary", C.c_int32),
                ("bc_field", C.c_int32 * 4), ("n_species", C.c_int32), ("device", C.c_int32),
                ("transport", C.c_int32),
                ("dx", C.c_double), ("dy", C.c_double), ("dt", C.c_double),
                ("x_grid_min_local", C.c_double), ("y_grid_min_local", C.c_double),
                ("x_min", C.c_double), ("x_max", C.c_double), ("y_max", C.c_double),
                ("x_min_local", C.c_double), ("x_max_local", C.c_double),
                ("nccl_unique_id", C.c_void_p), ("sendrecv", SENDRECV_FN), ("sendrecv_user", C.c_void_p),
                ("fabric", C.c_void_p), ("particle_capacity", C.c_int64)]


class SpeciesC(C.Structure):
    _fields_ = [("charge", C.c_double), ("mass", C.c_double), ("bc_particle", C.c_int32 * 4),
                ("immobile", C.c_int32), ("zero_current", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("n_particles", C.c_int64 * MAX_SPECIES), ("n_sent_left", C.c_int64),
                ("n_sent_right", C.c_int64), ("n_removed", C.c_int64), ("n_recv", C.c_int64),
                ("n_window_removed", C.c_int64), ("n_sorts", C.c_int64), ("kernel_launches", C.c_int64),
                ("ms_fields", C.c_double), ("ms_push", C.c_double), ("ms_bcs", C.c_double),
                ("ms_sort", C.c_double), ("ms_exchange", C.c_double), ("ms_push_kernel", C.c_double),
                ("n_push_kernel", C.c_int64)]


class SdfDesc(C.Structure):   # cylgpu_sdf_desc
    _fields_ = [("nx_global", C.c_int32), ("ny_global", C.c_int32), ("n_mode", C.c_int32), ("n_species", C.c_int32),
                ("nx_local", C.c_int32), ("cell_x_min", C.c_int32),
                ("step", C.c_int32), ("restart", C.c_int32), ("jobid1", C.c_int32), ("jobid2", C.c_int32),
                ("have_extents", C.c_int32), ("derived_mask", C.c_uint32), ("derived_sum", C.c_int32),
                ("derived_species", C.c_int32),
                ("time", C.c_double), ("x_min", C.c_double), ("dx", C.c_double), ("dy", C.c_double),
                ("species_name", C.c_char_p * MAX_SPECIES),
                ("npart_global", C.c_int64 * MAX_SPECIES), ("npart_offset", C.c_int64 * MAX_SPECIES),
                ("npart_local", C.c_int64 * MAX_SPECIES), ("part_extents", (C.c_double * 6) * MAX_SPECIES),
                ("n_constants", C.c_int32), ("constants_found", C.c_uint32),
                ("constant_id", C.c_char_p * 16), ("constant_name", C.c_char_p * 16),
                ("constant_value", C.c_double * 16)]


# every symbol include/cylgpu.h declares: name -> (restype, argtypes)
H = C.c_void_p
_DP = C.POINTER(C.c_double)
SYMBOLS = {
    "cylgpu_last_error": (C.c_char_p, []),
    "cylgpu_version": (C.c_int, []),
    "cylgpu_fabric_create": (C.c_void_p, [C.c_int]),
    "cylgpu_fabric_destroy": (None, [C.c_void_p]),
    "cylgpu_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "cylgpu_create": (C.c_int, [C.POINTER(Config), C.POINTER(H)]),
    "cylgpu_destroy": (C.c_int, [H]),
    "cylgpu_set_species": (C.c_int, [H, C.c_int, C.POINTER(SpeciesC)]),
    "cylgpu_set_dt": (C.c_int, [H, C.c_double]),
    "cylgpu_set_bc_field": (C.c_int, [H, C.POINTER(C.c_int32)]),
    "cylgpu_set_stream": (C.c_int, [H, C.c_void_p]),
    "cylgpu_synchronize": (C.c_int, [H]),
    "cylgpu_upload_field": (C.c_int, [H, C.c_int, C.c_void_p]),
    "cylgpu_download_field": (C.c_int, [H, C.c_int, C.c_void_p]),
    "cylgpu_upload_snapshot": (C.c_int, [H, C.c_int, C.c_void_p]),
    "cylgpu_download_snapshot": (C.c_int, [H, C.c_int, C.c_void_p]),
    "cylgpu_field_device_ptr": (C.c_void_p, [H, C.c_int]),
    "cylgpu_snapshot_field_boundaries": (C.c_int, [H]),
    "cylgpu_upload_particles": (C.c_int, [H, C.c_int, C.c_int64, C.c_void_p]),
    "cylgpu_append_particles": (C.c_int, [H, C.c_int, C.c_int64, C.c_void_p]),
    "cylgpu_download_particles": (C.c_int, [H, C.c_int, C.c_int64, C.c_void_p, C.POINTER(C.c_int64)]),
    "cylgpu_particle_count": (C.c_int, [H, C.c_int, C.POINTER(C.c_int64)]),
    "cylgpu_particle_cells": (C.c_int, [H, C.c_int, C.c_int64, C.c_void_p]),
    "cylgpu_particle_device_ptr": (C.c_void_p, [H, C.c_int, C.c_int]),
    "cylgpu_fields_half": (C.c_int, [H]),
    "cylgpu_push": (C.c_int, [H]),
    "cylgpu_push_host": (C.c_int, [H, C.POINTER(C.c_int64), C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int64)]),
    "cylgpu_set_host_chunk": (C.c_int, [H, C.c_int64]),
    "cylgpu_set_current_smoothing": (C.c_int, [H, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]),
    "cylgpu_current_finish": (C.c_int, [H]),
    "cylgpu_fields_final": (C.c_int, [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cylgpu_window_shift": (C.c_int, [H, C.POINTER(C.c_int64), C.POINTER(C.c_void_p), _DP]),
    "cylgpu_rng_init": (C.c_int, [H, C.c_int]),
    "cylgpu_rng_set_state": (C.c_int, [H, C.POINTER(C.c_int32), C.c_int, C.c_double]),
    "cylgpu_rng_get_state": (C.c_int, [H, C.POINTER(C.c_int32), C.POINTER(C.c_int), _DP]),
    "cylgpu_rng_flush_cache": (C.c_int, [H]),
    "cylgpu_rng_uniform": (C.c_int, [H, _DP]),
    "cylgpu_insert_particles": (C.c_int, [H, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_double, C.c_double, C.POINTER(C.c_int64)]),
    "cylgpu_update_e_field": (C.c_int, [H]),
    "cylgpu_update_b_field": (C.c_int, [H]),
    "cylgpu_efield_bcs": (C.c_int, [H]),
    "cylgpu_bfield_bcs": (C.c_int, [H, C.c_int]),
    "cylgpu_bfield_final_bcs": (C.c_int, [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cylgpu_particle_bcs": (C.c_int, [H]),
    "cylgpu_push_no_bcs": (C.c_int, [H]),
    "cylgpu_current_bcs": (C.c_int, [H]),
    "cylgpu_sort_particles": (C.c_int, [H]),
    "cylgpu_set_pusher": (C.c_int, [H, C.c_int]),
    "cylgpu_set_taylor_switch": (C.c_int, [H, C.c_double]),
    "cylgpu_set_sort_interval": (C.c_int, [H, C.c_int]),
    "cylgpu_set_push_variant": (C.c_int, [H, C.c_int]),
    "cylgpu_set_reference_quirks": (C.c_int, [H, C.c_int]),
    "cylgpu_shape": (C.c_int, []),
    "cylgpu_ghost_cells": (C.c_int, []),
    "cylgpu_number_density_modes": (C.c_int, [H, C.c_int, C.c_void_p]),
    "cylgpu_charge_density": (C.c_int, [H, C.c_int, C.c_void_p]),
    "cylgpu_insert_particles_device": (C.c_int, [H, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_double, C.c_double, C.c_uint64, C.c_uint64,
                                                 C.POINTER(C.c_int64)]),
    "cylgpu_philox4x32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "cylgpu_sdf_write_host": (C.c_int, [C.c_char_p, C.POINTER(SdfDesc), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_void_p)]),
    "cylgpu_sdf_derived_count": (C.c_int, [C.POINTER(SdfDesc)]),
    "cylgpu_sdf_read_host": (C.c_int, [C.c_char_p, C.POINTER(SdfDesc), C.POINTER(C.c_void_p), C.c_double, C.c_double,
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "cylgpu_sdf_dump": (C.c_int, [H, C.c_char_p, C.POINTER(SdfDesc)]),
    "cylgpu_sdf_load": (C.c_int, [H, C.c_char_p, C.POINTER(SdfDesc)]),
    "cylgpu_particle_moment": (C.c_int, [H, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cylgpu_energy": (C.c_int, [H, _DP]),
    "cylgpu_stats": (C.c_int, [H, C.POINTER(Stats)]),
    "cylgpu_reset_stats": (C.c_int, [H]),
    "cylgpu_set_timing": (C.c_int, [H, C.c_int]),
    "cylgpu_set_exchange_capacity": (C.c_int, [H, C.c_int64]),
    "cylgpu_insert_particles_host": (C.c_int, [H, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_double, C.c_double, C.c_void_p, C.c_int64, C.c_void_p]),
    "cylgpu_transport_info": (C.c_int, [H, C.c_void_p]),
    "cylgpu_load_x": (C.c_int, [H, C.c_void_p]),
    "cylgpu_calculate_breaks": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "cylgpu_driver_configure": (C.c_int, [H, C.c_void_p]),
    "cylgpu_driver_init_half_step": (C.c_int, [H]),
    "cylgpu_driver_step": (C.c_int, [H, C.c_int64]),
    "cylgpu_driver_get_state": (C.c_int, [H, C.c_void_p]),
    "cylgpu_driver_set_time": (C.c_int, [H, C.c_double, C.c_int64]),
}

class LaserC(C.Structure):          # include/cylgpu.h: cylgpu_laser
    _fields_ = [("boundary", C.c_int32), ("pad_", C.c_int32), ("amp", C.c_double), ("omega", C.c_double),
                ("pol_angle", C.c_double), ("t_start", C.c_double), ("t_end", C.c_double), ("t_centre", C.c_double),
                ("t_width", C.c_double), ("r_width", C.c_double), ("phase", C.c_double), ("phase_curv", C.c_double)]


class InsertProfileC(C.Structure):  # cylgpu_insert_profile
    _fields_ = [("npart_per_cell", C.c_double), ("density", C.c_double), ("temp", C.c_double * 3),
                ("drift", C.c_double * 3), ("density_min", C.c_double), ("density_max", C.c_double)]


class DriverConfig(C.Structure):    # cylgpu_driver_config
    _fields_ = [("cell_x_min", C.c_int32), ("move_window", C.c_int32), ("raw_bc_field", C.c_int32 * 4),
                ("bc_x_min_after_move", C.c_int32), ("bc_x_max_after_move", C.c_int32), ("n_lasers", C.c_int32),
                ("insert_mode", C.c_int32), ("insert_seed", C.c_uint64), ("x_grid_min", C.c_double),
                ("xb_min", C.c_double), ("window_v_x", C.c_double), ("window_start_time", C.c_double), ("window_stop_time", C.c_double),
                ("lasers", C.POINTER(LaserC)), ("insert", InsertProfileC * MAX_SPECIES),
                ("time", C.c_double), ("window_shift_fraction", C.c_double), ("step", C.c_int64),
                ("window_shifts_total", C.c_int64), ("window_started", C.c_int32), ("pad_", C.c_int32)]


class DriverState(C.Structure):     # cylgpu_driver_state
    _fields_ = [("time", C.c_double), ("step", C.c_int64), ("window_started", C.c_int32), ("pad_", C.c_int32),
                ("window_shift_fraction", C.c_double), ("window_shifts_total", C.c_int64),
                ("x_grid_min", C.c_double), ("x_min", C.c_double), ("x_max", C.c_double),
                ("x_grid_min_local", C.c_double), ("x_min_local", C.c_double), ("x_max_local", C.c_double),
                ("bc_field", C.c_int32 * 4), ("raw_bc_field", C.c_int32 * 4)]


_LIB = None


def load():
    """dlopen libcylgpu.so and bind every symbol; raises if the library or a symbol is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m cylindrical_epoch_b200.build [--shape=...]` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.cylgpu_ghost_cells() != NG:
        raise RuntimeError(f"{LIB_PATH} was built for ng = {lib.cylgpu_ghost_cells()}, CYL_SHAPE = {SHAPE} needs ng = {NG}")
    _LIB = lib
    return lib
