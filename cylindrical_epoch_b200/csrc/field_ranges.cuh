// field_ranges.cuh -- the column ranges of the communication-avoiding ("wide") field phases.  Plain inline
// functions, shared by the library (fields.cu / bcs.cu / api.cu) and by the kernel-emulation tests (tests/emul/),
// so that the CPU tests exercise the very range rules the GPU path uses.  Product code: no oracle here.
//
// The reference exchanges the ghost columns after every sweep (efield_bcs / bfield_bcs, boundary.F90:1355-1476:
// five field exchanges per step).  The arrays carry ng = 5 ghost columns but a sweep only reaches +-1, so a slab
// can advance the ghost column the NEXT sweep of the phase reads itself -- the same kernel on the
// same values the neighbour holds gives the same bits -- and exchange ONCE at the end of update_eb_fields_half
// and once at the end of update_eb_fields_final.  At the start of a phase all five ghost columns hold the
// neighbour's interior values (the closing exchange of the previous phase, the window's nine-array halo, or an
// uploaded reference state).  What the interior sweeps 0..nx read from beyond:
//   half:  nothing new: the reference's own sweeps cover column 0 (B(1) reads E(0)); its E exchange only feeds
//          B(0), a ghost column the closing exchange rewrites
//   final: the r_max line of Bx(nx) reads Br(nx+1)             -> the B sweep also does column nx+1
//          E(nx) reads Btheta(nx+1), whose r_max row is a line update (laser.f90:668-688, reading Er(nx),
//          Er(nx+1) and J(nx+1))                               -> that line also on column nx+1
//   final, with the moving window's shift_fields next (phase 2): the shift moves column nx+1 into the interior
//          and then exchanges all nine arrays itself, so the phase advances column nx+1 completely -- B sweep to
//          nx+2, both r_max lines one column further, E sweep to nx+1 -- and needs no exchange of its own
// Everything else a sweep reads in the ghost columns is what the phase started with.  The closing exchange then
// rewrites all five ghost columns, so the arrays at the end of each phase are the reference's, ghosts included.
// On a side without a neighbour (a domain boundary that is not periodic) the ranges are the reference's own.
#pragma once

struct FieldRanges {
  int e_lo, e_hi;       // bulk E sweep (reference: 0 .. nx)
  int b_lo, b_hi;       // bulk B sweep (reference: 0 .. nx)
  int obx_lo, obx_hi;   // r_max outflow, the Bx row (reference: 0 .. nx, 1 .. on x_min, .. nx-1 on x_max)
  int obt_lo, obt_hi;   // r_max outflow, the Btheta row (reference: 1 .. nx)
};

// phase 0: update_eb_fields_half, 1: update_eb_fields_final, 2: update_eb_fields_final right before shift_fields.  fill_l / fill_r: the halo fills the ghost columns
// of that side (a neighbour exists and the side is not a non-periodic domain boundary).
static inline FieldRanges field_ranges(int phase, bool wide, bool fill_l, bool fill_r, bool x_min_boundary,
                                       bool x_max_boundary, int nx) {
  FieldRanges R;
  R.e_lo = 0; R.e_hi = nx;
  R.b_lo = 0; R.b_hi = nx;
  R.obx_lo = x_min_boundary ? 1 : 0; R.obx_hi = x_max_boundary ? nx - 1 : nx;
  R.obt_lo = 1; R.obt_hi = nx;
  if (!wide) return R;
  if (phase == 1 && fill_r) { R.b_hi = nx + 1; R.obt_hi = nx + 1; }
  if (phase == 2 && fill_r) { R.b_hi = nx + 2; R.obx_hi = nx + 1; R.obt_hi = nx + 2; R.e_hi = nx + 1; }
  (void)fill_l;
  return R;
}
