"""Constants shared with the reference (constants.F90:55-72,171-201,526-548)."""
import os as _os

# The particle shape is a compile-time choice of the reference and of the library (csrc/shape.cuh): CYL_SHAPE =
# triangle (default) | tophat | bspline3 selects the library of that shape for the whole process (_lib.py).
# ng = jng = png + 2 follows the shape (constants.F90:524-545): the layout of every array that crosses the C-ABI.
SHAPE = _os.environ.get("CYL_SHAPE", "triangle") or "triangle"
NG = {"triangle": 5, "tophat": 4, "bspline3": 6}[SHAPE]
# boundary-condition codes
BC_PERIODIC, BC_OTHER, BC_SIMPLE_LASER, BC_SIMPLE_OUTFLOW, BC_OPEN = 1, 2, 3, 4, 5
BC_ZERO_GRADIENT, BC_CLAMP, BC_REFLECT, BC_CONDUCT, BC_THERMAL = 7, 8, 9, 10, 11
BC_CPML_LASER, BC_CPML_OUTFLOW, BC_ZERO_B = 12, 13, 16
BD_X_MIN, BD_X_MAX, BD_Y_MIN, BD_Y_MAX = 0, 1, 2, 3
# physical constants
PI = 3.141592653589793238462643383279503
Q0 = 1.602176565e-19
M0 = 9.10938291e-31
C_LIGHT = 2.99792458e8
KB = 1.3806488e-23
EPSILON0 = 8.854187817620389850536563031710750e-12

FIELD_NAMES = ["exm", "erm", "etm", "bxm", "brm", "btm", "jxm", "jrm", "jtm",
               "bxm_old", "brm_old", "btm_old", "jxm_old", "jrm_old", "jtm_old"]
SNAP_NAMES = [f"{n}_x_min" for n in FIELD_NAMES[:6]] + [f"{n}_x_max" for n in FIELD_NAMES[:6]]

TRANSPORT_NONE, TRANSPORT_NCCL, TRANSPORT_CALLBACK, TRANSPORT_FABRIC = 0, 1, 2, 3
