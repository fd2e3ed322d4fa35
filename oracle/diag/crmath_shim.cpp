// crmath_shim.cpp — DIAGNOSTIC ONLY, not part of the oracle's asserted mode.
//
// Wraps the CUDA path's correctly-rounded f64 trig (retto_b200/csrc/rt_fmath.h, host build) behind three C
// functions that oracle.set_libm(1) installs into the oracle through orc_set_trig_hooks.  With the hooks in
// place the oracle performs the same f64 trig as the kernels, so a second, non-asserted run can COUNT the boxes
// on which a glibc build of the reference and the CUDA path could disagree (tests print that count; expected 0).
// The oracle itself (retto_oracle.cpp) includes no product source and defaults to glibc.
#include "../../retto_b200/csrc/rt_fmath.h"

extern "C" {
double diag_cr_atan2(double y, double x) { return rtm::rt_atan2(y, x); }
void diag_cr_sincos(double a, double* s, double* c) { rtm::rt_sincos(a, s, c); }
double diag_cr_acos(double v) { return rtm::rt_acos(v); }
}
