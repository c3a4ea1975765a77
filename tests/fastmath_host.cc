// Host build of branson_b200/csrc/fastmath.cuh for tests/test_fastmath.py (same source as the device code; only the
// reciprocal seed differs).  g++ -O2 -ffp-contract=off -shared -fPIC -DBG_FASTMATH_HOST
#include "../branson_b200/csrc/fastmath.cuh"

extern "C" {
void fmh_exp(const double *x, double *y, long n) { for (long i = 0; i < n; ++i) y[i] = bg::fm_exp_flush(x[i]); }
void fmh_log(const double *x, double *y, long n) { for (long i = 0; i < n; ++i) y[i] = bg::fm_log_pos(x[i]); }
void fmh_log_u01_bits(const uint64_t *w, double *y, long n) {
  for (long i = 0; i < n; ++i) y[i] = bg::fm_log_pos_scaled((double)((w[i] >> 11) | 1ULL), -53);
}
void fmh_sincos(const double *x, double *s, double *c, long n) { for (long i = 0; i < n; ++i) bg::fm_sincos(x[i], &s[i], &c[i]); }
}
