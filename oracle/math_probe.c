/* TEST INFRASTRUCTURE: array entry points over include/pgd_math.h for tests/test_math.py (accuracy against double
 * libm) and tests/test_gpu_math.py (the sm_100a build gives the same bits). */
#include "../include/pgd_math.h"

void probe_sincos(const float* a, float* s, float* c, int n) { for (int i = 0; i < n; ++i) pgd_sincosf(a[i], s + i, c + i); }
void probe_atan2(const float* y, const float* x, float* r, int n) { for (int i = 0; i < n; ++i) r[i] = pgd_atan2f(y[i], x[i]); }
void probe_wrap(const float* a, float* r, int n) { for (int i = 0; i < n; ++i) r[i] = pgd_wrap_to_pi(a[i]); }
void probe_exp(const float* a, float* r, int n) { for (int i = 0; i < n; ++i) r[i] = pgd_expf(a[i]); }
void probe_pow10(const float* a, float* r, int n) { for (int i = 0; i < n; ++i) r[i] = pgd_pow10f(a[i]); }
void probe_tan(const float* a, float* r, int n) { for (int i = 0; i < n; ++i) r[i] = pgd_tanf(a[i]); }
void probe_asin(const float* a, float* r, int n) { for (int i = 0; i < n; ++i) r[i] = pgd_asinf(a[i]); }
void probe_log(const float* a, float* r, int n) { for (int i = 0; i < n; ++i) r[i] = pgd_logf(a[i]); }
void probe_noise(const float* p, float* r, int n, float sigma, float dropout, unsigned seed, unsigned call) {
  for (int i = 0; i < n; ++i) r[i] = pgd_lidar_noise(p[i], sigma, dropout, pgd_noise_key(seed, call, (unsigned)(i / 240), (unsigned)(i % 240)));
}
