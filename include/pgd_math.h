/* float32 elementary functions with ONE operation sequence for every build.
 *
 * The step is checked bit for bit against the CPU oracle (oracle/pgd_oracle.c).  IEEE add / mul / div / sqrt / fma
 * give the same bits on the host (gcc -ffp-contract=off) and on the device (nvcc -fmad=false); sinf / cosf / atan2f /
 * expf / powf / tanf do not: glibc and CUDA's libdevice are different 1-2 ulp approximations, and a last-bit
 * difference in a heading turns, hundreds of steps later, into a contact flag that fires one step apart.  Every
 * transcendental of the step therefore comes from this header: range reduction and polynomial evaluation are written
 * as explicit fmaf() chains, which both compilers keep as written.  Accuracy (tests/test_math.py, against double
 * libm over the ranges the step uses): sin / cos <= 1.5 ulp for |a| <= 64, atan2 <= 2 ulp, exp <= 1 ulp.
 *
 * No table, no branch on the value except the quadrant select: ~25 instructions for sincos where CUDA's accurate
 * sincosf carries a Payne-Hanek slow path (11 % of the executed instructions of the round-1 kernel).
 *
 * Reference semantics restated here: utils/math_utils.py:32-33 wrap_to_pi, numpy cos / sin / arctan2 (float64 there).
 */
#ifndef PGD_MATH_H
#define PGD_MATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define PGD_MATH_FN __host__ __device__ __forceinline__
#else
#define PGD_MATH_FN static inline
#endif

#define PGD_PI_F 3.14159265358979323846f
#define PGD_TWO_PI_F 6.28318530717958647692f

/* sin and cos of a (radians), |a| up to a few hundred.  Cody-Waite reduction by pi/2 in three parts (the products
 * k * part are exact for |k| < 2^9), then the classic minimax polynomials on [-pi/4, pi/4]. */
PGD_MATH_FN void pgd_sincosf(float a, float* s_out, float* c_out) {
  const float k = rintf(a * 0.63661977236758134308f);
  float r = fmaf(k, -1.5703125f, a);                       /* pi/2 = 1.5703125 + 4.837512969970703125e-4 + 7.549789954e-8 */
  r = fmaf(k, -4.837512969970703125e-4f, r);
  r = fmaf(k, -7.54978995489188e-8f, r);
  const float r2 = r * r;
  float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  sp = fmaf(sp, r2, -1.6666654611e-1f);
  const float sn = fmaf(sp * r2, r, r);
  float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  cp = fmaf(cp, r2, 4.166664568298827e-2f);
  const float cs = fmaf(cp * r2, r2, fmaf(r2, -0.5f, 1.0f));
  const int q = (int)k & 3;
  const float s = (q & 1) ? cs : sn, c = (q & 1) ? sn : cs;
  *s_out = (q & 2) ? -s : s;
  *c_out = ((q + 1) & 2) ? -c : c;
}

/* atan2(y, x): atan of min / max on [0, 1] by an odd minimax polynomial (relative error 1.5e-8 before rounding),
 * then the octant is unfolded.  atan2(0, 0) = 0. */
PGD_MATH_FN float pgd_atan2f(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float t = mx > 0.0f ? mn / mx : 0.0f;
  const float s = t * t;
  float p = 2.9744952975e-03f;
  p = fmaf(p, s, -1.6580837056e-02f);
  p = fmaf(p, s, 4.3553045658e-02f);
  p = fmaf(p, s, -7.5805442375e-02f);
  p = fmaf(p, s, 1.0678929485e-01f);
  p = fmaf(p, s, -1.4214208243e-01f);
  p = fmaf(p, s, 1.9994137434e-01f);
  p = fmaf(p, s, -3.3333166998e-01f);
  float r = fmaf(p * s, t, t);
  if (ay > ax) r = 1.57079632679489661923f - r;
  if (x < 0.0f) r = PGD_PI_F - r;
  return y < 0.0f ? -r : r;
}

/* ((x + pi) mod 2 pi) - pi for |x| of a few turns (math_utils.py:32-33); the result lies in [-pi, pi] and differs
 * from the reference's only in which end of the interval an exact odd multiple of pi lands on. */
PGD_MATH_FN float pgd_wrap_to_pi(float x) {
  const float k = rintf(x * 0.15915494309189533577f);
  float r = fmaf(k, -6.28125f, x);                          /* 2 pi = 6.28125 + 1.9350051879882812e-3 + 3.0199159819e-7 */
  r = fmaf(k, -1.9350051879882812e-3f, r);
  r = fmaf(k, -3.0199159819e-7f, r);
  return r;
}

/* e^x for x in [-80, 80] (the step only needs x = 0.01 * speed_km_h, i.e. [0, ~1]). */
PGD_MATH_FN float pgd_expf(float x) {
  x = fminf(fmaxf(x, -80.0f), 80.0f);
  const float k = rintf(x * 1.44269504088896341f);
  float r = fmaf(k, -0.693359375f, x);                      /* ln 2 = 0.693359375 - 2.12194440e-4 */
  r = fmaf(k, 2.12194440e-4f, r);
  float p = 1.9875691500e-4f;
  p = fmaf(p, r, 1.3981999507e-3f);
  p = fmaf(p, r, 8.3334519073e-3f);
  p = fmaf(p, r, 4.1665795894e-2f);
  p = fmaf(p, r, 1.6666665459e-1f);
  p = fmaf(p, r, 5.0000001201e-1f);
  const float e = fmaf(p * r, r, r) + 1.0f;
  const int32_t bits = ((int32_t)k + 127) << 23;            /* 2^k, k in [-116, 116] */
  float scale;
  memcpy(&scale, &bits, 4);
  return e * scale;
}

/* ln(x) for normal x > 0 (Box-Muller of the lidar noise): exponent split by bit operations, mantissa in [sqrt(1/2),
 * sqrt(2)), the classic degree-9 polynomial. */
PGD_MATH_FN float pgd_logf(float x) {
  int32_t bits;
  memcpy(&bits, &x, 4);
  int32_t e = ((bits >> 23) & 0xff) - 126;                 /* x = m * 2^e, m in [0.5, 1) */
  bits = (bits & 0x007fffff) | 0x3f000000;
  float m;
  memcpy(&m, &bits, 4);
  if (m < 0.707106781186547524f) {
    e -= 1;
    m = m + m - 1.0f;
  } else {
    m = m - 1.0f;
  }
  const float z = m * m;
  float p = 7.0376836292e-2f;
  p = fmaf(p, m, -1.1514610310e-1f);
  p = fmaf(p, m, 1.1676998740e-1f);
  p = fmaf(p, m, -1.2420140846e-1f);
  p = fmaf(p, m, 1.4249322787e-1f);
  p = fmaf(p, m, -1.6668057665e-1f);
  p = fmaf(p, m, 2.0000714765e-1f);
  p = fmaf(p, m, -2.4999993993e-1f);
  p = fmaf(p, m, 3.3333331174e-1f);
  float y = p * m * z;
  const float fe = (float)e;
  y = fmaf(fe, -2.12194440e-4f, y);
  y = fmaf(z, -0.5f, y);
  return fmaf(fe, 0.693359375f, m + y);
}

/* Counter-based random numbers for the lidar noise (obs/state_obs.py:172-182 draws from the process-global numpy
 * generator, so only the distribution can be matched): a 32-bit mix of (seed, call, environment, beam), the same on
 * the device and in the oracle. */
PGD_MATH_FN uint32_t pgd_mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
  return h;
}
PGD_MATH_FN uint32_t pgd_noise_key(uint32_t seed, uint32_t call, uint32_t env, uint32_t beam) {
  return pgd_mix32(pgd_mix32(pgd_mix32(seed ^ 0x9e3779b9u) + call) * 0x85ebca6bu + env * 1025u + beam);
}
/* clip(p + N(0, sigma), 0, 1), then 0 with probability `dropout` -- the two steps of _add_noise_to_cloud_points */
PGD_MATH_FN float pgd_lidar_noise(float p, float sigma, float dropout, uint32_t key) {
  if (sigma > 0.0f) {
    const uint32_t h1 = pgd_mix32(key + 0x68bc21ebu), h2 = pgd_mix32(key + 0x02e5be93u);
    const float u1 = (float)((h1 >> 8) + 1u) * (1.0f / 16777216.0f);  /* (0, 1] */
    const float u2 = (float)(h2 >> 8) * (1.0f / 16777216.0f);         /* [0, 1) */
    float sn, cs;
    pgd_sincosf(6.28318530717958647692f * u2, &sn, &cs);
    const float z = sqrtf(-2.0f * pgd_logf(u1)) * cs;
    p = fminf(fmaxf(p + sigma * z, 0.0f), 1.0f);
  }
  if (dropout > 0.0f) {
    const float u = (float)(pgd_mix32(key + 0x3c6ef372u) >> 8) * (1.0f / 16777216.0f);
    if (u < dropout) p = 0.0f;
  }
  return p;
}

/* x^10 by squaring (IDM free-road term (v / v0)^10, policy/idm_policy.py:254-262). */
PGD_MATH_FN float pgd_pow10f(float x) {
  const float x2 = x * x, x4 = x2 * x2, x8 = x4 * x4;
  return x8 * x2;
}

/* tan(a) for |a| <= 1.4 (steering angle): sin / cos of the functions above. */
PGD_MATH_FN float pgd_tanf(float a) {
  float s, c;
  pgd_sincosf(a, &s, &c);
  return s / c;
}

/* asin(x) for x in [0, 1] (only used for the conservative beam window of the lidar cull). */
PGD_MATH_FN float pgd_asinf(float x) { return pgd_atan2f(x, sqrtf(fmaxf(1.0f - x * x, 0.0f))); }

#endif
