/* Double-double arithmetic and (practically) correctly rounded sin / cos / atan / atan2 in float64.
 *
 * Why this exists: the reference builds its maps in Python float64 on glibc's libm, and several DISCRETE
 * decisions hang on the last bit of a trigonometric result -- e.g. the number of traffic spawn slots of a lane is
 * int(length / 10) (manager/traffic_manager.py:265) and intersection exits are exactly 30 m long up to rounding
 * (component/blocks/intersection.py:24, create_block_utils.py:16-59), so length / 10 lands on either side of 3.0.
 * CUDA's libdevice sin / cos / atan2 are 1-2 ulp functions; glibc's are < 0.55 ulp, i.e. they return the correctly
 * rounded value except in rare hard cases.  The device map generator therefore evaluates these functions to ~100
 * bits (double-double) and rounds once, which reproduces glibc bit-for-bit in all but those rare cases (rate
 * measured in tests/test_device_mapgen.py) -- and makes the SAME source give the SAME bits when it is compiled for
 * the host (tests, no GPU needed) and for sm_100a.
 *
 * Only +, -, *, /, sqrt and fma are used (all IEEE-exact on both targets); nothing here calls libm / libdevice
 * except as a first guess that is then corrected.
 */
#ifndef PGD_DD_CUH
#define PGD_DD_CUH
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define PGD_HD __host__ __device__
#else
#define PGD_HD
#endif

namespace pgdgen {

struct dd {
  double hi, lo;
};

PGD_HD inline dd two_sum(double a, double b) {
  double s = a + b;
  double bb = s - a;
  double e = (a - (s - bb)) + (b - bb);
  return dd{s, e};
}
PGD_HD inline dd quick_two_sum(double a, double b) {  // |a| >= |b|
  double s = a + b;
  return dd{s, b - (s - a)};
}
PGD_HD inline dd two_prod(double a, double b) {
  double p = a * b;
  return dd{p, fma(a, b, -p)};
}
PGD_HD inline dd dd_add(dd a, dd b) {
  dd s = two_sum(a.hi, b.hi);
  dd t = two_sum(a.lo, b.lo);
  s.lo += t.hi;
  s = quick_two_sum(s.hi, s.lo);
  s.lo += t.lo;
  return quick_two_sum(s.hi, s.lo);
}
PGD_HD inline dd dd_add_d(dd a, double b) {
  dd s = two_sum(a.hi, b);
  s.lo += a.lo;
  return quick_two_sum(s.hi, s.lo);
}
PGD_HD inline dd dd_neg(dd a) { return dd{-a.hi, -a.lo}; }
PGD_HD inline dd dd_sub(dd a, dd b) { return dd_add(a, dd_neg(b)); }
PGD_HD inline dd dd_mul(dd a, dd b) {
  dd p = two_prod(a.hi, b.hi);
  p.lo += a.hi * b.lo + a.lo * b.hi;
  return quick_two_sum(p.hi, p.lo);
}
PGD_HD inline dd dd_mul_d(dd a, double b) {
  dd p = two_prod(a.hi, b);
  p.lo += a.lo * b;
  return quick_two_sum(p.hi, p.lo);
}
PGD_HD inline dd dd_div(dd a, dd b) {
  double q1 = a.hi / b.hi;
  dd r = dd_sub(a, dd_mul_d(b, q1));
  double q2 = r.hi / b.hi;
  r = dd_sub(r, dd_mul_d(b, q2));
  double q3 = r.hi / b.hi;
  dd q = quick_two_sum(q1, q2);
  return dd_add_d(q, q3);
}
PGD_HD inline dd dd_div_d(dd a, double b) { return dd_div(a, dd{b, 0.0}); }

/* pi/2 to ~160 bits: three non-overlapping doubles */
#define PGD_PIO2_1 1.5707963267948966
#define PGD_PIO2_2 6.123233995736766e-17
#define PGD_PIO2_3 -1.4973849048591698e-33

/* sin and cos of x (|x| < ~1e5) as double-doubles. */
PGD_HD inline void dd_sincos(double x, dd* s_out, dd* c_out) {
  double kf = rint(x * 0.63661977236758134308);  // x * 2/pi
  // r = x - k*pi/2 ; the first subtraction is exact (two_prod + two_sum), the others lose < 2^-104 relative
  dd p1 = two_prod(kf, PGD_PIO2_1);
  dd r = two_sum(x, -p1.hi);
  r.lo -= p1.lo;
  r = quick_two_sum(r.hi, r.lo);
  r = dd_sub(r, two_prod(kf, PGD_PIO2_2));
  r = dd_sub(r, two_prod(kf, PGD_PIO2_3));
  // Taylor series on |r| <= pi/4 (+ a little), Horner form in r^2 with the signed inverse factorials as
  // double-double constants: 14 terms leave < 2^-106
  const double SC[14][2] = {
      {-0.16666666666666666, -9.25185853854297e-18},   {0.008333333333333333, 1.1564823173178714e-19},
      {-0.0001984126984126984, -1.7209558293420705e-22}, {2.7557319223985893e-06, -1.858393274046472e-22},
      {-2.505210838544172e-08, 1.448814070935912e-24},  {1.6059043836821613e-10, 1.2585294588752098e-26},
      {-7.647163731819816e-13, -7.03872877733453e-30},  {2.8114572543455206e-15, 1.6508842730861433e-31},
      {-8.22063524662433e-18, -2.2141894119604265e-34}, {1.9572941063391263e-20, -1.3643503830087908e-36},
      {-3.868170170630684e-23, 8.843177655482344e-40},  {6.446950284384474e-26, -1.9330404233703465e-42},
      {-9.183689863795546e-29, -1.4303150396787322e-45}, {1.1309962886447716e-31, 1.0498015412959506e-47}};
  const double CC[14][2] = {
      {-0.5, 0.0},                                      {0.041666666666666664, 2.3129646346357427e-18},
      {-0.001388888888888889, 5.300543954373577e-20},   {2.48015873015873e-05, 2.1511947866775882e-23},
      {-2.755731922398589e-07, -2.3767714622250297e-23}, {2.08767569878681e-09, -1.20734505911326e-25},
      {-1.1470745597729725e-11, -2.0655512752830745e-28}, {4.779477332387385e-14, 4.399205485834081e-31},
      {-1.5619206968586225e-16, -1.1910679660273754e-32}, {4.110317623312165e-19, 1.4412973378659527e-36},
      {-8.896791392450574e-22, 7.911402614872376e-38},  {1.6117375710961184e-24, -3.6846573564509766e-41},
      {-2.4795962632247976e-27, 1.2953730964765229e-43}, {3.279889237069838e-30, 1.5117542744029879e-46}};
  const dd r2 = dd_mul(r, r);
  dd ps = dd{SC[13][0], SC[13][1]}, pc = dd{CC[13][0], CC[13][1]};
  for (int k = 12; k >= 0; --k) {
    ps = dd_add(dd_mul(ps, r2), dd{SC[k][0], SC[k][1]});
    pc = dd_add(dd_mul(pc, r2), dd{CC[k][0], CC[k][1]});
  }
  const dd s = dd_add(r, dd_mul(dd_mul(ps, r2), r));   // r + r^3 * (...)
  const dd c = dd_add_d(dd_mul(pc, r2), 1.0);          // 1 + r^2 * (...)
  long long q = (long long)kf;
  switch (q & 3) {
    case 0: *s_out = s; *c_out = c; break;
    case 1: *s_out = c; *c_out = dd_neg(s); break;
    case 2: *s_out = dd_neg(s); *c_out = dd_neg(c); break;
    default: *s_out = dd_neg(c); *c_out = s; break;
  }
}

/* round-to-nearest of a normalised double-double is its high word */
PGD_HD inline double cr_sin(double x) {
  if (x == 0.0) return x;
  dd s, c;
  dd_sincos(x, &s, &c);
  return s.hi;
}
PGD_HD inline double cr_cos(double x) {
  dd s, c;
  dd_sincos(x, &s, &c);
  return c.hi;
}

/* atan2(y, x): first guess from the platform's atan2 (<= 2 ulp on either target), then ONE Newton step carried
 * out in double-double: tan(theta - a0) = (y cos a0 - x sin a0) / (x cos a0 + y sin a0). */
PGD_HD inline double cr_atan2(double y, double x) {
  if (y == 0.0 && x > 0.0) return y;  // +-0, like IEEE atan2
  if (x == 0.0 && y == 0.0) return atan2(y, x);
  double a0 = atan2(y, x);
  dd s, c;
  dd_sincos(a0, &s, &c);
  dd num = dd_sub(dd_mul_d(c, y), dd_mul_d(s, x));
  dd den = dd_add(dd_mul_d(c, x), dd_mul_d(s, y));
  dd delta = dd_div(num, den);  // ~1e-16: atan(delta) = delta up to 1e-48
  dd r = two_sum(a0, delta.hi);
  r.lo += delta.lo;
  return r.hi + r.lo;
}
PGD_HD inline double cr_atan(double v) { return cr_atan2(v, 1.0); }

/* Python's float % and // (Objects/floatobject.c float_rem / float_floor_div); numpy scalars behave the same
 * (npy_divmod). */
PGD_HD inline double py_mod(double x, double y) {
  double m = fmod(x, y);
  if (m != 0.0) {
    if ((y < 0) != (m < 0)) m += y;
  } else {
    m = copysign(0.0, y);
  }
  return m;
}
PGD_HD inline double py_floordiv(double x, double y) {
  double m = fmod(x, y);
  double div = (x - m) / y;
  if (m != 0.0 && ((y < 0) != (m < 0))) div -= 1.0;
  if (div != 0.0) {
    double f = floor(div);
    if (div - f > 0.5) f += 1.0;
    return f;
  }
  return copysign(0.0, x / y);
}

}  // namespace pgdgen
#endif
