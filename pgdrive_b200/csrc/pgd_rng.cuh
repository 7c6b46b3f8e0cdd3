/* The reference's random streams, restated for host + device: every stream is numpy's legacy
 * RandomState(MT19937) seeded with the first 8 bytes of sha512(str(seed)) as little-endian u32 words
 * (/root/reference/pgdrive/utils/random_utils.py:14-50,89-100), and the draws the reset path makes from it:
 * randint (masked rejection), random_sample (53-bit), choice with and without probabilities, list shuffle,
 * uniform(low, high) as used by Box.sample (utils/space.py:423-457).  The algorithms are numpy's published legacy
 * ones (numpy/random/mtrand.pyx, _mt19937.pyx, distributions.c; numpy is a dependency of the reference, pinned by
 * its setup.py only as "numpy"); parity is pinned by tests/test_device_mapgen.py against numpy itself.
 */
#ifndef PGD_RNG_CUH
#define PGD_RNG_CUH
#include <stdint.h>

#include "pgd_dd.cuh"

namespace pgdgen {

// ---------------------------------------------------------------------------------------------- sha512
PGD_HD inline uint64_t rotr64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }

/* first 8 digest bytes of sha512(decimal string of v), returned as the integer the reference builds from them:
 * lo u32 word + (hi u32 word << 32), both little-endian */
PGD_HD inline uint64_t hash_seed(uint64_t v) {
  const uint64_t K[80] = {
      0x428a2f98d728ae22ULL, 0x7137449123ef65cdULL, 0xb5c0fbcfec4d3b2fULL, 0xe9b5dba58189dbbcULL, 0x3956c25bf348b538ULL,
      0x59f111f1b605d019ULL, 0x923f82a4af194f9bULL, 0xab1c5ed5da6d8118ULL, 0xd807aa98a3030242ULL, 0x12835b0145706fbeULL,
      0x243185be4ee4b28cULL, 0x550c7dc3d5ffb4e2ULL, 0x72be5d74f27b896fULL, 0x80deb1fe3b1696b1ULL, 0x9bdc06a725c71235ULL,
      0xc19bf174cf692694ULL, 0xe49b69c19ef14ad2ULL, 0xefbe4786384f25e3ULL, 0x0fc19dc68b8cd5b5ULL, 0x240ca1cc77ac9c65ULL,
      0x2de92c6f592b0275ULL, 0x4a7484aa6ea6e483ULL, 0x5cb0a9dcbd41fbd4ULL, 0x76f988da831153b5ULL, 0x983e5152ee66dfabULL,
      0xa831c66d2db43210ULL, 0xb00327c898fb213fULL, 0xbf597fc7beef0ee4ULL, 0xc6e00bf33da88fc2ULL, 0xd5a79147930aa725ULL,
      0x06ca6351e003826fULL, 0x142929670a0e6e70ULL, 0x27b70a8546d22ffcULL, 0x2e1b21385c26c926ULL, 0x4d2c6dfc5ac42aedULL,
      0x53380d139d95b3dfULL, 0x650a73548baf63deULL, 0x766a0abb3c77b2a8ULL, 0x81c2c92e47edaee6ULL, 0x92722c851482353bULL,
      0xa2bfe8a14cf10364ULL, 0xa81a664bbc423001ULL, 0xc24b8b70d0f89791ULL, 0xc76c51a30654be30ULL, 0xd192e819d6ef5218ULL,
      0xd69906245565a910ULL, 0xf40e35855771202aULL, 0x106aa07032bbd1b8ULL, 0x19a4c116b8d2d0c8ULL, 0x1e376c085141ab53ULL,
      0x2748774cdf8eeb99ULL, 0x34b0bcb5e19b48a8ULL, 0x391c0cb3c5c95a63ULL, 0x4ed8aa4ae3418acbULL, 0x5b9cca4f7763e373ULL,
      0x682e6ff3d6b2b8a3ULL, 0x748f82ee5defb2fcULL, 0x78a5636f43172f60ULL, 0x84c87814a1f0ab72ULL, 0x8cc702081a6439ecULL,
      0x90befffa23631e28ULL, 0xa4506cebde82bde9ULL, 0xbef9a3f7b2c67915ULL, 0xc67178f2e372532bULL, 0xca273eceea26619cULL,
      0xd186b8c721c0c207ULL, 0xeada7dd6cde0eb1eULL, 0xf57d4f7fee6ed178ULL, 0x06f067aa72176fbaULL, 0x0a637dc5a2c898a6ULL,
      0x113f9804bef90daeULL, 0x1b710b35131c471bULL, 0x28db77f523047d84ULL, 0x32caab7b40c72493ULL, 0x3c9ebe0a15c9bebcULL,
      0x431d67c49c100d4cULL, 0x4cc5d4becb3e42b6ULL, 0x597f299cfc657e2aULL, 0x5fcb6fab3ad6faecULL, 0x6c44198c4a475817ULL};
  // decimal digits, most significant first
  unsigned char msg[24];
  int n = 0;
  {
    unsigned char tmp[24];
    int t = 0;
    if (v == 0) tmp[t++] = '0';
    while (v > 0) {
      tmp[t++] = (unsigned char)('0' + (v % 10));
      v /= 10;
    }
    while (t > 0) msg[n++] = tmp[--t];
  }
  uint64_t w[80];
  for (int i = 0; i < 16; ++i) w[i] = 0;
  for (int i = 0; i < n; ++i) w[i >> 3] |= (uint64_t)msg[i] << (56 - 8 * (i & 7));
  w[n >> 3] |= (uint64_t)0x80 << (56 - 8 * (n & 7));
  w[15] = (uint64_t)n * 8;
  for (int i = 16; i < 80; ++i) {
    uint64_t s0 = rotr64(w[i - 15], 1) ^ rotr64(w[i - 15], 8) ^ (w[i - 15] >> 7);
    uint64_t s1 = rotr64(w[i - 2], 19) ^ rotr64(w[i - 2], 61) ^ (w[i - 2] >> 6);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint64_t a = 0x6a09e667f3bcc908ULL, b = 0xbb67ae8584caa73bULL, c = 0x3c6ef372fe94f82bULL, d = 0xa54ff53a5f1d36f1ULL;
  uint64_t e = 0x510e527fade682d1ULL, f = 0x9b05688c2b3e6c1fULL, g = 0x1f83d9abfb41bd6bULL, h = 0x5be0cd19137e2179ULL;
  for (int i = 0; i < 80; ++i) {
    uint64_t S1 = rotr64(e, 14) ^ rotr64(e, 18) ^ rotr64(e, 41);
    uint64_t ch = (e & f) ^ (~e & g);
    uint64_t t1 = h + S1 + ch + K[i] + w[i];
    uint64_t S0 = rotr64(a, 28) ^ rotr64(a, 34) ^ rotr64(a, 39);
    uint64_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint64_t t2 = S0 + mj;
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  uint64_t h0 = 0x6a09e667f3bcc908ULL + a;  // big-endian bytes of h0 are digest[0..8)
  // digest bytes b0..b7 = h0 >> 56 ... h0 & 0xff ; lo = b0 | b1<<8 | b2<<16 | b3<<24 ; hi likewise from b4..b7
  uint32_t lo = 0, hi = 0;
  for (int i = 0; i < 4; ++i) {
    lo |= (uint32_t)((h0 >> (56 - 8 * i)) & 0xff) << (8 * i);
    hi |= (uint32_t)((h0 >> (24 - 8 * i)) & 0xff) << (8 * i);
  }
  return (uint64_t)lo | ((uint64_t)hi << 32);
}

// ---------------------------------------------------------------------------------------------- MT19937
struct MT {
  uint32_t mt[624];
  int pos;
};

PGD_HD inline void mt_init_genrand(MT* s, uint32_t seed) {
  s->mt[0] = seed;
  for (int i = 1; i < 624; ++i) s->mt[i] = 1812433253U * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
  s->pos = 624;
}

PGD_HD inline void mt_init_by_array(MT* s, const uint32_t* key, int len) {
  mt_init_genrand(s, 19650218U);
  int i = 1, j = 0;
  int k = 624 > len ? 624 : len;
  for (; k; --k) {
    s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1664525U)) + key[j] + (uint32_t)j;
    ++i; ++j;
    if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
    if (j >= len) j = 0;
  }
  for (k = 623; k; --k) {
    s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1566083941U)) - (uint32_t)i;
    ++i;
    if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
  }
  s->mt[0] = 0x80000000U;
  s->pos = 624;
}

/* One output.  The twist is done one word at a time, in place and in order, which yields the same sequence as the
 * textbook block regeneration (word k only needs the OLD words k + 1 and k + 397, both still untouched when k is
 * produced; past 227 it needs NEW words that are already there) -- most of the reference's streams are seeded, asked
 * for one or two numbers and dropped. */
PGD_HD inline uint32_t mt_next(MT* s) {
  if (s->pos >= 624) s->pos = 0;  // a new block of 624 starts
  const int k = s->pos++;
  uint32_t* mt = s->mt;
  const int k1 = (k == 623) ? 0 : k + 1;
  const int km = (k < 227) ? k + 397 : k - 227;
  uint32_t y = (mt[k] & 0x80000000U) | (mt[k1] & 0x7fffffffU);
  y = mt[km] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
  mt[k] = y;
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680U;
  y ^= (y << 15) & 0xefc60000U;
  y ^= (y >> 18);
  return y;
}

/* get_np_random(seed): RandomState().seed(u32 words of hash_seed(seed)) */
PGD_HD inline void mt_seeded(MT* s, uint64_t seed) {
  uint64_t big = hash_seed(seed);
  uint32_t words[2] = {(uint32_t)(big & 0xffffffffU), (uint32_t)(big >> 32)};
  int n = words[1] ? 2 : 1;  // big == 0 -> [0]
  mt_init_by_array(s, words, n);
}

/* random_sample(): 53-bit double in [0, 1) */
PGD_HD inline double mt_double(MT* s) {
  int32_t a = (int32_t)(mt_next(s) >> 5), b = (int32_t)(mt_next(s) >> 6);
  return (a * 67108864.0 + b) / 9007199254740992.0;
}

/* masked rejection sampling of [0, max] with 32-bit draws: randint(0, max + 1), random_interval(max) */
PGD_HD inline uint32_t mt_interval(MT* s, uint32_t max) {
  if (max == 0) return 0;  // no draw is consumed
  uint32_t mask = max;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  uint32_t v;
  while ((v = (mt_next(s) & mask)) > max) {
  }
  return v;
}
PGD_HD inline uint32_t mt_randint(MT* s, uint32_t high) { return mt_interval(s, high - 1); }  // randint(0, high)

/* choice(n, p=p): searchsorted(cumsum(p) / cumsum(p)[-1], random_sample(), side="right") */
PGD_HD inline int mt_choice_p(MT* s, const double* p, int n) {
  double cdf[16];
  double acc = 0.0;
  for (int i = 0; i < n; ++i) {
    acc = (i == 0) ? p[0] : acc + p[i];
    cdf[i] = acc;
  }
  double last = cdf[n - 1];
  for (int i = 0; i < n; ++i) cdf[i] /= last;
  double u = mt_double(s);
  int idx = 0;
  while (idx < n && cdf[idx] <= u) ++idx;
  return idx;
}

/* first random_sample() of a fresh get_np_random(q) stream: every Box of a parameter space is seeded with the same
 * integer q and draws once (base_class/base_runnable.py:81-88, utils/space.py:109-113) */
PGD_HD inline double first_sample_of(MT* tmp, uint64_t q) {
  mt_seeded(tmp, q);
  return mt_double(tmp);
}
/* float Box sample rounded to float32 and integer Box sample (utils/space.py:423-457) */
PGD_HD inline double box_f32(double low, double high, double u) {
  double lo = (double)(float)low, hi = (double)(float)high;
  return (double)(float)(lo + (hi - lo) * u);
}
PGD_HD inline int box_int(int low, int high, double u) {
  double lo = (double)low, hi = (double)(high + 1);
  return (int)floor(lo + (hi - lo) * u);
}

}  // namespace pgdgen
#endif
