// The gather's wire format for observation rows (include/pgdrive_b200.h: pgd_pack_rows / pgd_expand_rows /
// pgd_packed_row_words).  No counterpart in the reference (one engine per process, engine/engine_utils.py:8-15).
//
// A row is [head | 240 lidar beams]; most beams are exactly 1.0 ("no hit within 50 m").  Packed row, at a fixed stride of
// pgd_packed_row_words(obs_dim) floats (obs_dim + 8 rounded up to 32 words, so every packed row starts on a 128-byte
// line): the head unchanged, a 240-bit hit mask (8 words), then the values of the beams that are not 1.0, in beam
// order, then zeros up to the next 32-byte boundary.  Only what is written crosses NVLink.
//
// One warp per row, 8 rows per CTA.  The dense row (1 096 bytes, 8-byte aligned) is read and written with plain
// coalesced 4-byte accesses -- measured against a variant that moved whole 32-row tiles through shared memory with 16-byte
// accesses (profiles/r04c_rows_bench_*.json: the tiles were 25-45 % slower; a barrier per tile costs more than the
// partial sectors; r04f_rows_bench_v4.json: expanding PAIRS of rows -- 2 192 bytes = 137 x 16 -- with 16-byte stores of
// whole sectors was 20 % slower as well).  The PACKED row is what crosses NVLink: it is assembled in shared memory and leaves as 16-byte stores of
// whole 32-byte sectors -- one fully coalesced store instruction for a typical row instead of a dozen small ones.
#include <cuda_runtime.h>
#include <stdint.h>

#include "pgd_internal.h"

#define ROWS_WARPS 8

static __host__ __device__ inline int packed_words(int obs_dim) { return (obs_dim + 8 + 31) / 32 * 32; }

extern "C" int32_t pgd_packed_row_words(int32_t obs_dim) { return obs_dim < PGD_LIDAR_BEAMS ? -1 : packed_words(obs_dim); }

__global__ void __launch_bounds__(ROWS_WARPS * 32)
pgd_pack_rows_kernel(const float* __restrict__ dense, float* __restrict__ packed, int n_rows, int obs_dim) {
  extern __shared__ __align__(16) float rows_smem[];  // [8 warps][stride]: the packed row of each warp
  const int stride = packed_words(obs_dim), head = obs_dim - PGD_LIDAR_BEAMS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * ROWS_WARPS + warp;
  if (row >= n_rows) return;
  const float* src = dense + (size_t)row * obs_dim;
  float* out = rows_smem + warp * stride;
  for (int i = lane; i < head; i += 32) out[i] = __ldcs(src + i);
  float v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = c * 32 + lane < PGD_LIDAR_BEAMS ? __ldcs(src + head + c * 32 + lane) : 1.0f;
  int base = head + 8;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const bool hit = __float_as_uint(v[c]) != 0x3f800000u;  // the bit pattern: the row comes back exactly
    const unsigned mask = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) out[head + c] = __uint_as_float(mask);
    if (hit) out[base + __popc(mask & ((1u << lane) - 1u))] = v[c];
    base += __popc(mask);
  }
  const int padded = (base + 7) & ~7;  // whole 32-byte sectors; <= stride because stride is a multiple of 32 words
  for (int i = base + lane; i < padded; i += 32) out[i] = 0.0f;
  __syncwarp();
  float* dst = packed + (size_t)row * stride;
  if ((reinterpret_cast<uintptr_t>(packed) & 15) == 0) {
    for (int i = lane; i < (padded >> 2); i += 32)
      reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(out)[i];
  } else {
    for (int i = lane; i < padded; i += 32) dst[i] = out[i];
  }
}

__global__ void __launch_bounds__(ROWS_WARPS * 32)
pgd_expand_rows_kernel(const float* __restrict__ packed, float* __restrict__ dense, int n_rows, int obs_dim) {
  const int stride = packed_words(obs_dim), head = obs_dim - PGD_LIDAR_BEAMS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * ROWS_WARPS + warp;
  if (row >= n_rows) return;
  const float* src = packed + (size_t)row * stride;
  float* dst = dense + (size_t)row * obs_dim;
  const unsigned my_mask = lane < 8 ? __float_as_uint(__ldcs(src + head + lane)) : 0u;
  for (int i = lane; i < head; i += 32) dst[i] = __ldcs(src + i);
  int base = head + 8;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const unsigned mask = __shfl_sync(0xffffffffu, my_mask, c);
    const int beam = c * 32 + lane;
    float v = 1.0f;
    if ((mask >> lane) & 1u) v = src[min(base + __popc(mask & ((1u << lane) - 1u)), stride - 1)];
    if (beam < PGD_LIDAR_BEAMS) dst[head + beam] = v;
    base += __popc(mask);
  }
}

// Delta expansion.  Rank 0 expands step t's rows into the SAME whole-batch buffer that holds the rows of step t - depth,
// and a lidar beam that was 1.0 then and is 1.0 now needs no store: with `mask_state` = the hit masks of the rows the
// buffer holds (8 words per row, kept by the caller next to the buffer), a row costs its head, the beams that were or
// are hits, and the new mask -- about 400 bytes of traffic instead of 1 300 (r04g_rows_bench_delta.json).  `full` != 0
// writes every beam and only initialises the state (first use of a buffer).  The result is bit-identical to
// pgd_expand_rows as long as nobody else writes the buffer's rows in between.
__global__ void __launch_bounds__(ROWS_WARPS * 32)
pgd_expand_rows_delta_kernel(const float* __restrict__ packed, float* __restrict__ dense, uint32_t* __restrict__ mask_state,
                             int n_rows, int obs_dim, int full) {
  const int stride = packed_words(obs_dim), head = obs_dim - PGD_LIDAR_BEAMS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * ROWS_WARPS + warp;
  if (row >= n_rows) return;
  const float* src = packed + (size_t)row * stride;
  float* dst = dense + (size_t)row * obs_dim;
  unsigned new_mask = 0u, old_mask = 0u;
  if (lane < 8) {
    new_mask = __float_as_uint(__ldcs(src + head + lane));
    old_mask = full ? 0xffffffffu : mask_state[(size_t)row * 8 + lane];
  }
  for (int i = lane; i < head; i += 32) dst[i] = __ldcs(src + i);
  int base = head + 8;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const unsigned nm = __shfl_sync(0xffffffffu, new_mask, c);
    const unsigned touched = nm | __shfl_sync(0xffffffffu, old_mask, c);
    if (touched) {  // uniform over the warp
      const int beam = c * 32 + lane;
      float v = 1.0f;
      if ((nm >> lane) & 1u) v = src[min(base + __popc(nm & ((1u << lane) - 1u)), stride - 1)];
      if (((touched >> lane) & 1u) && beam < PGD_LIDAR_BEAMS) dst[head + beam] = v;
    }
    base += __popc(nm);
  }
  if (lane < 8 && (full || new_mask != old_mask)) mask_state[(size_t)row * 8 + lane] = new_mask;
}

static int rows_args_ok(const char* who, const void* a, const void* b, int32_t n_rows, int32_t obs_dim) {
  if (!a || !b || n_rows < 0 || obs_dim < PGD_LIDAR_BEAMS)
    return fail(-1, std::string(who) + ": null pointer, negative row count or rows shorter than the lidar");
  if (((uintptr_t)a & 3) || ((uintptr_t)b & 3)) return fail(-1, std::string(who) + ": pointers must be 4-byte aligned");
  return 0;
}

typedef void (*RowsKernel)(const float*, float*, int, int);
static int rows_launch(RowsKernel kernel, const float* a, float* b, int32_t n_rows, int32_t obs_dim, void* stream) {
  // shared memory: the packing kernel's staged rows (9 KB for 274-float rows); longer rows (detector fans) opt into more
  const size_t smem = kernel == pgd_pack_rows_kernel ? (size_t)ROWS_WARPS * packed_words(obs_dim) * sizeof(float) : 0;
  if (smem > 227 * 1024) return fail(-1, "packed rows: observation rows too long for shared memory");
  if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kernel<<<(n_rows + ROWS_WARPS - 1) / ROWS_WARPS, ROWS_WARPS * 32, smem, (cudaStream_t)stream>>>(a, b, n_rows, obs_dim);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int pgd_pack_rows(const float* dense_dev, float* packed_dev, int32_t n_rows, int32_t obs_dim, void* stream) {
  if (int rc = rows_args_ok("pgd_pack_rows", dense_dev, packed_dev, n_rows, obs_dim)) return rc;
  if (n_rows == 0) return 0;
  return rows_launch(pgd_pack_rows_kernel, dense_dev, packed_dev, n_rows, obs_dim, stream);
}

extern "C" int pgd_expand_rows(const float* packed_dev, float* dense_dev, int32_t n_rows, int32_t obs_dim, void* stream) {
  if (int rc = rows_args_ok("pgd_expand_rows", packed_dev, dense_dev, n_rows, obs_dim)) return rc;
  if (n_rows == 0) return 0;
  return rows_launch(pgd_expand_rows_kernel, packed_dev, dense_dev, n_rows, obs_dim, stream);
}

extern "C" int pgd_expand_rows_delta(const float* packed_dev, float* dense_dev, uint32_t* mask_state_dev, int32_t n_rows,
                                     int32_t obs_dim, int32_t full, void* stream) {
  if (int rc = rows_args_ok("pgd_expand_rows_delta", packed_dev, dense_dev, n_rows, obs_dim)) return rc;
  if (!mask_state_dev || ((uintptr_t)mask_state_dev & 3)) return fail(-1, "pgd_expand_rows_delta: mask state missing or unaligned");
  if (n_rows == 0) return 0;
  pgd_expand_rows_delta_kernel<<<(n_rows + ROWS_WARPS - 1) / ROWS_WARPS, ROWS_WARPS * 32, 0, (cudaStream_t)stream>>>(
      packed_dev, dense_dev, mask_state_dev, n_rows, obs_dim, full);
  CU(cudaGetLastError());
  return 0;
}
