// Role per warp / environment per lane: kernel wrapper around pgd_step.cuh .
//
// A CTA = R warps x 32 lanes advances 32 environments (see the header of pgd_step.cuh for the phases).  The 32
// observation rows are assembled in shared memory in their HBM layout and leave with ONE bulk (TMA) copy per CTA
// (cp.async.bulk.global.shared::cta, 35 KB at 274 floats per row); the destination may be a peer-mapped buffer on
// another GPU.  CTAs that are not full (partial reset, tail) fall back to coalesced per-row stores.
#include <stdlib.h>

#include "pgd_internal.h"

#ifdef PGS_PHASE_CLOCKS  // diagnostic build: cycles between the CTA barriers, summed over CTAs (thread 0 of each); inside
// phase X the cycles of traffic warp 1's lane 0 per part of an item (slots 16..)
__device__ unsigned long long g_pgs_clk[32];
__host__ __device__ __forceinline__ long long pgs_item_now() {
#ifdef __CUDA_ARCH__
  return clock64();
#else
  return 0;
#endif
}
__host__ __device__ __forceinline__ void pgs_item_add(int i, long long n) {
#ifdef __CUDA_ARCH__
  if (threadIdx.x == 32) atomicAdd(&g_pgs_clk[16 + i], (unsigned long long)n);
#endif
}
#define PGS_ITEM_CLK_BEGIN long long iclk_ = pgs_item_now();
#define PGS_ITEM_CLK(i)                      \
  do {                                      \
    const long long now_ = pgs_item_now();  \
    pgs_item_add(i, now_ - iclk_);          \
    iclk_ = now_;                           \
  } while (0)
#define PGS_ITEM_COUNT(i, n) pgs_item_add(i, n)
__host__ __device__ __forceinline__ void pgs_ego_add(int i, long long n) {
#ifdef __CUDA_ARCH__
  if (threadIdx.x == 0) atomicAdd(&g_pgs_clk[i], (unsigned long long)n);
#endif
}
#define PGS_EGO_CLK_BEGIN long long eclk_ = pgs_item_now();
#define PGS_EGO_CLK(i)                      \
  do {                                     \
    const long long now_ = pgs_item_now(); \
    pgs_ego_add(i, now_ - eclk_);          \
    eclk_ = now_;                          \
  } while (0)
#endif
#include "pgd_step.cuh"

using namespace pgdstep;

#ifndef PGS_ROLES
#define PGS_ROLES 4
#endif
#ifndef PGS_MIN_CTAS
#define PGS_MIN_CTAS 4
#endif
#ifndef PGS_OBS_EVICT_FIRST
#define PGS_OBS_EVICT_FIRST 1
#endif

#ifdef PGS_PHASE_CLOCKS
#define PGS_CLK(i)                                                              \
  do {                                                                         \
    if (threadIdx.x == 0) {                                                    \
      const long long now_ = clock64();                                        \
      atomicAdd(&g_pgs_clk[i], (unsigned long long)(now_ - clk_));              \
      clk_ = now_;                                                             \
    }                                                                          \
  } while (0)
extern "C" int pgd_debug_phase_clocks(unsigned long long* out, int reset) {
  cudaMemcpyFromSymbol(out, g_pgs_clk, sizeof(g_pgs_clk));
  if (reset) {
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(g_pgs_clk, z, sizeof(z));
  }
  return 0;
}
#else
#define PGS_CLK(i)
#endif

template <int V, int R>
__global__ void __launch_bounds__(R * 32, PGS_MIN_CTAS) pgd_step_kernel(Tables T, State S, PgdConfig cfg, int mode,
                                                                         uint32_t call_index,
                                                                         int env_begin, int env_end, int envs_per_cta,
                                                                         const float* __restrict__ actions,
                                                                         float* __restrict__ obs,
                                                                         float* __restrict__ reward,
                                                                         uint8_t* __restrict__ done,
                                                                         PgdInfo* __restrict__ info) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem<V, R>& sm = *reinterpret_cast<Smem<V, R>*>(smem_raw);
  float* rows = reinterpret_cast<float*>(smem_raw + smem_obs_offset<V, R>());
  const int obs_dim = obs_dim_of(cfg);
  unsigned char* tv = smem_raw + smem_tv_offset<V, R>(obs_dim);
  TrajPtr traj = reinterpret_cast<TrajPtr>(tv);
  VisPtr vis = reinterpret_cast<VisPtr>(tv);
  const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
  // lanes [0, envs_per_cta) of every warp carry an environment (pgd_launch_step: fewer than 32 when that makes the grid
  // a whole number of waves)
  const int env0 = env_begin + blockIdx.x * envs_per_cta;
  const int cta_end = env0 + envs_per_cta < env_end ? env0 + envs_per_cta : env_end;
#ifdef PGS_PHASE_CLOCKS
  long long clk_ = clock64();
#endif
  Thr<V, R> th;
  thread_init(th, T, S, cfg, mode, lane, role, env0 + lane, cta_end);
  phase_0(sm, th);
  if (!__syncthreads_or(th.valid)) return;  // reset pass: no environment of this CTA is marked
  PGS_CLK(0);
  phase_a(sm, th, S, cfg, actions, rows);
  PGS_CLK(1);
  __syncthreads();
  PGS_CLK(2);
  phase_b(sm, th, T, rows);
  __syncthreads();
  PGS_CLK(3);
  phase_x(sm, th, T, S, cfg, obs_dim, rows, traj, vis);
  PGS_CLK(4);
  PGS_CLK(5);
  PGS_CLK(6);
  __syncthreads();  // everything has moved, the ego's look-ups are done
  PGS_CLK(7);
  phase_f(sm, th, T, S, cfg, mode, obs_dim, rows, vis, reward, done, info);
  PGS_CLK(8);
  PGS_CLK(9);
  phase_l_fill(sm, role, lane, obs_dim, rows);
  __syncwarp();
  phase_l(sm, T, S, role, lane, env0, obs_dim, rows, vis);
  PGS_CLK(10);
  __syncwarp();
  phase_n(sm, cfg, call_index, role, lane, env0, obs_dim, rows);
  PGS_CLK(11);
  // ---- write-out -----------------------------------------------------------------------------------------------------
  const int n_rows = cta_end - env0;
  const int all = __syncthreads_and(lane >= n_rows || sm.wrote[lane]);
  float* dst = obs + (size_t)env0 * obs_dim;
  const uint32_t bytes = (uint32_t)(n_rows * obs_dim * sizeof(float));
  if (all && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && (bytes & 15) == 0) {
    if (threadIdx.x == 0) {
      const uint32_t src = (uint32_t)__cvta_generic_to_shared(rows);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#if PGS_OBS_EVICT_FIRST
      // the rows are written once and read by somebody else (the policy, the gather): do not let 72 MB of them per
      // step push the tables and the 26 MB of state that the next step re-reads out of the 126 MB L2
      uint64_t pol;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src),
                   "r"(bytes), "l"(pol)
                   : "memory");
#else
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
                   : "memory");
#endif
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    PGS_CLK(12);
  } else {
    for (int e = role; e < PGS_LANES; e += R) {
      if (!sm.wrote[e]) continue;
      const float* src = rows + (size_t)e * obs_dim;
      float* d = dst + (size_t)e * obs_dim;
      for (int c = lane; c < obs_dim; c += 32) __stcs(d + c, src[c]);
    }
  }
}

template <int V, int R>
static int launch_one(PgdHandle* h, const Tables& T, const State& S, int mode, int env_begin, int env_end,
                      const float* actions, float* obs, float* reward, uint8_t* done, PgdInfo* info, cudaStream_t st) {
  static int configured = 0;  // per instantiation: largest dynamic shared-memory size opted into so far
  const int smem = (int)smem_bytes<V, R>(obs_dim_of(h->cfg), h->cfg.decision_repeat);
  if (smem > configured) {
    CU(cudaFuncSetAttribute(pgd_step_kernel<V, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  // Environments per CTA: 32 (all lanes).  Fewer (an even number, so that the CTA's rows start on a 16-byte boundary)
  // would make the grid a whole number of waves -- 2 048 CTAs on 592 slots are 3.46 -- but the CTAs do not run in
  // lock-step and the step time follows the NUMBER of CTAs (profiles/r02y_sweep.log: 28 per CTA is 12 % slower).
  const int n = env_end - env_begin;
  int epc = PGS_LANES;
  const char* fixed = getenv("PGDRIVE_B200_ENVS_PER_CTA");  // experiments (tools/gpu_epc_call.sh)
  if (fixed) epc = atoi(fixed);
  if (epc < 2 || epc > PGS_LANES) epc = PGS_LANES;
  const int grid = (n + epc - 1) / epc;
  pgd_step_kernel<V, R><<<grid, R * 32, smem, st>>>(T, S, h->cfg, mode, h->call_index, env_begin, env_end, epc, actions, obs,
                                                       reward, done, info);
  return 0;
}

int pgd_launch_step(PgdHandle* h, int mode, int env_begin, int env_end, const float* actions, float* obs,
                       float* reward, uint8_t* done, PgdInfo* info, cudaStream_t st) {
  Tables T;
  T.maps = h->T.maps; T.lanes = h->T.lanes; T.roads = h->T.roads; T.boxes = h->T.boxes;
  T.cell_start = h->T.cell_start; T.cell_entries = h->T.cell_entries; T.episodes = h->T.episodes;
  T.slots = h->T.slots; T.route_nodes = h->T.route_nodes; T.route_roads = h->T.route_roads;
  State S;
  S.pose = (F4*)h->S.pose; S.ctrl = (F4*)h->S.ctrl; S.pidl = (F4*)h->S.pidl; S.nav = (I4*)h->S.nav;
  S.misc = (I4*)h->S.misc; S.envi = (I4*)h->S.envi; S.envf = (F4*)h->S.envf;
  if (h->timing && mode == 0) cudaEventRecord(h->ev0, st);
  const int V = h->cfg.num_slots;
  int rc;
#define PGS_LAUNCH(VV) rc = launch_one<VV, PGS_ROLES>(h, T, S, mode, env_begin, env_end, actions, obs, reward, done, info, st)
  if (V == 16) PGS_LAUNCH(16);
  else if (V == 24) PGS_LAUNCH(24);
  else PGS_LAUNCH(32);
#undef PGS_LAUNCH
  if (rc) return rc;
  if (h->timing && mode == 0) cudaEventRecord(h->ev1, st);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}
