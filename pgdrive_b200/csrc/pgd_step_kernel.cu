// Role per warp / environment per lane: kernel wrapper around pgd_step.cuh .
//
// A CTA = R warps x 32 lanes advances 32 environments (see the header of pgd_step.cuh for the phases).  The 32
// observation rows are assembled in shared memory in their HBM layout and leave with ONE bulk (TMA) copy per CTA
// (cp.async.bulk.global.shared::cta, 35 KB at 274 floats per row); the destination may be a peer-mapped buffer on
// another GPU.  CTAs that are not full (partial reset, tail) fall back to coalesced per-row stores.
#include "pgd_internal.h"
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
#define PGS_STAGE_MAX_LANES 210  // most lanes a 3-block map has

#ifdef PGS_PHASE_CLOCKS  // diagnostic build: cycles between the CTA barriers, summed over CTAs (thread 0 of each)
__device__ unsigned long long g_pgs_clk[16];
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
    unsigned long long z[16] = {0};
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
                                                                         int env_begin, int env_end,
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
  const int env0 = env_begin + blockIdx.x * PGS_LANES;
#ifdef PGS_PHASE_CLOCKS
  long long clk_ = clock64();
#endif
  Thr<V, R> th;
  thread_init(th, T, S, cfg, mode, lane, role, env0 + lane, env_end);
  phase_0(sm, th);
  if (!__syncthreads_or(th.valid)) return;  // reset pass: no environment of this CTA is marked
#ifdef PGS_STAGE_LANES
  // EXPERIMENT (north_star: "lane / segment geometry TMA-staged into shared memory per block"): when the 32
  // environments of the CTA play the same map, its lane table (<= 210 x 64 B) is brought into shared memory by one
  // bulk (TMA) copy and every lane access of the step reads it there.  Measured against the same build without the
  // staging in profiles/ (tools/kernel_variants.py: "plain" vs "stage", environments assigned to seeds in blocks of 32).
  {
    __shared__ __align__(8) unsigned long long stage_bar;
    PgdLane* stage = reinterpret_cast<PgdLane*>(smem_raw + smem_bytes<V, R>(obs_dim, cfg.decision_repeat));
    const int map0 = sm.ctx_map[0];
    const bool same = __syncthreads_and(th.valid && sm.ctx_map[lane] == map0 && th.mp.n_lanes <= PGS_STAGE_MAX_LANES);
    if (same) {
      const uint32_t bytes = (uint32_t)th.mp.n_lanes * (uint32_t)sizeof(PgdLane);
      const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&stage_bar);
      if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(stage)),
                     "l"(th.lanes), "r"(bytes), "r"(bar)
                     : "memory");
      }
      uint32_t done = 0;
      while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(bar)
                     : "memory");
      }
      th.lanes = stage;
      th.staged_lanes = stage;
    }
  }
#endif
  PGS_CLK(0);
  phase_a(sm, th, S, cfg, actions);
  PGS_CLK(1);
  __syncthreads();
  PGS_CLK(2);
  phase_b(sm, th, rows);
  __syncthreads();
  PGS_CLK(3);
  phase_c(sm, th, T, S, cfg, rows, traj);
  phase_c_traffic(sm, th, T, S, rows);
  PGS_CLK(4);
  __syncthreads();
  PGS_CLK(5);
  {  // pre-fill the rows with 1.0 = "no hit" (the IDM look-up data that shared this storage is dead now)
    float4* o4 = reinterpret_cast<float4*>(rows);
    const int n4 = PGS_LANES * obs_dim / 4;  // 32 rows: a multiple of 4 floats for every row length
    for (int i = threadIdx.x; i < n4; i += R * 32) o4[i] = make_float4(1.f, 1.f, 1.f, 1.f);
  }
  PGS_CLK(6);
  phase_d(sm, th, T, S, cfg, traj);
  phase_d_traffic(sm, th, T, S, cfg, traj);
  PGS_CLK(7);
  __syncthreads();
  PGS_CLK(8);
  phase_f(sm, th, T, S, cfg, mode, obs_dim, rows, vis, reward, done, info);
  PGS_CLK(9);
  __syncthreads();
  PGS_CLK(10);
  phase_l(sm, T, S, role, lane, env0, obs_dim, rows, vis);
  __syncwarp();
  phase_n(sm, cfg, call_index, role, lane, env0, obs_dim, rows);
  PGS_CLK(11);
  // ---- write-out -----------------------------------------------------------------------------------------------------
  const int all = __syncthreads_and(sm.wrote[lane]);
  float* dst = obs + (size_t)env0 * obs_dim;
  if (all && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
    if (threadIdx.x == 0) {
      const uint32_t bytes = (uint32_t)(PGS_LANES * obs_dim * sizeof(float));
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
#ifdef PGS_STAGE_LANES
  const int smem = (int)smem_bytes<V, R>(obs_dim_of(h->cfg), h->cfg.decision_repeat) + PGS_STAGE_MAX_LANES * (int)sizeof(PgdLane);
#else
  const int smem = (int)smem_bytes<V, R>(obs_dim_of(h->cfg), h->cfg.decision_repeat);
#endif
  if (smem > configured) {
    CU(cudaFuncSetAttribute(pgd_step_kernel<V, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  const int grid = (env_end - env_begin + PGS_LANES - 1) / PGS_LANES;
  pgd_step_kernel<V, R><<<grid, R * 32, smem, st>>>(T, S, h->cfg, mode, h->call_index, env_begin, env_end, actions, obs, reward,
                                                       done, info);
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
