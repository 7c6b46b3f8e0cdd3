// One thread per environment: kernel wrapper around pgd_step_v2.cuh (experimental second layout of the step,
// selected per handle with PgdConfig.layout = 1; the cooperative kernel of pgd_step.cu stays the default).
//
// A warp advances 32 environments per instruction.  Per-vehicle state is slot-major ([slot][env]) so that the warp's
// loads and stores of a slot are 32 consecutive 16-byte vectors; the thread's vehicles and its observation row live
// in thread-local arrays (lane-interleaved local memory, served by L1).  The rows are written out through a
// shared-memory transpose: 32 environments x 32 floats per tile, so that every global store instruction covers
// 32 consecutive floats of one row; the 240 lidar beams are evaluated right there (lidar_beam) and never stored
// in local memory.
//
// STATUS: the step function is checked bit for bit against the CPU oracle in its host build
// (tests/test_step_v2.py); the sm_100a build has not run on a GPU yet (tests/test_gpu_step_v2.py is opt-in).
#include "pgd_internal.h"
#include "pgd_step_v2.cuh"

using namespace pgdv2;

#define V2_CTA_THREADS 64
#define V2_HEAD (PGD_OBS_DIM - PGD_LIDAR_BEAMS + 2)                     /* 36: state, [vehicle size], navi, neighbours */
#define V2_HEAD_DET (2 * PGD_MAX_DETECTOR_BEAMS + 6 + 2 + 10 + 16)      /* 514: with both detector fans */

template <int V, int HEAD_CAP>
__global__ void __launch_bounds__(V2_CTA_THREADS) pgd_step_v2_kernel(Tables T, State S, PgdConfig cfg, int mode,
                                                                     int env_begin, int env_end,
                                                                     const float* __restrict__ actions,
                                                                     float* __restrict__ obs,
                                                                     float* __restrict__ reward,
                                                                     uint8_t* __restrict__ done,
                                                                     PgdInfo* __restrict__ info) {
  __shared__ float tile[V2_CTA_THREADS / 32][32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int env = env_begin + blockIdx.x * V2_CTA_THREADS + threadIdx.x;
  const int warp_env0 = env - lane;
  const bool valid = env < env_end;
  float row[HEAD_CAP];  // the row up to the lidar beams
  LidarCtx<V> lc;
  lc.n = 0;
  const int obs_dim = (cfg.n_side > 0 ? cfg.n_side : 2) + 6 + cfg.n_lane_line + (cfg.random_agent_model ? 2 : 0) + 10 +
                      16 + PGD_LIDAR_BEAMS;
  bool wrote = false;
  if (valid) {
    const I4 envi = S.envi[env];
    wrote = !(mode == 1 && envi.z != V2_DONE_PENDING_RESET);
    float r = 0.0f;
    uint8_t d = 0;
    PgdInfo inf;
    step_env<V>(T, S, cfg, mode, env, cfg.num_envs, actions ? actions + 2 * (size_t)env : nullptr, row, lc, &r, &d,
                info ? &inf : nullptr);
    if (wrote) {
      if (mode == 0) {
        reward[env] = r;
        done[env] = d;
      }
      if (info) info[env] = inf;
    }
  }
  // transposed write-out of the observation rows of the warp's 32 environments
  const unsigned wmask = __ballot_sync(0xffffffffu, wrote);
  if (wmask == 0) return;
  const int head = obs_dim - PGD_LIDAR_BEAMS;
  for (int c = 0; c < obs_dim; c += 32) {
    const int nc = obs_dim - c < 32 ? obs_dim - c : 32;
    if (wrote) {
#pragma unroll 1
      for (int k = 0; k < nc; ++k) tile[warp][lane][k] = (c + k < head) ? row[c + k] : lidar_beam<V>(lc, c + k - head);
    }
    __syncwarp();
    if (lane < nc)
#pragma unroll 4
      for (int r = 0; r < 32; ++r)
        if ((wmask >> r) & 1u) obs[(size_t)(warp_env0 + r) * obs_dim + c + lane] = tile[warp][r][lane];
    __syncwarp();
  }
}

int pgd_launch_step_v2(PgdHandle* h, int mode, int env_begin, int env_end, const float* actions, float* obs,
                       float* reward, uint8_t* done, PgdInfo* info, cudaStream_t st) {
  if (h->cfg.decision_repeat > V2_MAX_SUBSTEPS)
    return fail(-3, "the one-thread-per-environment layout supports decision_repeat <= 16");
  Tables T;
  T.maps = h->T.maps; T.lanes = h->T.lanes; T.roads = h->T.roads; T.boxes = h->T.boxes;
  T.cell_start = h->T.cell_start; T.cell_entries = h->T.cell_entries; T.episodes = h->T.episodes;
  T.slots = h->T.slots; T.route_nodes = h->T.route_nodes; T.route_roads = h->T.route_roads;
  State S;
  S.pose = (F4*)h->S.pose; S.ctrl = (F4*)h->S.ctrl; S.pidl = (F4*)h->S.pidl; S.nav = (I4*)h->S.nav;
  S.misc = (I4*)h->S.misc; S.envi = (I4*)h->S.envi; S.envf = (F4*)h->S.envf;
  const int grid = (env_end - env_begin + V2_CTA_THREADS - 1) / V2_CTA_THREADS;
  if (h->timing && mode == 0) cudaEventRecord(h->ev0, st);
  const bool det = h->cfg.n_side > 0 || h->cfg.n_lane_line > 0;  // detectors need the long row
#define V2_LAUNCH(VV, CAP)                                                                                   \
  pgd_step_v2_kernel<VV, CAP><<<grid, V2_CTA_THREADS, 0, st>>>(T, S, h->cfg, mode, env_begin, env_end, actions, \
                                                               obs, reward, done, info)
  if (h->cfg.num_slots == 16 && !det) V2_LAUNCH(16, V2_HEAD);
  else if (h->cfg.num_slots == 32 && !det) V2_LAUNCH(32, V2_HEAD);
  else if (h->cfg.num_slots == 16) V2_LAUNCH(16, V2_HEAD_DET);
  else V2_LAUNCH(32, V2_HEAD_DET);
#undef V2_LAUNCH
  if (h->timing && mode == 0) cudaEventRecord(h->ev1, st);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}
