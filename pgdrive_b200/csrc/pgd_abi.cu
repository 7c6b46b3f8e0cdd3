// C-ABI of libpgdrive_b200.so (include/pgdrive_b200.h): handle life cycle, table upload, reset / step entry points,
// state exchange, peer memory.  The step kernel itself lives in pgd_step.cuh (phases, host + device) and
// pgd_step_kernel.cu (kernel wrapper); the on-device reset path in pgd_mapgen.cu.
//
// State lives in HBM as structure-of-arrays, slot-major ([slot][env], 16-byte vectors): a warp of the step kernel
// holds 32 environments, so its loads / stores of a slot are 32 consecutive vectors.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <string>

#include "../../include/pgdrive_b200.h"
#include "pgd_internal.h"

#define DONE_PENDING_RESET 2
#define PGD_INTERNAL_V_POSE_SET 8  // PGS_V_POSE_SET of pgd_step.cuh

// marks environments for a forced reset on the given episode templates
__global__ void pgd_mark_reset_kernel(DevState S, const int32_t* env_ids, const int32_t* episode_ids, int n,
                                      int num_envs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int e = env_ids ? env_ids[i] : i;
  if (e < 0 || e >= num_envs) return;
  int4 v = S.envi[e];
  v.x = episode_ids[i];
  v.z = DONE_PENDING_RESET;
  S.envi[e] = v;
}

thread_local std::string g_pgd_err;

extern "C" const char* pgd_last_error(void) { return g_pgd_err.c_str(); }

extern "C" int pgd_create(const PgdConfig* cfg, int device, PgdHandle** out) {
  if (!cfg || !out) return fail(-1, "pgd_create: null argument");
  if (cfg->num_envs <= 0) return fail(-1, "pgd_create: num_envs must be positive");
  if (cfg->num_slots != 16 && cfg->num_slots != 24 && cfg->num_slots != 32)
    return fail(-1, "pgd_create: num_slots must be 16, 24 or 32");
  if (cfg->decision_repeat < 1 || cfg->decision_repeat > 8)
    return fail(-3, "pgd_create: decision_repeat must be in [1, 8]");
  if (cfg->n_side < 0 || cfg->n_side > PGD_MAX_DETECTOR_BEAMS || cfg->n_lane_line < 0 ||
      cfg->n_lane_line > PGD_MAX_DETECTOR_BEAMS)
    return fail(-1, "pgd_create: detector beam counts must be in [0, 240]");
  CU(cudaSetDevice(device));
  PgdHandle* h = new PgdHandle();
  memset(h, 0, sizeof(*h));
  h->cfg = *cfg;
  h->device = device;
  const size_t nv = (size_t)cfg->num_envs * cfg->num_slots, n = (size_t)cfg->num_envs;
  const size_t sizes[7] = {nv * 16, nv * 16, nv * 16, nv * 16, nv * 16, n * 16, n * 16};
  for (int i = 0; i < 7; ++i) {
    CU(cudaMalloc(&h->state_mem[i], sizes[i]));
    CU(cudaMemset(h->state_mem[i], 0, sizes[i]));
  }
  h->S.pose = (float4*)h->state_mem[0];
  h->S.ctrl = (float4*)h->state_mem[1];
  h->S.pidl = (float4*)h->state_mem[2];
  h->S.nav = (int4*)h->state_mem[3];
  h->S.misc = (int4*)h->state_mem[4];
  h->S.envi = (int4*)h->state_mem[5];
  h->S.envf = (float4*)h->state_mem[6];
  CU(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&h->own_stream2, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&h->ev_act, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&h->ev_last, cudaEventDisableTiming));
  CU(cudaEventCreate(&h->ev0));
  CU(cudaEventCreate(&h->ev1));
  *out = h;
  return 0;
}

extern "C" int pgd_destroy(PgdHandle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < 7; ++i) cudaFree(h->state_mem[i]);
  for (int i = 0; i < 10; ++i) cudaFree(h->table_mem[i]);
  cudaFree(h->d_ids); cudaFree(h->d_eps);
  pgd_hostpath_destroy(h);
  cudaStreamDestroy(h->own_stream);
  cudaStreamDestroy(h->own_stream2);
  cudaEventDestroy(h->ev_act);
  cudaEventDestroy(h->ev_last);
  cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1);
  delete h;
  return 0;
}

extern "C" int pgd_load_tables(PgdHandle* h, const PgdTables* t) {
  if (!h || !t) return fail(-1, "pgd_load_tables: null argument");
  CU(cudaSetDevice(h->device));
  for (int i = 0; i < t->n_episodes; ++i) {
    if (t->episodes[i].n_slots > h->cfg.num_slots)
      return fail(-3, "pgd_load_tables: an episode needs more vehicle slots than num_slots");
    if (t->episodes[i].n_groups > PGD_MAX_GROUPS) return fail(-3, "pgd_load_tables: too many trigger groups");
  }
  const void* src[10] = {t->maps, t->lanes, t->roads, t->boxes, t->cell_start, t->cell_entries,
                         t->episodes, t->slots, t->route_nodes, t->route_roads};
  const size_t bytes[10] = {(size_t)t->n_maps * sizeof(PgdMap), (size_t)t->n_lanes * sizeof(PgdLane),
                            (size_t)t->n_roads * sizeof(PgdRoad), (size_t)t->n_boxes * sizeof(PgdBox),
                            (size_t)t->n_cell_start * 4, (size_t)t->n_cell_entries * 4,
                            (size_t)t->n_episodes * sizeof(PgdEpisode), (size_t)t->n_slots * sizeof(PgdSlot),
                            (size_t)t->n_route * 4, (size_t)t->n_route * 4};
  CU(cudaDeviceSynchronize());
  for (int i = 0; i < 10; ++i) {
    cudaFree(h->table_mem[i]);
    h->table_mem[i] = nullptr;
    CU(cudaMalloc(&h->table_mem[i], bytes[i] ? bytes[i] : 16));
    if (bytes[i]) CU(cudaMemcpy(h->table_mem[i], src[i], bytes[i], cudaMemcpyHostToDevice));
  }
  h->T.maps = (const PgdMap*)h->table_mem[0];
  h->T.lanes = (const PgdLane*)h->table_mem[1];
  h->T.roads = (const PgdRoad*)h->table_mem[2];
  h->T.boxes = (const PgdBox*)h->table_mem[3];
  h->T.cell_start = (const int32_t*)h->table_mem[4];
  h->T.cell_entries = (const int32_t*)h->table_mem[5];
  h->T.episodes = (const PgdEpisode*)h->table_mem[6];
  h->T.slots = (const PgdSlot*)h->table_mem[7];
  h->T.route_nodes = (const int32_t*)h->table_mem[8];
  h->T.route_roads = (const int32_t*)h->table_mem[9];
  const int64_t counts[10] = {t->n_maps, t->n_lanes, t->n_roads, t->n_boxes, t->n_cell_start, t->n_cell_entries,
                              t->n_episodes, t->n_slots, t->n_route, t->n_route};
  for (int i = 0; i < 10; ++i) h->table_count[i] = counts[i];
  h->n_episodes = t->n_episodes;
  h->tables_loaded = true;
  return 0;
}

static int launch_step(PgdHandle* h, int mode, int env_begin, int env_end, const float* actions, float* obs,
                       float* reward, uint8_t* done, PgdInfo* info, cudaStream_t st) {
  return pgd_launch_step(h, mode, env_begin, env_end, actions, obs, reward, done, info, st);
}

// pgd_step_host runs on the handle's own (non-blocking) streams; pgd_reset / pgd_step run on the caller's.  Both touch
// the same simulator state, so the host-buffer step waits for whatever was enqueued last on the caller's stream.
static int note_caller_stream(PgdHandle* h, cudaStream_t st) {
  CU(cudaEventRecord(h->ev_last, st));
  h->have_last = true;
  return 0;
}

extern "C" int pgd_reset(PgdHandle* h, const int32_t* env_ids, const int32_t* episode_ids, int32_t n, float* obs_dev,
                         PgdInfo* info_dev, void* stream) {
  if (!h || !episode_ids || !obs_dev) return fail(-1, "pgd_reset: null argument");
  if (!h->tables_loaded) return fail(-3, "pgd_reset: no tables loaded");
  if (n <= 0 || n > h->cfg.num_envs) return fail(-1, "pgd_reset: n out of range");
  for (int i = 0; i < n; ++i) {
    if (episode_ids[i] < 0 || episode_ids[i] >= h->n_episodes) return fail(-1, "pgd_reset: episode id out of range");
    if (env_ids && (env_ids[i] < 0 || env_ids[i] >= h->cfg.num_envs)) return fail(-1, "pgd_reset: env id out of range");
  }
  CU(cudaSetDevice(h->device));
  h->call_index++;
  cudaStream_t st = (cudaStream_t)stream;
  if (n > h->scratch_cap) {
    cudaFree(h->d_ids); cudaFree(h->d_eps);
    CU(cudaMalloc(&h->d_ids, (size_t)n * 4));
    CU(cudaMalloc(&h->d_eps, (size_t)n * 4));
    h->scratch_cap = n;
  }
  // pageable-host copies: cudaMemcpyAsync returns after staging, so the caller's arrays may be reused
  if (env_ids) CU(cudaMemcpyAsync(h->d_ids, env_ids, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(h->d_eps, episode_ids, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  pgd_mark_reset_kernel<<<(n + 255) / 256, 256, 0, st>>>(h->S, env_ids ? h->d_ids : nullptr, h->d_eps, n,
                                                         h->cfg.num_envs);
  h->launches++;
  CU(cudaGetLastError());
  const int rc = launch_step(h, 1, 0, h->cfg.num_envs, nullptr, obs_dev, nullptr, nullptr, info_dev, st);
  if (rc) return rc;
  return note_caller_stream(h, st);
}

extern "C" int pgd_step(PgdHandle* h, const float* actions_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
                        PgdInfo* info_dev, void* stream) {
  if (!h || !actions_dev || !obs_dev || !reward_dev || !done_dev) return fail(-1, "pgd_step: null argument");
  if (!h->tables_loaded) return fail(-3, "pgd_step: no tables loaded");
  CU(cudaSetDevice(h->device));
  h->call_index++;
  const int rc = launch_step(h, 0, 0, h->cfg.num_envs, actions_dev, obs_dev, reward_dev, done_dev, info_dev,
                             (cudaStream_t)stream);
  if (rc) return rc;
  return note_caller_stream(h, (cudaStream_t)stream);
}

// the V per-slot records of one environment: strided by num_envs (state is [slot][env])
static cudaError_t copy_slots(PgdHandle* h, void* dev_base, int env, void* host, bool to_host) {
  const int V = h->cfg.num_slots;
  char* d = (char*)dev_base + (size_t)env * 16;
  const size_t pitch = (size_t)h->cfg.num_envs * 16;
  return to_host ? cudaMemcpy2D(host, 16, d, pitch, 16, V, cudaMemcpyDeviceToHost)
                 : cudaMemcpy2D(d, pitch, host, 16, 16, V, cudaMemcpyHostToDevice);
}

extern "C" int pgd_get_state(PgdHandle* h, int32_t env, PgdEnvState* out) {
  if (!h || !out) return fail(-1, "pgd_get_state: null argument");
  if (env < 0 || env >= h->cfg.num_envs) return fail(-1, "pgd_get_state: env out of range");
  CU(cudaSetDevice(h->device));
  CU(cudaDeviceSynchronize());
  const int V = h->cfg.num_slots;
  float4 pose[PGD_MAX_SLOTS], ctrl[PGD_MAX_SLOTS], pidl[PGD_MAX_SLOTS], envf;
  int4 nav[PGD_MAX_SLOTS], misc[PGD_MAX_SLOTS], envi;
  CU(copy_slots(h, h->S.pose, env, pose, true));
  CU(copy_slots(h, h->S.ctrl, env, ctrl, true));
  CU(copy_slots(h, h->S.pidl, env, pidl, true));
  CU(copy_slots(h, h->S.nav, env, nav, true));
  CU(copy_slots(h, h->S.misc, env, misc, true));
  CU(cudaMemcpy(&envi, h->S.envi + env, 16, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(&envf, h->S.envf + env, 16, cudaMemcpyDeviceToHost));
  memset(out, 0, sizeof(*out));
  out->episode = envi.x; out->next_group = envi.y; out->done = envi.z; out->ep_len = envi.w;
  out->prev_steer = envf.x; out->prev_throttle = envf.y; out->ep_reward = envf.z; out->energy = envf.w;
  for (int i = 0; i < V; ++i) {
    PgdVehState* s = &out->veh[i];
    s->x = pose[i].x; s->y = pose[i].y; s->heading = pose[i].z; s->speed = pose[i].w;
    s->steer = ctrl[i].x; s->throttle = ctrl[i].y; s->pid_hp = ctrl[i].z; s->pid_hi = ctrl[i].w;
    s->pid_lp = pidl[i].x; s->pid_li = pidl[i].y; s->target_speed = pidl[i].z; s->yaw_rate = pidl[i].w;
    s->lane = nav[i].x; s->ck0 = nav[i].y & 0xffff; s->ck1 = nav[i].y >> 16; s->rt_lane = nav[i].z;
    s->timer = nav[i].w; s->rnd_n = misc[i].x; s->airborne = misc[i].y; s->flags = misc[i].z & ~PGD_INTERNAL_V_POSE_SET;
  }
  return 0;
}

extern "C" int pgd_set_state(PgdHandle* h, int32_t env, const PgdEnvState* in) {
  if (!h || !in) return fail(-1, "pgd_set_state: null argument");
  if (env < 0 || env >= h->cfg.num_envs) return fail(-1, "pgd_set_state: env out of range");
  if (in->episode < 0 || in->episode >= h->n_episodes) return fail(-1, "pgd_set_state: episode out of range");
  CU(cudaSetDevice(h->device));
  CU(cudaDeviceSynchronize());
  const int V = h->cfg.num_slots;
  float4 pose[PGD_MAX_SLOTS], ctrl[PGD_MAX_SLOTS], pidl[PGD_MAX_SLOTS], envf;
  int4 nav[PGD_MAX_SLOTS], misc[PGD_MAX_SLOTS], envi;
  envi = make_int4(in->episode, in->next_group, in->done, in->ep_len);
  envf = make_float4(in->prev_steer, in->prev_throttle, in->ep_reward, in->energy);
  for (int i = 0; i < V; ++i) {
    const PgdVehState* s = &in->veh[i];
    pose[i] = make_float4(s->x, s->y, s->heading, s->speed);
    ctrl[i] = make_float4(s->steer, s->throttle, s->pid_hp, s->pid_hi);
    pidl[i] = make_float4(s->pid_lp, s->pid_li, s->target_speed, s->yaw_rate);
    nav[i] = make_int4(s->lane, s->ck0 | (s->ck1 << 16), s->rt_lane, s->timer);
    // a parked vehicle normally takes its pose from the episode template; this one was placed by the caller
    const bool parked = (s->flags & PGD_V_ALIVE) && !(s->flags & PGD_V_ACTIVE);
    misc[i] = make_int4(s->rnd_n, s->airborne, s->flags | (parked ? PGD_INTERNAL_V_POSE_SET : 0), 0);
  }
  CU(copy_slots(h, h->S.pose, env, pose, false));
  CU(copy_slots(h, h->S.ctrl, env, ctrl, false));
  CU(copy_slots(h, h->S.pidl, env, pidl, false));
  CU(copy_slots(h, h->S.nav, env, nav, false));
  CU(copy_slots(h, h->S.misc, env, misc, false));
  CU(cudaMemcpy(h->S.envi + env, &envi, 16, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->S.envf + env, &envf, 16, cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int pgd_peer_alloc(PgdHandle* h, uint64_t bytes, void** dev_ptr, unsigned char handle_out[64]) {
  if (!h || !dev_ptr || !handle_out) return fail(-1, "pgd_peer_alloc: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CU(cudaSetDevice(h->device));
  CU(cudaMalloc(dev_ptr, bytes));
  CU(cudaMemset(*dev_ptr, 0, bytes));
  cudaIpcMemHandle_t ipc;
  CU(cudaIpcGetMemHandle(&ipc, *dev_ptr));
  memcpy(handle_out, &ipc, 64);
  return 0;
}

extern "C" int pgd_peer_open(PgdHandle* h, const unsigned char handle[64], void** dev_ptr) {
  if (!h || !dev_ptr || !handle) return fail(-1, "pgd_peer_open: null argument");
  CU(cudaSetDevice(h->device));
  cudaIpcMemHandle_t ipc;
  memcpy(&ipc, handle, 64);
  CU(cudaIpcOpenMemHandle(dev_ptr, ipc, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int pgd_peer_release(PgdHandle* h, void* dev_ptr, int32_t is_owner) {
  if (!h || !dev_ptr) return fail(-1, "pgd_peer_release: null argument");
  CU(cudaSetDevice(h->device));
  CU(cudaDeviceSynchronize());
  if (is_owner) CU(cudaFree(dev_ptr));
  else CU(cudaIpcCloseMemHandle(dev_ptr));
  return 0;
}

// Sum of the 32-bit words of a device buffer into a 64-bit accumulator (*out_dev += sum): what a consumer of the gathered
// batch does at the least -- touch every byte once, at HBM speed (16-byte loads, grid sized to the SM count).
__global__ void __launch_bounds__(256) pgd_words_sum_kernel(const uint4* __restrict__ p, size_t n16,
                                                            unsigned long long* __restrict__ out) {
  unsigned long long acc = 0;
  const size_t step = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * step < n16; i += 4 * step) {  // four independent 16-byte loads in flight per thread
    const uint4 a = __ldcs(p + i), b = __ldcs(p + i + step), c = __ldcs(p + i + 2 * step), d = __ldcs(p + i + 3 * step);
    acc += ((unsigned long long)a.x + a.y + a.z + a.w) + ((unsigned long long)b.x + b.y + b.z + b.w) +
           ((unsigned long long)c.x + c.y + c.z + c.w) + ((unsigned long long)d.x + d.y + d.z + d.w);
  }
  for (; i < n16; i += step) {
    const uint4 v = __ldcs(p + i);
    acc += (unsigned long long)v.x + v.y + v.z + v.w;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

extern "C" int pgd_words_checksum(PgdHandle* h, const void* dev_ptr, uint64_t bytes, uint64_t* out_dev, void* stream) {
  if (!h || !dev_ptr || !out_dev) return fail(-1, "pgd_words_checksum: null argument");
  if ((bytes & 15) || ((uintptr_t)dev_ptr & 15)) return fail(-1, "pgd_words_checksum: pointer and size must be multiples of 16");
  CU(cudaSetDevice(h->device));
  pgd_words_sum_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>((const uint4*)dev_ptr, (size_t)(bytes / 16),
                                                                  (unsigned long long*)out_dev);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int64_t pgd_state_bytes_per_env(PgdHandle* h) { return h ? (int64_t)h->cfg.num_slots * 80 + 32 : 0; }
extern "C" int64_t pgd_launch_count(PgdHandle* h) { return h ? h->launches : 0; }
extern "C" int pgd_set_timing(PgdHandle* h, int32_t on) {
  if (!h) return fail(-1, "pgd_set_timing: null handle");
  h->timing = on;
  return 0;
}
extern "C" float pgd_last_kernel_ms(PgdHandle* h) {
  if (!h || !h->timing) return -1.0f;
  float ms = -1.0f;
  if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0f;
  if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0f;
  return ms;
}
