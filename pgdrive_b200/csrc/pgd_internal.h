// Internal (not installed) declarations shared by the translation units of libpgdrive_b200.so.
#ifndef PGD_INTERNAL_H
#define PGD_INTERNAL_H
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/pgdrive_b200.h"

struct DevTables {
  const PgdMap* maps;
  const PgdLane* lanes;
  const PgdRoad* roads;
  const PgdBox* boxes;
  const int32_t* cell_start;
  const int32_t* cell_entries;
  const PgdEpisode* episodes;
  const PgdSlot* slots;
  const int32_t* route_nodes;
  const int32_t* route_roads;
};

// SoA state, slot-major: index = slot * num_envs + env for the per-slot arrays, env for the per-env ones.
struct DevState {
  float4* pose;  // x, y, heading, speed
  float4* ctrl;  // steer, throttle, heading-PID last error, heading-PID summed error
  float4* pidl;  // lateral-PID last error, summed error, IDM target speed, yaw rate
  int4* nav;     // lane, ck0 | ck1 << 16, routing target lane, overtake timer
  int4* misc;    // rnd draws used, airborne sub-steps left, PGD_V_* flags, -
  int4* envi;    // episode, next trigger group, done, episode length
  float4* envf;  // previous steering, previous throttle, episode reward, episode energy
};

extern thread_local std::string g_pgd_err;

static inline int fail(int code, const std::string& msg) {
  g_pgd_err = msg;
  return code;
}

#define CU(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) return fail(-2, std::string(#call) + ": " + cudaGetErrorString(e_));   \
  } while (0)

struct PgdHandle {
  PgdConfig cfg;
  int device;
  DevTables T;
  void* table_mem[10];
  int64_t table_count[10];  // elements per table (maps, lanes, roads, boxes, cell_start, cell_entries, episodes, slots, route, route)
  int n_episodes;
  DevState S;
  void* state_mem[7];
  uint32_t call_index;  // API calls (reset / step) so far: one key component of the lidar-noise generator
  bool tables_loaded;
  int64_t launches;
  // reset scratch
  int32_t* d_ids;
  int32_t* d_eps;
  int scratch_cap;
  // the host-buffer step (pgd_hostpath.cu): transfer buffers, delta state, thread pool; its two streams
  void* hostpath;
  void* rowspath;     // pgd_rows_to_host: the same machinery for rows that are already in HBM
  int last_host_call;  // 1 pgd_step_host, 2 pgd_rows_to_host (pgd_host_transfer_bytes reports the last one)
  cudaStream_t own_stream, own_stream2;
  cudaEvent_t ev_act;
  cudaEvent_t ev_last;  // recorded after every enqueue on the caller's stream; the host-buffer step waits for it
  bool have_last;
  // timing
  int timing;
  cudaEvent_t ev0, ev1;
};

// pgd_step_kernel.cu: the fused environment step (mode 0) / the reset pass (mode 1) over environments [env_begin, env_end)
int pgd_launch_step(PgdHandle* h, int mode, int env_begin, int env_end, const float* actions, float* obs,
                    float* reward, uint8_t* done, PgdInfo* info, cudaStream_t st);
// pgd_hostpath.cu
void pgd_hostpath_destroy(PgdHandle* h);
#endif
