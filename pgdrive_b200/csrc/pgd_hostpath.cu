// pgd_step_host: the step through HOST buffers (include/pgdrive_b200.h) -- what `env.step(action)` of the reference is to
// its caller (envs/base_env.py:184-224): actions come from host memory, observations / rewards / dones / infos end there.
//
// The PCIe link is the bound of this path: a dense observation row is 1 096 bytes and the kernel makes 65 536 of them in
// 0.15 ms.  So rows cross the link PACKED, in the same spirit as the gather's wire format (pgd_rows.cu): per row the head
// and a 240-bit hit mask (168 bytes), and the values of the beams that are not 1.0 compacted per chunk -- 180 bytes
// per row instead of 1 096 in the steady state of the random policy -- and they are expanded on the host by a small
// pool of threads straight into the caller's array (page-locked or not).  The expansion is a DELTA: the caller's array
// still holds the rows of the previous step, so only the head and the beams that were or are hits are written (the
// handle remembers the hit masks it left there; another destination, or pgd_host_invalidate, forces a full expansion).
// The result is bit-identical to the dense copy (tests/test_gpu_step.py, tests/test_hostpath.py).
//
// Per step: the environments are cut into chunks; one stream runs their step + packing kernels back to back, a second
// one copies (the kernels of chunk c + 1 never wait for the copy of chunk c); per chunk ONE device-to-host copy carries
// [hit count | per-group offsets | rewards | dones | infos | head + mask rows | the first hits]; the host waits for
// chunk c, expands it with the pool while chunks c + 1.. are still computing / copying.
#include <cuda_runtime.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/pgdrive_b200.h"
#include "pgd_internal.h"

#define HP_GROUP 64           // rows per group: one CTA of the packing kernel, one work item of the host pool
#define HP_PACK_WARPS 16      // 4 rows per warp

// ---- device: dense rows -> [head + mask] rows + hit values compacted per group --------------------------------------
// Group g of a chunk reserves its segment of `hits` with one atomicAdd on the chunk's counter (segments are in arrival
// order; `seg[g]` says where each one starts) and writes its rows' hits into it in row order, beams in order.
__global__ void __launch_bounds__(HP_PACK_WARPS * 32)
pgd_pack_compact_kernel(const float* __restrict__ dense, int row_begin, int m, int obs_dim, int* __restrict__ total,
                        int* __restrict__ seg, float* __restrict__ base, float* __restrict__ hits) {
  __shared__ unsigned s_mask[HP_GROUP][8];
  __shared__ int s_cnt[HP_GROUP], s_off[HP_GROUP], s_warp[HP_GROUP / 32], s_seg;
  const int head = obs_dim - PGD_LIDAR_BEAMS, bw = head + 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = blockIdx.x * HP_GROUP;
  const int rows = min(HP_GROUP, m - r0);
  constexpr int PER_WARP = HP_GROUP / HP_PACK_WARPS;
  for (int k = 0; k < PER_WARP; ++k) {
    const int r = warp * PER_WARP + k;
    if (r >= rows) {
      if (lane == 0) s_cnt[r] = 0;
      continue;
    }
    const float* src = dense + (size_t)(row_begin + r0 + r) * obs_dim;
    float* out = base + (size_t)(r0 + r) * bw;
    for (int i = lane; i < head; i += 32) out[i] = src[i];
    unsigned mine = 0u;
    int cnt = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int beam = c * 32 + lane;
      const float v = beam < PGD_LIDAR_BEAMS ? src[head + beam] : 1.0f;
      const unsigned mask = __ballot_sync(0xffffffffu, __float_as_uint(v) != 0x3f800000u);
      if (lane == c) mine = mask;
      cnt += __popc(mask);
    }
    if (lane < 8) {
      out[head + lane] = __uint_as_float(mine);
      s_mask[r][lane] = mine;
    }
    if (lane == 0) s_cnt[r] = cnt;
  }
  __syncthreads();
  if (threadIdx.x < HP_GROUP) {  // exclusive scan of the group's counts
    const int v = s_cnt[threadIdx.x];
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    s_off[threadIdx.x] = inc - v;
    if (lane == 31) s_warp[warp] = inc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int sum = 0;
    for (int w = 0; w < HP_GROUP / 32; ++w) {
      const int t = s_warp[w];
      s_warp[w] = sum;
      sum += t;
    }
    s_seg = sum ? atomicAdd(total, sum) : 0;
    seg[blockIdx.x] = s_seg;
  }
  __syncthreads();
  for (int k = 0; k < PER_WARP; ++k) {
    const int r = warp * PER_WARP + k;
    if (r >= rows || s_cnt[r] == 0) continue;
    const float* src = dense + (size_t)(row_begin + r0 + r) * obs_dim;
    int o = s_seg + s_warp[r >> 5] + s_off[r];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const unsigned mask = s_mask[r][c];
      if (mask) {
        if ((mask >> lane) & 1u) hits[o + __popc(mask & ((1u << lane) - 1u))] = src[head + c * 32 + lane];
        o += __popc(mask);
      }
    }
  }
}

// ---- host: a pool of threads that spin while a step is in flight: pgd_hostpool.h (pure C++, sanitizer-checked) ------
#include "pgd_hostpool.h"

extern "C" int pgd_host_pool_selftest(int32_t workers, int32_t items, int32_t rounds) {
  if (workers < 1 || items < 0 || rounds < 0) return fail(-1, "pgd_host_pool_selftest: bad argument");
  HostPool pool(workers - 1);
  std::vector<std::atomic<int>> count((size_t)items);
  for (auto& c : count) c.store(0);
  for (int r = 0; r < rounds; ++r) {
    if (r % 7 == 0) pool.begin();  // sleep / wake cycles in between, as between steps
    const std::function<void(int)> job = [&](int i) { count[(size_t)i].fetch_add(1, std::memory_order_relaxed); };
    pool.run(items, job);
    if (r % 7 == 6) pool.end();
    for (int i = 0; i < items; ++i)
      if (count[(size_t)i].load() != r + 1) return fail(-4, "pgd_host_pool_selftest: an item ran twice or not at all");
  }
  pool.end();
  return 0;
}

// ---- one chunk of environments: where its pieces sit in the device / pinned-host transfer buffer ----------------------
struct HostChunk {
  int b, e, groups;                                            // environments [b, e)
  size_t off_total, off_seg, off_rew, off_done, off_info, off_base, off_hits, bytes;
  int expect;                                                  // hit values that travel with the first copy
  char *dev, *host;                                            // start of this chunk in the two buffers
  cudaEvent_t packed, ready;                                   // results on the device / in host memory
};

struct HostPath {
  std::vector<HostChunk> chunks;
  char *dev = nullptr, *host = nullptr;
  size_t bytes = 0, header_bytes = 0;
  uint32_t* mask = nullptr;   // [num_envs][8]: hit masks of the rows left in `last_obs`
  const float* last_obs = nullptr;
  bool state_valid = false;
  HostPool* pool = nullptr;
  float *h_act = nullptr, *d_act = nullptr;
  float* d_obs = nullptr;     // dense rows of the step kernel (device only)
  uint64_t last_h2d = 0, last_d2h = 0;  // bytes over PCIe in the last step
  int rows = 0, obs_dim = 0;            // pgd_rows_to_host: the batch this instance was sized for
  int first_hits = -1;        // >= 0 (PGDRIVE_B200_HOST_FIRST_HITS, tests): hit values per row in a chunk's first copy
};

static size_t up16(size_t v) { return (v + 15) & ~(size_t)15; }

// pure host code, exported for the CPU tests: expand rows [r_begin, r_end) of one chunk image into dense rows
extern "C" int pgd_host_expand_rows(const float* base, const float* hits, int32_t hit_offset, int32_t n_rows,
                                    int32_t obs_dim, float* dense, uint32_t* mask_state, int32_t full) {
  if (!base || !dense || !mask_state || n_rows < 0 || obs_dim < PGD_LIDAR_BEAMS)
    return fail(-1, "pgd_host_expand_rows: bad argument");
  const int head = obs_dim - PGD_LIDAR_BEAMS, bw = head + 8;
  int off = hit_offset;
  // The destination rows are a strided stream the hardware prefetcher does not follow (a 136-byte head every 1 096
  // bytes): ask for the lines of the row HP_AHEAD rows on while this one is written (98 -> 59 ns per row on one thread of
  // the development host; what is left is the core's DRAM bandwidth: 168 bytes read, 3 lines owned and written back).
  constexpr int HP_AHEAD = 8;
  const size_t head_bytes = (size_t)head * 4;
  for (int r = 0; r < n_rows; ++r) {
    const float* bp = base + (size_t)r * bw;
    float* dst = dense + (size_t)r * obs_dim;
    if (r + HP_AHEAD < n_rows) {
      const char* nd = (const char*)(dst + (size_t)HP_AHEAD * obs_dim);
      for (size_t b = 0; b < head_bytes + 64; b += 64) __builtin_prefetch(nd + b, 1, 0);
      __builtin_prefetch(mask_state + (size_t)(r + HP_AHEAD) * 8, 1, 0);
    }
    memcpy(dst, bp, head_bytes);
    uint32_t nm[8];
    memcpy(nm, bp + head, 32);
    uint32_t* om = mask_state + (size_t)r * 8;
    float* beams = dst + head;
    if (full) {
      for (int i = 0; i < PGD_LIDAR_BEAMS; ++i) beams[i] = 1.0f;
      for (int c = 0; c < 8; ++c)
        for (uint32_t m = nm[c]; m; m &= m - 1) beams[c * 32 + __builtin_ctz(m)] = hits[off++];
      memcpy(om, nm, 32);
    } else {
      uint32_t any = 0;
      for (int c = 0; c < 8; ++c) any |= nm[c] | om[c];
      if (any) {  // most rows: no beam was or is a hit, nothing but the head to write
        for (int c = 0; c < 8; ++c) {
          const uint32_t m = nm[c];
          for (uint32_t t = m | om[c]; t; t &= t - 1) {
            const int bit = __builtin_ctz(t);
            beams[c * 32 + bit] = ((m >> bit) & 1u) ? hits[off++] : 1.0f;
          }
        }
        memcpy(om, nm, 32);
      }
    }
  }
  return off - hit_offset;
}

// threads that expand rows (the caller's included): the CPUs this process may run on, at most 16
static int pool_threads(int n_rows, bool every_rank_calls) {
  int workers = (int)std::thread::hardware_concurrency();
  cpu_set_t set;  // a cpuset smaller than the machine
  if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) workers = CPU_COUNT(&set);
  const char* lws = getenv("LOCAL_WORLD_SIZE");  // torchrun: the ranks of this node share its CPUs
  if (lws && atoi(lws) > 1) {
    // every rank expands its own shard at the same time: share the CPUs; only one rank expands (the gathered batch on
    // rank 0): leave a CPU to each of the others, which wait in a barrier meanwhile
    workers = every_rank_calls ? workers / atoi(lws) : workers - (atoi(lws) - 1);
  }
  const char* w = getenv("PGDRIVE_B200_HOST_THREADS");
  if (w && atoi(w) > 0) workers = atoi(w);
  if (workers > 16) workers = 16;
  if (workers < 1) workers = 1;
  if (n_rows < 4096) workers = 1;  // a handful of rows: the caller's thread alone
  return workers;
}

static int hostpath_init(PgdHandle* h) {
  if (h->hostpath) return 0;
  HostPath* hp = new HostPath();
  const int n = h->cfg.num_envs;
  const int od = pgd_obs_dim(&h->cfg), bw = od - PGD_LIDAR_BEAMS + 8;
  int n_chunks = n >= 8192 ? 4 : 1;  // profiles/r04i_e2e_bench.jsonl, r04j: 4 chunks beat 8 and 16 at 65 536 environments
  const char* fixed = getenv("PGDRIVE_B200_HOST_CHUNKS");
  if (fixed && atoi(fixed) > 0) n_chunks = atoi(fixed);
  const char* fh = getenv("PGDRIVE_B200_HOST_FIRST_HITS");  // tests: 0 forces the second copy
  if (fh && atoi(fh) >= 0 && atoi(fh) <= PGD_LIDAR_BEAMS) hp->first_hits = atoi(fh);
  const int per = (n / n_chunks + HP_GROUP - 1) / HP_GROUP * HP_GROUP;
  size_t at = 0;
  for (int b = 0; b < n; b += per) {
    HostChunk c;
    memset(&c, 0, sizeof(c));
    c.b = b;
    c.e = b + per < n ? b + per : n;
    const size_t m = (size_t)(c.e - c.b);
    c.groups = (int)((m + HP_GROUP - 1) / HP_GROUP);
    c.off_total = 0;
    c.off_seg = 16;
    c.off_rew = c.off_seg + up16((size_t)c.groups * 4);
    c.off_done = c.off_rew + up16(m * 4);
    c.off_info = c.off_done + up16(m);
    c.off_base = c.off_info + up16(m * sizeof(PgdInfo));
    c.off_hits = c.off_base + up16(m * bw * 4);
    c.expect = hp->first_hits >= 0 ? (int)m * hp->first_hits : (int)m * 4;
    c.bytes = c.off_hits + up16(m * PGD_LIDAR_BEAMS * 4);
    c.dev = (char*)at;  // offsets for now
    at += (c.bytes + 255) & ~(size_t)255;
    hp->chunks.push_back(c);
  }
  hp->bytes = at;
  CU(cudaMalloc(&hp->dev, hp->bytes));
  CU(cudaMallocHost(&hp->host, hp->bytes));
  CU(cudaMemset(hp->dev, 0, hp->bytes));
  for (auto& c : hp->chunks) {
    const size_t o = (size_t)c.dev;
    c.dev = hp->dev + o;
    c.host = hp->host + o;
    CU(cudaEventCreateWithFlags(&c.ready, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c.packed, cudaEventDisableTiming));
  }
  CU(cudaMallocHost(&hp->h_act, (size_t)n * 8));
  CU(cudaMalloc(&hp->d_act, (size_t)n * 8));
  CU(cudaMalloc(&hp->d_obs, (size_t)n * od * 4));
  hp->mask = (uint32_t*)calloc((size_t)n * 8, 4);
  if (!hp->mask) return fail(-2, "pgd_step_host: out of host memory");
  hp->pool = new HostPool(pool_threads(n, true) - 1);
  h->hostpath = hp;
  return 0;
}

static void hostpath_free(HostPath* hp);

void pgd_hostpath_destroy(PgdHandle* h) {
  hostpath_free((HostPath*)h->hostpath);
  hostpath_free((HostPath*)h->rowspath);
  h->hostpath = h->rowspath = nullptr;
}

static void hostpath_free(HostPath* hp) {
  if (!hp) return;
  delete hp->pool;
  for (auto& c : hp->chunks) {
    cudaEventDestroy(c.ready);
    cudaEventDestroy(c.packed);
  }
  cudaFree(hp->dev);
  cudaFreeHost(hp->host);
  cudaFreeHost(hp->h_act);
  cudaFree(hp->d_act);
  cudaFree(hp->d_obs);
  free(hp->mask);
  delete hp;
}

extern "C" int pgd_host_transfer_bytes(PgdHandle* h, uint64_t* h2d, uint64_t* d2h) {
  if (!h || !h2d || !d2h) return fail(-1, "pgd_host_transfer_bytes: null argument");
  HostPath* hp = (HostPath*)h->hostpath;
  if (h->last_host_call == 2) hp = (HostPath*)h->rowspath;
  *h2d = hp ? hp->last_h2d : 0;
  *d2h = hp ? hp->last_d2h : 0;
  return 0;
}

extern "C" int pgd_host_invalidate(PgdHandle* h) {
  if (!h) return fail(-1, "pgd_host_invalidate: null handle");
  if (h->hostpath) ((HostPath*)h->hostpath)->state_valid = false;
  if (h->rowspath) ((HostPath*)h->rowspath)->state_valid = false;
  return 0;
}

extern "C" int pgd_step_host(PgdHandle* h, const float* actions, float* obs, float* reward, uint8_t* done,
                             PgdInfo* info) {
  if (!h || !actions || !obs || !reward || !done) return fail(-1, "pgd_step_host: null argument");
  if (!h->tables_loaded) return fail(-3, "pgd_step_host: no tables loaded");
  CU(cudaSetDevice(h->device));
  h->call_index++;
  h->last_host_call = 1;
  if (int rc = hostpath_init(h)) return rc;
  HostPath* hp = (HostPath*)h->hostpath;
  const size_t n = (size_t)h->cfg.num_envs;
  const int od = pgd_obs_dim(&h->cfg), head = od - PGD_LIDAR_BEAMS, bw = head + 8;
  cudaStream_t s0 = h->own_stream, s1 = h->own_stream2;
  if (h->have_last) CU(cudaStreamWaitEvent(s0, h->ev_last, 0));  // order after the caller-stream reset / step
  memcpy(hp->h_act, actions, n * 8);
  CU(cudaMemcpyAsync(hp->d_act, hp->h_act, n * 8, cudaMemcpyHostToDevice, s0));
  // the hit counters of all chunks (first word of every chunk's buffer): one strided memset
  const size_t pitch = hp->chunks.size() > 1 ? (size_t)(hp->chunks[1].dev - hp->chunks[0].dev) : 16;
  CU(cudaMemset2DAsync(hp->dev, pitch, 0, 4, hp->chunks.size(), s0));
  hp->last_h2d = n * 8;
  hp->last_d2h = 0;
  // With lidar noise every beam differs from 1.0 and the packed row is longer than the dense one: ship dense rows.
  const bool dense_rows = h->cfg.lidar_gaussian_noise > 0.0f || getenv("PGDRIVE_B200_HOST_DENSE") != nullptr;
  for (size_t k = 0; k < hp->chunks.size(); ++k) {
    HostChunk& c = hp->chunks[k];
    const int m = c.e - c.b;
    // the kernel indexes its outputs by global environment: hand it bases that put [b, e) into this chunk's buffer
    float* rew = (float*)(c.dev + c.off_rew) - c.b;
    uint8_t* dn = (uint8_t*)(c.dev + c.off_done) - c.b;
    PgdInfo* inf = info ? (PgdInfo*)(c.dev + c.off_info) - c.b : nullptr;
    if (int rc = pgd_launch_step(h, 0, c.b, c.e, hp->d_act, hp->d_obs, rew, dn, inf, s0)) return rc;
    if (!dense_rows) {
      pgd_pack_compact_kernel<<<c.groups, HP_PACK_WARPS * 32, 0, s0>>>(
          hp->d_obs, c.b, m, od, (int*)(c.dev + c.off_total), (int*)(c.dev + c.off_seg), (float*)(c.dev + c.off_base),
          (float*)(c.dev + c.off_hits));
      h->launches++;
    }
    CU(cudaEventRecord(c.packed, s0));
    CU(cudaStreamWaitEvent(s1, c.packed, 0));
    if (dense_rows) {
      CU(cudaMemcpyAsync(c.host, c.dev, c.off_base, cudaMemcpyDeviceToHost, s1));
      CU(cudaMemcpyAsync(obs + (size_t)c.b * od, hp->d_obs + (size_t)c.b * od, (size_t)m * od * 4,
                         cudaMemcpyDeviceToHost, s1));
      hp->last_d2h += c.off_base + (size_t)m * od * 4;
    } else {
      const size_t first_bytes = c.off_hits + up16((size_t)c.expect * 4);
      CU(cudaMemcpyAsync(c.host, c.dev, first_bytes, cudaMemcpyDeviceToHost, s1));
      hp->last_d2h += first_bytes;
    }
    CU(cudaEventRecord(c.ready, s1));
  }
  CU(cudaGetLastError());
  const bool full = !(hp->state_valid && hp->last_obs == obs);
  if (dense_rows) hp->state_valid = false;
  hp->pool->begin();
  int rc = 0;
  for (size_t k = 0; k < hp->chunks.size() && rc == 0; ++k) {
    HostChunk& c = hp->chunks[k];
    cudaError_t err;  // spin: the chunk is microseconds away and a blocking wait costs a wake-up
    while ((err = cudaEventQuery(c.ready)) == cudaErrorNotReady) cpu_relax();
    if (err != cudaSuccess) {
      rc = fail(-2, std::string("pgd_step_host: ") + cudaGetErrorString(err));
      break;
    }
    const int m = c.e - c.b;
    const float* c_hits = (const float*)(c.host + c.off_hits);
    if (!dense_rows) {
      const int total = *(const int*)(c.host + c.off_total);
      const int had = c.expect;
      // the next step's first copy carries a quarter more than this step needed (hits change slowly between steps)
      if (hp->first_hits < 0) c.expect = std::min(m * PGD_LIDAR_BEAMS, total + total / 4 + 64);
      if (total > had) {  // more hits than travelled with the first copy: fetch the rest
        const size_t have = (size_t)had * 4;
        err = cudaMemcpyAsync(c.host + c.off_hits + have, c.dev + c.off_hits + have, (size_t)total * 4 - have,
                              cudaMemcpyDeviceToHost, s1);
        hp->last_d2h += (size_t)total * 4 - have;
        if (err == cudaSuccess) err = cudaStreamSynchronize(s1);
        if (err != cudaSuccess) {
          rc = fail(-2, std::string("pgd_step_host: ") + cudaGetErrorString(err));
          break;
        }
      }
    }
    const std::function<void(int)> job = [&](int g) {
      const int r0 = g * HP_GROUP, rows = m - r0 < HP_GROUP ? m - r0 : HP_GROUP;
      const size_t env = (size_t)c.b + r0;
      memcpy(reward + env, c.host + c.off_rew + (size_t)r0 * 4, (size_t)rows * 4);
      memcpy(done + env, c.host + c.off_done + r0, (size_t)rows);
      if (info) memcpy(info + env, c.host + c.off_info + (size_t)r0 * sizeof(PgdInfo), (size_t)rows * sizeof(PgdInfo));
      if (!dense_rows)
        pgd_host_expand_rows((const float*)(c.host + c.off_base) + (size_t)r0 * bw, c_hits,
                             ((const int*)(c.host + c.off_seg))[g], rows, od, obs + env * od, hp->mask + env * 8,
                             full ? 1 : 0);
    };
    hp->pool->run(c.groups, job);
  }
  hp->pool->end();
  if (rc) {
    cudaStreamSynchronize(s0);
    cudaStreamSynchronize(s1);
    hp->state_valid = false;
    return rc;
  }
  if (dense_rows) {  // the dense copies went straight into the caller's array: wait for them
    CU(cudaStreamSynchronize(s0));
    CU(cudaStreamSynchronize(s1));
  } else {
    hp->last_obs = obs;
    hp->state_valid = true;
  }
  return 0;
}

// ---- any batch of observation rows in HBM -> host arrays, packed over PCIe ------------------------------------------
// The same transfer for rows that are already in device memory -- e.g. the whole gathered batch in rank 0's HBM
// (bench.py's end-to-end leg at N > 1: 577 MB dense per step at 8 GPUs).  `stream`: the stream the rows were written
// on; the call returns when the host arrays are valid.  Delta expansion as in pgd_step_host (same destination array as
// in the previous call with the same number of rows; pgd_host_invalidate resets both).
static int rowspath_init(PgdHandle* h, int n_rows, int od) {
  HostPath* old = (HostPath*)h->rowspath;
  if (old && old->rows == n_rows && old->obs_dim == od) return 0;
  hostpath_free(old);
  h->rowspath = nullptr;
  HostPath* hp = new HostPath();
  hp->rows = n_rows;
  hp->obs_dim = od;
  const int bw = od - PGD_LIDAR_BEAMS + 8;
  const int per = n_rows >= 4 * 16384 ? 16384 * ((n_rows / 16384 + 15) / 16) : n_rows >= 8192 ?
      (n_rows / 4 + HP_GROUP - 1) / HP_GROUP * HP_GROUP : n_rows;  // at most 16 chunks of >= 16 384 rows
  size_t at = 0;
  for (int b = 0; b < n_rows; b += per) {
    HostChunk c;
    memset(&c, 0, sizeof(c));
    c.b = b;
    c.e = b + per < n_rows ? b + per : n_rows;
    const size_t m = (size_t)(c.e - c.b);
    c.groups = (int)((m + HP_GROUP - 1) / HP_GROUP);
    c.off_seg = 16;
    c.off_rew = c.off_done = c.off_info = c.off_base = c.off_seg + up16((size_t)c.groups * 4);
    c.off_hits = c.off_base + up16(m * bw * 4);
    c.expect = (int)m * 4;
    c.bytes = c.off_hits + up16(m * PGD_LIDAR_BEAMS * 4);
    c.dev = (char*)at;
    at += (c.bytes + 255) & ~(size_t)255;
    hp->chunks.push_back(c);
  }
  hp->bytes = at;
  h->rowspath = hp;  // from here on pgd_hostpath_destroy releases whatever was allocated
  CU(cudaMalloc(&hp->dev, hp->bytes));
  CU(cudaMallocHost(&hp->host, hp->bytes));
  CU(cudaMemset(hp->dev, 0, hp->bytes));
  for (auto& c : hp->chunks) {
    const size_t o = (size_t)c.dev;
    c.dev = hp->dev + o;
    c.host = hp->host + o;
    CU(cudaEventCreateWithFlags(&c.ready, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c.packed, cudaEventDisableTiming));
  }
  hp->mask = (uint32_t*)calloc((size_t)n_rows * 8, 4);
  if (!hp->mask) return fail(-2, "pgd_rows_to_host: out of host memory");
  hp->pool = new HostPool(pool_threads(n_rows, false) - 1);
  return 0;
}

extern "C" int pgd_rows_to_host(PgdHandle* h, const float* obs_dev, const float* reward_dev, const uint8_t* done_dev,
                                int32_t n_rows, int32_t obs_dim, float* obs, float* reward, uint8_t* done, void* stream) {
  if (!h || !obs_dev || !obs || n_rows <= 0 || obs_dim < PGD_LIDAR_BEAMS)
    return fail(-1, "pgd_rows_to_host: null pointer, no rows or rows shorter than the lidar");
  if ((reward_dev == nullptr) != (reward == nullptr) || (done_dev == nullptr) != (done == nullptr))
    return fail(-1, "pgd_rows_to_host: reward / done need both their device and their host array");
  if ((uintptr_t)obs_dev & 3) return fail(-1, "pgd_rows_to_host: rows must be 4-byte aligned");
  CU(cudaSetDevice(h->device));
  h->last_host_call = 2;
  if (int rc = rowspath_init(h, n_rows, obs_dim)) return rc;
  HostPath* hp = (HostPath*)h->rowspath;
  const int od = obs_dim, bw = od - PGD_LIDAR_BEAMS + 8;
  cudaStream_t s0 = h->own_stream, s1 = h->own_stream2;
  CU(cudaEventRecord(h->ev_act, (cudaStream_t)stream));  // the rows are complete when the caller's stream gets here
  CU(cudaStreamWaitEvent(s0, h->ev_act, 0));
  CU(cudaStreamWaitEvent(s1, h->ev_act, 0));
  const size_t pitch = hp->chunks.size() > 1 ? (size_t)(hp->chunks[1].dev - hp->chunks[0].dev) : 16;
  CU(cudaMemset2DAsync(hp->dev, pitch, 0, 4, hp->chunks.size(), s0));
  hp->last_h2d = 0;
  hp->last_d2h = 0;
  for (size_t k = 0; k < hp->chunks.size(); ++k) {
    HostChunk& c = hp->chunks[k];
    const int m = c.e - c.b;
    pgd_pack_compact_kernel<<<c.groups, HP_PACK_WARPS * 32, 0, s0>>>(
        obs_dev, c.b, m, od, (int*)(c.dev + c.off_total), (int*)(c.dev + c.off_seg), (float*)(c.dev + c.off_base),
        (float*)(c.dev + c.off_hits));
    h->launches++;
    CU(cudaEventRecord(c.packed, s0));
    CU(cudaStreamWaitEvent(s1, c.packed, 0));
    const size_t first_bytes = c.off_hits + up16((size_t)c.expect * 4);
    CU(cudaMemcpyAsync(c.host, c.dev, first_bytes, cudaMemcpyDeviceToHost, s1));
    hp->last_d2h += first_bytes;
    CU(cudaEventRecord(c.ready, s1));
  }
  CU(cudaGetLastError());
  const bool full = !(hp->state_valid && hp->last_obs == obs);
  hp->pool->begin();
  int rc = 0;
  for (size_t k = 0; k < hp->chunks.size() && rc == 0; ++k) {
    HostChunk& c = hp->chunks[k];
    cudaError_t err;
    while ((err = cudaEventQuery(c.ready)) == cudaErrorNotReady) cpu_relax();
    const int m = c.e - c.b;
    const int total = err == cudaSuccess ? *(const int*)(c.host + c.off_total) : 0;
    const int had = c.expect;
    c.expect = std::min(m * PGD_LIDAR_BEAMS, total + total / 4 + 64);
    if (err == cudaSuccess && total > had) {  // more hits than travelled with the first copy: fetch the rest
      // (s1 also carries the later chunks' copies: the wait below covers them too, which is rare and harmless)
      const size_t have = (size_t)had * 4;
      err = cudaMemcpyAsync(c.host + c.off_hits + have, c.dev + c.off_hits + have, (size_t)total * 4 - have,
                            cudaMemcpyDeviceToHost, s1);
      hp->last_d2h += (size_t)total * 4 - have;
      if (err == cudaSuccess) err = cudaStreamSynchronize(s1);
    }
    if (err != cudaSuccess) {
      rc = fail(-2, std::string("pgd_rows_to_host: ") + cudaGetErrorString(err));
      break;
    }
    const float* c_hits = (const float*)(c.host + c.off_hits);
    const std::function<void(int)> job = [&](int g) {
      const int r0 = g * HP_GROUP, rows = m - r0 < HP_GROUP ? m - r0 : HP_GROUP;
      const size_t row = (size_t)c.b + r0;
      pgd_host_expand_rows((const float*)(c.host + c.off_base) + (size_t)r0 * bw, c_hits,
                           ((const int*)(c.host + c.off_seg))[g], rows, od, obs + row * od, hp->mask + row * 8,
                           full ? 1 : 0);
    };
    hp->pool->run(c.groups, job);
  }
  hp->pool->end();
  // rewards and dones are 5 bytes per row: plain copies behind the rows (enqueued last: a copy into pageable memory
  // blocks the calling thread until the stream gets there)
  cudaError_t err = cudaSuccess;
  if (rc == 0 && reward) err = cudaMemcpyAsync(reward, reward_dev, (size_t)n_rows * 4, cudaMemcpyDeviceToHost, s1);
  if (rc == 0 && err == cudaSuccess && done)
    err = cudaMemcpyAsync(done, done_dev, (size_t)n_rows, cudaMemcpyDeviceToHost, s1);
  hp->last_d2h += (reward ? (size_t)n_rows * 4 : 0) + (done ? (size_t)n_rows : 0);
  if (err == cudaSuccess) err = cudaStreamSynchronize(s1);  // (after an error: whatever is still in flight)
  cudaStreamSynchronize(s0);
  if (rc == 0 && err != cudaSuccess) rc = fail(-2, std::string("pgd_rows_to_host: ") + cudaGetErrorString(err));
  hp->state_valid = rc == 0;
  hp->last_obs = obs;
  return rc;
}
