// The reset path on the device: one warp per seed runs the block search, builds the map, tabulates its collision
// primitives and samples the episode template (pgd_mapgen.cuh), writing straight into the tables the step kernel
// reads.  The algorithm is a sequential search with data-dependent back-tracking, so lane 0 of each warp carries it
// and warps (= seeds) are the parallel dimension: 1 000 seeds occupy every SM of a B200 with ~7 warps.
#include <vector>

#include "pgd_internal.h"
#include "pgd_mapgen.cuh"

using namespace pgdgen;

struct GenBuffers {
  // scratch, per map
  GLane* lanes;
  GRoad* roads;
  GBlock* blocks;
  GBox* boxes;
  int32_t* queue;
  int32_t* cand;
  MT* mt;
  // outputs (fixed stride per map)
  PgdMap* maps;
  PgdLane* out_lanes;
  PgdRoad* out_roads;
  PgdBox* out_boxes;
  int32_t* cell_start;
  int32_t* cell_entries;
  PgdEpisode* episodes;
  PgdSlot* slots;
  int32_t* route_nodes;
  int32_t* route_roads;
  int32_t* counts;  // [n][8]
  int32_t* status;  // [n]
};

__global__ void __launch_bounds__(32) pgd_mapgen_kernel(const int32_t* __restrict__ seeds, int n, PgdGenConfig cfg,
                                                        PgdGenCaps caps, GenBuffers B) {
  const int m = blockIdx.x;
  if (m >= n || threadIdx.x != 0) return;
  GenScratch s;
  s.lanes = B.lanes + (size_t)m * caps.lanes;
  s.roads = B.roads + (size_t)m * caps.roads;
  s.blocks = B.blocks + (size_t)m * caps.blocks;
  s.boxes = B.boxes + (size_t)m * caps.boxes;
  s.queue = B.queue + (size_t)m * 2 * caps.queue;
  s.cand = B.cand + (size_t)m * (3 * caps.cand + 4 * caps.roads);
  s.mt = B.mt + (size_t)m * 3;
  GenOut o;
  o.lane_off = m * caps.lanes;
  o.road_off = m * caps.roads;
  o.box_off = m * caps.boxes;
  o.cell_off = m * caps.cells;
  o.entry_off = m * caps.entries;
  o.slot_off = m * PGD_MAX_SLOTS;
  o.route_off = m * caps.route;
  o.map_id = m;
  o.map = B.maps + m;
  o.lanes = B.out_lanes + o.lane_off;
  o.roads = B.out_roads + o.road_off;
  o.boxes = B.out_boxes + o.box_off;
  o.cell_start = B.cell_start + o.cell_off;
  o.cell_entries = B.cell_entries + o.entry_off;
  o.episode = B.episodes + m;
  o.slots = B.slots + o.slot_off;
  o.route_nodes = B.route_nodes + o.route_off;
  o.route_roads = B.route_roads + o.route_off;
  o.counts = B.counts + (size_t)m * 8;
  o.sequence = nullptr;
  B.status[m] = generate_one((uint64_t)seeds[m], cfg, caps, s, o);
}

static const char* gen_error_text(int code) {
  static const char* text[] = {"ok", "lane pool full", "road pool full", "box table full", "grid cell table full",
                               "grid entry table full", "search queue full", "route table full",
                               "spawn candidate list full", "more than 32 vehicle slots",
                               "block search could not finish", "road lookup failed", "can not set a destination",
                               "more than 11 traffic trigger groups", "too many blocks", "unsupported generator config"};
  return (code >= 0 && code <= 15) ? text[code] : "unknown";
}

extern "C" int pgd_generate_tables(PgdHandle* h, const int32_t* seeds, int32_t n, const PgdGenConfig* gen,
                                   const PgdGenCaps* caps, int32_t* status_out, int32_t* counts_out, void* stream) {
  if (!h || !seeds || !gen || !caps || !status_out || n <= 0) return fail(-1, "pgd_generate_tables: bad argument");
  if (caps->blocks < gen->block_num + 1 || caps->lanes <= 0 || caps->roads <= 0 || caps->boxes <= 0 ||
      caps->cells <= 0 || caps->entries <= 0 || caps->queue <= 0 || caps->route <= 0 || caps->cand <= 0)
    return fail(-1, "pgd_generate_tables: capacities must be positive and hold block_num + 1 blocks");
  if ((int64_t)n * caps->entries > INT32_MAX || (int64_t)n * caps->boxes > INT32_MAX)
    return fail(-1, "pgd_generate_tables: too many maps for 32-bit table offsets");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t N = (size_t)n;
  // outputs: become the handle's tables
  const size_t out_bytes[10] = {N * sizeof(PgdMap), N * caps->lanes * sizeof(PgdLane), N * caps->roads * sizeof(PgdRoad),
                                N * caps->boxes * sizeof(PgdBox), N * caps->cells * 4, N * caps->entries * 4,
                                N * sizeof(PgdEpisode), N * PGD_MAX_SLOTS * sizeof(PgdSlot), N * caps->route * 4,
                                N * caps->route * 4};
  const int64_t out_count[10] = {n, (int64_t)n * caps->lanes, (int64_t)n * caps->roads, (int64_t)n * caps->boxes,
                                 (int64_t)n * caps->cells, (int64_t)n * caps->entries, n, (int64_t)n * PGD_MAX_SLOTS,
                                 (int64_t)n * caps->route, (int64_t)n * caps->route};
  void* out[10] = {nullptr};
  // scratch
  const size_t scr_bytes[9] = {N * caps->lanes * sizeof(GLane), N * caps->roads * sizeof(GRoad),
                               N * caps->blocks * sizeof(GBlock), N * caps->boxes * sizeof(GBox),
                               N * 2 * caps->queue * 4, N * (3 * (size_t)caps->cand + 4 * (size_t)caps->roads) * 4,
                               N * 3 * sizeof(MT), N * 8 * 4 + N * 4, N * 4};
  void* scr[9] = {nullptr};
  auto release = [&](bool keep_out) {
    for (int i = 0; i < 9; ++i) cudaFree(scr[i]);
    if (!keep_out)
      for (int i = 0; i < 10; ++i) cudaFree(out[i]);
  };
  for (int i = 0; i < 10; ++i) {
    if (cudaMalloc(&out[i], out_bytes[i]) != cudaSuccess || cudaMemsetAsync(out[i], 0, out_bytes[i], st) != cudaSuccess) {
      release(false);
      return fail(-2, "pgd_generate_tables: out of device memory for the tables");
    }
  }
  for (int i = 0; i < 9; ++i) {
    if (cudaMalloc(&scr[i], scr_bytes[i]) != cudaSuccess) {
      release(false);
      return fail(-2, "pgd_generate_tables: out of device memory for the generator's scratch");
    }
  }
  GenBuffers B;
  B.lanes = (GLane*)scr[0]; B.roads = (GRoad*)scr[1]; B.blocks = (GBlock*)scr[2]; B.boxes = (GBox*)scr[3];
  B.queue = (int32_t*)scr[4]; B.cand = (int32_t*)scr[5]; B.mt = (MT*)scr[6];
  B.counts = (int32_t*)scr[7]; B.status = B.counts + N * 8;
  int32_t* d_seeds = (int32_t*)scr[8];
  B.maps = (PgdMap*)out[0]; B.out_lanes = (PgdLane*)out[1]; B.out_roads = (PgdRoad*)out[2];
  B.out_boxes = (PgdBox*)out[3]; B.cell_start = (int32_t*)out[4]; B.cell_entries = (int32_t*)out[5];
  B.episodes = (PgdEpisode*)out[6]; B.slots = (PgdSlot*)out[7]; B.route_nodes = (int32_t*)out[8];
  B.route_roads = (int32_t*)out[9];
  cudaError_t e = cudaMemcpyAsync(d_seeds, seeds, N * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(scr[7], 0xff, scr_bytes[7], st);
  if (e == cudaSuccess) {
    pgd_mapgen_kernel<<<n, 32, 0, st>>>(d_seeds, n, *gen, *caps, B);
    h->launches++;
    e = cudaGetLastError();
  }
  std::vector<int32_t> counts(N * 8);
  if (e == cudaSuccess) e = cudaMemcpyAsync(status_out, B.status, N * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(counts.data(), B.counts, N * 8 * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    release(false);
    return fail(-2, std::string("pgd_generate_tables: ") + cudaGetErrorString(e));
  }
  if (counts_out) memcpy(counts_out, counts.data(), N * 8 * 4);
  for (int i = 0; i < n; ++i) {
    if (status_out[i] != 0) {
      release(false);
      return fail(-4, "pgd_generate_tables: seed " + std::to_string(seeds[i]) + ": " + gen_error_text(status_out[i]));
    }
  }
  for (int i = 0; i < n; ++i) {
    if (counts[(size_t)i * 8 + 5] > h->cfg.num_slots) {
      release(false);
      return fail(-3, "pgd_generate_tables: seed " + std::to_string(seeds[i]) + " needs " +
                          std::to_string(counts[(size_t)i * 8 + 5]) + " vehicle slots; the handle has " +
                          std::to_string(h->cfg.num_slots));
    }
  }
  // install
  CU(cudaDeviceSynchronize());
  for (int i = 0; i < 10; ++i) {
    cudaFree(h->table_mem[i]);
    h->table_mem[i] = out[i];
    h->table_count[i] = out_count[i];
  }
  h->T.maps = B.maps; h->T.lanes = B.out_lanes; h->T.roads = B.out_roads; h->T.boxes = B.out_boxes;
  h->T.cell_start = B.cell_start; h->T.cell_entries = B.cell_entries; h->T.episodes = B.episodes;
  h->T.slots = B.slots; h->T.route_nodes = B.route_nodes; h->T.route_roads = B.route_roads;
  h->n_episodes = n;
  h->tables_loaded = true;
  release(true);
  return 0;
}

// Replace map / episode `index` of tables made by pgd_generate_tables (fixed stride per map: caps) with a host-built
// single-seed table set (offsets from 0, as pgdrive_b200/tables.py makes them).
extern "C" int pgd_patch_tables(PgdHandle* h, int32_t index, const PgdTables* t, const PgdGenCaps* caps) {
  if (!h || !t || !caps) return fail(-1, "pgd_patch_tables: null argument");
  if (!h->tables_loaded || index < 0 || index >= h->n_episodes) return fail(-1, "pgd_patch_tables: index out of range");
  if (h->table_count[1] != (int64_t)h->n_episodes * caps->lanes || h->table_count[7] != (int64_t)h->n_episodes * PGD_MAX_SLOTS)
    return fail(-1, "pgd_patch_tables: the handle's tables were not generated with these capacities");
  if (t->n_maps != 1 || t->n_episodes != 1) return fail(-1, "pgd_patch_tables: expects exactly one map and one episode");
  if (t->n_lanes > caps->lanes || t->n_roads > caps->roads || t->n_boxes > caps->boxes || t->n_cell_start > caps->cells ||
      t->n_cell_entries > caps->entries || t->n_slots > PGD_MAX_SLOTS || t->n_route > caps->route)
    return fail(-3, "pgd_patch_tables: the map does not fit the per-map capacities");
  if (t->episodes[0].n_slots > h->cfg.num_slots) return fail(-3, "pgd_patch_tables: more vehicle slots than the handle has");
  CU(cudaSetDevice(h->device));
  CU(cudaDeviceSynchronize());
  const size_t m = (size_t)index;
  PgdMap mp = t->maps[0];
  mp.lane_off = (int32_t)(m * caps->lanes); mp.road_off = (int32_t)(m * caps->roads); mp.box_off = (int32_t)(m * caps->boxes);
  mp.cell_off = (int32_t)(m * caps->cells); mp.entry_off = (int32_t)(m * caps->entries);
  PgdEpisode ep = t->episodes[0];
  ep.map = index;
  ep.slot_off = (int32_t)(m * PGD_MAX_SLOTS);
  std::vector<PgdSlot> slots(t->slots, t->slots + t->n_slots);
  for (auto& sl : slots) sl.route_off += (int32_t)(m * caps->route);
  CU(cudaMemcpy((PgdMap*)h->table_mem[0] + m, &mp, sizeof(mp), cudaMemcpyHostToDevice));
  CU(cudaMemcpy((PgdLane*)h->table_mem[1] + mp.lane_off, t->lanes, (size_t)t->n_lanes * sizeof(PgdLane), cudaMemcpyHostToDevice));
  CU(cudaMemcpy((PgdRoad*)h->table_mem[2] + mp.road_off, t->roads, (size_t)t->n_roads * sizeof(PgdRoad), cudaMemcpyHostToDevice));
  CU(cudaMemcpy((PgdBox*)h->table_mem[3] + mp.box_off, t->boxes, (size_t)t->n_boxes * sizeof(PgdBox), cudaMemcpyHostToDevice));
  CU(cudaMemcpy((int32_t*)h->table_mem[4] + mp.cell_off, t->cell_start, (size_t)t->n_cell_start * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy((int32_t*)h->table_mem[5] + mp.entry_off, t->cell_entries, (size_t)t->n_cell_entries * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy((PgdEpisode*)h->table_mem[6] + m, &ep, sizeof(ep), cudaMemcpyHostToDevice));
  CU(cudaMemcpy((PgdSlot*)h->table_mem[7] + ep.slot_off, slots.data(), slots.size() * sizeof(PgdSlot), cudaMemcpyHostToDevice));
  CU(cudaMemcpy((int32_t*)h->table_mem[8] + m * caps->route, t->route_nodes, (size_t)t->n_route * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy((int32_t*)h->table_mem[9] + m * caps->route, t->route_roads, (size_t)t->n_route * 4, cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int pgd_table_sizes(PgdHandle* h, int64_t sizes[9]) {
  if (!h || !sizes) return fail(-1, "pgd_table_sizes: null argument");
  if (!h->tables_loaded) return fail(-3, "pgd_table_sizes: no tables loaded");
  for (int i = 0; i < 9; ++i) sizes[i] = h->table_count[i];
  return 0;
}

extern "C" int pgd_download_tables(PgdHandle* h, PgdTables* dst) {
  if (!h || !dst) return fail(-1, "pgd_download_tables: null argument");
  if (!h->tables_loaded) return fail(-3, "pgd_download_tables: no tables loaded");
  CU(cudaSetDevice(h->device));
  CU(cudaDeviceSynchronize());
  const int64_t have[10] = {dst->n_maps, dst->n_lanes, dst->n_roads, dst->n_boxes, dst->n_cell_start,
                            dst->n_cell_entries, dst->n_episodes, dst->n_slots, dst->n_route, dst->n_route};
  void* to[10] = {(void*)dst->maps, (void*)dst->lanes, (void*)dst->roads, (void*)dst->boxes, (void*)dst->cell_start,
                  (void*)dst->cell_entries, (void*)dst->episodes, (void*)dst->slots, (void*)dst->route_nodes,
                  (void*)dst->route_roads};
  const size_t elem[10] = {sizeof(PgdMap), sizeof(PgdLane), sizeof(PgdRoad), sizeof(PgdBox), 4, 4, sizeof(PgdEpisode),
                           sizeof(PgdSlot), 4, 4};
  for (int i = 0; i < 10; ++i) {
    if (have[i] != h->table_count[i]) return fail(-1, "pgd_download_tables: buffer sizes must equal pgd_table_sizes");
    if (have[i] > 0) CU(cudaMemcpy(to[i], h->table_mem[i], (size_t)have[i] * elem[i], cudaMemcpyDeviceToHost));
  }
  return 0;
}
