/* C-ABI of the B200 batched driving simulator: the drop-in boundary for PGDriveEnv's reset/step.
 *
 * The reference's only native boundary is the Cython module pgdrive.cutils (cutils.pyx:60-142,
 * cutils_perceive: one lidar sweep that calls back into Python-wrapped Bullet once per ray).  A C
 * replacement for that op alone cannot be faster than its per-ray call-backs, so the boundary is moved
 * up to the environment step.  Each entry point names the reference interface it replaces
 * (paths under /root/reference/pgdrive).
 *
 * Conventions: every function returns 0 on success or a negative error code; pgd_last_error() gives
 * the message of the calling thread's last failure.  One host thread per handle; handles are
 * independent (one per GPU).  The *_dev calls enqueue their work on the caller's CUDA stream (cudaStream_t
 * passed as void*, NULL = default stream) and do not synchronise; pgd_step_host runs on streams the handle
 * owns, ordered after everything the *_dev calls enqueued before it, and returns when its results are valid.
 * Pointers are borrowed for the duration of the call only.  No torch types appear here.
 */
#ifndef PGDRIVE_B200_H
#define PGDRIVE_B200_H
#include <stdint.h>

#include "pgd_tables.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct PgdHandle PgdHandle;

/* engine/engine_utils.py:8-15 initialize_engine + envs/base_env.py:166-178 lazy_init: allocate the
 * structure-of-arrays state of cfg->num_envs environments x cfg->num_slots vehicle slots on `device`. */
int pgd_create(const PgdConfig* cfg, int device, PgdHandle** out);

/* envs/base_env.py:402-407 close() */
int pgd_destroy(PgdHandle* h);

/* manager/map_manager.py:98-155 update_map (map cache) + manager/traffic_manager.py:239-290: upload
 * maps and per-seed episode templates (host pointers; copied to the device before returning). */
int pgd_load_tables(PgdHandle* h, const PgdTables* tables);

/* envs/base_env.py:269-301 reset(force_seed): environments env_ids[i] (host int32[n]) restart on episode
 * template episode_ids[i]; their rows of obs_dev [num_envs, 274] f32 (and info_dev, may be NULL) are
 * rewritten with the reset observation.  env_ids == NULL means environments 0..n-1. */
int pgd_reset(PgdHandle* h, const int32_t* env_ids, const int32_t* episode_ids, int32_t n, float* obs_dev,
              PgdInfo* info_dev, void* stream);

/* envs/base_env.py:184-224,303-344 step(): one decision step of every environment in ONE kernel launch.
 * actions_dev [num_envs, 2] f32 (steering, throttle); outputs obs_dev [num_envs, 274] f32,
 * reward_dev [num_envs] f32, done_dev [num_envs] u8, info_dev [num_envs] PgdInfo (may be NULL). */
int pgd_step(PgdHandle* h, const float* actions_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
             PgdInfo* info_dev, void* stream);

/* Same step with HOST buffers (the call a gym user makes; envs/base_env.py:184-224 `step`): actions are staged through
 * pinned memory and the call returns when the results are valid in the caller's arrays (page-locked or not).
 * Observation rows cross PCIe packed -- head + 240-bit hit mask per row, the beams that are not 1.0 compacted per chunk
 * -- and a pool of host threads expands them into `obs`.  When `obs` is the array of the previous call, only the head
 * and the beams that were or are hits are written (the handle remembers the hit masks it left there), so the caller
 * must not modify `obs` between calls -- or call pgd_host_invalidate, after which the next step rewrites every beam.
 * The result is bit-identical to pgd_step + a dense copy.  PGDRIVE_B200_HOST_DENSE=1 (or lidar noise, which leaves no
 * beam at 1.0) ships dense rows instead; PGDRIVE_B200_HOST_THREADS / PGDRIVE_B200_HOST_CHUNKS override the pool size
 * (default: the CPUs of the process -- divided by LOCAL_WORLD_SIZE under torchrun --, at most 16) and the number of chunks per step. */
int pgd_step_host(PgdHandle* h, const float* actions, float* obs, float* reward, uint8_t* done, PgdInfo* info);
int pgd_host_invalidate(PgdHandle* h);
int pgd_host_transfer_bytes(PgdHandle* h, uint64_t* h2d, uint64_t* d2h);  /* bytes over PCIe in the last pgd_step_host */
/* The same transfer for observation rows that are already in device memory -- e.g. the whole gathered batch in rank 0's
 * HBM: packed over PCIe, expanded by the host threads into `obs` as a delta against the previous call with the same
 * destination and row count.  `reward_dev` / `done_dev` (and their host arrays) may be NULL.  `stream`: the stream
 * the rows were written on; returns when the host arrays are valid. */
int pgd_rows_to_host(PgdHandle* h, const float* obs_dev, const float* reward_dev, const uint8_t* done_dev, int32_t n_rows,
                     int32_t obs_dim, float* obs, float* reward, uint8_t* done, void* stream);
/* The host half of that path, pure CPU code (exported for tests): expand `n_rows` rows of [head | 8 mask words] (`base`)
 * plus their hit values, `hits[hit_offset...]` in row and beam order, into dense rows; `mask_state` (8 words per row)
 * holds the masks of the rows `dense` held before and receives the new ones; `full` != 0 rewrites every beam.
 * Returns the number of hit values consumed, or a negative error code. */
int pgd_host_expand_rows(const float* base, const float* hits, int32_t hit_offset, int32_t n_rows, int32_t obs_dim,
                         float* dense, uint32_t* mask_state, int32_t full);
/* Self-test of the host thread pool (no GPU): `rounds` jobs of `items` items on `workers` threads; 0 when every item
 * of every job ran exactly once. */
int pgd_host_pool_selftest(int32_t workers, int32_t items, int32_t rounds);

/* component/vehicle/base_vehicle.py:683-698 get_state / set_state, for the whole environment (debugging
 * and parity tests; synchronous). */
int pgd_get_state(PgdHandle* h, int32_t env, PgdEnvState* out);
int pgd_set_state(PgdHandle* h, int32_t env, const PgdEnvState* in);

/* The reset path ON THE DEVICE: component/algorithm/BIG.py:67-151 (block search with retry / back-tracking),
 * component/blocks/*.py (block builders), component/map/pg_map.py:48-71 (rebuild from the block sequence),
 * component/blocks/base_block.py:181-463 (static collision primitives), manager/traffic_manager.py:239-290 +
 * component/vehicle_module/navigation.py:99-153 (traffic slots, vehicle parameters, routes), on the reference's own
 * random streams (utils/random_utils.py:14-50: sha512-seeded MT19937, numpy legacy draws).  One warp generates one
 * seed; map i / episode i of the installed tables belong to seeds[i].  seeds, status_out [n] and counts_out [n * 8]
 * (lanes, roads, boxes, cells + 1, grid entries, slots, route entries, blocks; may be NULL) are HOST pointers; the
 * tables never leave the device.  Returns -4 (and still fills status_out) when a seed could not be generated, -3
 * when an episode needs more vehicle slots than the handle has.  Synchronises `stream`. */
int pgd_generate_tables(PgdHandle* h, const int32_t* seeds, int32_t n, const PgdGenConfig* gen, const PgdGenCaps* caps,
                        int32_t* status_out, int32_t* counts_out, void* stream);

/* Replace map / episode `index` of tables made by pgd_generate_tables with a host-built table set of ONE seed (the
 * reference's map cache filled from another source: manager/map_manager.py:134-155).  Used for the seeds on which the
 * reference's own result hangs on the last bit of a libm call (pgdrive_b200/devgen_ties.json): those are built by the
 * reference-pinned host path and patched in, so the device tables equal the host tables on every seed. */
int pgd_patch_tables(PgdHandle* h, int32_t index, const PgdTables* tables, const PgdGenCaps* caps);

/* Element counts of the handle's device tables, in the order of PgdTables (maps, lanes, roads, boxes, cell_start,
 * cell_entries, episodes, slots, route), and a device -> host copy of them into caller-owned buffers of those
 * sizes (component/map/base_map.py:103-118 save_map is the nearest reference interface; used by tests and to cache
 * generated maps). */
int pgd_table_sizes(PgdHandle* h, int64_t sizes[9]);
int pgd_download_tables(PgdHandle* h, PgdTables* dst);

/* Cross-process peer memory (one process per GPU, NVLink / NVSwitch).  The owner allocates a device buffer and
 * exports a 64-byte handle; every other rank opens it and passes `base + its row offset` as obs_dev / reward_dev /
 * done_dev of pgd_step, so the step kernel stores its results straight into the owner's HBM over NVLink: the gather
 * of the observation batch to rank 0 (BASELINE.json north_star) is fused into the kernel and no collective moves
 * payload.  (The reference has no multi-process path at all: engine/engine_utils.py:8-15.) */
int pgd_peer_alloc(PgdHandle* h, uint64_t bytes, void** dev_ptr, unsigned char handle_out[64]);
int pgd_peer_open(PgdHandle* h, const unsigned char handle[64], void** dev_ptr);
int pgd_peer_release(PgdHandle* h, void* dev_ptr, int32_t is_owner);

/* The gather's wire format for observation rows (no counterpart in the reference).  A row is [head | 240 lidar beams] and
 * most beams are exactly 1.0 ("no hit within 50 m").  pgd_pack_rows writes, per row and at a fixed stride of
 * pgd_packed_row_words(obs_dim) floats (obs_dim + 8 rounded up to a multiple of 32: 288 for 274, every packed row on its
 * own 128-byte lines), the head unchanged, a 240-bit hit mask (8 words), the values of the beams that are not 1.0, in beam
 * order, and zeros up to the next 32-byte boundary; pgd_expand_rows restores the rows bit for bit.  `packed_dev` of pgd_pack_rows may be peer memory (pgd_peer_open): a
 * rank packs its local rows straight into rank 0's HBM and only the bytes written cross NVLink -- about a quarter of
 * the 1 096-byte row --, rank 0 expands them into the whole-batch buffer.  Device pointers, asynchronous on `stream`. */
int pgd_pack_rows(const float* dense_dev, float* packed_dev, int32_t n_rows, int32_t obs_dim, void* stream);
int pgd_expand_rows(const float* packed_dev, float* dense_dev, int32_t n_rows, int32_t obs_dim, void* stream);
/* Delta expansion into a buffer that still holds the rows of an earlier step: `mask_state_dev` (8 uint32 per row, owned by
 * the caller, one per buffer) remembers the hit masks of the rows the buffer holds, and only the head, the beams that
 * were or are hits, and the new mask are written -- a third of the traffic.  `full` != 0 writes every beam and
 * initialises the state (first use of a buffer).  Bit-identical to pgd_expand_rows as long as nobody else writes the
 * buffer's rows in between. */
int pgd_expand_rows_delta(const float* packed_dev, float* dense_dev, uint32_t* mask_state_dev, int32_t n_rows,
                          int32_t obs_dim, int32_t full, void* stream);
int32_t pgd_packed_row_words(int32_t obs_dim);   /* stride of a packed row in floats; -1 when obs_dim < 240 */

/* measurement helpers */
/* *out_dev += sum of the 32-bit words of [dev_ptr, dev_ptr + bytes) (both multiples of 16), 64-bit accumulator: the
 * stand-in for a consumer that reads the whole gathered batch every step (bench.py), exact and order-independent so
 * that the consumer's checksum can be compared with the producers'. */
int pgd_words_checksum(PgdHandle* h, const void* dev_ptr, uint64_t bytes, uint64_t* out_dev, void* stream);
int64_t pgd_state_bytes_per_env(PgdHandle* h);   /* bytes of simulator state kept per environment */
int64_t pgd_launch_count(PgdHandle* h);          /* kernels launched by this handle so far */
int pgd_set_timing(PgdHandle* h, int32_t on);    /* bracket the step kernel with CUDA events */
float pgd_last_kernel_ms(PgdHandle* h);          /* duration of the last timed step kernel (syncs) */

const char* pgd_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
