/* TEST INFRASTRUCTURE ONLY.  Host (g++) build of the one-thread-per-environment step
 * (pgdrive_b200/csrc/pgd_step_v2.cuh is written for host + device): runs the environments one after the other so
 * that the step's logic can be checked against the independent CPU oracle (oracle/pgd_oracle.c) in a container
 * without a GPU.  The product never loads this library.
 */
#include <stdlib.h>
#include <string.h>

#include "../pgdrive_b200/csrc/pgd_step_v2.cuh"

using namespace pgdv2;

struct HostV2 {
  Tables T;
  State S;
  PgdConfig cfg;
};

template <int V>
static void run_v(HostV2* h, int mode, const float* actions, float* obs, float* reward, uint8_t* done, PgdInfo* info) {
  const int n = h->cfg.num_envs, od = pgd_obs_dim(&h->cfg), head = od - PGD_LIDAR_BEAMS;
  for (int e = 0; e < n; ++e) {
    if (mode == 1 && h->S.envi[e].z != V2_DONE_PENDING_RESET) continue;
    LidarCtx<V> lc;
    float* row = obs + (size_t)e * od;
    step_env<V>(h->T, h->S, h->cfg, mode, e, n, actions ? actions + 2 * e : nullptr, row, lc,
                reward ? reward + e : nullptr, done ? done + e : nullptr, info ? info + e : nullptr);
    for (int i = 0; i < PGD_LIDAR_BEAMS; ++i) row[head + i] = lidar_beam<V>(lc, i);
  }
}

static void run(HostV2* h, int mode, const float* actions, float* obs, float* reward, uint8_t* done, PgdInfo* info) {
  if (h->cfg.num_slots == 16) run_v<16>(h, mode, actions, obs, reward, done, info);
  else run_v<32>(h, mode, actions, obs, reward, done, info);
}

extern "C" {

void* v2h_create(const PgdTables* t, const PgdConfig* cfg) {
  HostV2* h = (HostV2*)calloc(1, sizeof(HostV2));
  h->cfg = *cfg;
  h->T.maps = t->maps; h->T.lanes = t->lanes; h->T.roads = t->roads; h->T.boxes = t->boxes;
  h->T.cell_start = t->cell_start; h->T.cell_entries = t->cell_entries; h->T.episodes = t->episodes;
  h->T.slots = t->slots; h->T.route_nodes = t->route_nodes; h->T.route_roads = t->route_roads;
  const size_t nv = (size_t)cfg->num_envs * cfg->num_slots, n = (size_t)cfg->num_envs;
  h->S.pose = (F4*)calloc(nv, 16); h->S.ctrl = (F4*)calloc(nv, 16); h->S.pidl = (F4*)calloc(nv, 16);
  h->S.nav = (I4*)calloc(nv, 16); h->S.misc = (I4*)calloc(nv, 16);
  h->S.envi = (I4*)calloc(n, 16); h->S.envf = (F4*)calloc(n, 16);
  return h;
}

void v2h_destroy(void* p) {
  HostV2* h = (HostV2*)p;
  free(h->S.pose); free(h->S.ctrl); free(h->S.pidl); free(h->S.nav); free(h->S.misc); free(h->S.envi); free(h->S.envf);
  free(h);
}

/* pgd_reset: environments env_ids[i] restart on episode_ids[i]; their observation rows are rewritten */
void v2h_reset(void* p, const int32_t* env_ids, const int32_t* episode_ids, int n, float* obs, PgdInfo* info) {
  HostV2* h = (HostV2*)p;
  for (int i = 0; i < n; ++i) {
    int e = env_ids ? env_ids[i] : i;
    h->S.envi[e].x = episode_ids[i];
    h->S.envi[e].z = V2_DONE_PENDING_RESET;
  }
  run(h, 1, nullptr, obs, nullptr, nullptr, info);
}

/* pgd_get_state for the slot-major layout (exchange format of include/pgd_tables.h) */
void v2h_get_state(void* p, int env, PgdEnvState* out) {
  HostV2* h = (HostV2*)p;
  const int n = h->cfg.num_envs, V = h->cfg.num_slots;
  memset(out, 0, sizeof(*out));
  const I4 ei = h->S.envi[env];
  const F4 ef = h->S.envf[env];
  out->episode = ei.x; out->next_group = ei.y; out->done = ei.z; out->ep_len = ei.w;
  out->prev_steer = ef.x; out->prev_throttle = ef.y; out->ep_reward = ef.z; out->energy = ef.w;
  const int n_slots = h->T.episodes[ei.x].n_slots;
  for (int i = 0; i < n_slots && i < V; ++i) {
    const size_t gi = (size_t)i * n + env;
    const F4 po = h->S.pose[gi], c = h->S.ctrl[gi], l = h->S.pidl[gi];
    const I4 nv = h->S.nav[gi], m = h->S.misc[gi];
    PgdVehState* s = &out->veh[i];
    s->x = po.x; s->y = po.y; s->heading = po.z; s->speed = po.w;
    s->steer = c.x; s->throttle = c.y; s->pid_hp = c.z; s->pid_hi = c.w;
    s->pid_lp = l.x; s->pid_li = l.y; s->target_speed = l.z; s->yaw_rate = l.w;
    s->lane = nv.x; s->ck0 = nv.y & 0xffff; s->ck1 = nv.y >> 16; s->rt_lane = nv.z; s->timer = nv.w;
    s->rnd_n = m.x; s->airborne = m.y; s->flags = m.z;
  }
}

void v2h_step(void* p, const float* actions, float* obs, float* reward, uint8_t* done, PgdInfo* info) {
  run((HostV2*)p, 0, actions, obs, reward, done, info);
}
}
