/* TEST INFRASTRUCTURE ONLY.  Host (g++) build of the role-per-warp step (pgdrive_b200/csrc/pgd_step.cuh is
 * written for host + device): every CTA of 32 environments is emulated by running the phases in order over all
 * (role, lane) pairs -- the loops stand for the CTA barriers -- so that the step's logic, including its
 * shared-memory exchanges, is checked against the independent CPU oracle (oracle/pgd_oracle.c) without a GPU.
 * The product never loads this library. */
#include <stdlib.h>
#include <string.h>

#include "../pgdrive_b200/csrc/pgd_step.cuh"

using namespace pgdstep;

struct HostStep {
  Tables T;
  State S;
  PgdConfig cfg;
  int roles;
  int envs_per_cta;  // lanes of every warp that carry an environment (the kernel launcher picks 2..32)
  uint32_t call_index;  // API calls so far (pgd_abi.cu keeps the same count): lidar-noise key
};

template <int V, int R>
static void run_vr(HostStep* h, int mode, const float* actions, float* obs, float* reward, uint8_t* done, PgdInfo* info) {
  typedef Smem<V, R> SM;
  const int n = h->cfg.num_envs, od = obs_dim_of(h->cfg);
  const size_t bytes = smem_bytes<V, R>(od, h->cfg.decision_repeat);
  unsigned char* raw = (unsigned char*)aligned_alloc(128, (bytes + 127) / 128 * 128);
  SM& sm = *reinterpret_cast<SM*>(raw);
  float* rows = reinterpret_cast<float*>(raw + smem_obs_offset<V, R>());
  TrajPtr traj = reinterpret_cast<TrajPtr>(raw + smem_tv_offset<V, R>(od));
  VisPtr vis = reinterpret_cast<VisPtr>(raw + smem_tv_offset<V, R>(od));
  static Thr<V, R> th[R][PGS_LANES];
  const int epc = h->envs_per_cta;
  for (int env0 = 0; env0 < n; env0 += epc) {
    memset(raw, 0xff, bytes);  // shared memory starts as garbage on the device
    bool any = false;
    const int cta_end = env0 + epc < n ? env0 + epc : n;
    for (int r = 0; r < R; ++r)
      for (int l = 0; l < PGS_LANES; ++l) {
        thread_init(th[r][l], h->T, h->S, h->cfg, mode, l, r, env0 + l, cta_end);
        any = any || th[r][l].valid;
      }
    if (!any) continue;
    for (int r = 0; r < R; ++r) for (int l = 0; l < PGS_LANES; ++l) phase_0(sm, th[r][l]);
#define ALL(call) for (int r = 0; r < R; ++r) for (int l = 0; l < PGS_LANES; ++l) { Thr<V, R>& t = th[r][l]; (void)t; call; }
    ALL(phase_a(sm, t, h->S, h->cfg, actions, rows));
    ALL(phase_b(sm, t, h->T, rows));
    // phase X: role 0 (all its lanes) first -- trajectories, lanes --, then the traffic roles' first half, then (named
    // barrier PGS_BAR_TRAFFIC on the device) their second half
    for (int l = 0; l < PGS_LANES; ++l) phase_x(sm, th[0][l], h->T, h->S, h->cfg, od, rows, traj, vis);
    for (int r = 1; r < R; ++r)
      for (int l = 0; l < PGS_LANES; ++l) {
        phase_x_items(sm, th[r][l], h->T, h->S, h->cfg, rows, traj);
        phase_x_parked(sm, th[r][l], h->T, h->S, h->cfg, traj);
      }
    memset(rows, 0xff, (size_t)PGS_LANES * od * 4);  // what is left of IDM's data is garbage to the observation
    memset(traj, 0xff, smem_bytes<V, R>(od, h->cfg.decision_repeat) - smem_tv_offset<V, R>(od));
    for (int r = 1; r < R; ++r)
      for (int l = 0; l < PGS_LANES; ++l) phase_x_tail(sm, th[r][l], h->T, h->cfg, od, rows, vis);
    ALL(phase_f(sm, t, h->T, h->S, h->cfg, mode, od, rows, vis, reward, done, info));
    ALL(phase_l_fill(sm, r, l, od, rows));
    ALL(phase_l(sm, h->T, h->S, r, l, env0, od, rows, vis));
    ALL(phase_n(sm, h->cfg, h->call_index, r, l, env0, od, rows));
#undef ALL
    for (int l = 0; l < PGS_LANES; ++l)
      if (sm.wrote[l]) memcpy(obs + (size_t)(env0 + l) * od, rows + (size_t)l * od, (size_t)od * 4);
  }
  free(raw);
}

template <int V>
static void run_v(HostStep* h, int mode, const float* actions, float* obs, float* reward, uint8_t* done, PgdInfo* info) {
  switch (h->roles) {
    case 2: run_vr<V, 2>(h, mode, actions, obs, reward, done, info); break;
    case 3: run_vr<V, 3>(h, mode, actions, obs, reward, done, info); break;
    case 4: run_vr<V, 4>(h, mode, actions, obs, reward, done, info); break;
    case 6: run_vr<V, 6>(h, mode, actions, obs, reward, done, info); break;
    default: run_vr<V, 8>(h, mode, actions, obs, reward, done, info); break;
  }
}

static void run(HostStep* h, int mode, const float* actions, float* obs, float* reward, uint8_t* done, PgdInfo* info) {
  if (h->cfg.num_slots == 16) run_v<16>(h, mode, actions, obs, reward, done, info);
  else if (h->cfg.num_slots == 24) run_v<24>(h, mode, actions, obs, reward, done, info);
  else run_v<32>(h, mode, actions, obs, reward, done, info);
}

extern "C" {

void* sth_create(const PgdTables* t, const PgdConfig* cfg, int roles) {
  HostStep* h = (HostStep*)calloc(1, sizeof(HostStep));
  h->cfg = *cfg;
  h->roles = roles;
  h->envs_per_cta = PGS_LANES;
  h->T.maps = t->maps; h->T.lanes = t->lanes; h->T.roads = t->roads; h->T.boxes = t->boxes;
  h->T.cell_start = t->cell_start; h->T.cell_entries = t->cell_entries; h->T.episodes = t->episodes;
  h->T.slots = t->slots; h->T.route_nodes = t->route_nodes; h->T.route_roads = t->route_roads;
  const size_t nv = (size_t)cfg->num_envs * cfg->num_slots, n = (size_t)cfg->num_envs;
  h->S.pose = (F4*)calloc(nv, 16); h->S.ctrl = (F4*)calloc(nv, 16); h->S.pidl = (F4*)calloc(nv, 16);
  h->S.nav = (I4*)calloc(nv, 16); h->S.misc = (I4*)calloc(nv, 16);
  h->S.envi = (I4*)calloc(n, 16); h->S.envf = (F4*)calloc(n, 16);
  return h;
}

void sth_set_envs_per_cta(void* p, int n) { ((HostStep*)p)->envs_per_cta = n >= 1 && n <= PGS_LANES ? n : PGS_LANES; }

void sth_destroy(void* p) {
  HostStep* h = (HostStep*)p;
  free(h->S.pose); free(h->S.ctrl); free(h->S.pidl); free(h->S.nav); free(h->S.misc); free(h->S.envi); free(h->S.envf);
  free(h);
}

/* pgd_reset: environments env_ids[i] restart on episode_ids[i]; their observation rows are rewritten */
void sth_reset(void* p, const int32_t* env_ids, const int32_t* episode_ids, int n, float* obs, PgdInfo* info) {
  HostStep* h = (HostStep*)p;
  for (int i = 0; i < n; ++i) {
    int e = env_ids ? env_ids[i] : i;
    h->S.envi[e].x = episode_ids[i];
    h->S.envi[e].z = PGS_DONE_PENDING_RESET;
  }
  h->call_index++;
  run(h, 1, nullptr, obs, nullptr, nullptr, info);
}

/* pgd_get_state for the slot-major layout (exchange format of include/pgd_tables.h) */
void sth_get_state(void* p, int env, PgdEnvState* out) {
  HostStep* h = (HostStep*)p;
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

void sth_step(void* p, const float* actions, float* obs, float* reward, uint8_t* done, PgdInfo* info) {
  ((HostStep*)p)->call_index++;
  run((HostStep*)p, 0, actions, obs, reward, done, info);
}
}
