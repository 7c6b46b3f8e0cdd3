// Fused environment step for sm_100a: one launch advances every environment by one decision step.
//
// Thread mapping: V threads (V = vehicle slots per env, 16 or 32) cooperate on one environment; thread s owns
// vehicle slot s (slot 0 = ego).  A 128-thread CTA therefore holds 8 (V=16) or 4 (V=32) environments.
// State lives in HBM as structure-of-arrays over the flat index env * V + slot (16-byte vectors, so a warp
// reads/writes 512 contiguous bytes per field); per-environment poses are mirrored in shared memory for the
// all-pairs phases (IDM neighbour search, chassis contacts, lidar).  Map tables are read-only and shared by all
// environments on the same seed, i.e. L2-resident.
//
// Phases (reference call stack):
//   A  load state, or copy the episode template when the env is being reset   base_env.py:269-301
//   B  ego action clip + traffic trigger                                       env_input_policy.py:17-26, traffic_manager.py:71-89
//   C  IDM / PID action of every awake traffic vehicle                         idm_policy.py:83-353
//   D  5 physics sub-steps + chassis contact                                   base_engine.py:206-232, collision_callback.py:7-35
//   E  localisation, checkpoints, line / sidewalk contacts                     navigation.py:155-344, base_vehicle.py:615-644
//   F  observation (state, navi, 4 neighbours, 240-beam lidar), reward, done   state_obs.py:58-170, pgdrive_env.py:162-258
//   G  store state
#include <cuda_runtime.h>
#include <limits.h>
#include <math_constants.h>
#include <stdint.h>
#include <string.h>

#include <string>

#include "../../include/pgd_math.h"
#include "../../include/pgdrive_b200.h"
#include "pgd_internal.h"

#define PI_F 3.14159265358979323846f
#define TWO_PI_F 6.28318530717958647692f
#define GRAVITY 9.81f
#define LIDAR_RANGE 50.0f
#define MAX_SPEED_KMH 80.0f
#define IDM_MAX_LONG 30.0f
#define IDM_NORMAL_SPEED 30.0f
#define IDM_CREEP_SPEED 5.0f
#define IDM_SAFE_DIST 15.0f
#define IDM_LANE_CHANGE_FREQ 50
#define IDM_SPEED_INCREASE 10.0f
#define IDM_MAX_SPEED 100.0f
#define YAW_TAU 0.1f

#ifndef CTA_THREADS
#define CTA_THREADS 128
#endif
// Phase barriers keep the warps of a CTA inside the same code region, so one instruction-cache fill serves all
// of them (the kernel body is larger than the 32 KB L1.5 instruction cache).
#ifndef PHASE_BARRIERS
#define PHASE_BARRIERS 1
#endif
#if PHASE_BARRIERS
#define PHASE_SYNC() __syncthreads()
#else
#define PHASE_SYNC()
#endif
#ifndef MIN_CTAS_PER_SM
#define MIN_CTAS_PER_SM 4
#endif
#define DONE_PENDING_RESET 2
#define PGD_OBS_CAP_DET (2 * PGD_MAX_DETECTOR_BEAMS + 6 + 10 + 16 + PGD_LIDAR_BEAMS)  /* 752 */

// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float clipf(float a, float lo, float hi) { return fminf(fmaxf(a, lo), hi); }

__device__ __noinline__ float2 sincos2(float a) {  // (sin, cos); single out-of-line copy of the accurate sincosf
  float s, c;
  pgd_sincosf(a, &s, &c);
  return make_float2(s, c);
}
#define SINCOS(angle, s_out, c_out)          \
  do {                                       \
    const float2 sc_ = sincos2(angle);       \
    (s_out) = sc_.x;                         \
    (c_out) = sc_.y;                         \
  } while (0)

// Out-of-line on purpose: fmodf / atan2f / sincosf expand to hundreds of instructions each and were inlined at a
// dozen call sites; the kernel then missed the instruction cache for 64 % of its issue slots (profiles/r01d).
__device__ __noinline__ float wrap_to_pi(float x) {
  return pgd_wrap_to_pi(x);
}

struct Lane {  // registers copy of a PgdLane (4 x 16 B loads)
  float sx, sy, ex, ey, ax, ay, length, width, radius, ph0, dir, heading;
  int road, idx, kind;
};

__device__ __forceinline__ Lane load_lane(const PgdLane* p) {
  const float4* q = reinterpret_cast<const float4*>(p);
  float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
  int4 d = __ldg(reinterpret_cast<const int4*>(q + 3));
  Lane l;
  l.sx = a.x; l.sy = a.y; l.ex = a.z; l.ey = a.w;
  l.ax = b.x; l.ay = b.y; l.length = b.z; l.width = b.w;
  l.radius = c.x; l.ph0 = c.y; l.dir = c.z; l.heading = c.w;
  l.road = d.x; l.idx = d.y; l.kind = d.z;
  return l;
}

__device__ __noinline__ float2 arc_local(float cx, float cy, float ph0, float dir, float radius, float x, float y) {
  float dx = x - cx, dy = y - cy;
  float phi = pgd_atan2f(dy, dx);
  phi = ph0 + wrap_to_pi(phi - ph0);
  float r = sqrtf(dx * dx + dy * dy);
  return make_float2(dir * (phi - ph0) * radius, dir * (radius - r));
}

__device__ __forceinline__ void lane_local(const Lane& l, float x, float y, float& lon, float& lat) {
  if (l.kind == PGD_LANE_STRAIGHT) {
    float dx = x - l.sx, dy = y - l.sy;
    lon = dx * l.ax + dy * l.ay;
    lat = dx * -l.ay + dy * l.ax;
  } else {
    const float2 r = arc_local(l.ax, l.ay, l.ph0, l.dir, l.radius, x, y);
    lon = r.x;
    lat = r.y;
  }
}

__device__ __forceinline__ void lane_position(const Lane& l, float lon, float lat, float& x, float& y) {
  if (l.kind == PGD_LANE_STRAIGHT) {
    x = l.sx + lon * l.ax + lat * -l.ay;
    y = l.sy + lon * l.ay + lat * l.ax;
  } else {
    float phi = l.dir * lon / l.radius + l.ph0;
    float r = l.radius - lat * l.dir;
    float s, c;
    SINCOS(phi, s, c);
    x = l.ax + r * c;
    y = l.ay + r * s;
  }
}

__device__ __forceinline__ float lane_heading_at(const Lane& l, float lon) {
  if (l.kind == PGD_LANE_STRAIGHT) return l.heading;
  float phi = l.dir * lon / l.radius + l.ph0;
  return phi + PI_F / 2 * l.dir;
}

__device__ __forceinline__ bool precedes(float ex, float ey, float sx, float sy) {
  float dx = ex - sx, dy = ey - sy;
  return dx * dx + dy * dy < 1e-2f;  // norm < 0.1 (abs_lane.py:114-119); lane ends either coincide or are metres apart
}

struct Rect {
  float cx, cy, ux, uy, hl, hw;
};

__device__ __forceinline__ bool rect_overlap(const Rect& a, const Rect& b) {
  float dx = b.cx - a.cx, dy = b.cy - a.cy;
  float c = fabsf(a.ux * b.ux + a.uy * b.uy);
  float s = fabsf(a.ux * b.uy - a.uy * b.ux);
  if (fabsf(dx * a.ux + dy * a.uy) > a.hl + b.hl * c + b.hw * s) return false;
  if (fabsf(-dx * a.uy + dy * a.ux) > a.hw + b.hl * s + b.hw * c) return false;
  if (fabsf(dx * b.ux + dy * b.uy) > b.hl + a.hl * c + a.hw * s) return false;
  if (fabsf(-dx * b.uy + dy * b.ux) > b.hw + a.hl * s + a.hw * c) return false;
  return true;
}

__device__ __forceinline__ float ray_rect(float ox, float oy, float dx, float dy, const Rect& r) {
  float px = ox - r.cx, py = oy - r.cy;
  float lo0 = px * r.ux + py * r.uy, lo1 = -px * r.uy + py * r.ux;
  float ld0 = dx * r.ux + dy * r.uy, ld1 = -dx * r.uy + dy * r.ux;
  if (fabsf(lo0) <= r.hl && fabsf(lo1) <= r.hw) return 1.0f;  // origin inside: Bullet's convex cast reports no hit
  float t0 = 0.0f, t1 = 1.0f;
  if (fabsf(ld0) < 1e-12f) {
    if (fabsf(lo0) > r.hl) return 1.0f;
  } else {
    float inv = 1.0f / ld0;
    float ta = (-r.hl - lo0) * inv, tb = (r.hl - lo0) * inv;
    t0 = fmaxf(t0, fminf(ta, tb));
    t1 = fminf(t1, fmaxf(ta, tb));
    if (t0 > t1) return 1.0f;
  }
  if (fabsf(ld1) < 1e-12f) {
    if (fabsf(lo1) > r.hw) return 1.0f;
  } else {
    float inv = 1.0f / ld1;
    float ta = (-r.hw - lo1) * inv, tb = (r.hw - lo1) * inv;
    t0 = fmaxf(t0, fminf(ta, tb));
    t1 = fminf(t1, fmaxf(ta, tb));
    if (t0 > t1) return 1.0f;
  }
  return t0;
}

__device__ __forceinline__ void project(float hx, float hy, float vx, float vy, float& fwd, float& side) {
  const float n = 1.0f + 1e-6f;
  fwd = (vx * hx + vy * hy) / n;
  side = (vx * -hy + vy * hx) / n;
}

__device__ __forceinline__ float pid(float& p_err, float& i_err, float kp, float ki, float kd, float err) {
  i_err += err;
  float d = err - p_err;
  p_err = err;
  return -kp * p_err - ki * i_err - kd * d;
}

// Per-environment shared mirror of what the all-pairs phases need from every slot.
template <int V, int OBS_CAP>
struct EnvShared {
  float x[V], y[V], h[V], v[V];    // pose at the start of the step (IDM) / current (contacts, lidar)
  float ux[V], uy[V];              // heading unit vector
  float hl[V], hw[V];              // chassis half extents
  int lane[V];
  int alive[V];
  float sx[V], sy[V], ex[V], ey[V], llen[V];  // start / end / length of the lane each vehicle is on
  float olong[V];                  // longitudinal coordinate of each vehicle on its own lane (start of step)
  int blo[V], bn[V];               // lidar: first beam index and beam count each chassis can intersect
  int croad[V], nroad[V];          // localisation: current / next route road of each moving vehicle
  int qlane[8];                    // ego queries: lane ids
  float qlon[8], qlat[8];          // ego queries: Frenet results
  float d2[V];                     // squared centre distance to the ego (neighbour ranking)
  int ired[4];
};

struct Sub {  // what one physics sub-step needs, hoisted out of the sub-step loop
  float accel;      // >0: engine acceleration [m/s^2]; else brake
  float brake_dv;   // speed removed per sub-step when braking
  float sb;         // sin(slip angle) of the kinematic bicycle
  float mu_g;
  float lr;
};

// ------------------------------------------------------------------------------------------------------------------
template <int V, int OBS_CAP>
__global__ void __launch_bounds__(CTA_THREADS, MIN_CTAS_PER_SM)
pgd_step_kernel(DevTables T, DevState S, PgdConfig cfg, int mode, int env_begin, int env_end,
                const float2* __restrict__ actions,
                float* __restrict__ obs, float* __restrict__ reward, uint8_t* __restrict__ done,
                PgdInfo* __restrict__ info) {
  constexpr int ENVS_PER_CTA = CTA_THREADS / V;
  __shared__ EnvShared<V, OBS_CAP> sh_all[ENVS_PER_CTA];
  // The observation rows of the CTA's environments are assembled back to back (stride = the row length), i.e. in
  // the layout they have in HBM, and leave with ONE bulk (TMA) shared -> global copy per CTA.
  __shared__ __align__(16) float obs_rows[ENVS_PER_CTA * OBS_CAP];
  const int slot = threadIdx.x % V;
  const int env_in_cta = threadIdx.x / V;
  const int env_raw = env_begin + blockIdx.x * ENVS_PER_CTA + env_in_cta;  // this launch covers [env_begin, env_end)
  const bool env_valid = env_raw < env_end;
  const int env = env_valid ? env_raw : env_end - 1;  // clamp: surplus threads shadow the last env, no stores
  EnvShared<V, OBS_CAP>& sh = sh_all[env_in_cta];
  const unsigned lane_id = threadIdx.x & 31;
  const unsigned group_mask = (V == 32) ? 0xffffffffu : (0xffffu << (lane_id & 16));
  const int gi = env * V + slot;

  // ---- phase A: load ------------------------------------------------------------------------------------------
  int4 envi = S.envi[env];
  float4 envf = S.envf[env];
  const bool pending = envi.z == DONE_PENDING_RESET;
  bool fresh;  // this call (re)starts the episode instead of stepping it
  if (mode == 1) {
    fresh = pending;
  } else {
    fresh = pending || (cfg.auto_reset && envi.z == 1);
  }
  const bool skip = !env_valid || (mode == 1 && !pending);  // nothing is written for skipped envs

  const PgdEpisode* ep = T.episodes + envi.x;
  const int map_id = __ldg(&ep->map);
  const int n_slots = __ldg(&ep->n_slots);
  const int n_groups = __ldg(&ep->n_groups);
  const PgdMap mp = T.maps[map_id];
  const PgdLane* lanes = T.lanes + mp.lane_off;
  const PgdRoad* roads = T.roads + mp.road_off;
  const PgdBox* boxes = T.boxes + mp.box_off;
  const bool has_slot = slot < n_slots;
  const PgdSlot* tpl = T.slots + __ldg(&ep->slot_off) + (has_slot ? slot : 0);
  // template constants of this slot
  const float4 t0 = __ldg(reinterpret_cast<const float4*>(tpl));       // x, y, heading, length
  const float4 t1 = __ldg(reinterpret_cast<const float4*>(tpl) + 1);   // width, mass, lf, lr
  const float4 t2 = __ldg(reinterpret_cast<const float4*>(tpl) + 2);   // max_engine, max_brake, max_steer, friction
  const int4 t3 = __ldg(reinterpret_cast<const int4*>(tpl) + 3);       // lane, type, group, drop_substeps
  const int4 t4 = __ldg(reinterpret_cast<const int4*>(tpl) + 4);       // overtake_timer, route_off, route_len, pad
  const int route_off = t4.y, route_len = t4.z;
  const int32_t* rnodes = T.route_nodes + route_off;
  const int32_t* rroads = T.route_roads + route_off;

  float x, y, h, v, yaw_rate, steer, throttle, hp, hi, lp, li, target_speed;
  int lane, ck0, ck1, rt_lane, timer, rnd_n, airborne, vflags;
  if (fresh) {
    x = t0.x; y = t0.y; h = t0.z; v = 0.0f; yaw_rate = 0.0f;
    steer = throttle = hp = hi = lp = li = 0.0f;
    target_speed = IDM_NORMAL_SPEED;
    lane = t3.x; ck0 = 0; ck1 = route_len > 2 ? 1 : 0; rt_lane = -1;
    timer = t4.x; rnd_n = 0; airborne = t3.w;
    vflags = has_slot ? (PGD_V_ALIVE | PGD_V_ON_LANE | (slot == 0 ? PGD_V_ACTIVE : 0)) : 0;
    envi.y = 0; envi.z = 0; envi.w = 0;
    envf = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    float4 p = S.pose[gi], c = S.ctrl[gi], q = S.pidl[gi];
    int4 n = S.nav[gi], m = S.misc[gi];
    x = p.x; y = p.y; h = p.z; v = p.w;
    steer = c.x; throttle = c.y; hp = c.z; hi = c.w;
    lp = q.x; li = q.y; target_speed = q.z; yaw_rate = q.w;
    lane = n.x; ck0 = n.y & 0xffff; ck1 = n.y >> 16; rt_lane = n.z; timer = n.w;
    rnd_n = m.x; airborne = m.y; vflags = m.z;
  }
  const float half_l = t0.w * 0.5f, half_w = t1.x * 0.5f;
  bool alive = (vflags & PGD_V_ALIVE) != 0;
  bool active = (vflags & PGD_V_ACTIVE) != 0;

  // publish start-of-step poses; (hc, hs) = heading unit vector, kept current through the sub-steps
  float hs, hc;
  {
    float s, c;
    SINCOS(h, s, c);
    hs = s; hc = c;
    sh.x[slot] = x; sh.y[slot] = y; sh.h[slot] = h; sh.v[slot] = v;
    sh.ux[slot] = c; sh.uy[slot] = s;
    sh.hl[slot] = half_l; sh.hw[slot] = half_w;
    sh.lane[slot] = lane; sh.alive[slot] = alive;
  }
  __syncwarp(group_mask);

  const float last_x = sh.x[0], last_y = sh.y[0], last_h = sh.h[0];
  int crash = 0;

  const bool stepping = !fresh && !skip;
  PHASE_SYNC();
  if (stepping) {
    // ---- phase B: ego action + traffic trigger --------------------------------------------------------------
    if (slot == 0) {
      float2 a = actions[env];
      envf.x = steer;  // last_current_action[0] after the push
      envf.y = throttle;
      steer = clipf(a.x, -1.0f, 1.0f);  // fminf/fmaxf drop NaN -> -1, like the compiled cutils_clip
      throttle = clipf(a.y, -1.0f, 1.0f);
    }
    if (envi.y < n_groups) {
      const int ego_road = __ldg(&lanes[sh.lane[0]].road);
      if (ego_road == __ldg(&ep->trigger_road[envi.y])) {
        if (has_slot && t3.z == envi.y) active = true;
        envi.y += 1;
      }
    }

    // what the IDM neighbour search reads from every vehicle -- the geometry of the lane it is on and its
    // longitudinal coordinate there -- is only needed when some traffic vehicle of this environment is awake
    if (__ballot_sync(group_mask, alive && active && slot != 0) & group_mask) {
      if (alive) {
        const Lane l = load_lane(lanes + lane);
        sh.sx[slot] = l.sx; sh.sy[slot] = l.sy; sh.ex[slot] = l.ex; sh.ey[slot] = l.ey; sh.llen[slot] = l.length;
        float lon, lat;
        lane_local(l, x, y, lon, lat);
        sh.olong[slot] = lon;
      }
      __syncwarp(group_mask);
    }
  }
  PHASE_SYNC();
  if (stepping) {
    // ---- phase C: IDM ------------------------------------------------------------------------------------------
    if (alive && active && slot != 0) {
      const int cur_road_id = __ldg(&rroads[ck0]);
      const PgdRoad cur_road = roads[cur_road_id];
      bool ok;
      if (rt_lane < 0) {
        rt_lane = lane;
        ok = __ldg(&lanes[rt_lane].road) == cur_road_id;
      } else if (__ldg(&lanes[rt_lane].road) != cur_road_id) {
        ok = false;
        const float rex = __ldg(&lanes[rt_lane].ex), rey = __ldg(&lanes[rt_lane].ey);
        for (int k = 0; k < cur_road.n_lanes; ++k) {
          const PgdLane* c = lanes + cur_road.first_lane + k;
          if (precedes(rex, rey, __ldg(&c->sx), __ldg(&c->sy))) {
            rt_lane = cur_road.first_lane + k;
            ok = true;
            break;
          }
        }
      } else if (__ldg(&lanes[lane].road) == cur_road_id && rt_lane != lane) {
        rt_lane = lane;
        timer = reinterpret_cast<const uint8_t*>(tpl)[80 + (rnd_n % PGD_N_RND25)];
        rnd_n++;
        ok = true;
      } else {
        ok = true;
      }
      // front / back search on the routing lane and (when routed) its two neighbours
      const Lane rl = load_lane(lanes + rt_lane);
      int cand[3] = {-1, rt_lane, -1};
      if (ok) {
        const PgdRoad rr = roads[rl.road];
        if (rl.idx > 0) cand[0] = rr.first_lane + rl.idx - 1;
        if (rl.idx + 1 < rr.n_lanes) cand[2] = rr.first_lane + rl.idx + 1;
      }
      int front[3], back[3];
      float fdist[3], bdist[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        front[i] = back[i] = -1;
        fdist[i] = bdist[i] = IDM_MAX_LONG;
        if (cand[i] < 0) continue;
        const Lane l = (i == 1) ? rl : load_lane(lanes + cand[i]);
        float cur_long, lat;
        lane_local(l, x, y, cur_long, lat);
        const float left_long = l.length - cur_long;
        bool found_front = false, found_back = false;
        for (int j = 0; j < n_slots; ++j) {
          if (j == slot || !sh.alive[j]) continue;
          const float ox = sh.x[j], oy = sh.y[j];
          const float ddx = ox - x, ddy = oy - y;
          if (!(ddx * ddx + ddy * ddy < LIDAR_RANGE * LIDAR_RANGE)) continue;
          // every branch needs only the neighbour's longitudinal coordinate ON ITS OWN LANE (published by its
          // thread): same lane -> directly; following lane -> + what is left of mine; preceding lane -> its remainder
          if (sh.lane[j] == cand[i]) {
            const float lg = sh.olong[j] - cur_long;
            if (fdist[i] > lg && lg > 0.0f) { fdist[i] = lg; front[i] = j; found_front = true; }
            if (lg < 0.0f && fabsf(lg) < bdist[i]) { bdist[i] = fabsf(lg); back[i] = j; found_back = true; }
          } else if (!found_front && precedes(l.ex, l.ey, sh.sx[j], sh.sy[j])) {
            const float lg = sh.olong[j] + left_long;
            if (fdist[i] > lg && lg > 0.0f) { fdist[i] = lg; front[i] = j; }
          } else if (!found_back && precedes(sh.ex[j], sh.ey[j], l.sx, l.sy)) {
            const float lg = sh.llen[j] - sh.olong[j] + cur_long;
            if (bdist[i] > lg) { bdist[i] = lg; back[i] = j; }
          }
        }
      }
      int front_obj = front[1], steer_lane = rt_lane;
      float front_dist = fdist[1];
      if (ok) {  // lane_change_policy
        const int n_cur = cur_road.n_lanes;
        int lo = 0, hi_idx = n_cur - 1;
        bool decided = false;
        const int idx = rl.idx;
        if (ck0 != ck1) {
          const PgdRoad nxt = roads[__ldg(&rroads[ck1])];
          const int diff = n_cur - nxt.n_lanes;
          if (diff > 0) {
            const PgdLane* c0 = lanes + cur_road.first_lane;
            const PgdLane* n0 = lanes + nxt.first_lane;
            if (precedes(__ldg(&c0->ex), __ldg(&c0->ey), __ldg(&n0->sx), __ldg(&n0->sy))) {
              lo = 0; hi_idx = nxt.n_lanes - 1;
            } else {
              lo = diff; hi_idx = n_cur - 1;
            }
            if (idx < lo || idx > hi_idx) {
              decided = true;
              const int side = idx > hi_idx ? 0 : 2;
              if (bdist[side] < IDM_SAFE_DIST || fdist[side] < 5.0f) {
                target_speed = IDM_CREEP_SPEED;
              } else {
                target_speed = IDM_NORMAL_SPEED;
                front_obj = front[side];
                front_dist = fdist[side];
                steer_lane = cur_road.first_lane + idx + (side == 0 ? -1 : 1);
              }
            }
          }
        }
        if (!decided) {
          const float my_speed = clipf(v * 3.6f, 0.0f, 100000.0f);
          if (fabsf(my_speed - IDM_NORMAL_SPEED) > 3.0f && front[1] >= 0 &&
              fabsf(clipf(sh.v[front[1]] * 3.6f, 0.0f, 100000.0f) - IDM_NORMAL_SPEED) > 3.0f &&
              timer > IDM_LANE_CHANGE_FREQ) {
            float side_speed[3] = {0.f, 0.f, 0.f};
            bool side_ok[3] = {false, false, false};
#pragma unroll
            for (int sd = 0; sd < 3; sd += 2) {
              if (front[sd] >= 0) {
                side_speed[sd] = clipf(sh.v[front[sd]] * 3.6f, 0.0f, 100000.0f);
                side_ok[sd] = true;
              } else if (cand[sd] >= 0 && fdist[sd] > IDM_SAFE_DIST && bdist[sd] > IDM_SAFE_DIST) {
                side_speed[sd] = IDM_MAX_SPEED;
                side_ok[sd] = true;
              }
            }
            const float front_speed = clipf(sh.v[front[1]] * 3.6f, 0.0f, 100000.0f);
            if (side_ok[0] && side_speed[0] - front_speed > IDM_SPEED_INCREASE && idx - 1 >= lo && idx - 1 <= hi_idx) {
              decided = true;
              front_obj = front[0]; front_dist = fdist[0];
              steer_lane = cur_road.first_lane + idx - 1;
            } else if (side_ok[2] && side_speed[2] - front_speed > IDM_SPEED_INCREASE && idx + 1 >= lo &&
                       idx + 1 <= hi_idx) {
              decided = true;
              front_obj = front[2]; front_dist = fdist[2];
              steer_lane = cur_road.first_lane + idx + 1;
            }
          }
        }
        if (!decided) {
          target_speed = IDM_NORMAL_SPEED;
          timer += 1;
        }
      }
      // steering_control
      {
        const Lane tl = (steer_lane == rt_lane) ? rl : load_lane(lanes + steer_lane);
        float lon, lat;
        lane_local(tl, x, y, lon, lat);
        const float lane_heading = lane_heading_at(tl, lon + 1.0f);
        float st = pid(hp, hi, 1.7f, 0.01f, 3.5f, wrap_to_pi(lane_heading - h));
        st += pid(lp, li, 0.3f, 0.002f, 0.05f, -lat);
        steer = st;
      }
      // acceleration
      {
        const float sp = clipf(v * 3.6f, 0.0f, 100000.0f);
        float acc = 1.0f - pgd_pow10f(fmaxf(sp, 0.0f) / target_speed);
        if (front_obj >= 0) {
          const float hx = sh.ux[slot], hy = sh.uy[slot];
          const float fs = clipf(sh.v[front_obj] * 3.6f, 0.0f, 100000.0f);
          const float dvx = sp * hx - fs * sh.ux[front_obj], dvy = sp * hy - fs * sh.uy[front_obj];
          const float dv = dvx * hx + dvy * hy;
          const float d_star = 10.0f + sp * 1.5f + sp * dv / (2.0f * sqrtf(5.0f));
          float d = front_dist;
          if (!(fabsf(d) > 1e-2f)) d = d > 0.0f ? 1e-2f : -1e-2f;
          const float ratio = d_star / d;
          acc -= ratio * ratio;
        }
        throttle = acc;
      }
    }

  }
  PHASE_SYNC();
  if (stepping) {
    // ---- phase D: physics sub-steps + chassis contact ----------------------------------------------------------
    Sub sub;
    {
      sub.mu_g = t2.w * GRAVITY;
      sub.lr = t1.w;
      const bool overspeed = clipf(v * 3.6f, 0.0f, 100000.0f) > MAX_SPEED_KMH;
      if (throttle > 0.0f && !overspeed) {
        sub.accel = fminf(4.0f * t2.x * throttle / t1.y, sub.mu_g);
        sub.brake_dv = 0.0f;
      } else {
        sub.accel = 0.0f;
        const float imp = throttle >= 0.0f ? 2.0f : -throttle * t2.y;
        sub.brake_dv = fminf(4.0f * imp / t1.y, sub.mu_g * cfg.dt);
      }
      const float delta = clipf(-steer * t2.z, -1.4f, 1.4f);
      const float tb = t1.w / (t1.z + t1.w) * pgd_tanf(delta);
      sub.sb = tb / sqrtf(1.0f + tb * tb);
    }
    // A vehicle at rest with no yaw rate and no engine force (parked traffic, a braking ego) is a fixed point of the
    // sub-step below: speed = max(0 - dv, 0) = 0, yaw stays 0, the pose does not move.  Skipping it is exact.
    const bool at_rest = v == 0.0f && yaw_rate == 0.0f && !(sub.accel > 0.0f);
    for (int k = 0; k < cfg.decision_repeat; ++k) {
      if (alive) {
        if (airborne > 0) {
          airborne--;
        } else if (!at_rest) {
          float speed = v;
          if (sub.accel > 0.0f) speed += sub.accel * cfg.dt;
          else speed = fmaxf(speed - sub.brake_dv, 0.0f);
          float yaw = yaw_rate + (speed * sub.sb / sub.lr - yaw_rate) * (cfg.dt / YAW_TAU);
          if (speed * fabsf(yaw) > sub.mu_g) yaw = copysignf(sub.mu_g / speed, yaw);
          const float sb = speed > 1e-3f ? clipf(yaw * sub.lr / speed, -1.0f, 1.0f) : 0.0f;
          const float cb = sqrtf(fmaxf(1.0f - sb * sb, 0.0f));
          x += speed * (hc * cb - hs * sb) * cfg.dt;
          y += speed * (hs * cb + hc * sb) * cfg.dt;
          float nh = h + yaw * cfg.dt;
          if (nh > PI_F) nh -= TWO_PI_F;
          if (nh < -PI_F) nh += TWO_PI_F;
          yaw_rate = yaw;
          if (nh != h) SINCOS(nh, hs, hc);
          h = nh;
          v = speed;
        }
      }
      // contact of every chassis with the ego's
      if (slot == 0) { sh.x[0] = x; sh.y[0] = y; sh.ux[0] = hc; sh.uy[0] = hs; }
      __syncwarp(group_mask);
      if (alive && slot != 0) {
        // rectangles whose centres are further apart than their half-diagonals add up to cannot touch
        const float ddx = x - sh.x[0], ddy = y - sh.y[0];
        const float reach = sh.hl[0] + sh.hw[0] + half_l + half_w;
        if (ddx * ddx + ddy * ddy <= reach * reach) {
          Rect me = {x, y, hc, hs, half_l, half_w};
          Rect eg = {sh.x[0], sh.y[0], sh.ux[0], sh.uy[0], sh.hl[0], sh.hw[0]};
          if (rect_overlap(eg, me)) crash = 1;
        }
      }
      __syncwarp(group_mask);
    }
    crash = (__ballot_sync(group_mask, crash) & group_mask) != 0;
  }

  PHASE_SYNC();
  // ---- phase E: after_step -------------------------------------------------------------------------------------
  const bool moving = alive && active;  // vehicles that get an after_step: the ego always, traffic once awake
  // publish end-of-step poses and what localisation needs from each moving vehicle
  sh.x[slot] = x; sh.y[slot] = y; sh.h[slot] = h; sh.v[slot] = v;
  sh.ux[slot] = hc; sh.uy[slot] = hs;
  sh.croad[slot] = moving ? __ldg(&rroads[ck0]) : -1;
  sh.nroad[slot] = (moving && ck0 != ck1) ? __ldg(&rroads[ck1]) : -1;
  __syncwarp(group_mask);
  const float ex_ = sh.x[0], ey_ = sh.y[0], eux = sh.ux[0], euy = sh.uy[0], eh = sh.h[0];
  uint32_t flags = 0;
  {
    // The V threads scan one vehicle's bucket together (entry k = b0 + slot, + V, ...): point-in-rectangle against
    // lane surfaces for localisation and, for the ego, chassis-vs-rectangle for line ghosts and sidewalks.  The
    // candidate kept per category is the one with the lowest table index, as a sequential scan would find first.
    const unsigned need_all = __ballot_sync(group_mask, moving);
    unsigned need = (V == 32) ? need_all : ((need_all >> (lane_id & 16)) & 0xffffu);
    const Rect er = {ex_, ey_, eux, euy, sh.hl[0], sh.hw[0]};
    const int32_t* ent = T.cell_entries + mp.entry_off;
    int my_any = INT_MAX, my_cur = INT_MAX, my_next = INT_MAX;
    while (need) {
      const int t = __ffs(need) - 1;
      need &= need - 1;
      const float vx = sh.x[t], vy = sh.y[t], vc = sh.ux[t], vs = sh.uy[t];
      const int cur_road = sh.croad[t], next_road = sh.nroad[t];
      int b_any = INT_MAX, b_cur = INT_MAX, b_next = INT_MAX;
      const int cx = (int)floorf((vx - mp.x0) * mp.inv_cell), cy = (int)floorf((vy - mp.y0) * mp.inv_cell);
      if (cx >= 0 && cy >= 0 && cx < mp.nx && cy < mp.ny) {
        const int cell = mp.cell_off + cy * mp.nx + cx;
        const int b0 = __ldg(&T.cell_start[cell]), b1 = __ldg(&T.cell_start[cell + 1]);
        for (int k = b0 + slot; k < b1; k += V) {
          const int b = __ldg(&ent[k]);
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(boxes + b));
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(boxes + b) + 1);  // hl, hw, kind, lane
          const int kind = __float_as_int(g1.z);
          if (kind == PGD_BOX_LANE) {
            const float dx = vx - g0.x, dy = vy - g0.y;
            if (!(fabsf(dx * g0.z + dy * g0.w) <= g1.x && fabsf(-dx * g0.w + dy * g0.z) <= g1.y)) continue;
            // keep lanes whose direction at the vehicle makes an acute angle with its heading (scene_utils.py:158-170).
            // The lane direction needs no trigonometry: a straight lane's is its unit vector, an arc's is the tangent
            // dir * (-dy, dx) / r at the vehicle's bearing from the centre (r > 0 does not change the sign).
            const float4* lq = reinterpret_cast<const float4*>(lanes + __float_as_int(g1.w));
            const float4 lb = __ldg(lq + 1);                                  // ax, ay, length, width
            const int4 ld = __ldg(reinterpret_cast<const int4*>(lq + 3));     // road, idx, kind, pad
            float dot;
            if (ld.z == PGD_LANE_STRAIGHT) {
              dot = lb.x * vc + lb.y * vs;
            } else {
              const float ldir = __ldg(reinterpret_cast<const float*>(lq + 2) + 2);
              dot = ldir * ((vx - lb.x) * vs - (vy - lb.y) * vc);
            }
            if (!(dot > 0.0f)) continue;
            b_any = min(b_any, b);
            if (ld.x == cur_road) b_cur = min(b_cur, b);
            if (ld.x == next_road) b_next = min(b_next, b);
          } else if (t == 0) {
            const Rect r = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y};
            if (!rect_overlap(er, r)) continue;
            flags |= kind == PGD_BOX_WHITE ? PGD_F_ON_WHITE
                   : kind == PGD_BOX_YELLOW ? PGD_F_ON_YELLOW
                   : kind == PGD_BOX_BROKEN ? PGD_F_ON_BROKEN : PGD_F_CRASH_SIDEWALK;
          }
        }
      }
#pragma unroll
      for (int o = V / 2; o > 0; o >>= 1) {
        b_any = min(b_any, __shfl_xor_sync(group_mask, b_any, o, V));
        b_cur = min(b_cur, __shfl_xor_sync(group_mask, b_cur, o, V));
        b_next = min(b_next, __shfl_xor_sync(group_mask, b_next, o, V));
      }
      if (slot == t) { my_any = b_any; my_cur = b_cur; my_next = b_next; }
    }
#pragma unroll
    for (int o = V / 2; o > 0; o >>= 1) flags |= __shfl_xor_sync(group_mask, flags, o, V);
    if (moving) {
      const int nb = my_cur != INT_MAX ? my_cur : (my_next != INT_MAX ? my_next : my_any);
      bool on_lane = nb != INT_MAX;
      if (on_lane) lane = __ldg(&boxes[nb].lane);
      if (ck0 != ck1) {  // _update_target_checkpoints
        const Lane l = load_lane(lanes + lane);
        float lon, lat;
        lane_local(l, x, y, lon, lat);
        const int start = __ldg(&roads[l.road].start_node);
        if (lon < 5.0f) {
          for (int j = ck1; j < route_len - 1; ++j) {
            if (__ldg(&rnodes[j]) == start) {
              ck0 = j;
              ck1 = (j + 1 == route_len - 1) ? j : j + 1;
              break;
            }
          }
        }
      }
      vflags = on_lane ? (vflags | PGD_V_ON_LANE) : (vflags & ~PGD_V_ON_LANE);
      if (slot != 0 && !on_lane) alive = false;  // traffic_manager.py:91-109
    }
  }
  sh.alive[slot] = alive;
  __syncwarp(group_mask);

  PHASE_SYNC();
  // ---- phase F: observation, reward, done -----------------------------------------------------------------------
  // Staged row; copied to obs[env] at the end of phase F.  Layout (obs/state_obs.py): [side beams | left, right],
  // 6 state values, [lane-line beams], 10 navi, 16 neighbours, 240 lidar.  `ob` is positioned so that the indices of
  // the detector-less layout (state 2..7, navi 8..17, neighbours 18..33, lidar 34..) address the tail of the row.
  const int n_first = cfg.n_side > 0 ? cfg.n_side : 2;
  const int obs_dim = n_first + 6 + cfg.n_lane_line + 10 + 16 + PGD_LIDAR_BEAMS;
  float* const row = obs_rows + env_in_cta * obs_dim;
  float* const st = row + n_first - 2;                   // st[2..7]  = state values
  float* const ob = row + n_first + cfg.n_lane_line - 2;  // ob[8..]   = navi, neighbours, lidar
  // lidar: beam i = slot + V * k.  Each chassis first publishes the (conservative) arc of beams that can reach it:
  // it lies inside the disc of radius half-diagonal around its centre, so only beams within asin(hd / d) of its
  // bearing and only chassis closer than 50 m + hd matter.  The cull never changes a result (the reference's own
  // angular mask has the same property, tests/test_component/test_detector_mask.py:308-311).
  {
    int blo = 0, bn = -1;
    if (alive && slot != 0) {
      const float dx = x - ex_, dy = y - ey_;
      const float d2 = dx * dx + dy * dy;
      const float hd = sqrtf(half_l * half_l + half_w * half_w);
      const float reach = LIDAR_RANGE + hd;
      if (d2 < reach * reach) {
        const float d = sqrtf(d2);
        bn = PGD_LIDAR_BEAMS;
        if (d > hd * 1.001f) {
          const float per_rad = (float)PGD_LIDAR_BEAMS / TWO_PI_F;
          const float c = (pgd_atan2f(dy, dx) - eh) * per_rad;
          const float w = pgd_asinf(fminf(hd / d, 1.0f)) * per_rad;
          const int n = (int)ceilf(2.0f * w) + 3;
          if (n < PGD_LIDAR_BEAMS) {
            bn = n;
            blo = ((int)floorf(c - w) - 1) % PGD_LIDAR_BEAMS;
            if (blo < 0) blo += PGD_LIDAR_BEAMS;
          }
        }
      }
    }
    sh.blo[slot] = blo;
    sh.bn[slot] = bn;
    __syncwarp(group_mask);
    const unsigned near_all = __ballot_sync(group_mask, bn >= 0);
    const unsigned near = (V == 32) ? near_all : ((near_all >> (lane_id & 16)) & 0xffffu);
    if (!skip) {
      for (int i = slot; i < PGD_LIDAR_BEAMS; i += V) {
        float best = 1.0f;
        unsigned m = near;
        if (m) {
          const float ang = (float)i * (TWO_PI_F / (float)PGD_LIDAR_BEAMS) + eh;
          float s, c;
          SINCOS(ang, s, c);
          const float dx = c * LIDAR_RANGE, dy = s * LIDAR_RANGE;
          while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            int rel = i - sh.blo[j];
            if (rel < 0) rel += PGD_LIDAR_BEAMS;
            if (rel > sh.bn[j]) continue;
            const Rect r = {sh.x[j], sh.y[j], sh.ux[j], sh.uy[j], sh.hl[j], sh.hw[j]};
            best = fminf(best, ray_rect(ex_, ey_, dx, dy, r));
          }
        }
        ob[34 + i] = best;
      }
    }
  }
  // side / lane-line detectors (distance_detector.py:137-152): ray fans against the line ghosts of the map.  A beam
  // looks up the bucket of a point every 8 m along itself; buckets list every box within 4 m of them, so each ghost
  // the beam can touch is found in at least one of them (found twice is harmless: the result is a minimum).
  if (OBS_CAP > PGD_OBS_DIM && !skip) {
    const int n_rays = cfg.n_side + cfg.n_lane_line;
    const int32_t* ent = T.cell_entries + mp.entry_off;
    for (int rI = slot; rI < n_rays; rI += V) {
      const bool side = rI < cfg.n_side;
      const int i = side ? rI : rI - cfg.n_side;
      const int n = side ? cfg.n_side : cfg.n_lane_line;
      const float dist = side ? cfg.side_distance : cfg.lane_line_distance;
      const float ang = (float)i * (TWO_PI_F / (float)n) + PI_F / 2 + eh;
      float sn, cs;
      SINCOS(ang, sn, cs);
      const float dx = cs * dist, dy = sn * dist;
      float best = 1.0f;
      for (float sd = 4.0f; sd - 4.0f < dist; sd += 8.0f) {
        if (best * dist < sd - 4.0f) break;  // already hit something nearer than what the remaining buckets cover
        const float px = ex_ + cs * sd, py = ey_ + sn * sd;
        const int cx = (int)floorf((px - mp.x0) * mp.inv_cell), cy = (int)floorf((py - mp.y0) * mp.inv_cell);
        if (cx < 0 || cy < 0 || cx >= mp.nx || cy >= mp.ny) continue;
        const int cell = mp.cell_off + cy * mp.nx + cx;
        const int b0 = __ldg(&T.cell_start[cell]), b1 = __ldg(&T.cell_start[cell + 1]);
        for (int k = b0; k < b1; ++k) {
          const int b = __ldg(&ent[k]);
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(boxes + b) + 1);  // hl, hw, kind, lane
          const int kind = __float_as_int(g1.z);
          if (!(kind == PGD_BOX_WHITE || kind == PGD_BOX_YELLOW || (!side && kind == PGD_BOX_BROKEN))) continue;
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(boxes + b));
          const Rect r = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y};
          best = fminf(best, ray_rect(ex_, ey_, dx, dy, r));
        }
      }
      if (side) row[i] = best;
      else row[n_first + 6 + i] = best;
    }
  }

  PHASE_SYNC();
  // the 4 nearest vehicles inside the 50 m cylinder (lidar.py:55-77): every chassis ranks itself by centre distance
  // (ties -> lower slot, like a stable selection) and the 4 best write their own features
  {
    float myd2 = CUDART_INF_F;
    if (alive && slot != 0) {
      const float dx = x - ex_, dy = y - ey_;
      const float d2 = dx * dx + dy * dy;
      if (d2 < LIDAR_RANGE * LIDAR_RANGE) myd2 = d2;
    }
    sh.d2[slot] = myd2;
    __syncwarp(group_mask);
    const bool in_range = myd2 < CUDART_INF_F;
    const unsigned near_all = __ballot_sync(group_mask, in_range);
    const int n_near = __popc((V == 32) ? near_all : ((near_all >> (lane_id & 16)) & 0xffffu));
    if (!skip) {
      if (in_range) {
        int rank = 0;
        for (int j = 1; j < V; ++j) {
          const float o = sh.d2[j];
          rank += (o < myd2 || (o == myd2 && j < slot)) ? 1 : 0;
        }
        if (rank < 4) {
          const float esp = clipf(sh.v[0] * 3.6f, 0.0f, 100000.0f);
          float pf, ps, vf, vs;
          project(eux, euy, x - ex_, y - ey_, pf, ps);
          const float ws = clipf(v * 3.6f, 0.0f, 100000.0f);
          project(eux, euy, ws * hc - esp * eux, ws * hs - esp * euy, vf, vs);
          float4 q;
          q.x = clipf((pf / LIDAR_RANGE + 1.0f) / 2.0f, 0.0f, 1.0f);
          q.y = clipf((ps / LIDAR_RANGE + 1.0f) / 2.0f, 0.0f, 1.0f);
          q.z = clipf((vf / MAX_SPEED_KMH + 1.0f) / 2.0f, 0.0f, 1.0f);
          q.w = clipf((vs / MAX_SPEED_KMH + 1.0f) / 2.0f, 0.0f, 1.0f);
          float* o4 = ob + 18 + 4 * rank;
          o4[0] = q.x; o4[1] = q.y; o4[2] = q.z; o4[3] = q.w;
        }
      }
      if (slot < 4 && slot >= n_near) {
        float* o4 = ob + 18 + 4 * slot;
        o4[0] = o4[1] = o4[2] = o4[3] = 0.0f;
      }
    }
  }
  // Ego bookkeeping.  Its seven independent geometric queries (four Frenet projections, two checkpoint features,
  // the heading difference) are spread over threads 0..6 of the group so that they issue side by side.
  if (slot == 0) {
    const int cur_road_id = __ldg(&rroads[ck0]);
    const PgdRoad cur_road = roads[cur_road_id];
    const PgdRoad fr = roads[__ldg(&rroads[route_len - 2])];
    const int el_road = __ldg(&lanes[lane].road);
    const bool use_ego_lane = el_road == cur_road_id;
    const int reward_lane = use_ego_lane ? lane : cur_road.first_lane;
    sh.qlane[0] = cur_road.first_lane;                 // to_left / to_right
    sh.qlane[1] = fr.first_lane + fr.n_lanes - 1;      // arrive_destination
    sh.qlane[2] = reward_lane;                         // reward: longitudinal of the last position
    sh.qlane[3] = reward_lane;                         //         and of the current one
    sh.qlane[4] = cur_road.first_lane;                 // navigation info of the current road
    sh.qlane[5] = __ldg(&roads[__ldg(&rroads[ck1])].first_lane);  // and of the next one
    sh.qlane[6] = cur_road.first_lane + cur_road.n_lanes - 1;     // heading_diff: right-most reference lane
    sh.ired[0] = cur_road.n_lanes;
    sh.ired[1] = use_ego_lane ? 0 : (__ldg(&roads[el_road].negative) ? -1 : 1);
    sh.ired[2] = cur_road_id;
  }
  __syncwarp(group_mask);
  if (slot < 7 && !skip) {
    const int n_ref = sh.ired[0];
    const Lane l = load_lane(lanes + sh.qlane[slot]);
    if (slot < 4) {
      const bool last = slot == 2;
      float lon, lat;
      lane_local(l, last ? last_x : ex_, last ? last_y : ey_, lon, lat);
      sh.qlon[slot] = lon;
      sh.qlat[slot] = lat;
      if (slot == 1) sh.qlon[4] = l.length;  // final lane length
    } else if (slot < 6) {  // navigation.py:213-260
      const float later_middle = ((float)n_ref / 2.0f - 0.5f) * mp.lane_width;
      float px, py;
      lane_position(l, l.length, later_middle, px, py);
      float dx = px - ex_, dy = py - ey_;
      const float dn = sqrtf(dx * dx + dy * dy);
      if (dn > 50.0f) { dx = dx / dn * 50.0f; dy = dy / dn * 50.0f; }
      float ph, ps;
      project(eux, euy, dx, dy, ph, ps);
      float bend = 0.0f, dir = 0.0f, angle = 0.0f;
      if (l.kind == PGD_LANE_ARC) {
        bend = l.radius / (60.0f + (float)n_ref * mp.lane_width);
        dir = l.dir;
        angle = l.length / l.radius;
      }
      float* q = ob + 8 + 5 * (slot - 4);
      q[0] = clipf((ph / 50.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
      q[1] = clipf((ps / 50.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
      q[2] = clipf(bend, 0.0f, 1.0f);
      q[3] = clipf((dir + 1.0f) / 2.0f, 0.0f, 1.0f);
      q[4] = clipf((angle * (180.0f / PI_F) / 135.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
    } else {  // heading_diff (base_vehicle.py:433-458)
      float lx, ly;
      if (l.kind == PGD_LANE_STRAIGHT) { lx = -l.ay; ly = l.ax; }
      else if (l.dir < 0.0f) { lx = ex_ - l.ax; ly = ey_ - l.ay; }
      else { lx = l.ax - ex_; ly = l.ay - ey_; }
      const float ln = sqrtf(lx * lx + ly * ly);
      st[2] = ln > 0.0f ? clipf((eux * lx + euy * ly) / ln, -1.0f, 1.0f) / 2.0f + 0.5f : 0.0f;
    }
  }
  __syncwarp(group_mask);
  if (slot == 0) {
    const bool on_lane = (vflags & PGD_V_ON_LANE) != 0;
    if (on_lane) flags |= PGD_F_ON_LANE;
    if (crash) flags |= PGD_F_CRASH_VEHICLE;
    const int n_ref = sh.ired[0];
    const float to_left = sh.qlat[0] + mp.lane_width / 2.0f;
    const float to_right = mp.lane_width * (float)n_ref - to_left;
    if (to_left < 0.0f || to_right < 0.0f) flags |= PGD_F_OUT_OF_ROUTE;
    {
      const float lon = sh.qlon[1], lat = sh.qlat[1], flen = sh.qlon[4];
      if (flen - 5.0f < lon && lon < flen + 5.0f && mp.lane_width / 2.0f >= lat &&
          lat >= (0.5f - (float)n_ref) * mp.lane_width)
        flags |= PGD_F_ARRIVE_DEST;
    }
    bool out_of_road = (flags & (PGD_F_ON_YELLOW | PGD_F_ON_WHITE | PGD_F_CRASH_SIDEWALK)) || !on_lane;
    if (cfg.out_of_route_done && (flags & PGD_F_OUT_OF_ROUTE)) out_of_road = true;
    if (out_of_road) flags |= PGD_F_OUT_OF_ROAD;

    const float sp = clipf(v * 3.6f, 0.0f, 100000.0f);
    float o8[8];
    o8[0] = clipf(to_left / 18.0f, 0.0f, 1.0f);
    o8[1] = clipf(to_right / 18.0f, 0.0f, 1.0f);
    o8[3] = clipf((sp + 1.0f) / (MAX_SPEED_KMH + 1.0f), 0.0f, 1.0f);
    o8[4] = clipf((steer / 60.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
    o8[5] = clipf((envf.x + 1.0f) / 2.0f, 0.0f, 1.0f);
    o8[6] = clipf((envf.y + 1.0f) / 2.0f, 0.0f, 1.0f);
    // yaw rate = arccos(clip(cos(heading change), 0, 1)) / 0.1, evaluated as min(|change|, pi/2) (well-conditioned)
    o8[7] = clipf(fminf(fabsf(wrap_to_pi(h - last_h)), PI_F / 2) / 0.1f, 0.0f, 1.0f);
    // reward / cost / done
    float r = 0.0f, step_reward = 0.0f, cost = 0.0f, step_energy = 0.0f;
    int is_done = 0;
    if (!fresh) {
      const float sign = sh.ired[1] == 0 ? 1.0f : (float)sh.ired[1];
      const float long_last = sh.qlon[2], long_now = sh.qlon[3], lat_now = sh.qlat[3];
      float lateral_factor = 1.0f;
      if (cfg.use_lateral) lateral_factor = clipf(1.0f - 2.0f * fabsf(lat_now) / mp.lane_width, 0.0f, 1.0f);
      r += cfg.driving_reward * (long_now - long_last) * lateral_factor * sign;
      r += cfg.speed_reward * (sp / MAX_SPEED_KMH) * sign;
      step_reward = r;
      if (flags & PGD_F_ARRIVE_DEST) r = cfg.success_reward;
      else if (out_of_road) r = -cfg.out_of_road_penalty;
      else if (crash) r = -cfg.crash_vehicle_penalty;
      if (out_of_road) cost = cfg.out_of_road_cost;
      else if (crash) cost = cfg.crash_vehicle_cost;
      is_done = ((flags & PGD_F_ARRIVE_DEST) || out_of_road || crash) ? 1 : 0;
      const float ddx = last_x - x, ddy = last_y - y;
      step_energy = 3.25f * pgd_expf(0.01f * sp) * (sqrtf(ddx * ddx + ddy * ddy) / 1000.0f) / 100.0f * 1000.0f;
      envf.w += step_energy;
      envf.z += r;
      envi.w += 1;
      if (cfg.horizon > 0 && envi.w >= cfg.horizon) { is_done = 1; flags |= PGD_F_MAX_STEP; }
      if (envi.z == 1) is_done = 1;  // sticky
      envi.z = is_done;
    } else {
      flags |= PGD_F_WAS_RESET;
    }
    if (!skip) {
      if (cfg.n_side <= 0) { row[0] = o8[0]; row[1] = o8[1]; }
#pragma unroll
      for (int k = 3; k < 8; ++k) st[k] = o8[k];
      if (mode == 0) {
        reward[env] = r;
        done[env] = (uint8_t)is_done;
      }
      if (info) {
        PgdInfo inf;
        inf.velocity = sp; inf.steering = steer; inf.acceleration = throttle;
        inf.step_energy = step_energy; inf.episode_energy = envf.w;
        inf.step_reward = step_reward; inf.episode_reward = envf.z; inf.cost = cost;
        inf.episode_length = envi.w; inf.flags = flags;
        info[env] = inf;
      }
      S.envi[env] = envi;
      S.envf[env] = envf;
    }
  }

  // Write the staged rows out.  Full CTA (every environment valid and stepped, 16-byte aligned destination): one
  // elected thread issues a single cp.async.bulk of ENVS_PER_CTA rows (8 768 B at V = 16) -- the destination may be
  // local HBM or a peer-mapped gather buffer on another GPU.  Otherwise each group streams its own row with
  // coalesced 8-byte stores.
  const size_t cta_row0 = (size_t)(env_begin + blockIdx.x * ENVS_PER_CTA) * obs_dim;
  const unsigned cta_bytes = (unsigned)(ENVS_PER_CTA * obs_dim * sizeof(float));
  const bool bulk_possible = (cta_bytes % 16u) == 0 && ((reinterpret_cast<uintptr_t>(obs + cta_row0)) & 15u) == 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk copy
  const bool bulk = __syncthreads_and(env_valid && !skip && bulk_possible);
  if (bulk) {
    if (threadIdx.x == 0) {
      const unsigned src = (unsigned)__cvta_generic_to_shared(obs_rows);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   :: "l"(obs + cta_row0), "r"(src), "r"(cta_bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  } else if (!skip) {
    if ((obs_dim & 1) == 0) {
      float2* dst = reinterpret_cast<float2*>(obs + (size_t)env * obs_dim);
      const float2* src2 = reinterpret_cast<const float2*>(row);
#pragma unroll 3
      for (int k = slot; k < obs_dim / 2; k += V) dst[k] = src2[k];
    } else {  // odd row length: rows are only 4-byte aligned
      float* dst = obs + (size_t)env * obs_dim;
      for (int k = slot; k < obs_dim; k += V) dst[k] = row[k];
    }
  }

  // ---- phase G: store --------------------------------------------------------------------------------------------
  if (!skip) {
    vflags = (vflags & PGD_V_ON_LANE) | (alive ? PGD_V_ALIVE : 0) | (active ? PGD_V_ACTIVE : 0);
    S.pose[gi] = make_float4(x, y, h, v);
    S.ctrl[gi] = make_float4(steer, throttle, hp, hi);
    S.pidl[gi] = make_float4(lp, li, target_speed, yaw_rate);
    S.nav[gi] = make_int4(lane, ck0 | (ck1 << 16), rt_lane, timer);
    S.misc[gi] = make_int4(rnd_n, airborne, vflags, 0);
  }
  if (bulk && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

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

// ===================================================================================================================
// host side: C-ABI
// ===================================================================================================================
thread_local std::string g_pgd_err;

extern "C" const char* pgd_last_error(void) { return g_pgd_err.c_str(); }

extern "C" int pgd_create(const PgdConfig* cfg, int device, PgdHandle** out) {
  if (!cfg || !out) return fail(-1, "pgd_create: null argument");
  if (cfg->num_envs <= 0) return fail(-1, "pgd_create: num_envs must be positive");
  if (cfg->layout < 0 || cfg->layout > 2) return fail(-1, "pgd_create: layout must be 0, 1 or 2");
  if (cfg->num_slots != 16 && cfg->num_slots != 32 && !(cfg->num_slots == 24 && cfg->layout == 2))
    return fail(-1, "pgd_create: num_slots must be 16, 24 (layout 2) or 32");
  if (cfg->random_agent_model && cfg->layout == 0)
    return fail(-3, "pgd_create: random_agent_model is not in the cooperative layout (layout = 0)");
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
  cudaFreeHost(h->h_act); cudaFreeHost(h->h_obs); cudaFreeHost(h->h_rew); cudaFreeHost(h->h_done);
  cudaFreeHost(h->h_info);
  cudaFree(h->d_act); cudaFree(h->d_obs); cudaFree(h->d_rew); cudaFree(h->d_done); cudaFree(h->d_info);
  cudaStreamDestroy(h->own_stream);
  cudaStreamDestroy(h->own_stream2);
  cudaEventDestroy(h->ev_act);
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
  if (h->cfg.layout == 1) return pgd_launch_step_v2(h, mode, env_begin, env_end, actions, obs, reward, done, info, st);
  if (h->cfg.layout == 2) return pgd_launch_step_v3(h, mode, env_begin, env_end, actions, obs, reward, done, info, st);
  const int V = h->cfg.num_slots;
  const int envs_per_cta = CTA_THREADS / V;
  const int grid = (env_end - env_begin + envs_per_cta - 1) / envs_per_cta;
  if (h->timing && mode == 0) cudaEventRecord(h->ev0, st);
  const bool det = h->cfg.n_side > 0 || h->cfg.n_lane_line > 0;  // detectors need the large staged row
#define PGD_LAUNCH(VV, CAP)                                                                                        \
  pgd_step_kernel<VV, CAP><<<grid, CTA_THREADS, 0, st>>>(h->T, h->S, h->cfg, mode, env_begin, env_end,            \
                                                          (const float2*)actions, obs, reward, done, info)
  if (V == 16 && !det) PGD_LAUNCH(16, PGD_OBS_DIM);
  else if (V == 32 && !det) PGD_LAUNCH(32, PGD_OBS_DIM);
#if CTA_THREADS <= 128
  else if (V == 16) PGD_LAUNCH(16, PGD_OBS_CAP_DET);
  else PGD_LAUNCH(32, PGD_OBS_CAP_DET);
#else  // experiment builds with larger CTAs: the detector rows do not fit static shared memory
  else return fail(-3, "this build has no side / lane-line detector kernels");
#endif
#undef PGD_LAUNCH
  if (h->timing && mode == 0) cudaEventRecord(h->ev1, st);
  h->launches++;
  CU(cudaGetLastError());
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
  return launch_step(h, 1, 0, h->cfg.num_envs, nullptr, obs_dev, nullptr, nullptr, info_dev, st);
}

extern "C" int pgd_step(PgdHandle* h, const float* actions_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
                        PgdInfo* info_dev, void* stream) {
  if (!h || !actions_dev || !obs_dev || !reward_dev || !done_dev) return fail(-1, "pgd_step: null argument");
  if (!h->tables_loaded) return fail(-3, "pgd_step: no tables loaded");
  CU(cudaSetDevice(h->device));
  return launch_step(h, 0, 0, h->cfg.num_envs, actions_dev, obs_dev, reward_dev, done_dev, info_dev,
                     (cudaStream_t)stream);
}

static int ensure_staging(PgdHandle* h) {
  if (h->h_act) return 0;
  const size_t n = (size_t)h->cfg.num_envs;
  CU(cudaMallocHost(&h->h_act, n * 8));
  const size_t od = (size_t)pgd_obs_dim(&h->cfg);
  CU(cudaMallocHost(&h->h_obs, n * od * 4));
  CU(cudaMallocHost(&h->h_rew, n * 4));
  CU(cudaMallocHost(&h->h_done, n));
  CU(cudaMallocHost(&h->h_info, n * sizeof(PgdInfo)));
  CU(cudaMalloc(&h->d_act, n * 8));
  CU(cudaMalloc(&h->d_obs, n * od * 4));
  CU(cudaMalloc(&h->d_rew, n * 4));
  CU(cudaMalloc(&h->d_done, n));
  CU(cudaMalloc(&h->d_info, n * sizeof(PgdInfo)));
  return 0;
}

static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

extern "C" int pgd_step_host(PgdHandle* h, const float* actions, float* obs, float* reward, uint8_t* done,
                             PgdInfo* info) {
  if (!h || !actions || !obs || !reward || !done) return fail(-1, "pgd_step_host: null argument");
  if (!h->tables_loaded) return fail(-3, "pgd_step_host: no tables loaded");
  CU(cudaSetDevice(h->device));
  int rc = ensure_staging(h);
  if (rc) return rc;
  const size_t n = (size_t)h->cfg.num_envs;
  const size_t od = (size_t)pgd_obs_dim(&h->cfg);
  cudaStream_t st = h->own_stream;
  // Page-locked caller buffers are DMA targets themselves; pageable ones go through the handle's pinned staging.
  const bool direct = is_pinned(obs) && is_pinned(reward) && is_pinned(done) && (!info || is_pinned(info));
  float* o_dst = direct ? obs : h->h_obs;
  float* r_dst = direct ? reward : h->h_rew;
  uint8_t* d_dst = direct ? done : h->h_done;
  PgdInfo* i_dst = direct ? info : h->h_info;
  memcpy(h->h_act, actions, n * 8);
  CU(cudaMemcpyAsync(h->d_act, h->h_act, n * 8, cudaMemcpyHostToDevice, st));
  // The step is cut into chunks of environments on two streams so that the device-to-host copy of one chunk (the
  // PCIe-bound part: 1.1 KB per environment) overlaps the kernel of the next.
  const int chunks = n >= 8192 ? 4 : 1;
  if (chunks > 1) {
    CU(cudaEventRecord(h->ev_act, st));
    CU(cudaStreamWaitEvent(h->own_stream2, h->ev_act, 0));
  }
  const int per = (int)((n / chunks + 31) / 32 * 32);
  for (int c = 0; c < chunks; ++c) {
    const int b = c * per, e = (c == chunks - 1) ? (int)n : (c + 1) * per;
    cudaStream_t cs = (c & 1) ? h->own_stream2 : st;
    rc = launch_step(h, 0, b, e, h->d_act, h->d_obs, h->d_rew, h->d_done, info ? h->d_info : nullptr, cs);
    if (rc) return rc;
    const size_t m = (size_t)(e - b);
    CU(cudaMemcpyAsync(o_dst + (size_t)b * od, h->d_obs + (size_t)b * od, m * od * 4,
                       cudaMemcpyDeviceToHost, cs));
    CU(cudaMemcpyAsync(r_dst + b, h->d_rew + b, m * 4, cudaMemcpyDeviceToHost, cs));
    CU(cudaMemcpyAsync(d_dst + b, h->d_done + b, m, cudaMemcpyDeviceToHost, cs));
    if (info) CU(cudaMemcpyAsync(i_dst + b, h->d_info + b, m * sizeof(PgdInfo), cudaMemcpyDeviceToHost, cs));
  }
  CU(cudaStreamSynchronize(st));
  if (chunks > 1) CU(cudaStreamSynchronize(h->own_stream2));
  if (!direct) {
    memcpy(obs, h->h_obs, n * od * 4);
    memcpy(reward, h->h_rew, n * 4);
    memcpy(done, h->h_done, n);
    if (info) memcpy(info, h->h_info, n * sizeof(PgdInfo));
  }
  return 0;
}

// the V per-slot records of one environment: contiguous in the default layout ([env][slot]), strided by num_envs in
// the one-thread-per-environment layout ([slot][env])
static cudaError_t copy_slots(PgdHandle* h, void* dev_base, int env, void* host, bool to_host) {
  const int V = h->cfg.num_slots;
  if (h->cfg.layout == 0) {
    char* d = (char*)dev_base + (size_t)env * V * 16;
    return to_host ? cudaMemcpy(host, d, (size_t)V * 16, cudaMemcpyDeviceToHost)
                   : cudaMemcpy(d, host, (size_t)V * 16, cudaMemcpyHostToDevice);
  }
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
    s->timer = nav[i].w; s->rnd_n = misc[i].x; s->airborne = misc[i].y; s->flags = misc[i].z;
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
    misc[i] = make_int4(s->rnd_n, s->airborne, s->flags, 0);
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
