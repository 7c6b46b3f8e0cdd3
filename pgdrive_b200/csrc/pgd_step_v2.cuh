/* Environment step, ONE THREAD PER ENVIRONMENT (experimental second layout of the step; not the default).
 *
 * Why: the cooperative kernel (pgd_step.cu: 16 threads per environment) runs every ego-only stretch of the step --
 * action, five physics sub-steps, reward bookkeeping -- on 2 of a warp's 32 lanes and is bound by instruction fetch
 * (profiles/r01k_experiments.md).  Here a thread walks through its environment's vehicles sequentially, so a warp
 * executes 32 environments per instruction: utilisation depends on how alike the environments are, not on how many
 * vehicles one environment has.  State is slot-major ([slot][env]) so that the 32 lanes of a warp read and write
 * consecutive 16-byte vectors.
 *
 * The arithmetic is the cooperative kernel's, expression for expression (it is what tests pin against the oracle):
 * phases A-G of pgd_step.cu, with its shared-memory exchanges replaced by loops over a thread-local vehicle array.
 * The file is written for host AND device: oracle/step_v2_host.cpp compiles it with g++ so that its logic can be
 * checked against the CPU oracle without a GPU (tests/test_step_v2.py).
 *
 * Reference call stack (paths under /root/reference/pgdrive): envs/base_env.py:184-224,303-344 (step),
 * policy/idm_policy.py:83-353 (IDM), engine/base_engine.py:206-232 (sub-steps), vehicle_module/navigation.py:155-344,
 * utils/scene_utils.py:138-185 (localisation), component/vehicle/base_vehicle.py:615-644 (line / sidewalk contacts),
 * cutils.pyx:60-142 + vehicle_module/lidar.py:55-77 (lidar, neighbours), obs/state_obs.py:58-170 (observation),
 * envs/pgdrive_env.py:162-258 (reward / cost / done).
 */
#ifndef PGD_STEP_V2_CUH
#define PGD_STEP_V2_CUH
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/pgd_math.h"
#include "../../include/pgd_tables.h"

#ifdef __CUDACC__
#define V2_HD __host__ __device__ __forceinline__
// one out-of-line copy on the device: the accurate sincosf / atan2f / fmodf expand to hundreds of instructions each,
// and inlined at every call site they made the kernel 10 k instructions (the cooperative kernel learnt the same
// lesson, profiles/README.md r01d -> r01g)
#define V2_HD_OUTLINE __host__ __device__ __noinline__
#else
#define V2_HD inline
#define V2_HD_OUTLINE inline
#endif
#define V2_SINCOS(a, s, c) pgdv2::sincos_hd((a), &(s), &(c))
#define V2_LDG(p) pgdv2::ldg(p)

namespace pgdv2 {

#define V2_PI 3.14159265358979323846f
#define V2_TWO_PI 6.28318530717958647692f
#define V2_GRAVITY 9.81f
#define V2_LIDAR_RANGE 50.0f
#define V2_MAX_SPEED_KMH 80.0f
#define V2_IDM_MAX_LONG 30.0f
#define V2_IDM_NORMAL_SPEED 30.0f
#define V2_IDM_CREEP_SPEED 5.0f
#define V2_IDM_SAFE_DIST 15.0f
#define V2_IDM_LANE_CHANGE_FREQ 50
#define V2_IDM_SPEED_INCREASE 10.0f
#define V2_IDM_MAX_SPEED 100.0f
#define V2_YAW_TAU 0.1f
#define V2_DONE_PENDING_RESET 2
#define V2_HDG_VALID (1 << 30) /* thread-local bit of Veh.vflags: (hc, hs) computed; never stored */
#define V2_MAX_SUBSTEPS 16  /* decision_repeat supported by this layout (default 5) */

struct alignas(16) F4 { float x, y, z, w; };  // one 16-byte load / store
struct alignas(16) I4 { int x, y, z, w; };

struct Tables {  // device (or host) pointers to the tables of include/pgd_tables.h
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

struct State {  // slot-major: per-slot arrays are indexed slot * num_envs + env, per-env arrays by env
  F4* pose;   // x, y, heading, speed
  F4* ctrl;   // steer, throttle, heading-PID last error, heading-PID summed error
  F4* pidl;   // lateral-PID last error, summed error, IDM target speed, yaw rate
  I4* nav;    // lane, ck0 | ck1 << 16, routing target lane, overtake timer
  I4* misc;   // rnd draws used, airborne sub-steps left, PGD_V_* flags, -
  I4* envi;   // episode, next trigger group, done, episode length
  F4* envf;   // previous steering, previous throttle, episode reward, episode energy
};

V2_HD_OUTLINE void sincos_hd(float a, float* s, float* c) {
  pgd_sincosf(a, s, c);
}

template <class T_>
V2_HD T_ ldg(const T_* p) {  // read-only table data
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}

/* A whole table record (PgdLane 64 B, PgdBox / PgdRoad 32 B, PgdMap 64 B: multiples of 16, 16-byte aligned) with
 * 16-byte read-only loads on the device. */
template <class T_>
V2_HD T_ load_rec(const T_* p) {
#ifdef __CUDA_ARCH__
  static_assert(sizeof(T_) % 16 == 0, "table records are multiples of 16 bytes");
  T_ out;
  const uint4* src = reinterpret_cast<const uint4*>(p);
  uint4* dst = reinterpret_cast<uint4*>(&out);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T_) / 16); ++i) dst[i] = __ldg(src + i);
  return out;
#else
  return *p;
#endif
}

V2_HD float clipf(float a, float lo, float hi) { return fminf(fmaxf(a, lo), hi); }

V2_HD_OUTLINE float wrap_to_pi(float x) {
  return pgd_wrap_to_pi(x);
}

V2_HD_OUTLINE void arc_local(float cx, float cy, float ph0, float dir, float radius, float x, float y, float* lon,
                             float* lat) {
  float dx = x - cx, dy = y - cy;
  float phi = pgd_atan2f(dy, dx);
  phi = ph0 + wrap_to_pi(phi - ph0);
  float r = sqrtf(dx * dx + dy * dy);
  *lon = dir * (phi - ph0) * radius;
  *lat = dir * (radius - r);
}

V2_HD void lane_local(const PgdLane& l, float x, float y, float& lon, float& lat) {
  if (l.kind == PGD_LANE_STRAIGHT) {
    float dx = x - l.sx, dy = y - l.sy;
    lon = dx * l.ax + dy * l.ay;
    lat = dx * -l.ay + dy * l.ax;
  } else {
    arc_local(l.ax, l.ay, l.ph0, l.dir, l.radius, x, y, &lon, &lat);
  }
}

V2_HD void lane_position(const PgdLane& l, float lon, float lat, float& x, float& y) {
  if (l.kind == PGD_LANE_STRAIGHT) {
    x = l.sx + lon * l.ax + lat * -l.ay;
    y = l.sy + lon * l.ay + lat * l.ax;
  } else {
    float phi = l.dir * lon / l.radius + l.ph0;
    float r = l.radius - lat * l.dir;
    float s, c;
    V2_SINCOS(phi, s, c);
    x = l.ax + r * c;
    y = l.ay + r * s;
  }
}

V2_HD float lane_heading_at(const PgdLane& l, float lon) {
  if (l.kind == PGD_LANE_STRAIGHT) return l.heading;
  float phi = l.dir * lon / l.radius + l.ph0;
  return phi + V2_PI / 2 * l.dir;
}

V2_HD bool precedes(float ex, float ey, float sx, float sy) {
  float dx = ex - sx, dy = ey - sy;
  return dx * dx + dy * dy < 1e-2f;
}

struct Rect { float cx, cy, ux, uy, hl, hw; };

V2_HD_OUTLINE bool rect_overlap(const Rect& a, const Rect& b) {
  float dx = b.cx - a.cx, dy = b.cy - a.cy;
  float c = fabsf(a.ux * b.ux + a.uy * b.uy);
  float s = fabsf(a.ux * b.uy - a.uy * b.ux);
  if (fabsf(dx * a.ux + dy * a.uy) > a.hl + b.hl * c + b.hw * s) return false;
  if (fabsf(-dx * a.uy + dy * a.ux) > a.hw + b.hl * s + b.hw * c) return false;
  if (fabsf(dx * b.ux + dy * b.uy) > b.hl + a.hl * c + a.hw * s) return false;
  if (fabsf(-dx * b.uy + dy * b.ux) > b.hw + a.hl * s + a.hw * c) return false;
  return true;
}

V2_HD float ray_rect(float ox, float oy, float dx, float dy, const Rect& r) {
  float px = ox - r.cx, py = oy - r.cy;
  float lo0 = px * r.ux + py * r.uy, lo1 = -px * r.uy + py * r.ux;
  float ld0 = dx * r.ux + dy * r.uy, ld1 = -dx * r.uy + dy * r.ux;
  if (fabsf(lo0) <= r.hl && fabsf(lo1) <= r.hw) return 1.0f;
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

V2_HD_OUTLINE void project(float hx, float hy, float vx, float vy, float& fwd, float& side) {
  const float n = 1.0f + 1e-6f;
  fwd = (vx * hx + vy * hy) / n;
  side = (vx * -hy + vy * hx) / n;
}

V2_HD float pid(float& p_err, float& i_err, float kp, float ki, float kd, float err) {
  i_err += err;
  float d = err - p_err;
  p_err = err;
  return -kp * p_err - ki * i_err - kd * d;
}

struct Veh {  // one vehicle of the thread's environment
  float x, y, h, v, yaw, steer, throttle, hp, hi, lp, li, tspeed;
  float hc, hs, hl, hw;
  int lane, ck0, ck1, rt_lane, timer, rnd_n, airborne, vflags;
};

struct Sub {
  float accel, brake_dv, sb, mu_g, lr;
};

V2_HD_OUTLINE void substep(Veh& q, const Sub& sub, float dt) {
  float speed = q.v;
  if (sub.accel > 0.0f) speed += sub.accel * dt;
  else speed = fmaxf(speed - sub.brake_dv, 0.0f);
  float yaw = q.yaw + (speed * sub.sb / sub.lr - q.yaw) * (dt / V2_YAW_TAU);
  if (speed * fabsf(yaw) > sub.mu_g) yaw = copysignf(sub.mu_g / speed, yaw);
  const float sb = speed > 1e-3f ? clipf(yaw * sub.lr / speed, -1.0f, 1.0f) : 0.0f;
  const float cb = sqrtf(fmaxf(1.0f - sb * sb, 0.0f));
  q.x += speed * (q.hc * cb - q.hs * sb) * dt;
  q.y += speed * (q.hs * cb + q.hc * sb) * dt;
  float nh = q.h + yaw * dt;
  if (nh > V2_PI) nh -= V2_TWO_PI;
  if (nh < -V2_PI) nh += V2_TWO_PI;
  q.yaw = yaw;
  if (nh != q.h) V2_SINCOS(nh, q.hs, q.hc);
  q.h = nh;
  q.v = speed;
}

/* What the 240 lidar beams of one environment depend on: the ego's pose and, per chassis in range, its rectangle and
 * the (conservative, exact) arc of beams that can reach it.  The beams themselves are evaluated when the observation
 * row is written out, so they never sit in thread-local memory. */
template <int V>
struct LidarCtx {
  float ex, ey, eh;
  int n;
  float cx[V], cy[V], ux[V], uy[V], hl[V], hw[V];
  int blo[V], bn[V];
};

template <int V>
V2_HD_OUTLINE float lidar_beam(const LidarCtx<V>& lc, int i) {
  float best = 1.0f;
  bool have_dir = false;
  float dx = 0.0f, dy = 0.0f;
  for (int j = 0; j < lc.n; ++j) {
    int rel = i - lc.blo[j];
    if (rel < 0) rel += PGD_LIDAR_BEAMS;
    if (rel > lc.bn[j]) continue;
    if (!have_dir) {
      const float ang = (float)i * (V2_TWO_PI / (float)PGD_LIDAR_BEAMS) + lc.eh;
      float sn, cs;
      V2_SINCOS(ang, sn, cs);
      dx = cs * V2_LIDAR_RANGE;
      dy = sn * V2_LIDAR_RANGE;
      have_dir = true;
    }
    const Rect r = {lc.cx[j], lc.cy[j], lc.ux[j], lc.uy[j], lc.hl[j], lc.hw[j]};
    best = fminf(best, ray_rect(lc.ex, lc.ey, dx, dy, r));
  }
  return best;
}

V2_HD_OUTLINE void ensure_heading(Veh& q) {  // heading unit vector of a parked vehicle, on first use
  if (!(q.vflags & V2_HDG_VALID)) {
    V2_SINCOS(q.h, q.hs, q.hc);
    q.vflags |= V2_HDG_VALID;
  }
}

/* One environment, one decision step (mode 0) or the reset pass (mode 1: only environments marked pending are
 * touched).  V = vehicle slots.  obs receives the row up to the lidar beams (34 floats without detectors); the
 * beams are described by lc and evaluated by lidar_beam(). */
template <int V>
V2_HD void step_env(const Tables& T, const State& S, const PgdConfig& cfg, int mode, int env, int num_envs,
                    const float* action, float* obs, LidarCtx<V>& lc, float* reward, uint8_t* done, PgdInfo* info) {
  I4 envi = S.envi[env];
  F4 envf = S.envf[env];
  const bool pending = envi.z == V2_DONE_PENDING_RESET;
  bool fresh;
  if (mode == 1) fresh = pending;
  else fresh = pending || (cfg.auto_reset && envi.z == 1);
  if (mode == 1 && !pending) return;
  const bool stepping = !fresh;

  const PgdEpisode* ep = T.episodes + envi.x;
  const PgdMap mp = load_rec(T.maps + V2_LDG(&ep->map));
  const int n_slots = V2_LDG(&ep->n_slots);
  const int n_groups = V2_LDG(&ep->n_groups);
  const PgdLane* lanes = T.lanes + mp.lane_off;
  const PgdRoad* roads = T.roads + mp.road_off;
  const PgdBox* boxes = T.boxes + mp.box_off;
  const PgdSlot* tpl = T.slots + V2_LDG(&ep->slot_off);

  // ---- phase A: load ------------------------------------------------------------------------------------------
  Veh veh[V];
#ifdef V2_POISON  // host debugging: thread-local arrays start as garbage on the device; results must not depend on it
  memset(veh, 0xff, sizeof(veh));
  memset(&lc, 0xff, sizeof(lc));
#endif
  uint32_t was_parked = 0;  // slots that entered this step as parked traffic
  uint32_t untouched = 0;   // slots whose stored state stays as it is
  uint32_t drop_ran = 0;    // parked slots whose drop counter changed
  if (!fresh) {  // the flag words of all slots first: 16 independent 16-byte loads in flight
#pragma unroll
    for (int s = 0; s < V; ++s) {
      if (s < n_slots) {
        const I4 m = S.misc[(size_t)s * num_envs + env];
        veh[s].rnd_n = m.x; veh[s].airborne = m.y; veh[s].vflags = m.z;
      }
    }
  }
  #pragma unroll 1
  for (int s = 0; s < n_slots; ++s) {
    Veh& q = veh[s];
    const PgdSlot& t = tpl[s];
    if (fresh) {
      q.x = t.x; q.y = t.y; q.h = t.heading; q.v = 0.0f; q.yaw = 0.0f;
      q.steer = q.throttle = q.hp = q.hi = q.lp = q.li = 0.0f;
      q.tspeed = V2_IDM_NORMAL_SPEED;
      q.lane = t.lane; q.ck0 = 0; q.ck1 = t.route_len > 2 ? 1 : 0; q.rt_lane = -1;
      q.timer = t.overtake_timer; q.rnd_n = 0; q.airborne = t.drop_substeps;
      q.vflags = PGD_V_ALIVE | PGD_V_ON_LANE | (s == 0 ? PGD_V_ACTIVE : 0);
    } else {
      const size_t gi = (size_t)s * num_envs + env;
      if (!(q.vflags & PGD_V_ALIVE)) {
        untouched |= 1u << s;  // removed earlier: nothing reads it, nothing is stored
      } else if (q.vflags & PGD_V_ACTIVE) {
        const F4 p = S.pose[gi], c = S.ctrl[gi], l = S.pidl[gi];
        const I4 n = S.nav[gi];
        q.x = p.x; q.y = p.y; q.h = p.z; q.v = p.w;
        q.steer = c.x; q.throttle = c.y; q.hp = c.z; q.hi = c.w;
        q.lp = l.x; q.li = l.y; q.tspeed = l.z; q.yaw = l.w;
        q.lane = n.x; q.ck0 = n.y & 0xffff; q.ck1 = n.y >> 16; q.rt_lane = n.z; q.timer = n.w;
      } else {
        // Traffic that has not been woken yet has never been touched by IDM, physics (it is at rest) or
        // localisation: everything but its drop counter still has the value the reset gave it, so it is taken from
        // the episode template (L2-resident, shared by all environments on the seed) instead of from the state.
        q.x = t.x; q.y = t.y; q.h = t.heading; q.v = 0.0f; q.yaw = 0.0f;
        q.steer = q.throttle = q.hp = q.hi = q.lp = q.li = 0.0f;
        q.tspeed = V2_IDM_NORMAL_SPEED;
        q.lane = t.lane; q.ck0 = 0; q.ck1 = t.route_len > 2 ? 1 : 0; q.rt_lane = -1;
        q.timer = t.overtake_timer;
      }
    }
    q.hl = t.length * 0.5f;
    q.hw = t.width * 0.5f;
    // heading unit vector: now for vehicles that move, on first use for parked ones (most are never looked at)
    if ((q.vflags & (PGD_V_ALIVE | PGD_V_ACTIVE)) == (PGD_V_ALIVE | PGD_V_ACTIVE)) ensure_heading(q);
    if (!(q.vflags & PGD_V_ACTIVE)) was_parked |= 1u << s;
  }
  if (fresh) {
    envi.y = 0; envi.z = 0; envi.w = 0;
    envf.x = envf.y = envf.z = envf.w = 0.0f;
  }
  const float last_x = veh[0].x, last_y = veh[0].y, last_h = veh[0].h;
  int crash = 0;

  if (stepping) {
    // ---- phase B: ego action + traffic trigger --------------------------------------------------------------
    {
      Veh& ego = veh[0];
      envf.x = ego.steer;
      envf.y = ego.throttle;
      ego.steer = clipf(action[0], -1.0f, 1.0f);  // fminf / fmaxf drop NaN -> -1, like the compiled cutils_clip
      ego.throttle = clipf(action[1], -1.0f, 1.0f);
    }
    if (envi.y < n_groups) {
      const int ego_road = V2_LDG(&lanes[veh[0].lane].road);
      if (ego_road == V2_LDG(&ep->trigger_road[envi.y])) {
        #pragma unroll 1
        for (int s = 1; s < n_slots; ++s)
          if (tpl[s].group == envi.y) {
            veh[s].vflags |= PGD_V_ACTIVE;
            ensure_heading(veh[s]);
          }
        envi.y += 1;
      }
    }
    // ---- phase C: IDM ------------------------------------------------------------------------------------------
    bool any_awake = false;
    #pragma unroll 1
    for (int s = 1; s < n_slots; ++s)
      any_awake = any_awake || ((veh[s].vflags & (PGD_V_ALIVE | PGD_V_ACTIVE)) == (PGD_V_ALIVE | PGD_V_ACTIVE));
    if (any_awake) {
      float olong[V], lsx[V], lsy[V], lex[V], ley[V], llen[V];
      #pragma unroll 1
      for (int s = 0; s < n_slots; ++s) {
        if (!(veh[s].vflags & PGD_V_ALIVE)) continue;
        ensure_heading(veh[s]);
        const PgdLane l = load_rec(lanes + (veh[s].lane));
        lsx[s] = l.sx; lsy[s] = l.sy; lex[s] = l.ex; ley[s] = l.ey; llen[s] = l.length;
        float lon, lat;
        lane_local(l, veh[s].x, veh[s].y, lon, lat);
        olong[s] = lon;
      }
      #pragma unroll 1
      for (int s = 1; s < n_slots; ++s) {
        Veh& q = veh[s];
        if ((q.vflags & (PGD_V_ALIVE | PGD_V_ACTIVE)) != (PGD_V_ALIVE | PGD_V_ACTIVE)) continue;
        const PgdSlot& t = tpl[s];
        const int32_t* rroads = T.route_roads + t.route_off;
        const int cur_road_id = V2_LDG(&rroads[q.ck0]);
        const PgdRoad cur_road = load_rec(roads + (cur_road_id));
        bool ok;
        if (q.rt_lane < 0) {
          q.rt_lane = q.lane;
          ok = V2_LDG(&lanes[q.rt_lane].road) == cur_road_id;
        } else if (V2_LDG(&lanes[q.rt_lane].road) != cur_road_id) {
          ok = false;
          const float rex = V2_LDG(&lanes[q.rt_lane].ex), rey = V2_LDG(&lanes[q.rt_lane].ey);
          #pragma unroll 1
          for (int k = 0; k < cur_road.n_lanes; ++k) {
            const PgdLane* c = lanes + cur_road.first_lane + k;
            if (precedes(rex, rey, V2_LDG(&c->sx), V2_LDG(&c->sy))) {
              q.rt_lane = cur_road.first_lane + k;
              ok = true;
              break;
            }
          }
        } else if (V2_LDG(&lanes[q.lane].road) == cur_road_id && q.rt_lane != q.lane) {
          q.rt_lane = q.lane;
          q.timer = t.rnd25[q.rnd_n % PGD_N_RND25];
          q.rnd_n++;
          ok = true;
        } else {
          ok = true;
        }
        const PgdLane rl = load_rec(lanes + (q.rt_lane));
        int cand[3] = {-1, q.rt_lane, -1};
        if (ok) {
          const PgdRoad rr = load_rec(roads + (rl.road));
          if (rl.idx > 0) cand[0] = rr.first_lane + rl.idx - 1;
          if (rl.idx + 1 < rr.n_lanes) cand[2] = rr.first_lane + rl.idx + 1;
        }
        int front[3], back[3];
        float fdist[3], bdist[3];
        for (int i = 0; i < 3; ++i) {
          front[i] = back[i] = -1;
          fdist[i] = bdist[i] = V2_IDM_MAX_LONG;
          if (cand[i] < 0) continue;
          const PgdLane l = (i == 1) ? rl : load_rec(lanes + cand[i]);
          float cur_long, lat;
          lane_local(l, q.x, q.y, cur_long, lat);
          const float left_long = l.length - cur_long;
          bool found_front = false, found_back = false;
          #pragma unroll 1
          for (int j = 0; j < n_slots; ++j) {
            if (j == s || !(veh[j].vflags & PGD_V_ALIVE)) continue;
            const float ddx = veh[j].x - q.x, ddy = veh[j].y - q.y;
            if (!(ddx * ddx + ddy * ddy < V2_LIDAR_RANGE * V2_LIDAR_RANGE)) continue;
            if (veh[j].lane == cand[i]) {
              const float lg = olong[j] - cur_long;
              if (fdist[i] > lg && lg > 0.0f) { fdist[i] = lg; front[i] = j; found_front = true; }
              if (lg < 0.0f && fabsf(lg) < bdist[i]) { bdist[i] = fabsf(lg); back[i] = j; found_back = true; }
            } else if (!found_front && precedes(l.ex, l.ey, lsx[j], lsy[j])) {
              const float lg = olong[j] + left_long;
              if (fdist[i] > lg && lg > 0.0f) { fdist[i] = lg; front[i] = j; }
            } else if (!found_back && precedes(lex[j], ley[j], l.sx, l.sy)) {
              const float lg = llen[j] - olong[j] + cur_long;
              if (bdist[i] > lg) { bdist[i] = lg; back[i] = j; }
            }
          }
        }
        int front_obj = front[1], steer_lane = q.rt_lane;
        float front_dist = fdist[1];
        if (ok) {  // lane_change_policy
          const int n_cur = cur_road.n_lanes;
          int lo = 0, hi_idx = n_cur - 1;
          bool decided = false;
          const int idx = rl.idx;
          if (q.ck0 != q.ck1) {
            const PgdRoad nxt = load_rec(roads + (V2_LDG(&rroads[q.ck1])));
            const int diff = n_cur - nxt.n_lanes;
            if (diff > 0) {
              const PgdLane* c0 = lanes + cur_road.first_lane;
              const PgdLane* n0 = lanes + nxt.first_lane;
              if (precedes(V2_LDG(&c0->ex), V2_LDG(&c0->ey), V2_LDG(&n0->sx), V2_LDG(&n0->sy))) {
                lo = 0; hi_idx = nxt.n_lanes - 1;
              } else {
                lo = diff; hi_idx = n_cur - 1;
              }
              if (idx < lo || idx > hi_idx) {
                decided = true;
                const int side = idx > hi_idx ? 0 : 2;
                if (bdist[side] < V2_IDM_SAFE_DIST || fdist[side] < 5.0f) {
                  q.tspeed = V2_IDM_CREEP_SPEED;
                } else {
                  q.tspeed = V2_IDM_NORMAL_SPEED;
                  front_obj = front[side];
                  front_dist = fdist[side];
                  steer_lane = cur_road.first_lane + idx + (side == 0 ? -1 : 1);
                }
              }
            }
          }
          if (!decided) {
            const float my_speed = clipf(q.v * 3.6f, 0.0f, 100000.0f);
            if (fabsf(my_speed - V2_IDM_NORMAL_SPEED) > 3.0f && front[1] >= 0 &&
                fabsf(clipf(veh[front[1]].v * 3.6f, 0.0f, 100000.0f) - V2_IDM_NORMAL_SPEED) > 3.0f &&
                q.timer > V2_IDM_LANE_CHANGE_FREQ) {
              float side_speed[3] = {0.f, 0.f, 0.f};
              bool side_ok[3] = {false, false, false};
              for (int sd = 0; sd < 3; sd += 2) {
                if (front[sd] >= 0) {
                  side_speed[sd] = clipf(veh[front[sd]].v * 3.6f, 0.0f, 100000.0f);
                  side_ok[sd] = true;
                } else if (cand[sd] >= 0 && fdist[sd] > V2_IDM_SAFE_DIST && bdist[sd] > V2_IDM_SAFE_DIST) {
                  side_speed[sd] = V2_IDM_MAX_SPEED;
                  side_ok[sd] = true;
                }
              }
              const float front_speed = clipf(veh[front[1]].v * 3.6f, 0.0f, 100000.0f);
              if (side_ok[0] && side_speed[0] - front_speed > V2_IDM_SPEED_INCREASE && idx - 1 >= lo &&
                  idx - 1 <= hi_idx) {
                decided = true;
                front_obj = front[0]; front_dist = fdist[0];
                steer_lane = cur_road.first_lane + idx - 1;
              } else if (side_ok[2] && side_speed[2] - front_speed > V2_IDM_SPEED_INCREASE && idx + 1 >= lo &&
                         idx + 1 <= hi_idx) {
                decided = true;
                front_obj = front[2]; front_dist = fdist[2];
                steer_lane = cur_road.first_lane + idx + 1;
              }
            }
          }
          if (!decided) {
            q.tspeed = V2_IDM_NORMAL_SPEED;
            q.timer += 1;
          }
        }
        {  // steering_control
          const PgdLane tl = (steer_lane == q.rt_lane) ? rl : load_rec(lanes + steer_lane);
          float lon, lat;
          lane_local(tl, q.x, q.y, lon, lat);
          const float lane_heading = lane_heading_at(tl, lon + 1.0f);
          float st = pid(q.hp, q.hi, 1.7f, 0.01f, 3.5f, wrap_to_pi(lane_heading - q.h));
          st += pid(q.lp, q.li, 0.3f, 0.002f, 0.05f, -lat);
          q.steer = st;
        }
        {  // acceleration
          const float sp = clipf(q.v * 3.6f, 0.0f, 100000.0f);
          float acc = 1.0f - pgd_pow10f(fmaxf(sp, 0.0f) / q.tspeed);
          if (front_obj >= 0) {
            const float hx = q.hc, hy = q.hs;
            const float fs = clipf(veh[front_obj].v * 3.6f, 0.0f, 100000.0f);
            const float dvx = sp * hx - fs * veh[front_obj].hc, dvy = sp * hy - fs * veh[front_obj].hs;
            const float dv = dvx * hx + dvy * hy;
            const float d_star = 10.0f + sp * 1.5f + sp * dv / (2.0f * sqrtf(5.0f));
            float d = front_dist;
            if (!(fabsf(d) > 1e-2f)) d = d > 0.0f ? 1e-2f : -1e-2f;
            const float ratio = d_star / d;
            acc -= ratio * ratio;
          }
          q.throttle = acc;
        }
      }
    }
    // ---- phase D: physics sub-steps + chassis contact ----------------------------------------------------------
    // the ego first, remembering its pose after every sub-step; then every other vehicle integrates itself and
    // tests its chassis against those poses (same order of events as sub-step-by-sub-step for all)
    F4 ego_traj[V2_MAX_SUBSTEPS];
    {
      const int ns = cfg.decision_repeat < V2_MAX_SUBSTEPS ? cfg.decision_repeat : V2_MAX_SUBSTEPS;
      float ego_travel = -1.0f;  // not known until the ego (slot 0, always alive) has been integrated
      #pragma unroll 1
      for (int s = 0; s < n_slots; ++s) {
        Veh& q = veh[s];
        if (!(q.vflags & PGD_V_ALIVE)) continue;
        const float reach = veh[0].hl + veh[0].hw + q.hl + q.hw;
        if (s > 0 && ego_travel < 0.0f) {  // first vehicle after the ego: how far did the ego get from its start pose?
          float m2 = 0.0f;
          #pragma unroll 1
          for (int k = 0; k < ns; ++k) {
            const float ex = ego_traj[k].x - last_x, ey = ego_traj[k].y - last_y;
            m2 = fmaxf(m2, ex * ex + ey * ey);
          }
          ego_travel = sqrtf(m2) * 1.001f + 1e-3f;
        }
        // A vehicle at rest with no yaw rate and no engine force (parked traffic, a braking ego) is a fixed point of
        // the sub-step: speed = max(0 - dv, 0) = 0, yaw stays 0, the pose does not move.  With v = 0 it cannot be
        // over the speed limit, so "no engine force" is just throttle <= 0 and the force model is not needed at all;
        // only its drop counter runs and its (fixed) chassis is tested against the ego's pose of every sub-step.
        if (q.v == 0.0f && q.yaw == 0.0f && !(q.throttle > 0.0f)) {
          if (q.airborne > 0) drop_ran |= 1u << s;
          q.airborne = q.airborne > ns ? q.airborne - ns : 0;
          if (s == 0) {
            #pragma unroll 1
            for (int k = 0; k < ns; ++k) {
              ego_traj[k].x = q.x; ego_traj[k].y = q.y; ego_traj[k].z = q.hc; ego_traj[k].w = q.hs;
            }
          } else {
            // the ego stays within ego_travel of where it started the step: a parked chassis further away than
            // reach + ego_travel cannot be touched in any sub-step (triangle inequality; the margin covers rounding)
            const float ddx0 = q.x - last_x, ddy0 = q.y - last_y;
            const float far = reach + ego_travel;
            if (ddx0 * ddx0 + ddy0 * ddy0 > far * far) continue;
            #pragma unroll 1
            for (int k = 0; k < ns; ++k) {
              const float ddx = q.x - ego_traj[k].x, ddy = q.y - ego_traj[k].y;
              if (ddx * ddx + ddy * ddy <= reach * reach) {
                ensure_heading(q);
                Rect me = {q.x, q.y, q.hc, q.hs, q.hl, q.hw};
                Rect eg = {ego_traj[k].x, ego_traj[k].y, ego_traj[k].z, ego_traj[k].w, veh[0].hl, veh[0].hw};
                if (rect_overlap(eg, me)) crash = 1;
              }
            }
          }
          continue;
        }
        const PgdSlot& t = tpl[s];
        Sub sub;
        sub.mu_g = t.friction * V2_GRAVITY;
        sub.lr = t.lr;
        const bool overspeed = clipf(q.v * 3.6f, 0.0f, 100000.0f) > V2_MAX_SPEED_KMH;
        if (q.throttle > 0.0f && !overspeed) {
          sub.accel = fminf(4.0f * t.max_engine * q.throttle / t.mass, sub.mu_g);
          sub.brake_dv = 0.0f;
        } else {
          sub.accel = 0.0f;
          const float imp = q.throttle >= 0.0f ? 2.0f : -q.throttle * t.max_brake;
          sub.brake_dv = fminf(4.0f * imp / t.mass, sub.mu_g * cfg.dt);
        }
        const float delta = clipf(-q.steer * t.max_steer, -1.4f, 1.4f);
        const float tb = t.lr / (t.lf + t.lr) * pgd_tanf(delta);
        sub.sb = tb / sqrtf(1.0f + tb * tb);
        const bool at_rest = q.v == 0.0f && q.yaw == 0.0f && !(sub.accel > 0.0f);
        #pragma unroll 1
        for (int k = 0; k < ns; ++k) {
          if (q.airborne > 0) q.airborne--;
          else if (!at_rest) substep(q, sub, cfg.dt);
          if (s == 0) {
            ego_traj[k].x = q.x; ego_traj[k].y = q.y; ego_traj[k].z = q.hc; ego_traj[k].w = q.hs;
          } else {
            const float ddx = q.x - ego_traj[k].x, ddy = q.y - ego_traj[k].y;
            if (ddx * ddx + ddy * ddy <= reach * reach) {
              Rect me = {q.x, q.y, q.hc, q.hs, q.hl, q.hw};
              Rect eg = {ego_traj[k].x, ego_traj[k].y, ego_traj[k].z, ego_traj[k].w, veh[0].hl, veh[0].hw};
              if (rect_overlap(eg, me)) crash = 1;
            }
          }
        }
      }
    }
  }

  // ---- phase E: after_step -------------------------------------------------------------------------------------
  uint32_t flags = 0;
  {
    const Veh& ego = veh[0];
    const Rect er = {ego.x, ego.y, ego.hc, ego.hs, ego.hl, ego.hw};
    const int32_t* ent = T.cell_entries + mp.entry_off;
    #pragma unroll 1
    for (int s = 0; s < n_slots; ++s) {
      Veh& q = veh[s];
      if ((q.vflags & (PGD_V_ALIVE | PGD_V_ACTIVE)) != (PGD_V_ALIVE | PGD_V_ACTIVE)) continue;
      const PgdSlot& t = tpl[s];
      const int32_t* rroads = T.route_roads + t.route_off;
      const int32_t* rnodes = T.route_nodes + t.route_off;
      const int cur_road = V2_LDG(&rroads[q.ck0]);
      const int next_road = q.ck0 != q.ck1 ? V2_LDG(&rroads[q.ck1]) : -1;
      int b_any = INT_MAX, b_cur = INT_MAX, b_next = INT_MAX;
      const int cx = (int)floorf((q.x - mp.x0) * mp.inv_cell), cy = (int)floorf((q.y - mp.y0) * mp.inv_cell);
      if (cx >= 0 && cy >= 0 && cx < mp.nx && cy < mp.ny) {
        const int cell = mp.cell_off + cy * mp.nx + cx;
        const int b0 = V2_LDG(&T.cell_start[cell]), b1 = V2_LDG(&T.cell_start[cell + 1]);
        // entries are fetched four at a time (indices, then records) so that their latencies overlap
        for (int k0 = b0; k0 < b1; k0 += 4) {
          int bb[4];
          PgdBox gg[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) bb[j] = (k0 + j < b1) ? V2_LDG(&ent[k0 + j]) : -1;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (bb[j] >= 0) gg[j] = load_rec(boxes + bb[j]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
          if (bb[j] < 0) continue;
          const int b = bb[j];
          const PgdBox& g = gg[j];
          if (g.kind == PGD_BOX_LANE) {
            const float dx = q.x - g.cx, dy = q.y - g.cy;
            if (!(fabsf(dx * g.ux + dy * g.uy) <= g.hl && fabsf(-dx * g.uy + dy * g.ux) <= g.hw)) continue;
            const PgdLane* l = lanes + g.lane;
            float dot;
            if (V2_LDG(&l->kind) == PGD_LANE_STRAIGHT) {
              dot = V2_LDG(&l->ax) * q.hc + V2_LDG(&l->ay) * q.hs;
            } else {
              dot = V2_LDG(&l->dir) * ((q.x - V2_LDG(&l->ax)) * q.hs - (q.y - V2_LDG(&l->ay)) * q.hc);
            }
            if (!(dot > 0.0f)) continue;
            const int lroad = V2_LDG(&l->road);
            if (b < b_any) b_any = b;
            if (lroad == cur_road && b < b_cur) b_cur = b;
            if (lroad == next_road && b < b_next) b_next = b;
          } else if (s == 0) {
            const Rect r = {g.cx, g.cy, g.ux, g.uy, g.hl, g.hw};
            if (!rect_overlap(er, r)) continue;
            flags |= g.kind == PGD_BOX_WHITE ? PGD_F_ON_WHITE
                   : g.kind == PGD_BOX_YELLOW ? PGD_F_ON_YELLOW
                   : g.kind == PGD_BOX_BROKEN ? PGD_F_ON_BROKEN : PGD_F_CRASH_SIDEWALK;
          }
          }
        }
      }
      const int nb = b_cur != INT_MAX ? b_cur : (b_next != INT_MAX ? b_next : b_any);
      const bool on_lane = nb != INT_MAX;
      if (on_lane) q.lane = V2_LDG(&boxes[nb].lane);
      if (q.ck0 != q.ck1) {  // _update_target_checkpoints
        const PgdLane l = load_rec(lanes + (q.lane));
        float lon, lat;
        lane_local(l, q.x, q.y, lon, lat);
        const int start = V2_LDG(&roads[l.road].start_node);
        if (lon < 5.0f) {
          #pragma unroll 1
          for (int j = q.ck1; j < t.route_len - 1; ++j) {
            if (V2_LDG(&rnodes[j]) == start) {
              q.ck0 = j;
              q.ck1 = (j + 1 == t.route_len - 1) ? j : j + 1;
              break;
            }
          }
        }
      }
      q.vflags = on_lane ? (q.vflags | PGD_V_ON_LANE) : (q.vflags & ~PGD_V_ON_LANE);
      if (s != 0 && !on_lane) q.vflags &= ~PGD_V_ALIVE;  // traffic_manager.py:91-109
    }
  }

  // ---- phase F: observation, reward, done -----------------------------------------------------------------------
  {
    const Veh ego = veh[0];
    const PgdSlot& t0 = tpl[0];
    const int32_t* rroads = T.route_roads + t0.route_off;
    // row layout (obs/state_obs.py): [side beams | left, right], 6 state values, [lane-line beams], 10 navi,
    // 16 neighbours, 240 lidar; st / ob are positioned so that the indices of the detector-less layout
    // (state 2..7, navi 8..17, neighbours 18..33, lidar 34..) address the right part of the row
    const int n_first = cfg.n_side > 0 ? cfg.n_side : 2;
    float* const st = obs + n_first - 2;
    const int n_extra = cfg.random_agent_model ? 2 : 0;
    float* const ob = obs + n_first + cfg.n_lane_line + n_extra - 2;
    if (n_extra) {  // obs/state_obs.py:103-105: LENGTH / MAX_LENGTH, WIDTH / MAX_WIDTH (base_vehicle.py:83-84)
      obs[n_first + 6 + cfg.n_lane_line] = clipf(t0.length / 10.0f, 0.0f, 1.0f);
      obs[n_first + 6 + cfg.n_lane_line + 1] = clipf(t0.width / 2.5f, 0.0f, 1.0f);
    }
    // lidar: every chassis in range publishes its rectangle and the arc of beams that can reach it (exact cull, see
    // pgd_step.cu phase F); beam values are computed by lidar_beam() at write-out
    lc.ex = ego.x; lc.ey = ego.y; lc.eh = ego.h;
    lc.n = 0;
    #pragma unroll 1
    for (int s = 1; s < n_slots; ++s) {
      Veh& q = veh[s];
      if (!(q.vflags & PGD_V_ALIVE)) continue;
      const float dx = q.x - ego.x, dy = q.y - ego.y;
      const float d2 = dx * dx + dy * dy;
      const float hd = sqrtf(q.hl * q.hl + q.hw * q.hw);
      const float reach = V2_LIDAR_RANGE + hd;
      if (!(d2 < reach * reach)) continue;
      ensure_heading(q);
      const float d = sqrtf(d2);
      int blo = 0, bn = PGD_LIDAR_BEAMS;
      if (d > hd * 1.001f) {
        const float per_rad = (float)PGD_LIDAR_BEAMS / V2_TWO_PI;
        const float c = (pgd_atan2f(dy, dx) - ego.h) * per_rad;
        const float w = pgd_asinf(fminf(hd / d, 1.0f)) * per_rad;
        const int n = (int)ceilf(2.0f * w) + 3;
        if (n < PGD_LIDAR_BEAMS) {
          bn = n;
          blo = ((int)floorf(c - w) - 1) % PGD_LIDAR_BEAMS;
          if (blo < 0) blo += PGD_LIDAR_BEAMS;
        }
      }
      const int j = lc.n++;
      lc.cx[j] = q.x; lc.cy[j] = q.y; lc.ux[j] = q.hc; lc.uy[j] = q.hs; lc.hl[j] = q.hl; lc.hw[j] = q.hw;
      lc.blo[j] = blo; lc.bn[j] = bn;
    }
    // side / lane-line detectors (distance_detector.py:137-152): ray fans against the line ghosts of the map; a beam
    // looks up the bucket of a point every 8 m along itself (buckets list every box within 4 m of them)
    if (cfg.n_side > 0 || cfg.n_lane_line > 0) {
      const int n_rays = cfg.n_side + cfg.n_lane_line;
      const int32_t* ent = T.cell_entries + mp.entry_off;
      #pragma unroll 1
      for (int rI = 0; rI < n_rays; ++rI) {
        const bool side = rI < cfg.n_side;
        const int i = side ? rI : rI - cfg.n_side;
        const int n = side ? cfg.n_side : cfg.n_lane_line;
        const float dist = side ? cfg.side_distance : cfg.lane_line_distance;
        const float ang = (float)i * (V2_TWO_PI / (float)n) + V2_PI / 2 + ego.h;
        float sn, cs;
        V2_SINCOS(ang, sn, cs);
        const float dx = cs * dist, dy = sn * dist;
        float best = 1.0f;
        for (float sd = 4.0f; sd - 4.0f < dist; sd += 8.0f) {
          if (best * dist < sd - 4.0f) break;
          const float px = ego.x + cs * sd, py = ego.y + sn * sd;
          const int cx = (int)floorf((px - mp.x0) * mp.inv_cell), cy = (int)floorf((py - mp.y0) * mp.inv_cell);
          if (cx < 0 || cy < 0 || cx >= mp.nx || cy >= mp.ny) continue;
          const int cell = mp.cell_off + cy * mp.nx + cx;
          const int b0 = V2_LDG(&T.cell_start[cell]), b1 = V2_LDG(&T.cell_start[cell + 1]);
          #pragma unroll 1
          for (int k = b0; k < b1; ++k) {
            const PgdBox g = load_rec(boxes + V2_LDG(&ent[k]));
            if (!(g.kind == PGD_BOX_WHITE || g.kind == PGD_BOX_YELLOW || (!side && g.kind == PGD_BOX_BROKEN))) continue;
            const Rect r = {g.cx, g.cy, g.ux, g.uy, g.hl, g.hw};
            best = fminf(best, ray_rect(ego.x, ego.y, dx, dy, r));
          }
        }
        if (side) obs[i] = best;
        else obs[n_first + 6 + i] = best;
      }
    }
    // the 4 nearest vehicles inside the 50 m cylinder (ties -> lower slot)
    {
      float d2s[V];
      #pragma unroll 1
      for (int s = 1; s < n_slots; ++s) {
        d2s[s] = INFINITY;
        if (!(veh[s].vflags & PGD_V_ALIVE)) continue;
        const float dx = veh[s].x - ego.x, dy = veh[s].y - ego.y;
        const float d2 = dx * dx + dy * dy;
        if (d2 < V2_LIDAR_RANGE * V2_LIDAR_RANGE) d2s[s] = d2;
      }
      const float esp = clipf(ego.v * 3.6f, 0.0f, 100000.0f);
      #pragma unroll 1
      for (int rank = 0; rank < 4; ++rank) {
        int best = -1;
        #pragma unroll 1
        for (int s = 1; s < n_slots; ++s)
          if (d2s[s] < INFINITY && (best < 0 || d2s[s] < d2s[best])) best = s;
        float* o4 = ob + 18 + 4 * rank;
        if (best < 0) {
          o4[0] = o4[1] = o4[2] = o4[3] = 0.0f;
          continue;
        }
        d2s[best] = INFINITY;
        Veh& q = veh[best];
        ensure_heading(q);
        float pf, ps, vf, vs;
        project(ego.hc, ego.hs, q.x - ego.x, q.y - ego.y, pf, ps);
        const float ws = clipf(q.v * 3.6f, 0.0f, 100000.0f);
        project(ego.hc, ego.hs, ws * q.hc - esp * ego.hc, ws * q.hs - esp * ego.hs, vf, vs);
        o4[0] = clipf((pf / V2_LIDAR_RANGE + 1.0f) / 2.0f, 0.0f, 1.0f);
        o4[1] = clipf((ps / V2_LIDAR_RANGE + 1.0f) / 2.0f, 0.0f, 1.0f);
        o4[2] = clipf((vf / V2_MAX_SPEED_KMH + 1.0f) / 2.0f, 0.0f, 1.0f);
        o4[3] = clipf((vs / V2_MAX_SPEED_KMH + 1.0f) / 2.0f, 0.0f, 1.0f);
      }
    }
    // ego bookkeeping
    const int cur_road_id = V2_LDG(&rroads[ego.ck0]);
    const PgdRoad cur_road = load_rec(roads + (cur_road_id));
    const PgdRoad fr = load_rec(roads + (V2_LDG(&rroads[t0.route_len - 2])));
    const int el_road = V2_LDG(&lanes[ego.lane].road);
    const bool use_ego_lane = el_road == cur_road_id;
    const int reward_lane = use_ego_lane ? ego.lane : cur_road.first_lane;
    const int n_ref = cur_road.n_lanes;
    const int sign_i = use_ego_lane ? 0 : (V2_LDG(&roads[el_road].negative) ? -1 : 1);
    float qlon0, qlat0, qlon1, qlat1, long_last, lat_last, long_now, lat_now;
    lane_local(load_rec(lanes + cur_road.first_lane), ego.x, ego.y, qlon0, qlat0);
    const PgdLane final_lane = load_rec(lanes + (fr.first_lane + fr.n_lanes - 1));
    lane_local(final_lane, ego.x, ego.y, qlon1, qlat1);
    {
      const PgdLane rl = load_rec(lanes + (reward_lane));
      lane_local(rl, last_x, last_y, long_last, lat_last);
      lane_local(rl, ego.x, ego.y, long_now, lat_now);
    }
    #pragma unroll 1
    for (int c = 0; c < 2; ++c) {  // navigation.py:213-260
      const PgdLane l = load_rec(lanes + (c == 0 ? cur_road.first_lane : V2_LDG(&roads[V2_LDG(&rroads[ego.ck1])].first_lane)));
      const float later_middle = ((float)n_ref / 2.0f - 0.5f) * mp.lane_width;
      float px, py;
      lane_position(l, l.length, later_middle, px, py);
      float dx = px - ego.x, dy = py - ego.y;
      const float dn = sqrtf(dx * dx + dy * dy);
      if (dn > 50.0f) { dx = dx / dn * 50.0f; dy = dy / dn * 50.0f; }
      float ph, ps;
      project(ego.hc, ego.hs, dx, dy, ph, ps);
      float bend = 0.0f, dir = 0.0f, angle = 0.0f;
      if (l.kind == PGD_LANE_ARC) {
        bend = l.radius / (60.0f + (float)n_ref * mp.lane_width);
        dir = l.dir;
        angle = l.length / l.radius;
      }
      float* q = ob + 8 + 5 * c;
      q[0] = clipf((ph / 50.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
      q[1] = clipf((ps / 50.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
      q[2] = clipf(bend, 0.0f, 1.0f);
      q[3] = clipf((dir + 1.0f) / 2.0f, 0.0f, 1.0f);
      q[4] = clipf((angle * (180.0f / V2_PI) / 135.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
    }
    {  // heading_diff (base_vehicle.py:433-458)
      const PgdLane l = load_rec(lanes + (cur_road.first_lane + cur_road.n_lanes - 1));
      float lx, ly;
      if (l.kind == PGD_LANE_STRAIGHT) { lx = -l.ay; ly = l.ax; }
      else if (l.dir < 0.0f) { lx = ego.x - l.ax; ly = ego.y - l.ay; }
      else { lx = l.ax - ego.x; ly = l.ay - ego.y; }
      const float ln = sqrtf(lx * lx + ly * ly);
      st[2] = ln > 0.0f ? clipf((ego.hc * lx + ego.hs * ly) / ln, -1.0f, 1.0f) / 2.0f + 0.5f : 0.0f;
    }
    const bool on_lane = (ego.vflags & PGD_V_ON_LANE) != 0;
    if (on_lane) flags |= PGD_F_ON_LANE;
    if (crash) flags |= PGD_F_CRASH_VEHICLE;
    const float to_left = qlat0 + mp.lane_width / 2.0f;
    const float to_right = mp.lane_width * (float)n_ref - to_left;
    if (to_left < 0.0f || to_right < 0.0f) flags |= PGD_F_OUT_OF_ROUTE;
    {
      const float flen = final_lane.length;
      if (flen - 5.0f < qlon1 && qlon1 < flen + 5.0f && mp.lane_width / 2.0f >= qlat1 &&
          qlat1 >= (0.5f - (float)n_ref) * mp.lane_width)
        flags |= PGD_F_ARRIVE_DEST;
    }
    bool out_of_road = (flags & (PGD_F_ON_YELLOW | PGD_F_ON_WHITE | PGD_F_CRASH_SIDEWALK)) || !on_lane;
    if (cfg.out_of_route_done && (flags & PGD_F_OUT_OF_ROUTE)) out_of_road = true;
    if (out_of_road) flags |= PGD_F_OUT_OF_ROAD;

    const float sp = clipf(ego.v * 3.6f, 0.0f, 100000.0f);
    if (cfg.n_side <= 0) {
      obs[0] = clipf(to_left / 18.0f, 0.0f, 1.0f);
      obs[1] = clipf(to_right / 18.0f, 0.0f, 1.0f);
    }
    st[3] = clipf((sp + 1.0f) / (V2_MAX_SPEED_KMH + 1.0f), 0.0f, 1.0f);
    st[4] = clipf((ego.steer / 60.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
    st[5] = clipf((envf.x + 1.0f) / 2.0f, 0.0f, 1.0f);
    st[6] = clipf((envf.y + 1.0f) / 2.0f, 0.0f, 1.0f);
    st[7] = clipf(fminf(fabsf(wrap_to_pi(ego.h - last_h)), V2_PI / 2) / 0.1f, 0.0f, 1.0f);
    float r = 0.0f, step_reward = 0.0f, cost = 0.0f, step_energy = 0.0f;
    int is_done = 0;
    if (!fresh) {
      const float sign = sign_i == 0 ? 1.0f : (float)sign_i;
      float lateral_factor = 1.0f;
      if (cfg.use_lateral) lateral_factor = clipf(1.0f - 2.0f * fabsf(lat_now) / mp.lane_width, 0.0f, 1.0f);
      r += cfg.driving_reward * (long_now - long_last) * lateral_factor * sign;
      r += cfg.speed_reward * (sp / V2_MAX_SPEED_KMH) * sign;
      step_reward = r;
      if (flags & PGD_F_ARRIVE_DEST) r = cfg.success_reward;
      else if (out_of_road) r = -cfg.out_of_road_penalty;
      else if (crash) r = -cfg.crash_vehicle_penalty;
      if (out_of_road) cost = cfg.out_of_road_cost;
      else if (crash) cost = cfg.crash_vehicle_cost;
      is_done = ((flags & PGD_F_ARRIVE_DEST) || out_of_road || crash) ? 1 : 0;
      const float ddx = last_x - ego.x, ddy = last_y - ego.y;
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
    if (mode == 0) {
      *reward = r;
      *done = (uint8_t)is_done;
    }
    if (info) {
      info->velocity = sp; info->steering = ego.steer; info->acceleration = ego.throttle;
      info->step_energy = step_energy; info->episode_energy = envf.w;
      info->step_reward = step_reward; info->episode_reward = envf.z; info->cost = cost;
      info->episode_length = envi.w; info->flags = flags;
    }
    S.envi[env] = envi;
    S.envf[env] = envf;
  }

  // ---- phase G: store --------------------------------------------------------------------------------------------
  // Only what can have changed is written back: parked traffic has at most run its drop counter, vehicles removed in
  // an earlier step are not touched at all.  A freshly started episode writes every slot once.
  #pragma unroll 1
  for (int s = 0; s < V; ++s) {
    const size_t gi = (size_t)s * num_envs + env;
    if (s < n_slots) {
      const Veh& q = veh[s];
      if ((untouched >> s) & 1u) continue;
      const int vflags = q.vflags & ~V2_HDG_VALID;
      const I4 m = {q.rnd_n, q.airborne, vflags, 0};
      const bool still_parked = ((was_parked >> s) & 1u) && !(vflags & PGD_V_ACTIVE);
      if (still_parked && !fresh) {
        if ((drop_ran >> s) & 1u) S.misc[gi] = m;
        continue;
      }
      F4 p = {q.x, q.y, q.h, q.v}, c = {q.steer, q.throttle, q.hp, q.hi}, l = {q.lp, q.li, q.tspeed, q.yaw};
      I4 n = {q.lane, q.ck0 | (q.ck1 << 16), q.rt_lane, q.timer};
      S.pose[gi] = p; S.ctrl[gi] = c; S.pidl[gi] = l; S.nav[gi] = n; S.misc[gi] = m;
    } else if (fresh) {  // unused slots of a freshly started episode: clear the flags once
      I4 m = {0, 0, 0, 0};
      S.misc[gi] = m;
    }
  }
}

}  // namespace pgdv2
#endif
