/* Environment step, ROLE PER WARP / ENVIRONMENT PER LANE.
 *
 * A CTA advances 32 environments with R warps.  Lane l of every warp works on environment l of the CTA; the warp
 * index is the ROLE: warp 0 carries the 32 egos, warps 1..R-1 carry the traffic slots (slot s belongs to role
 * 1 + (s - 1) % (R - 1)).  An instruction of warp 0 therefore advances 32 egos (the round-1 kernel advanced 2 per
 * instruction, its one-thread-per-environment variant 32 but with 2 048 warps on the whole GPU), and a traffic warp
 * only spends instructions on vehicles that are awake.  What the vehicles of an environment need from each other --
 * poses for IDM's neighbour search, the ego's trajectory through the sub-steps for the chassis contacts, final poses
 * for lidar and the neighbour features -- goes through shared memory, structure-of-arrays with the environment as
 * the fastest index (conflict-free), between CTA barriers:
 *
 *   A  publish start-of-step public state of every slot; ego action; traffic trigger        (all roles)
 *   B  lanes with awake traffic: every vehicle's coordinate on its own lane (IDM look-up data)
 *   C  role 0: ego sub-steps -> trajectory;  traffic roles: IDM / PID of their awake vehicles
 *   D  role 0: ego localisation + line / sidewalk contacts;  traffic: sub-steps, chassis contact against the ego
 *      trajectory, localisation, removal; final poses published; observation rows pre-filled with 1.0
 *   F  role-parallel ego bookkeeping: reward / done / state (0), navigation info (1), 4 nearest vehicles (2),
 *      lidar windows (3); detector ray fans over all roles
 *   L  lidar as a scatter: every (visible chassis, beam of its window) pair is one work item; lanes = beams
 *   W  the 32 rows leave the CTA with ONE bulk (TMA) shared -> global copy (pgd_step_kernel.cu)
 *
 * State in HBM is slot-major ([slot][env]) so that a warp's loads / stores of a slot are 32 consecutive 16-byte
 * vectors.  Parked traffic is never loaded beyond its flag word (its pose is the episode template's).
 *
 * The arithmetic is the oracle's, expression for expression.  The file compiles for the host too
 * (oracle/step_host.cpp runs the phases in order over all (role, lane) pairs), so the whole step is checked bit
 * for bit against the independent CPU oracle without a GPU (tests/test_step_host.py).
 *
 * Reference call stack (paths under /root/reference/pgdrive): envs/base_env.py:184-224,303-344 (step),
 * policy/idm_policy.py:83-353 (IDM), engine/base_engine.py:206-232 (sub-steps), vehicle_module/navigation.py:155-344,
 * utils/scene_utils.py:138-185 (localisation), component/vehicle/base_vehicle.py:615-644 (line / sidewalk contacts),
 * cutils.pyx:60-142 + vehicle_module/lidar.py:55-77 (lidar, neighbours), obs/state_obs.py:58-170 (observation),
 * envs/pgdrive_env.py:162-258 (reward / cost / done).
 */
#ifndef PGD_STEP_CUH
#define PGD_STEP_CUH
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/pgd_math.h"
#include "../../include/pgd_tables.h"

#ifdef __CUDACC__
#define PGS_HD __host__ __device__ __forceinline__
#ifdef PGS_OUTLINE_HELPERS  // experiment: sincos / arc projection as real calls (one copy each)
#define PGS_HD_OUTLINE __host__ __device__ __noinline__
#else
#define PGS_HD_OUTLINE __host__ __device__ __forceinline__
#endif
#else
#define PGS_HD inline
#define PGS_HD_OUTLINE inline
#endif

#ifndef PGS_ITEM_CLK  // diagnostic build of the kernel only (pgd_step_kernel.cu, PGS_PHASE_CLOCKS)
#define PGS_ITEM_CLK_BEGIN
#define PGS_ITEM_CLK(i)
#define PGS_ITEM_COUNT(i, n)
#define PGS_EGO_CLK_BEGIN
#define PGS_EGO_CLK(i)
#endif

namespace pgdstep {

#define PGS_PI 3.14159265358979323846f
#define PGS_TWO_PI 6.28318530717958647692f
#define PGS_GRAVITY 9.81f
#define PGS_LIDAR_RANGE 50.0f
#define PGS_MAX_SPEED_KMH 80.0f
#define PGS_IDM_MAX_LONG 30.0f
#define PGS_IDM_NORMAL_SPEED 30.0f
#define PGS_IDM_CREEP_SPEED 5.0f
#define PGS_IDM_SAFE_DIST 15.0f
#define PGS_IDM_LANE_CHANGE_FREQ 50
#define PGS_IDM_SPEED_INCREASE 10.0f
#define PGS_IDM_MAX_SPEED 100.0f
#define PGS_YAW_TAU 0.1f
#define PGS_DONE_PENDING_RESET 2
#define PGS_MAX_SUBSTEPS 8 /* decision_repeat supported (default 5) */
#define PGS_LANES 32       /* environments per CTA = lanes of a warp */
#define PGS_MAX_ROLES 8
/* internal bit of the per-slot flag word: pgd_set_state wrote this parked vehicle's pose, read it from the state
 * instead of the episode template (masked out of pgd_get_state) */
#define PGS_V_POSE_SET 8

struct alignas(16) F4 { float x, y, z, w; };
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

#if defined(__CUDA_ARCH__) && defined(PGS_TABLE_EVICT_LAST)
// read-only table data with an L2 evict_last policy: the tables (7 MB for 100 maps) are re-read by every CTA of every
// step while 100+ MB of state and observation rows stream through the L2 in between
__device__ __forceinline__ uint64_t table_policy() {
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint4 ldg16_keep(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p), "l"(table_policy()));
  return v;
}
#endif

template <class T_>
PGS_HD T_ ldg(const T_* p) {
#if defined(__CUDA_ARCH__) && !defined(PGS_PLAIN_LOADS)
  return __ldg(p);
#else
  return *p;  // PGS_PLAIN_LOADS (map-staging experiment): the pointer may lead into shared memory
#endif
}

/* A whole table record (PgdLane 64 B, PgdBox / PgdRoad 32 B, PgdMap 64 B) with 16-byte read-only loads. */
template <class T_>
PGS_HD T_ load_rec(const T_* p) {
#ifdef __CUDA_ARCH__
  static_assert(sizeof(T_) % 16 == 0, "table records are multiples of 16 bytes");
  T_ out;
  const uint4* src = reinterpret_cast<const uint4*>(p);
  uint4* dst = reinterpret_cast<uint4*>(&out);
#pragma unroll
#if defined(PGS_PLAIN_LOADS)
  for (int i = 0; i < (int)(sizeof(T_) / 16); ++i) dst[i] = src[i];
#elif defined(PGS_TABLE_EVICT_LAST)
  for (int i = 0; i < (int)(sizeof(T_) / 16); ++i) dst[i] = ldg16_keep(src + i);
#else
  for (int i = 0; i < (int)(sizeof(T_) / 16); ++i) dst[i] = __ldg(src + i);
#endif
  return out;
#else
  return *p;
#endif
}

PGS_HD float clipf(float a, float lo, float hi) { return fminf(fmaxf(a, lo), hi); }

// out of line (one copy; ~45 instructions), results by value so that nothing goes through local memory
struct SinCos { float s, c; };
PGS_HD_OUTLINE SinCos sincos_hd(float a) {
  SinCos r;
  pgd_sincosf(a, &r.s, &r.c);
  return r;
}
#define PGS_SINCOS(a, s_out, c_out)                 \
  do {                                              \
    const pgdstep::SinCos sc_ = pgdstep::sincos_hd(a); \
    (s_out) = sc_.s;                                \
    (c_out) = sc_.c;                                \
  } while (0)

struct LonLat { float lon, lat; };
#ifdef PGS_OUTLINE_ARC
__host__ __device__ __noinline__
#else
PGS_HD_OUTLINE
#endif
LonLat arc_local(float cx, float cy, float ph0, float dir, float radius, float x, float y) {
  float dx = x - cx, dy = y - cy;
  float phi = pgd_atan2f(dy, dx);
  phi = ph0 + pgd_wrap_to_pi(phi - ph0);
  float r = sqrtf(dx * dx + dy * dy);
  LonLat o;
  o.lon = dir * (phi - ph0) * radius;
  o.lat = dir * (radius - r);
  return o;
}

PGS_HD void lane_local(const PgdLane& l, float x, float y, float& lon, float& lat) {
  if (l.kind == PGD_LANE_STRAIGHT) {
    float dx = x - l.sx, dy = y - l.sy;
    lon = dx * l.ax + dy * l.ay;
    lat = dx * -l.ay + dy * l.ax;
  } else {
    const LonLat o = arc_local(l.ax, l.ay, l.ph0, l.dir, l.radius, x, y);
    lon = o.lon;
    lat = o.lat;
  }
}

PGS_HD void lane_position(const PgdLane& l, float lon, float lat, float& x, float& y) {
  if (l.kind == PGD_LANE_STRAIGHT) {
    x = l.sx + lon * l.ax + lat * -l.ay;
    y = l.sy + lon * l.ay + lat * l.ax;
  } else {
    float phi = l.dir * lon / l.radius + l.ph0;
    float r = l.radius - lat * l.dir;
    float s, c;
    PGS_SINCOS(phi, s, c);
    x = l.ax + r * c;
    y = l.ay + r * s;
  }
}

PGS_HD float lane_heading_at(const PgdLane& l, float lon) {
  if (l.kind == PGD_LANE_STRAIGHT) return l.heading;
  float phi = l.dir * lon / l.radius + l.ph0;
  return phi + PGS_PI / 2 * l.dir;
}

PGS_HD bool precedes(float ex, float ey, float sx, float sy) {  // abs_lane.py:114-119 (norm < 0.1)
  float dx = ex - sx, dy = ey - sy;
  return dx * dx + dy * dy < 1e-2f;
}

struct Rect { float cx, cy, ux, uy, hl, hw; };

PGS_HD bool rect_overlap(const Rect& a, const Rect& b) {
  float dx = b.cx - a.cx, dy = b.cy - a.cy;
  float c = fabsf(a.ux * b.ux + a.uy * b.uy);
  float s = fabsf(a.ux * b.uy - a.uy * b.ux);
  if (fabsf(dx * a.ux + dy * a.uy) > a.hl + b.hl * c + b.hw * s) return false;
  if (fabsf(-dx * a.uy + dy * a.ux) > a.hw + b.hl * s + b.hw * c) return false;
  if (fabsf(dx * b.ux + dy * b.uy) > b.hl + a.hl * c + a.hw * s) return false;
  if (fabsf(-dx * b.uy + dy * b.ux) > b.hw + a.hl * s + a.hw * c) return false;
  return true;
}

PGS_HD float ray_rect(float ox, float oy, float dx, float dy, const Rect& r) {
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

PGS_HD void project(float hx, float hy, float vx, float vy, float& fwd, float& side) {  // base_vehicle.py:460-475
  const float n = 1.0f + 1e-6f;
  fwd = (vx * hx + vy * hy) / n;
  side = (vx * -hy + vy * hx) / n;
}

PGS_HD float pid(float& p_err, float& i_err, float kp, float ki, float kd, float err) {  // PID_controller.py
  i_err += err;
  float d = err - p_err;
  p_err = err;
  return -kp * p_err - ki * i_err - kd * d;
}

PGS_HD float kmh(float v) { return clipf(v * 3.6f, 0.0f, 100000.0f); }  // base_vehicle.py:395-401

struct Veh {  // one vehicle, in registers while its owner works on it
  float x, y, h, v, yaw, steer, throttle, hp, hi, lp, li, tspeed;
  float hc, hs, hl, hw;
  int lane, ck0, ck1, rt_lane, timer, rnd_n, airborne, vflags;
};

struct Sub {  // what one physics sub-step needs, hoisted out of the sub-step loop
  float accel;     // > 0: engine acceleration [m/s^2]; else brake
  float brake_dv;  // speed removed per sub-step when braking
  float k_yaw;     // kinematic yaw rate per unit speed: sin(slip angle) / lr
  float k_relax;   // dt / tau of the yaw-rate relaxation
  float mu_g, lr;
};

PGS_HD void substep(Veh& q, const Sub& sub, float dt) {  // planar stand-in for BulletVehicle
  float speed = q.v;
  if (sub.accel > 0.0f) speed += sub.accel * dt;
  else speed = fmaxf(speed - sub.brake_dv, 0.0f);
  // yaw rate relaxes towards the kinematic-bicycle value (tyre relaxation + yaw inertia, tau = 0.1 s) ...
  float yaw = q.yaw + (speed * sub.k_yaw - q.yaw) * sub.k_relax;
  // ... and the tyres cannot give more than mu * g of lateral acceleration
  if (speed * fabsf(yaw) > sub.mu_g) yaw = copysignf(sub.mu_g / speed, yaw);
  const float sb = speed > 1e-3f ? clipf(yaw * sub.lr / speed, -1.0f, 1.0f) : 0.0f;
  const float cb = sqrtf(fmaxf(1.0f - sb * sb, 0.0f));
  q.x += speed * (q.hc * cb - q.hs * sb) * dt;
  q.y += speed * (q.hs * cb + q.hc * sb) * dt;
  float nh = q.h + yaw * dt;
  if (nh > PGS_PI) nh -= PGS_TWO_PI;
  if (nh < -PGS_PI) nh += PGS_TWO_PI;
  q.yaw = yaw;
  if (nh != q.h) PGS_SINCOS(nh, q.hs, q.hc);
  q.h = nh;
  q.v = speed;
}

PGS_HD Sub make_sub(const Veh& q, const PgdSlot& t, float dt) {  // base_vehicle.py:343-376
  Sub sub;
  sub.mu_g = t.friction * PGS_GRAVITY;
  sub.lr = t.lr;
  sub.k_relax = dt / PGS_YAW_TAU;
  const bool overspeed = kmh(q.v) > PGS_MAX_SPEED_KMH;
  if (q.throttle > 0.0f && !overspeed) {  // engine force on 4 wheels; Bullet ignores the brake when it is non-zero
    sub.accel = fminf(4.0f * t.max_engine * q.throttle / t.mass, sub.mu_g);
    sub.brake_dv = 0.0f;
  } else {  // per-wheel brake impulse: 2.0 idle, |throttle| * max_brake_force when braking
    sub.accel = 0.0f;
    const float imp = q.throttle >= 0.0f ? 2.0f : -q.throttle * t.max_brake;
    sub.brake_dv = fminf(4.0f * imp / t.mass, sub.mu_g * dt);
  }
  const float delta = clipf(-q.steer * t.max_steer, -1.4f, 1.4f);  // +steering = left = heading decreases
  const float tb = t.lr / (t.lf + t.lr) * pgd_tanf(delta);
  sub.k_yaw = tb / sqrtf(1.0f + tb * tb) / t.lr;
  return sub;
}

// ---- shared memory ------------------------------------------------------------------------------------------------
template <int V>
struct Pub {  // public per-slot state, [slot][lane]; lf = lane << 8 | PGD_V_* flags
  float x[V][PGS_LANES], y[V][PGS_LANES], hc[V][PGS_LANES], hs[V][PGS_LANES], v[V][PGS_LANES];
  int lf[V][PGS_LANES];
};
PGS_HD int lf_pack(int lane, int fl) { return (lane << 8) | (fl & 0xff); }
PGS_HD int lf_lane(int lf) { return lf >> 8; }

template <int V>
struct IdmPub {  // what IDM reads of OTHER vehicles (phases A, B -> X); shares its storage with the observation rows
  // coordinate of every vehicle on its own lane and that lane's ends / length (phase B)
  float olong[V][PGS_LANES], lsx[V][PGS_LANES], lsy[V][PGS_LANES], lex[V][PGS_LANES], ley[V][PGS_LANES],
      llen[V][PGS_LANES];
  // start-of-step copy of the public state (phase A): phase X moves the vehicles in Smem::p while other vehicles'
  // IDM still has to see where everybody was when the step began
  Pub<V> start;
};

/* Dynamic shared memory: [Smem, fixed part][rows: 32 x obs_dim floats in HBM layout; phases A .. X: IdmPub][tv: phase
 * X the ego's pose (x, y, cos, sin) after every sub-step, F4 [ns][32]; phases F, L the visible chassis, int [V][32]:
 * slot | first beam << 8 | beam count << 16]. */
typedef F4 (*TrajPtr)[PGS_LANES];
typedef int (*VisPtr)[PGS_LANES];
template <int V, int R>
struct Smem {
  Pub<V> p;
  F4 efin[PGS_LANES];                    // ego pose after the sub-steps (template pose when the episode restarts)
  int ego_lane[PGS_LANES];               // the ego's lane after its localisation (phase X, role 0)
  I4 envi[PGS_LANES];                    // per-environment counters / sums after phase A, parked here until phase F
  F4 envf[PGS_LANES];                    // (role 0 would carry them through phase X in registers, i.e. in local memory)
  float last_h[PGS_LANES];               // the ego's heading at the start of the step
  int amask[PGS_LANES], pmask[PGS_LANES], crash[PGS_LANES];  // awake / parked traffic slots, chassis contact (or-ed in)
  float last_x[PGS_LANES], last_y[PGS_LANES], ego_travel[PGS_LANES], ego_h[PGS_LANES], ego_v[PGS_LANES],
      ego_hl[PGS_LANES], ego_hw[PGS_LANES];
  int ego_ck[PGS_LANES], n_vis[PGS_LANES], wrote[PGS_LANES];
  // table context of each environment (role 0 writes it before phase A): after phase A nobody keeps map offsets or
  // table pointers in registers, they are read here where they are needed
  float cx_x0[PGS_LANES], cx_y0[PGS_LANES], cx_inv_cell[PGS_LANES], cx_lane_width[PGS_LANES];
  int cx_nx[PGS_LANES], cx_ny[PGS_LANES], cx_cell_off[PGS_LANES], cx_entry_off[PGS_LANES];
  int cx_lane_off[PGS_LANES], cx_road_off[PGS_LANES], cx_box_off[PGS_LANES], cx_slot_off[PGS_LANES], cx_n_slots[PGS_LANES];
  int n_work;
  uint16_t work[(V - 1) * PGS_LANES];    // awake traffic of the whole CTA: lane | slot << 8, by environment then slot
};

template <int V, int R>
PGS_HD constexpr size_t smem_obs_offset() { return (sizeof(Smem<V, R>) + 127) / 128 * 128; }
template <int V, int R>
PGS_HD size_t smem_tv_offset(int obs_dim) {
  const size_t rows = (size_t)PGS_LANES * obs_dim * sizeof(float);
  return smem_obs_offset<V, R>() + ((rows > sizeof(IdmPub<V>) ? rows : sizeof(IdmPub<V>)) + 127) / 128 * 128;
}
template <int V, int R>
PGS_HD size_t smem_bytes(int obs_dim, int substeps) {
  const size_t traj = (size_t)substeps * PGS_LANES * sizeof(F4), vis = (size_t)V * PGS_LANES * sizeof(int);
  return smem_tv_offset<V, R>(obs_dim) + (traj > vis ? traj : vis);
}

template <int V, int R>
struct Thr {  // what a thread keeps across the phases
  static constexpr int MAXOWN = (V - 1 + R - 2) / (R - 1);  // slots a traffic role publishes
  int lane, role, env, num_envs;
  bool valid, fresh, stepping;
  I4 envi;
  F4 envf;
  PgdMap mp;
  const PgdEpisode* ep;
  const PgdLane* lanes;
  const PgdRoad* roads;
  const PgdBox* boxes;
  const PgdSlot* tpl;
  int n_slots, n_groups, trig;
  Veh ego;         // role 0
  uint32_t flags;  // role 0: PGD_F_* of the ego
  // computed during phase X by warps that would otherwise wait, written to the observation row in phase F (the rows'
  // storage holds IDM's data until phase X ends), role 0: keep[0..9] navigation info, keep[10] heading difference,
  // keep[11..14] lateral distances (obs 0, 1), driving reward, route sign
  float keep[15];
  // loads that depend on nothing but the environment index, issued before the table look-ups they overlap with
  F4 pre_pose, pre_ctrl, pre_pidl;  // role 0: the ego's record
  I4 pre_nav, pre_misc;             // role 0 (all roles: pre_nav.x = the ego's lane, for the trigger test)
  // traffic roles: PGD_V_* flags (5 bits each) of the slots they publish, packed so that phase A's loop over them needs no
  // register array, i.e. no unrolling (the kernel is as large as the instruction cache)
  static constexpr int FLW = (MAXOWN + 11) / 12;  // 12 slots per word; one word for the 4 warps the kernel is built with
  uint64_t pre_fl[FLW];
};

PGS_HD int imin(int a, int b) { return a < b ? a : b; }
PGS_HD void smem_or(int* p, int v) {  // shared-memory word combined by several roles
#ifdef __CUDA_ARCH__
  if (v) atomicOr(p, v);
#else
  *p |= v;
#endif
}
PGS_HD void smem_min(int* p, int v) {
#ifdef __CUDA_ARCH__
  atomicMin(p, v);
#else
  if (v < *p) *p = v;
#endif
}
PGS_HD int ctz32(uint32_t m) {
#ifdef __CUDA_ARCH__
  return __ffs((int)m) - 1;
#else
  return __builtin_ctz(m);
#endif
}
PGS_HD uint32_t drop_low(uint32_t m, int n) {  // clear the n lowest set bits
  for (int i = 0; i < n && m; ++i) m &= m - 1;
  return m;
}

PGS_HD void veh_from_template(Veh& q, const PgdSlot& t, int s) {
  q.x = t.x; q.y = t.y; q.h = t.heading; q.v = 0.0f; q.yaw = 0.0f;
  q.steer = q.throttle = q.hp = q.hi = q.lp = q.li = 0.0f;
  q.tspeed = PGS_IDM_NORMAL_SPEED;
  q.lane = t.lane; q.ck0 = 0; q.ck1 = t.route_len > 2 ? 1 : 0; q.rt_lane = -1;
  q.timer = t.overtake_timer; q.rnd_n = 0; q.airborne = t.drop_substeps;
  q.vflags = PGD_V_ALIVE | PGD_V_ON_LANE | ((s == 0 || t.group == PGD_GROUP_AWAKE) ? PGD_V_ACTIVE : 0);
}

PGS_HD void veh_unpack(Veh& q, const F4& p, const F4& c, const F4& l, const I4& n, const I4& m) {
  q.x = p.x; q.y = p.y; q.h = p.z; q.v = p.w;
  q.steer = c.x; q.throttle = c.y; q.hp = c.z; q.hi = c.w;
  q.lp = l.x; q.li = l.y; q.tspeed = l.z; q.yaw = l.w;
  q.lane = n.x; q.ck0 = n.y & 0xffff; q.ck1 = n.y >> 16; q.rt_lane = n.z; q.timer = n.w;
  q.rnd_n = m.x; q.airborne = m.y; q.vflags = m.z;
}

PGS_HD void veh_load(Veh& q, const State& S, size_t gi) {
  const F4 p = S.pose[gi], c = S.ctrl[gi], l = S.pidl[gi];
  const I4 n = S.nav[gi], m = S.misc[gi];
  veh_unpack(q, p, c, l, n, m);
}

PGS_HD void veh_store(const Veh& q, const State& S, size_t gi) {
  const F4 p = {q.x, q.y, q.h, q.v}, c = {q.steer, q.throttle, q.hp, q.hi}, l = {q.lp, q.li, q.tspeed, q.yaw};
  const I4 n = {q.lane, q.ck0 | (q.ck1 << 16), q.rt_lane, q.timer}, m = {q.rnd_n, q.airborne, q.vflags, 0};
  S.pose[gi] = p; S.ctrl[gi] = c; S.pidl[gi] = l; S.nav[gi] = n; S.misc[gi] = m;
}

template <int V, int R>
PGS_HD IdmPub<V>& idm_of(float* obs) { return *reinterpret_cast<IdmPub<V>*>(obs); }
template <int V, int R>
PGS_HD const IdmPub<V>& idm_of(const float* obs) { return *reinterpret_cast<const IdmPub<V>*>(obs); }

// ---- thread set-up ---------------------------------------------------------------------------------------------------
template <int V, int R>
PGS_HD void thread_init(Thr<V, R>& th, const Tables& T, const State& S, const PgdConfig& cfg, int mode, int lane,
                       int role, int env, int env_end) {
  th.lane = lane; th.role = role; th.env = env; th.num_envs = cfg.num_envs;
  th.valid = env < env_end;
  th.fresh = th.stepping = false;
  th.trig = -1;
  th.flags = 0;
  if (!th.valid) return;
  th.envi = S.envi[env];
  // everything below that only needs the environment index is requested now, so that it is in flight while the
  // episode -> map -> template look-ups (dependent loads) run
  th.pre_nav = S.nav[env];  // slot 0
#pragma unroll
  for (int w = 0; w < Thr<V, R>::FLW; ++w) th.pre_fl[w] = 0;
  if (role == 0) {
    th.envf = S.envf[env];
    th.pre_pose = S.pose[env]; th.pre_ctrl = S.ctrl[env]; th.pre_pidl = S.pidl[env]; th.pre_misc = S.misc[env];
  } else {
#pragma unroll
    for (int k = 0; k < Thr<V, R>::MAXOWN; ++k) {
      const int s = role + k * (R - 1);
      I4 m = {0, 0, 0, 0};
      if (s < V) m = S.misc[(size_t)s * cfg.num_envs + env];
      th.pre_fl[k / 12] |= (uint64_t)(m.z & 31) << (5 * (k % 12));
    }
  }
  const bool pending = th.envi.z == PGS_DONE_PENDING_RESET;
  if (mode == 1) {
    th.fresh = pending;
    if (!pending) { th.valid = false; return; }  // the reset pass only touches environments marked for it
  } else {
    th.fresh = pending || (cfg.auto_reset && th.envi.z == 1);
  }
  th.stepping = !th.fresh;
  th.ep = T.episodes + th.envi.x;
  th.mp = load_rec(T.maps + ldg(&th.ep->map));
  th.n_slots = ldg(&th.ep->n_slots);
  th.n_groups = ldg(&th.ep->n_groups);
  th.lanes = T.lanes + th.mp.lane_off;
  th.roads = T.roads + th.mp.road_off;
  th.boxes = T.boxes + th.mp.box_off;
  th.tpl = T.slots + ldg(&th.ep->slot_off);
  if (th.fresh) {
    th.envi.y = 0; th.envi.z = 0; th.envi.w = 0;
    th.envf.x = th.envf.y = th.envf.z = th.envf.w = 0.0f;
  }
}

// ---- before phase A (same barrier interval as thread_init): clear the words that several roles combine into -------
template <int V, int R>
PGS_HD void phase_0(Smem<V, R>& sm, const Thr<V, R>& th) {
  if (th.role != 0) return;
  const int ln = th.lane;
  sm.wrote[ln] = th.valid ? 1 : 0;
  sm.amask[ln] = sm.pmask[ln] = sm.crash[ln] = 0;
  sm.n_vis[ln] = 0;
  if (th.valid) {
    const PgdMap& mp = th.mp;
    sm.cx_x0[ln] = mp.x0; sm.cx_y0[ln] = mp.y0; sm.cx_inv_cell[ln] = mp.inv_cell; sm.cx_lane_width[ln] = mp.lane_width;
    sm.cx_nx[ln] = mp.nx; sm.cx_ny[ln] = mp.ny; sm.cx_cell_off[ln] = mp.cell_off; sm.cx_entry_off[ln] = mp.entry_off;
    sm.cx_lane_off[ln] = mp.lane_off; sm.cx_road_off[ln] = mp.road_off; sm.cx_box_off[ln] = mp.box_off;
    sm.cx_slot_off[ln] = ldg(&th.ep->slot_off);
    sm.cx_n_slots[ln] = th.n_slots;
  }
}

/* Table context of environment e of the CTA (any thread may work on any environment after phase A). */
template <int V, int R>
PGS_HD const PgdLane* lanes_of(const Smem<V, R>& sm, const Tables& T, int e) { return T.lanes + sm.cx_lane_off[e]; }
template <int V, int R>
PGS_HD const PgdRoad* roads_of(const Smem<V, R>& sm, const Tables& T, int e) { return T.roads + sm.cx_road_off[e]; }
template <int V, int R>
PGS_HD const PgdBox* boxes_of(const Smem<V, R>& sm, const Tables& T, int e) { return T.boxes + sm.cx_box_off[e]; }
template <int V, int R>
PGS_HD const PgdSlot* slots_of(const Smem<V, R>& sm, const Tables& T, int e) { return T.slots + sm.cx_slot_off[e]; }

// ---- phase A: publish start-of-step state; ego action; traffic trigger ------------------------------------------
template <int V, int R>
PGS_HD void phase_a(Smem<V, R>& sm, Thr<V, R>& th, const State& S, const PgdConfig& cfg, const float* actions,
                   float* obs) {
  const int ln = th.lane;
  if (!th.valid) return;
  Pub<V>& P = sm.p;
  Pub<V>& P0 = idm_of<V, R>(obs).start;  // what IDM reads while phase X moves the vehicles in P
  // TrafficManager.before_step (traffic_manager.py:71-89): the next group wakes when the ego is on its trigger road.
  // Every role evaluates the (cheap) test itself instead of waiting for role 0.
  if (th.stepping && th.envi.y < th.n_groups) {
    if (ldg(&th.lanes[th.pre_nav.x].road) == ldg(&th.ep->trigger_road[th.envi.y])) th.trig = th.envi.y;
  }
  if (th.role == 0) {
    Veh& q = th.ego;
    const PgdSlot& t = th.tpl[0];
    if (th.fresh) veh_from_template(q, t, 0);
    else veh_unpack(q, th.pre_pose, th.pre_ctrl, th.pre_pidl, th.pre_nav, th.pre_misc);
    q.hl = t.length * 0.5f;
    q.hw = t.width * 0.5f;
    PGS_SINCOS(q.h, q.hs, q.hc);
    sm.last_x[ln] = q.x; sm.last_y[ln] = q.y;
    sm.ego_hl[ln] = q.hl; sm.ego_hw[ln] = q.hw;
    sm.ego_ck[ln] = q.ck0 | (q.ck1 << 16);
    P.x[0][ln] = q.x; P.y[0][ln] = q.y; P.hc[0][ln] = q.hc; P.hs[0][ln] = q.hs; P.v[0][ln] = q.v;
    P.lf[0][ln] = lf_pack(q.lane, q.vflags);
    P0.x[0][ln] = q.x; P0.y[0][ln] = q.y; P0.hc[0][ln] = q.hc; P0.hs[0][ln] = q.hs; P0.v[0][ln] = q.v;
    P0.lf[0][ln] = lf_pack(q.lane, q.vflags);
    if (th.stepping) {  // EnvInputPolicy.act (env_input_policy.py:17-26): clip; fminf / fmaxf turn NaN into -1
      const float a0 = clipf(actions[2 * (size_t)th.env], -1.0f, 1.0f);
      const float a1 = clipf(actions[2 * (size_t)th.env + 1], -1.0f, 1.0f);
      th.envf.y = q.throttle;  // last_current_action[0] after the push (base_vehicle.py:248)
      if (cfg.increment_steering) {  // _set_incremental_action (base_vehicle.py:351-358); q.hp = last raw action
        th.envf.x = q.hp;
        q.hp = a0;
        q.steer = clipf(q.steer + a0 * 0.05f, -1.0f, 1.0f);
      } else {
        th.envf.x = q.steer;
        q.steer = a0;
      }
      q.throttle = a1;
      if (th.trig >= 0) th.envi.y += 1;
    }
    sm.envi[ln] = th.envi;
    sm.envf[ln] = th.envf;
    sm.last_h[ln] = th.fresh ? q.h : th.pre_pose.z;
    return;
  }
  uint32_t amask = 0, pmask = 0;
#pragma unroll 1
  for (int k = 0; k < Thr<V, R>::MAXOWN; ++k) {
    const int s = th.role + k * (R - 1);
    if (s >= V) break;
    const size_t gi = (size_t)s * th.num_envs + th.env;
    if (s >= th.n_slots) {
      if (th.fresh) {  // unused slots of a freshly started episode: clear the flags once
        const I4 m = {0, 0, 0, 0};
        S.misc[gi] = m;
      }
      continue;
    }
    const PgdSlot& t = th.tpl[s];
    float x, y, h, v = 0.0f;
    int lane, fl;
    if (th.fresh) {
      Veh q;
      veh_from_template(q, t, s);
      veh_store(q, S, gi);
      x = q.x; y = q.y; h = q.h; lane = q.lane; fl = q.vflags;
    } else {
      fl = (int)(th.pre_fl[Thr<V, R>::FLW == 1 ? 0 : k / 12] >> (5 * (k % 12))) & 31;
      if (!(fl & PGD_V_ALIVE)) {
        P.lf[s][ln] = 0;
        P0.lf[s][ln] = 0;
        continue;
      }
      if (fl & PGD_V_ACTIVE) {
        const F4 p = S.pose[gi];
        x = p.x; y = p.y; h = p.z; v = p.w;
        lane = S.nav[gi].x;
      } else {
        // Traffic that has not been woken yet has never been touched by IDM, physics (it is at rest) or localisation:
        // its pose is the episode template's (L2-resident, shared by all environments on the seed).
        if (fl & PGS_V_POSE_SET) {  // placed by pgd_set_state
          const F4 p = S.pose[gi];
          x = p.x; y = p.y; h = p.z;
          lane = S.nav[gi].x;
        } else {
          x = t.x; y = t.y; h = t.heading; lane = t.lane;
        }
        if (t.group == th.trig) fl |= PGD_V_ACTIVE;
      }
    }
    float sn, cs;
    PGS_SINCOS(h, sn, cs);
    P.x[s][ln] = x; P.y[s][ln] = y; P.hc[s][ln] = cs; P.hs[s][ln] = sn; P.v[s][ln] = v;
    P.lf[s][ln] = lf_pack(lane, fl);
    P0.x[s][ln] = x; P0.y[s][ln] = y; P0.hc[s][ln] = cs; P0.hs[s][ln] = sn; P0.v[s][ln] = v;
    P0.lf[s][ln] = lf_pack(lane, fl);
    // an episode that restarts in this call does not act: its awake traffic (traffic_mode "respawn") is no work item
    if ((fl & PGD_V_ACTIVE) && th.stepping) amask |= 1u << s;
    else pmask |= 1u << s;
  }
  smem_or(&sm.amask[ln], (int)amask);
  smem_or(&sm.pmask[ln], (int)pmask);
}

template <int V, int R>
PGS_HD uint32_t env_amask(const Smem<V, R>& sm, int ln) { return (uint32_t)sm.amask[ln]; }  // alive + awake traffic
template <int V, int R>
PGS_HD uint32_t env_pmask(const Smem<V, R>& sm, int ln) { return (uint32_t)sm.pmask[ln]; }  // alive + parked traffic

/* The awake traffic of the CTA's 32 environments as ONE list (lane | slot << 8, by environment, then slot), so that
 * phases C and D deal vehicles to ALL traffic threads of the CTA instead of to the threads of their own environment:
 * a warp's lanes then run the same number of vehicles whatever the spread between environments.  Built by one warp
 * (an exclusive scan of the per-environment counts); `lane` = this thread's lane in that warp. */
template <int V, int R>
PGS_HD void build_work_list(Smem<V, R>& sm, int lane) {
#ifdef __CUDA_ARCH__
  const uint32_t m = (uint32_t)sm.amask[lane];
  const int c = __popc(m);
  int incl = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += up;
  }
  int at = incl - c;
  for (uint32_t mm = m; mm; mm &= mm - 1) sm.work[at++] = (uint16_t)(lane | (ctz32(mm) << 8));
  if (lane == 31) sm.n_work = incl;
#else
  if (lane != 0) return;  // the host build runs the "warp" as one loop
  int at = 0;
  for (int e = 0; e < PGS_LANES; ++e)
    for (uint32_t mm = (uint32_t)sm.amask[e]; mm; mm &= mm - 1) sm.work[at++] = (uint16_t)(e | (ctz32(mm) << 8));
  sm.n_work = at;
#endif
}

// ---- phase B: IDM look-up data (only environments with awake traffic) --------------------------------------------
template <int V, int R>
PGS_HD void phase_b(Smem<V, R>& sm, Thr<V, R>& th, const Tables& T, float* obs) {
  if (th.role == 1) build_work_list(sm, th.lane);  // every lane of the warp takes part (invalid lanes count 0)
  if (!th.valid || !th.stepping) return;
  const int ln = th.lane;
  if (!env_amask(sm, ln)) return;
  const Pub<V>& P = sm.p;
  IdmPub<V>& I = idm_of<V, R>(obs);
  // every alive vehicle (the ego and parked traffic included: they are obstacles), spread evenly over all roles
  const uint32_t alive = env_amask(sm, ln) | env_pmask(sm, ln) | 1u;
#pragma unroll 1
  for (uint32_t m = drop_low(alive, th.role); m; m = drop_low(m, R)) {
    const int s = ctz32(m);
    const PgdLane l = load_rec(lanes_of(sm, T, ln) + lf_lane(P.lf[s][ln]));
    I.lsx[s][ln] = l.sx; I.lsy[s][ln] = l.sy; I.lex[s][ln] = l.ex; I.ley[s][ln] = l.ey; I.llen[s][ln] = l.length;
    float lon, lat;
    lane_local(l, P.x[s][ln], P.y[s][ln], lon, lat);
    I.olong[s][ln] = lon;
  }
}

// ---- IDM / PID action of one awake traffic vehicle (idm_policy.py:190-353) --------------------------------------
template <int V, int R>
PGS_HD void idm_act(const Smem<V, R>& sm, const Tables& T, const float* obs, int ln, uint32_t alive, Veh& q, int s,
                   int& cur_road_out, int& next_road_out) {
  const IdmPub<V>& I = idm_of<V, R>(obs);
  const Pub<V>& P = I.start;  // everybody's pose at the start of the step
  const PgdSlot& t = slots_of(sm, T, ln)[s];
  const PgdLane* lanes = lanes_of(sm, T, ln);
  const PgdRoad* roads = roads_of(sm, T, ln);
  const int32_t* rroads = T.route_roads + t.route_off;
  const int cur_road_id = ldg(&rroads[q.ck0]);
  const int next_road_id = q.ck0 != q.ck1 ? ldg(&rroads[q.ck1]) : -1;
  cur_road_out = cur_road_id;
  next_road_out = next_road_id;
  const PgdRoad cur_road = load_rec(roads + cur_road_id);
  bool ok;  // move_to_next_road (:222-242)
  if (q.rt_lane < 0) {
    q.rt_lane = q.lane;
    ok = ldg(&lanes[q.rt_lane].road) == cur_road_id;
  } else if (ldg(&lanes[q.rt_lane].road) != cur_road_id) {
    ok = false;
    const float rex = ldg(&lanes[q.rt_lane].ex), rey = ldg(&lanes[q.rt_lane].ey);
#pragma unroll 1
    for (int k = 0; k < cur_road.n_lanes; ++k) {
      const PgdLane* c = lanes + cur_road.first_lane + k;
      if (precedes(rex, rey, ldg(&c->sx), ldg(&c->sy))) {
        q.rt_lane = cur_road.first_lane + k;
        ok = true;
        break;
      }
    }
  } else if (ldg(&lanes[q.lane].road) == cur_road_id && q.rt_lane != q.lane) {
    q.rt_lane = q.lane;
    q.timer = t.rnd25[q.rnd_n % PGD_N_RND25];
    q.rnd_n++;
    ok = true;
  } else {
    ok = true;
  }
  const PgdLane rl = load_rec(lanes + q.rt_lane);
  int cand[3] = {-1, q.rt_lane, -1};
  if (ok) {
    const PgdRoad rr = load_rec(roads + rl.road);
    if (rl.idx > 0) cand[0] = rr.first_lane + rl.idx - 1;
    if (rl.idx + 1 < rr.n_lanes) cand[2] = rr.first_lane + rl.idx + 1;
  }
  int front[3], back[3];
  float fdist[3], bdist[3];
  const uint32_t others = alive & ~(1u << s);
  // FrontBackObjects.get_find_front_back_objs (:83-133) for the three candidate lanes in ONE pass over the other vehicles
  // (per candidate lane the reference's loop, vehicle by vehicle in ascending slot order)
  bool on[3], found_front[3], found_back[3];
  float cur_long[3], left_long[3], lsx[3], lsy[3], lex[3], ley[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    front[i] = back[i] = -1;
    fdist[i] = bdist[i] = PGS_IDM_MAX_LONG;
    found_front[i] = found_back[i] = false;
    on[i] = cand[i] >= 0;
    cur_long[i] = left_long[i] = lsx[i] = lsy[i] = lex[i] = ley[i] = 0.0f;
    if (!on[i]) continue;
    const PgdLane l = (i == 1) ? rl : load_rec(lanes + cand[i]);
    float lat;
    lane_local(l, q.x, q.y, cur_long[i], lat);
    left_long[i] = l.length - cur_long[i];
    lsx[i] = l.sx; lsy[i] = l.sy; lex[i] = l.ex; ley[i] = l.ey;
  }
#pragma unroll 1
  for (uint32_t m = others; m; m &= m - 1) {  // ascending slot order, like the oracle's object list
    const int j = ctz32(m);
    const float ddx = P.x[j][ln] - q.x, ddy = P.y[j][ln] - q.y;
    if (!(ddx * ddx + ddy * ddy < PGS_LIDAR_RANGE * PGS_LIDAR_RANGE)) continue;
    const int lane_j = lf_lane(P.lf[j][ln]);
    const float olong_j = I.olong[j][ln];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (!on[i]) continue;
      if (lane_j == cand[i]) {
        const float lg = olong_j - cur_long[i];
        if (fdist[i] > lg && lg > 0.0f) { fdist[i] = lg; front[i] = j; found_front[i] = true; }
        if (lg < 0.0f && fabsf(lg) < bdist[i]) { bdist[i] = fabsf(lg); back[i] = j; found_back[i] = true; }
      } else if (!found_front[i] && precedes(lex[i], ley[i], I.lsx[j][ln], I.lsy[j][ln])) {
        const float lg = olong_j + left_long[i];
        if (fdist[i] > lg && lg > 0.0f) { fdist[i] = lg; front[i] = j; }
      } else if (!found_back[i] && precedes(I.lex[j][ln], I.ley[j][ln], lsx[i], lsy[i])) {
        const float lg = I.llen[j][ln] - olong_j + cur_long[i];
        if (bdist[i] > lg) { bdist[i] = lg; back[i] = j; }
      }
    }
  }
  int front_obj = front[1], steer_lane = q.rt_lane;
  float front_dist = fdist[1];
  if (ok) {  // lane_change_policy (:281-353)
    const int n_cur = cur_road.n_lanes;
    int lo = 0, hi_idx = n_cur - 1;
    bool decided = false;
    const int idx = rl.idx;
    if (q.ck0 != q.ck1) {
      const PgdRoad nxt = load_rec(roads + next_road_id);
      const int diff = n_cur - nxt.n_lanes;
      if (diff > 0) {
        const PgdLane* c0 = lanes + cur_road.first_lane;
        const PgdLane* n0 = lanes + nxt.first_lane;
        if (precedes(ldg(&c0->ex), ldg(&c0->ey), ldg(&n0->sx), ldg(&n0->sy))) {
          lo = 0; hi_idx = nxt.n_lanes - 1;
        } else {
          lo = diff; hi_idx = n_cur - 1;
        }
        if (idx < lo || idx > hi_idx) {
          decided = true;
          const int side = idx > hi_idx ? 0 : 2;
          if (bdist[side] < PGS_IDM_SAFE_DIST || fdist[side] < 5.0f) {
            q.tspeed = PGS_IDM_CREEP_SPEED;
          } else {
            q.tspeed = PGS_IDM_NORMAL_SPEED;
            front_obj = front[side];
            front_dist = fdist[side];
            steer_lane = cur_road.first_lane + idx + (side == 0 ? -1 : 1);
          }
        }
      }
    }
    if (!decided) {
      const float my_speed = kmh(q.v);
      if (fabsf(my_speed - PGS_IDM_NORMAL_SPEED) > 3.0f && front[1] >= 0 &&
          fabsf(kmh(P.v[front[1]][ln]) - PGS_IDM_NORMAL_SPEED) > 3.0f && q.timer > PGS_IDM_LANE_CHANGE_FREQ) {
        float side_speed[3] = {0.f, 0.f, 0.f};
        bool side_ok[3] = {false, false, false};
#pragma unroll
        for (int sd = 0; sd < 3; sd += 2) {
          if (front[sd] >= 0) {
            side_speed[sd] = kmh(P.v[front[sd]][ln]);
            side_ok[sd] = true;
          } else if (cand[sd] >= 0 && fdist[sd] > PGS_IDM_SAFE_DIST && bdist[sd] > PGS_IDM_SAFE_DIST) {
            side_speed[sd] = PGS_IDM_MAX_SPEED;
            side_ok[sd] = true;
          }
        }
        const float front_speed = kmh(P.v[front[1]][ln]);
        if (side_ok[0] && side_speed[0] - front_speed > PGS_IDM_SPEED_INCREASE && idx - 1 >= lo && idx - 1 <= hi_idx) {
          decided = true;
          front_obj = front[0]; front_dist = fdist[0];
          steer_lane = cur_road.first_lane + idx - 1;
        } else if (side_ok[2] && side_speed[2] - front_speed > PGS_IDM_SPEED_INCREASE && idx + 1 >= lo &&
                   idx + 1 <= hi_idx) {
          decided = true;
          front_obj = front[2]; front_dist = fdist[2];
          steer_lane = cur_road.first_lane + idx + 1;
        }
      }
    }
    if (!decided) {
      q.tspeed = PGS_IDM_NORMAL_SPEED;
      q.timer += 1;
    }
  }
  {  // steering_control (:244-252)
    const PgdLane tl = (steer_lane == q.rt_lane) ? rl : load_rec(lanes + steer_lane);
    float lon, lat;
    lane_local(tl, q.x, q.y, lon, lat);
    const float lane_heading = lane_heading_at(tl, lon + 1.0f);
    float st = pid(q.hp, q.hi, 1.7f, 0.01f, 3.5f, pgd_wrap_to_pi(lane_heading - q.h));
    st += pid(q.lp, q.li, 0.3f, 0.002f, 0.05f, -lat);
    q.steer = st;
  }
  {  // acceleration (:254-271), speeds in km/h as in the reference
    const float sp = kmh(q.v);
    float acc = 1.0f - pgd_pow10f(fmaxf(sp, 0.0f) / q.tspeed);
    if (front_obj >= 0) {
      const float hx = q.hc, hy = q.hs;
      const float fs = kmh(P.v[front_obj][ln]);
      const float dvx = sp * hx - fs * P.hc[front_obj][ln], dvy = sp * hy - fs * P.hs[front_obj][ln];
      const float dv = dvx * hx + dvy * hy;
      const float d_star = 10.0f + sp * 1.5f + sp * dv / (2.0f * sqrtf(5.0f));
      float d = front_dist;
      if (!(fabsf(d) > 1e-2f)) d = d > 0.0f ? 1e-2f : -1e-2f;  // not_zero
      const float ratio = d_star / d;
      acc -= ratio * ratio;
    }
    q.throttle = acc;
  }
}

// ---- localisation through the bucket grid (navigation.py:155-344, scene_utils.py:138-185) ------------------------
struct ScanOut {  // lowest box id (and its lane) over the point, per class of road; PGD_F_* contacts of the chassis
  int b_any, b_cur, b_next, l_any, l_cur, l_next;
  uint32_t flags;
};
struct GridRef { float x0, y0, inv_cell; int nx, ny, cell_off, entry_off; };  // bucket grid of one map
template <int V, int R>
PGS_HD GridRef grid_of(const Smem<V, R>& sm, int e) {
  const GridRef g = {sm.cx_x0[e], sm.cx_y0[e], sm.cx_inv_cell[e], sm.cx_nx[e], sm.cx_ny[e], sm.cx_cell_off[e],
                     sm.cx_entry_off[e]};
  return g;
}

/* The bucket of (x, y): lane-surface boxes that contain the point and run along the heading; with EGO also the chassis
 * against line ghosts and sidewalks (base_vehicle.py:615-644). */
template <bool EGO>
PGS_HD void bucket_scan(const GridRef& gr, const PgdLane* lanes, const PgdBox* boxes, const Tables& T, float x, float y,
                       float hc, float hs, float hl, float hw, int cur_road, int next_road, ScanOut& out) {
  out.b_any = out.b_cur = out.b_next = INT_MAX;
  out.l_any = out.l_cur = out.l_next = -1;
  out.flags = 0;
  const int32_t* ent = T.cell_entries + gr.entry_off;
  const Rect er = {x, y, hc, hs, hl, hw};
  const int cx = (int)floorf((x - gr.x0) * gr.inv_cell), cy = (int)floorf((y - gr.y0) * gr.inv_cell);
  if (!(cx >= 0 && cy >= 0 && cx < gr.nx && cy < gr.ny)) return;
  const int cell = gr.cell_off + cy * gr.nx + cx;
  const int b0 = ldg(&T.cell_start[cell]), b1 = ldg(&T.cell_start[cell + 1]);
  // entries are fetched four at a time (indices, then records) so that their latencies overlap; inside a cell the
  // lane-surface boxes come first, everything else carries PGD_ENTRY_NOT_LANE: traffic stops there
  for (int k0 = b0; k0 < b1; k0 += 4) {
    int bb[4];
    PgdBox gg[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bb[j] = (k0 + j < b1) ? ldg(&ent[k0 + j]) : -1;
      if (!EGO && bb[j] >= PGD_ENTRY_NOT_LANE) bb[j] = -1;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (bb[j] >= 0) gg[j] = load_rec(boxes + (bb[j] & PGD_ENTRY_ID_MASK));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (bb[j] < 0) continue;
      const int b = bb[j] & PGD_ENTRY_ID_MASK;
      const PgdBox& g = gg[j];
      if (g.kind == PGD_BOX_LANE) {
        const float dx = x - g.cx, dy = y - g.cy;
        if (!(fabsf(dx * g.ux + dy * g.uy) <= g.hl && fabsf(-dx * g.uy + dy * g.ux) <= g.hw)) continue;
        const PgdLane* l = lanes + g.lane;
        float dot;  // lane direction . heading > 0, written without trigonometry
        if (ldg(&l->kind) == PGD_LANE_STRAIGHT) {
          dot = ldg(&l->ax) * hc + ldg(&l->ay) * hs;
        } else {
          dot = ldg(&l->dir) * ((x - ldg(&l->ax)) * hs - (y - ldg(&l->ay)) * hc);
        }
        if (!(dot > 0.0f)) continue;
        const int lroad = ldg(&l->road);
        if (b < out.b_any) { out.b_any = b; out.l_any = g.lane; }
        if (lroad == cur_road && b < out.b_cur) { out.b_cur = b; out.l_cur = g.lane; }
        if (lroad == next_road && b < out.b_next) { out.b_next = b; out.l_next = g.lane; }
      } else if (EGO) {
        const Rect r = {g.cx, g.cy, g.ux, g.uy, g.hl, g.hw};
        if (!rect_overlap(er, r)) continue;
        out.flags |= g.kind == PGD_BOX_WHITE ? PGD_F_ON_WHITE
                   : g.kind == PGD_BOX_YELLOW ? PGD_F_ON_YELLOW
                   : g.kind == PGD_BOX_BROKEN ? PGD_F_ON_BROKEN : PGD_F_CRASH_SIDEWALK;
      }
    }
    // traffic stops at the flagged part (or the end) of the cell -- and at the first box on its current road: the ids
    // ascend within the cell (include/pgd_tables.h) and a box on the current road wins over every other class
    if (!EGO && (bb[3] < 0 || out.b_cur != INT_MAX)) break;
  }
}

/* What follows the scan: lane choice (current road, then next road, then any; lowest box id), checkpoint update
 * (_update_target_checkpoints), on-lane flag. */
template <int V, int R>
PGS_HD void after_scan(const Smem<V, R>& sm, const Tables& T, int e, const PgdSlot& t, const ScanOut& sc, float x,
                      float y, int& lane, int& ck0, int& ck1, bool& on_lane) {
  const int32_t* rnodes = T.route_nodes + t.route_off;
  const int nl = sc.b_cur != INT_MAX ? sc.l_cur : (sc.b_next != INT_MAX ? sc.l_next : sc.l_any);
  on_lane = nl >= 0;
  if (on_lane) lane = nl;
  if (ck0 != ck1) {
    const PgdLane l = load_rec(lanes_of(sm, T, e) + lane);
    float lon, lat;
    lane_local(l, x, y, lon, lat);
    const int start = ldg(&roads_of(sm, T, e)[l.road].start_node);
    if (lon < 5.0f) {
#pragma unroll 1
      for (int j = ck1; j < t.route_len - 1; ++j) {
        if (ldg(&rnodes[j]) == start) {
          ck0 = j;
          ck1 = (j + 1 == t.route_len - 1) ? j : j + 1;
          break;
        }
      }
    }
  }
}

/* cur_road / next_road: the roads of the vehicle's two checkpoints, looked up by idm_act (the checkpoints only change
 * here, afterwards). */
template <int V, int R>
PGS_HD void localise_traffic(const Smem<V, R>& sm, const Tables& T, int e, const PgdSlot& t, int cur_road, int next_road,
                            Veh& q) {
  ScanOut sc;
  bucket_scan<false>(grid_of(sm, e), lanes_of(sm, T, e), boxes_of(sm, T, e), T, q.x, q.y, q.hc, q.hs, q.hl, q.hw, cur_road,
                     next_road, sc);
  bool on_lane;
  after_scan(sm, T, e, t, sc, q.x, q.y, q.lane, q.ck0, q.ck1, on_lane);
  q.vflags = on_lane ? (q.vflags | PGD_V_ON_LANE) : (q.vflags & ~(PGD_V_ON_LANE | PGD_V_ALIVE));  // traffic_manager.py:91-109
}

// ---- phase X: everything that moves -------------------------------------------------------------------------------
/* Role 0: the ego's sub-steps (trajectory -> shared memory, announced to the traffic warps through named barrier 1),
 * then its localisation (navigation.py:155-211) and line / sidewalk contacts (base_vehicle.py:615-644), then the table
 * look-ups of the reward and of the navigation features (phase F only adds what depends on the traffic: the crash
 * flags).
 * Traffic roles: work items = the CTA's awake traffic in batches, dealt round-robin -- IDM / PID against the start-of-step
 * copy of everybody's pose, then the sub-steps with the chassis test against the ego's trajectory, localisation, removal,
 * and the final pose into Smem::p; then the parked vehicles of the thread's own environment against the ego's
 * trajectory; then, once all traffic warps are there (named barrier 2), the ego's neighbour features and lidar windows.
 * An item touches nothing another item reads, so the order in which the warps work does not matter. */
PGS_HD void named_arrive(int id, int threads) {  // all 32 lanes of the warp
#ifdef __CUDA_ARCH__
  __syncwarp();
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
#else
  (void)id; (void)threads;
#endif
}
PGS_HD void named_wait(int id, int threads) {  // all 32 lanes of the warp
#ifdef __CUDA_ARCH__
  __syncwarp();
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
#else
  (void)id; (void)threads;
#endif
}
template <int V, int R>
PGS_HD void reward_lookups(const Smem<V, R>& sm, Thr<V, R>& th, const Tables& T, const PgdConfig& cfg);
template <int V, int R>
PGS_HD void navi_lookups(const Smem<V, R>& sm, Thr<V, R>& th, const Tables& T);
template <int V, int R>
PGS_HD void task_neighbours(const Smem<V, R>& sm, const Thr<V, R>& th, const Tables& T, const PgdConfig& cfg, float* obs);
template <int V, int R>
PGS_HD void task_lidar_windows(Smem<V, R>& sm, const Thr<V, R>& th, const Tables& T, VisPtr vis);
#define PGS_BAR_TRAJ 1  // role 0 arrives, the R - 1 traffic warps wait: the egos' trajectories are in shared memory
#define PGS_BAR_TRAFFIC 2  // the R - 1 traffic warps among themselves: every vehicle has its final pose

template <int V, int R>
PGS_HD void phase_x_ego(Smem<V, R>& sm, Thr<V, R>& th, const Tables& T, const State& S, const PgdConfig& cfg,
                        TrajPtr traj) {
  const int ln = th.lane;
  const int ns = cfg.decision_repeat < PGS_MAX_SUBSTEPS ? cfg.decision_repeat : PGS_MAX_SUBSTEPS;
  Veh& q = th.ego;
  PGS_EGO_CLK_BEGIN
  if (th.valid) {
    if (th.stepping) {  // 5 x doPhysics(0.02) of the ego (base_engine.py:206-232), remembering every pose
      // A vehicle at rest with no yaw rate and no engine force is a fixed point of the sub-step (speed = max(0 - dv, 0)
      // = 0, the pose does not move); only its drop counter runs.
      const bool parked = q.v == 0.0f && q.yaw == 0.0f && !(q.throttle > 0.0f);
      Sub sub;
      if (!parked) sub = make_sub(q, slots_of(sm, T, ln)[0], cfg.dt);
      float m2 = 0.0f;
#pragma unroll 1
      for (int k = 0; k < ns; ++k) {
        if (q.airborne > 0) q.airborne--;  // placed 1 m above the road: no wheel contact while it drops
        else if (!parked) substep(q, sub, cfg.dt);
        const F4 p = {q.x, q.y, q.hc, q.hs};
        traj[k][ln] = p;
        const float ex = q.x - sm.last_x[ln], ey = q.y - sm.last_y[ln];
        m2 = fmaxf(m2, ex * ex + ey * ey);
      }
      sm.ego_travel[ln] = sqrtf(m2) * 1.001f + 1e-3f;  // how far the ego gets from its start pose within the step
    }
    const F4 fin = {q.x, q.y, q.hc, q.hs};
    sm.efin[ln] = fin;
    sm.ego_h[ln] = q.h;
    sm.ego_v[ln] = q.v;
  }
  named_arrive(PGS_BAR_TRAJ, R * 32);
  PGS_EGO_CLK(13);
  if (th.valid) {
    // the record goes home now (phase F only adds what localisation changes: lane, checkpoints, on-lane flag), so that
    // the warp does not carry it through the rest of the step
    veh_store(q, S, (size_t)th.env);
    // ego after_step, first half (navigation.py:155-211, base_vehicle.py:615-644)
    const PgdSlot& t0 = slots_of(sm, T, ln)[0];
    const int32_t* rroads = T.route_roads + t0.route_off;
    const int cur_road = ldg(&rroads[q.ck0]);
    const int next_road = q.ck0 != q.ck1 ? ldg(&rroads[q.ck1]) : -1;
    ScanOut sc;
    bucket_scan<true>(grid_of(sm, ln), lanes_of(sm, T, ln), boxes_of(sm, T, ln), T, q.x, q.y, q.hc, q.hs, q.hl, q.hw,
                      cur_road, next_road, sc);
    PGS_EGO_CLK(14);
    bool on_lane;  // the start-of-step lane stays when no lane box is under the vehicle
    after_scan(sm, T, ln, t0, sc, q.x, q.y, q.lane, q.ck0, q.ck1, on_lane);
    q.vflags = on_lane ? (q.vflags | PGD_V_ON_LANE) : (q.vflags & ~PGD_V_ON_LANE);
    th.flags = sc.flags | (on_lane ? PGD_F_ON_LANE : 0);
    sm.ego_lane[ln] = q.lane;
    sm.ego_ck[ln] = q.ck0 | (q.ck1 << 16);
  }
  PGS_EGO_CLK(15);
  if (th.valid) {
    reward_lookups(sm, th, T, cfg);
    navi_lookups(sm, th, T);
  }
}

/* Parked traffic of the thread's own environment (the slots this role published in phase A, whose drop counters it
 * already holds): only the drop counter runs and the (fixed) chassis is tested against the ego's pose of every
 * sub-step. */
template <int V, int R>
PGS_HD void phase_x_parked(Smem<V, R>& sm, Thr<V, R>& th, const Tables& T, const State& S, const PgdConfig& cfg,
                          TrajPtr traj) {
  if (!th.valid || th.role == 0 || !th.stepping) return;
  const int ln = th.lane;
  const Pub<V>& P = sm.p;
  const float ehl = sm.ego_hl[ln], ehw = sm.ego_hw[ln];
  const int ns = cfg.decision_repeat < PGS_MAX_SUBSTEPS ? cfg.decision_repeat : PGS_MAX_SUBSTEPS;
  int crash = 0;
  const uint32_t pmask = env_pmask(sm, ln);
#pragma unroll 1  // (one copy of the body: the kernel is as large as the instruction cache)
  for (int k = 0; k < Thr<V, R>::MAXOWN; ++k) {
    const int s = th.role + k * (R - 1);
    if (s >= sm.cx_n_slots[ln]) break;
    if (!((pmask >> s) & 1u)) continue;
    const size_t gi = (size_t)s * th.num_envs + th.env;
    const PgdSlot& t = slots_of(sm, T, ln)[s];
    const I4 m = S.misc[gi];  // (the thread's own write in phase A when the episode restarted)
    int air = m.y, fl = m.z;
    bool dirty = false;
    if (air > 0) {
      air = air > ns ? air - ns : 0;
      dirty = true;
    }
    const float px = P.x[s][ln], py = P.y[s][ln];
    const float ddx0 = px - sm.last_x[ln], ddy0 = py - sm.last_y[ln];
    const float far = ehl + ehw + 8.0f + sm.ego_travel[ln];  // no chassis has an 8 m half-diagonal
    if (!(ddx0 * ddx0 + ddy0 * ddy0 > far * far)) {  // else: triangle inequality; the margin covers rounding
      const float hl = t.length * 0.5f, hw = t.width * 0.5f;
      const float reach = ehl + ehw + hl + hw;
      bool hit = false;
#pragma unroll 1
      for (int kk = 0; kk < ns; ++kk) {
        const F4 e = traj[kk][ln];
        const float ddx = px - e.x, ddy = py - e.y;
        if (ddx * ddx + ddy * ddy <= reach * reach) {
          const Rect me = {px, py, P.hc[s][ln], P.hs[s][ln], hl, hw};
          const Rect eg = {e.x, e.y, e.z, e.w, ehl, ehw};
          if (rect_overlap(eg, me)) hit = true;
        }
      }
      if (hit) {  // collision_callback.py:13-30
        if (t.type >= PGD_TYPE_OBJECT) {
          if (!(fl & PGD_V_CRASHED)) {  // COST_ONCE: an object is charged the first time it is touched
            crash |= 2;
            fl |= PGD_V_CRASHED;
            dirty = true;
          }
        } else {
          crash |= 1;
        }
      }
    }
    if (dirty) {
      const I4 m2 = {0, air, fl, 0};  // parked: no IDM draw yet
      S.misc[gi] = m2;
    }
  }
  smem_or(&sm.crash[ln], crash);
}

template <int V, int R>
PGS_HD void phase_x_items(Smem<V, R>& sm, const Thr<V, R>& th, const Tables& T, const State& S, const PgdConfig& cfg,
                         const float* obs, TrajPtr traj) {
  Pub<V>& P = sm.p;
  const int ns = cfg.decision_repeat < PGS_MAX_SUBSTEPS ? cfg.decision_repeat : PGS_MAX_SUBSTEPS;
  const int n_work = sm.n_work;
  bool traj_ready = false;  // warp-uniform
  // Batches of `per` vehicles, dealt round-robin to the traffic warps.  A list that fits one pass is split evenly over
  // them: the time of a batch is the union of its lanes' paths through IDM and localisation, whatever their number.
  const int per = n_work >= (R - 1) * PGS_LANES ? PGS_LANES : (n_work + R - 2) / (R - 1);
  PGS_ITEM_CLK_BEGIN
  PGS_ITEM_COUNT(8, n_work);
#pragma unroll 1
  for (int base = (th.role - 1) * per; base < n_work; base += (R - 1) * per) {
    PGS_ITEM_COUNT(10, 1);
    PGS_ITEM_CLK(0);
    const bool have = th.lane < per && base + th.lane < n_work;
    int e = 0, s = 0, cur_road = 0, next_road = -1;
    size_t gi = 0;
    Veh q;
    if (have) {
      const int w = sm.work[base + th.lane];
      e = w & 0xff; s = w >> 8;
      gi = (size_t)s * th.num_envs + (th.env - th.lane + e);
      const uint32_t alive = env_amask(sm, e) | env_pmask(sm, e) | 1u;
      veh_load(q, S, gi);
      q.vflags = P.lf[s][e] & 0xff;  // a vehicle woken in this step carries ACTIVE only in shared memory so far
      q.hc = P.hc[s][e]; q.hs = P.hs[s][e];
      PGS_ITEM_CLK(1);
      idm_act(sm, T, obs, e, alive, q, s, cur_road, next_road);
    }
    PGS_ITEM_CLK(2);
    if (!traj_ready) {  // the whole warp, once
      named_wait(PGS_BAR_TRAJ, R * 32);
      traj_ready = true;
    }
    if (have) {
      const PgdSlot& t = slots_of(sm, T, e)[s];
      const float ehl = sm.ego_hl[e], ehw = sm.ego_hw[e];
      q.hl = t.length * 0.5f; q.hw = t.width * 0.5f;
      const float reach = ehl + ehw + q.hl + q.hw;
      const bool at_rest = q.v == 0.0f && q.yaw == 0.0f && !(q.throttle > 0.0f);
      Sub sub;
      if (!at_rest) sub = make_sub(q, t, cfg.dt);
      int crash = 0;
      PGS_ITEM_CLK(3);
#pragma unroll 1
      for (int k = 0; k < ns; ++k) {
        if (q.airborne > 0) q.airborne--;
        else if (!at_rest) substep(q, sub, cfg.dt);
        const F4 eg4 = traj[k][e];
        const float ddx = q.x - eg4.x, ddy = q.y - eg4.y;
        if (ddx * ddx + ddy * ddy <= reach * reach) {
          const Rect me = {q.x, q.y, q.hc, q.hs, q.hl, q.hw};
          const Rect eg = {eg4.x, eg4.y, eg4.z, eg4.w, ehl, ehw};
          if (rect_overlap(eg, me)) crash = 1;
        }
      }
      PGS_ITEM_CLK(4);
      localise_traffic(sm, T, e, t, cur_road, next_road, q);
      PGS_ITEM_CLK(5);
      veh_store(q, S, gi);
      P.x[s][e] = q.x; P.y[s][e] = q.y; P.hc[s][e] = q.hc; P.hs[s][e] = q.hs; P.v[s][e] = q.v;
      P.lf[s][e] = lf_pack(q.lane, q.vflags);
      smem_or(&sm.crash[e], crash);
    }
    PGS_ITEM_CLK(6);
  }
  if (!traj_ready) named_wait(PGS_BAR_TRAJ, R * 32);
}

template <int V, int R>
PGS_HD void phase_x_tail(Smem<V, R>& sm, Thr<V, R>& th, const Tables& T, const PgdConfig& cfg, int obs_dim,
                        float* obs_rows, VisPtr vis) {
  // the traffic has its final poses, IDM's data and the ego's trajectory (whose storage the rows and the list of visible
  // chassis take over) are dead: what only needs the traffic and the ego's final pose starts now, while role 0 is
  // still busy with the ego's look-ups
  if (th.valid) {
    if (th.role == 1) task_neighbours(sm, th, T, cfg, obs_rows + (size_t)th.lane * obs_dim);
    if (th.role == (R > 2 ? 2 : 1)) task_lidar_windows(sm, th, T, vis);
  }
}

/* Phase X of one thread.  (The host emulation runs the traffic roles' two halves one after the other over all of them,
 * oracle/step_host.cpp: named barrier PGS_BAR_TRAFFIC.) */
template <int V, int R>
PGS_HD void phase_x(Smem<V, R>& sm, Thr<V, R>& th, const Tables& T, const State& S, const PgdConfig& cfg, int obs_dim,
                   float* obs_rows, TrajPtr traj, VisPtr vis) {
  if (th.role == 0) {
    phase_x_ego(sm, th, T, S, cfg, traj);
    return;
  }
  phase_x_items(sm, th, T, S, cfg, obs_rows, traj);
  phase_x_parked(sm, th, T, S, cfg, traj);
  named_wait(PGS_BAR_TRAFFIC, (R - 1) * 32);
  phase_x_tail(sm, th, T, cfg, obs_dim, obs_rows, vis);
}

// ---- phase F: ego bookkeeping, one task per role ---------------------------------------------------------------------
PGS_HD int obs_dim_of(const PgdConfig& cfg) {
  return (cfg.n_side > 0 ? cfg.n_side : 2) + 6 + cfg.n_lane_line + (cfg.random_agent_model ? 2 : 0) + 10 + 16 +
         PGD_LIDAR_BEAMS;
}

/* Role 0 while the traffic moves (phase X): the table look-ups of the ego's bookkeeping -- route distances, arrival,
 * reward geometry (base_vehicle.py:383-388,738-745; pgdrive_env.py:162-258).  What phase F needs of them stays in
 * registers: th.flags, th.keep[11..14]. */
template <int V, int R>
PGS_HD void reward_lookups(const Smem<V, R>& sm, Thr<V, R>& th, const Tables& T, const PgdConfig& cfg) {
  const int ln = th.lane;
  const float ex = sm.efin[ln].x, ey = sm.efin[ln].y;
  const int lane = sm.ego_lane[ln], ck0 = sm.ego_ck[ln] & 0xffff;
  const PgdSlot& t0 = slots_of(sm, T, ln)[0];
  const PgdLane* lanes = lanes_of(sm, T, ln);
  const PgdRoad* roads = roads_of(sm, T, ln);
  const float lane_width = sm.cx_lane_width[ln];
  const int32_t* rroads = T.route_roads + t0.route_off;
  uint32_t flags = th.flags;
  const bool on_lane = (flags & PGD_F_ON_LANE) != 0;
  const float last_x = sm.last_x[ln], last_y = sm.last_y[ln];
  const int cur_road_id = ldg(&rroads[ck0]);
  const PgdRoad cur_road = load_rec(roads + cur_road_id);
  const PgdRoad fr = load_rec(roads + ldg(&rroads[t0.route_len - 2]));
  const int el_road = ldg(&lanes[lane].road);
  const bool use_ego_lane = el_road == cur_road_id;
  const int reward_lane = use_ego_lane ? lane : cur_road.first_lane;
  const int n_ref = cur_road.n_lanes;
  const int sign_i = use_ego_lane ? 0 : (ldg(&roads[el_road].negative) ? -1 : 1);
  float qlon0, qlat0, qlon1, qlat1, long_last = 0.0f, lat_last, long_now = 0.0f, lat_now = 0.0f;
  lane_local(load_rec(lanes + cur_road.first_lane), ex, ey, qlon0, qlat0);
  const PgdLane final_lane = load_rec(lanes + (fr.first_lane + fr.n_lanes - 1));
  lane_local(final_lane, ex, ey, qlon1, qlat1);
  if (!th.fresh) {
    const PgdLane rl = load_rec(lanes + reward_lane);
    lane_local(rl, last_x, last_y, long_last, lat_last);
    lane_local(rl, ex, ey, long_now, lat_now);
  }
  const float to_left = qlat0 + lane_width / 2.0f;  // base_vehicle.py:383-388
  const float to_right = lane_width * (float)n_ref - to_left;
  if (to_left < 0.0f || to_right < 0.0f) flags |= PGD_F_OUT_OF_ROUTE;
  {  // arrive_destination (base_vehicle.py:738-745)
    const float flen = final_lane.length;
    if (flen - 5.0f < qlon1 && qlon1 < flen + 5.0f && lane_width / 2.0f >= qlat1 &&
        qlat1 >= (0.5f - (float)n_ref) * lane_width)
      flags |= PGD_F_ARRIVE_DEST;
  }
  bool out_of_road = (flags & (PGD_F_ON_YELLOW | PGD_F_ON_WHITE | PGD_F_CRASH_SIDEWALK)) || !on_lane;
  if (cfg.out_of_route_done && (flags & PGD_F_OUT_OF_ROUTE)) out_of_road = true;
  if (out_of_road) flags |= PGD_F_OUT_OF_ROAD;
  const float sign = sign_i == 0 ? 1.0f : (float)sign_i;
  float lateral_factor = 1.0f;
  if (cfg.use_lateral) lateral_factor = clipf(1.0f - 2.0f * fabsf(lat_now) / lane_width, 0.0f, 1.0f);
  th.flags = flags;
  th.keep[11] = to_left;
  th.keep[12] = to_right;
  th.keep[13] = cfg.driving_reward * (long_now - long_last) * lateral_factor * sign;
  th.keep[14] = sign;
}

/* task 0 (role 0): what the traffic decides -- crash flags -- then reward / cost / done, the ego's state features, info,
 * and the words of the ego's record that localisation changed */
template <int V, int R>
PGS_HD void task_reward(Smem<V, R>& sm, Thr<V, R>& th, const Tables& T, const State& S, const PgdConfig& cfg, int mode,
                       float* obs, float* reward, uint8_t* done, PgdInfo* info) {
  const int ln = th.lane;
  struct {  // what is left of the ego in registers (phase_x_ego stored the record)
    float x, y, h, v, steer, throttle;
  } ego = {sm.efin[ln].x, sm.efin[ln].y, sm.ego_h[ln], sm.ego_v[ln], th.ego.steer, th.ego.throttle};
  const float last_h = sm.last_h[ln];
  th.envi = sm.envi[ln];
  th.envf = sm.envf[ln];
  const int n_first = cfg.n_side > 0 ? cfg.n_side : 2;
  float* const st = obs + n_first - 2;
  if (cfg.random_agent_model) {  // obs/state_obs.py:103-105: LENGTH / MAX_LENGTH, WIDTH / MAX_WIDTH (base_vehicle.py:83-84)
    const PgdSlot& t0 = slots_of(sm, T, ln)[0];
    obs[n_first + 6 + cfg.n_lane_line] = clipf(t0.length / 10.0f, 0.0f, 1.0f);
    obs[n_first + 6 + cfg.n_lane_line + 1] = clipf(t0.width / 2.5f, 0.0f, 1.0f);
  }
  const int crash = sm.crash[ln] & 1, crash_object = (sm.crash[ln] >> 1) & 1;
  const float last_x = sm.last_x[ln], last_y = sm.last_y[ln];
  uint32_t flags = th.flags;
  if (crash) flags |= PGD_F_CRASH_VEHICLE;
  if (crash_object) flags |= PGD_F_CRASH_OBJECT;
  const bool out_of_road = (flags & PGD_F_OUT_OF_ROAD) != 0;
  const float to_left = th.keep[11], to_right = th.keep[12];

  const float sp = kmh(ego.v);
  if (cfg.n_side <= 0) {
    obs[0] = clipf(to_left / 18.0f, 0.0f, 1.0f);
    obs[1] = clipf(to_right / 18.0f, 0.0f, 1.0f);
  }
  st[3] = clipf((sp + 1.0f) / (PGS_MAX_SPEED_KMH + 1.0f), 0.0f, 1.0f);
  st[4] = clipf((ego.steer / 60.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
  st[5] = clipf((th.envf.x + 1.0f) / 2.0f, 0.0f, 1.0f);
  st[6] = clipf((th.envf.y + 1.0f) / 2.0f, 0.0f, 1.0f);
  // yaw rate: arccos(clip(cos(angle between headings), 0, 1)) / 0.1 (state_obs.py:87-94) = min(|wrapped heading
  // change|, pi/2) / 0.1 without the ill-conditioned arccos
  st[7] = clipf(fminf(fabsf(pgd_wrap_to_pi(ego.h - last_h)), PGS_PI / 2) / 0.1f, 0.0f, 1.0f);
  float r = 0.0f, step_reward = 0.0f, cost = 0.0f, step_energy = 0.0f;
  int is_done = 0;
  if (!th.fresh) {  // envs/pgdrive_env.py:162-258
    const float sign = th.keep[14];
    r += th.keep[13];  // driving_reward * (long_now - long_last) * lateral_factor * sign
    r += cfg.speed_reward * (sp / PGS_MAX_SPEED_KMH) * sign;
    step_reward = r;
    if (flags & PGD_F_ARRIVE_DEST) r = cfg.success_reward;
    else if (out_of_road) r = -cfg.out_of_road_penalty;
    else if (crash) r = -cfg.crash_vehicle_penalty;
    else if (crash_object) r = -cfg.crash_object_penalty;
    if (out_of_road) cost = cfg.out_of_road_cost;
    else if (crash) cost = cfg.crash_vehicle_cost;
    else if (crash_object) cost = cfg.crash_object_cost;
    is_done = ((flags & PGD_F_ARRIVE_DEST) || out_of_road || crash || crash_object) ? 1 : 0;
    // SafePGDriveEnv.done_function (safe_pgdrive_env.py:44-51), literally: a step with a crash never ends the episode
    if (cfg.safe_rl_env && (crash || crash_object)) is_done = 0;
    const float ddx = last_x - ego.x, ddy = last_y - ego.y;  // base_vehicle.py:278-290
    step_energy = 3.25f * pgd_expf(0.01f * sp) * (sqrtf(ddx * ddx + ddy * ddy) / 1000.0f) / 100.0f * 1000.0f;
    th.envf.w += step_energy;
    th.envf.z += r;
    th.envi.w += 1;
    if (cfg.horizon > 0 && th.envi.w >= cfg.horizon) { is_done = 1; flags |= PGD_F_MAX_STEP; }
    if (th.envi.z == 1) is_done = 1;  // done is sticky (base_env.py:315-316)
    th.envi.z = is_done;
  } else {
    flags |= PGD_F_WAS_RESET;
  }
  if (mode == 0) {
    reward[th.env] = r;
    done[th.env] = (uint8_t)is_done;
  }
  if (info) {
    PgdInfo inf;
    inf.velocity = sp; inf.steering = ego.steer; inf.acceleration = ego.throttle;
    inf.step_energy = step_energy; inf.episode_energy = th.envf.w;
    inf.step_reward = step_reward; inf.episode_reward = th.envf.z; inf.cost = cost;
    inf.episode_length = th.envi.w; inf.flags = flags;
    info[th.env] = inf;
  }
  S.envi[th.env] = th.envi;
  S.envf[th.env] = th.envf;
  S.nav[th.env].x = sm.ego_lane[ln];
  S.nav[th.env].y = sm.ego_ck[ln];
  S.misc[th.env].z = th.ego.vflags;
}

/* Role 0 at the end of phase X: navigation info of the two checkpoints (navigation.py:213-260) + heading_diff
 * (base_vehicle.py:433-458) -> th.keep[0..10]; task_navi writes them to the row in phase F. */
template <int V, int R>
PGS_HD void navi_lookups(const Smem<V, R>& sm, Thr<V, R>& th, const Tables& T) {
  const int ln = th.lane;
  const F4 e = sm.efin[ln];
  const float ex = e.x, ey = e.y, ehc = e.z, ehs = e.w;
  const int ck0 = sm.ego_ck[ln] & 0xffff, ck1 = sm.ego_ck[ln] >> 16;  // after the ego's localisation
  const PgdLane* lanes = lanes_of(sm, T, ln);
  const PgdRoad* roads = roads_of(sm, T, ln);
  const float lane_width = sm.cx_lane_width[ln];
  const int32_t* rroads = T.route_roads + slots_of(sm, T, ln)[0].route_off;
  const PgdRoad cur_road = load_rec(roads + ldg(&rroads[ck0]));
  const int n_ref = cur_road.n_lanes;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const PgdLane l = load_rec(lanes + (c == 0 ? cur_road.first_lane : ldg(&roads[ldg(&rroads[ck1])].first_lane)));
    const float later_middle = ((float)n_ref / 2.0f - 0.5f) * lane_width;
    float px, py;
    lane_position(l, l.length, later_middle, px, py);
    float dx = px - ex, dy = py - ey;
    const float dn = sqrtf(dx * dx + dy * dy);
    if (dn > 50.0f) { dx = dx / dn * 50.0f; dy = dy / dn * 50.0f; }
    float ph, ps;
    project(ehc, ehs, dx, dy, ph, ps);
    float bend = 0.0f, dir = 0.0f, angle = 0.0f;
    if (l.kind == PGD_LANE_ARC) {
      bend = l.radius / (60.0f + (float)n_ref * lane_width);
      dir = l.dir;
      angle = l.length / l.radius;
    }
    th.keep[5 * c + 0] = clipf((ph / 50.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
    th.keep[5 * c + 1] = clipf((ps / 50.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
    th.keep[5 * c + 2] = clipf(bend, 0.0f, 1.0f);
    th.keep[5 * c + 3] = clipf((dir + 1.0f) / 2.0f, 0.0f, 1.0f);
    th.keep[5 * c + 4] = clipf((angle * (180.0f / PGS_PI) / 135.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
  }
  {  // heading_diff against the right-most reference lane
    const PgdLane l = load_rec(lanes + (cur_road.first_lane + cur_road.n_lanes - 1));
    float lx, ly;
    if (l.kind == PGD_LANE_STRAIGHT) { lx = -l.ay; ly = l.ax; }
    else if (l.dir < 0.0f) { lx = ex - l.ax; ly = ey - l.ay; }
    else { lx = l.ax - ex; ly = l.ay - ey; }
    const float lnm = sqrtf(lx * lx + ly * ly);
    th.keep[10] = lnm > 0.0f ? clipf((ehc * lx + ehs * ly) / lnm, -1.0f, 1.0f) / 2.0f + 0.5f : 0.0f;
  }
}

template <int V, int R>
PGS_HD void task_navi(const Thr<V, R>& th, const PgdConfig& cfg, float* obs) {
  const int n_first = cfg.n_side > 0 ? cfg.n_side : 2;
  float* const st = obs + n_first - 2;
  float* const ob = obs + n_first + cfg.n_lane_line + (cfg.random_agent_model ? 2 : 0) - 2;
#pragma unroll
  for (int i = 0; i < 10; ++i) ob[8 + i] = th.keep[i];
  st[2] = th.keep[10];
}

/* task 2: the 4 nearest vehicles inside the 50 m cylinder (lidar.py:55-77; ties -> lower slot) */
template <int V, int R>
PGS_HD void task_neighbours(const Smem<V, R>& sm, const Thr<V, R>& th, const Tables& T, const PgdConfig& cfg,
                           float* obs) {
  const int ln = th.lane;
  const Pub<V>& P = sm.p;
  const F4 e = sm.efin[ln];
  const float ex = e.x, ey = e.y, ehc = e.z, ehs = e.w;
  const float esp = kmh(sm.ego_v[ln]);
  const int n_first = cfg.n_side > 0 ? cfg.n_side : 2;
  float* const ob = obs + n_first + cfg.n_lane_line + (cfg.random_agent_model ? 2 : 0) - 2;
  uint32_t cand = 0;  // alive vehicles inside the cylinder
#pragma unroll 1
  for (int s = 1; s < sm.cx_n_slots[ln]; ++s) {
    if (!(P.lf[s][ln] & PGD_V_ALIVE)) continue;
    const float dx = P.x[s][ln] - ex, dy = P.y[s][ln] - ey;
    if (dx * dx + dy * dy < PGS_LIDAR_RANGE * PGS_LIDAR_RANGE && ldg(&slots_of(sm, T, ln)[s].type) < PGD_TYPE_OBJECT)
      cand |= 1u << s;  // get_surrounding_vehicles (lidar.py:46-54): cones and barriers are not vehicles
  }
#pragma unroll 1
  for (int rank = 0; rank < 4; ++rank) {
    int best = -1;
    float best_d2 = INFINITY;
#pragma unroll 1
    for (uint32_t m = cand; m; m &= m - 1) {
      const int s = ctz32(m);
      const float dx = P.x[s][ln] - ex, dy = P.y[s][ln] - ey;
      const float d2 = dx * dx + dy * dy;
      if (best < 0 || d2 < best_d2) { best = s; best_d2 = d2; }
    }
    float* o4 = ob + 18 + 4 * rank;
    if (best < 0) {
      o4[0] = o4[1] = o4[2] = o4[3] = 0.0f;
      continue;
    }
    cand &= ~(1u << best);
    float pf, ps, vf, vs;
    project(ehc, ehs, P.x[best][ln] - ex, P.y[best][ln] - ey, pf, ps);
    const float ws = kmh(P.v[best][ln]);
    project(ehc, ehs, ws * P.hc[best][ln] - esp * ehc, ws * P.hs[best][ln] - esp * ehs, vf, vs);
    o4[0] = clipf((pf / PGS_LIDAR_RANGE + 1.0f) / 2.0f, 0.0f, 1.0f);
    o4[1] = clipf((ps / PGS_LIDAR_RANGE + 1.0f) / 2.0f, 0.0f, 1.0f);
    o4[2] = clipf((vf / PGS_MAX_SPEED_KMH + 1.0f) / 2.0f, 0.0f, 1.0f);
    o4[3] = clipf((vs / PGS_MAX_SPEED_KMH + 1.0f) / 2.0f, 0.0f, 1.0f);
  }
}

/* task 3: which chassis the lidar can reach, and the (conservative) arc of beams that can hit each */
template <int V, int R>
PGS_HD void task_lidar_windows(Smem<V, R>& sm, const Thr<V, R>& th, const Tables& T, VisPtr vis) {
  const int ln = th.lane;
  const Pub<V>& P = sm.p;
  const F4 e = sm.efin[ln];
  const float ex = e.x, ey = e.y, eh = sm.ego_h[ln];
  int n = 0;
#pragma unroll 1
  for (int s = 1; s < sm.cx_n_slots[ln]; ++s) {
    if (!(P.lf[s][ln] & PGD_V_ALIVE)) continue;
    const float dx = P.x[s][ln] - ex, dy = P.y[s][ln] - ey;
    const float d2 = dx * dx + dy * dy;
    if (!(d2 < (PGS_LIDAR_RANGE + 8.0f) * (PGS_LIDAR_RANGE + 8.0f))) continue;  // no chassis has an 8 m half-diagonal
    const PgdSlot& t = slots_of(sm, T, ln)[s];
    const float hl = t.length * 0.5f, hw = t.width * 0.5f;
    const float hd = sqrtf(hl * hl + hw * hw);
    const float reach = PGS_LIDAR_RANGE + hd;
    if (!(d2 < reach * reach)) continue;
    const float d = sqrtf(d2);
    int blo = 0, bn = PGD_LIDAR_BEAMS - 1;
    if (d > hd * 1.001f) {
      const float per_rad = (float)PGD_LIDAR_BEAMS / PGS_TWO_PI;
      const float c = (pgd_atan2f(dy, dx) - eh) * per_rad;
      const float w = pgd_asinf(fminf(hd / d, 1.0f)) * per_rad;
      const int nn = (int)ceilf(2.0f * w) + 3;
      if (nn < PGD_LIDAR_BEAMS) {
        bn = nn;
        blo = ((int)floorf(c - w) - 1) % PGD_LIDAR_BEAMS;
        if (blo < 0) blo += PGD_LIDAR_BEAMS;
      }
    }
    vis[n++][ln] = s | (blo << 8) | (bn << 16);
  }
  sm.n_vis[ln] = n;
}

/* side / lane-line detectors (distance_detector.py:137-152): ray fans against the line ghosts of the map; a beam
 * looks up the bucket of a point every 2 PGD_GRID_MARGIN along itself (buckets list every box within one margin).  The rays of an
 * environment are spread over the roles. */
template <int V, int R>
PGS_HD void task_detectors(const Smem<V, R>& sm, const Thr<V, R>& th, const Tables& T, const PgdConfig& cfg,
                          float* obs) {
  const int ln = th.lane;
  const F4 e = sm.efin[ln];
  const float ex = e.x, ey = e.y, eh = sm.ego_h[ln];
  const int n_first = cfg.n_side > 0 ? cfg.n_side : 2;
  const int n_rays = cfg.n_side + cfg.n_lane_line;
  const GridRef mp = grid_of(sm, ln);
  const int32_t* ent = T.cell_entries + mp.entry_off;
#pragma unroll 1
  for (int rI = th.role; rI < n_rays; rI += R) {
    const bool side = rI < cfg.n_side;
    const int i = side ? rI : rI - cfg.n_side;
    const int n = side ? cfg.n_side : cfg.n_lane_line;
    const float dist = side ? cfg.side_distance : cfg.lane_line_distance;
    const float ang = (float)i * (PGS_TWO_PI / (float)n) + PGS_PI / 2 + eh;
    float sn, cs;
    PGS_SINCOS(ang, sn, cs);
    const float dx = cs * dist, dy = sn * dist;
    float best = 1.0f;
    const float gm = (float)PGD_GRID_MARGIN;
    for (float sd = gm; sd - gm < dist; sd += 2.0f * gm) {
      if (best * dist < sd - gm) break;
      const float px = ex + cs * sd, py = ey + sn * sd;
      const int cx = (int)floorf((px - mp.x0) * mp.inv_cell), cy = (int)floorf((py - mp.y0) * mp.inv_cell);
      if (cx < 0 || cy < 0 || cx >= mp.nx || cy >= mp.ny) continue;
      const int cell = mp.cell_off + cy * mp.nx + cx;
      const int b0 = ldg(&T.cell_start[cell]), b1 = ldg(&T.cell_start[cell + 1]);
#pragma unroll 1
      for (int k = b0; k < b1; ++k) {
        const int en = ldg(&ent[k]);
        if (en < PGD_ENTRY_NOT_LANE) continue;  // lane-surface boxes are no ray targets
        const PgdBox g = load_rec(boxes_of(sm, T, ln) + (en & PGD_ENTRY_ID_MASK));
        if (!(g.kind == PGD_BOX_WHITE || g.kind == PGD_BOX_YELLOW || (!side && g.kind == PGD_BOX_BROKEN))) continue;
        const Rect r = {g.cx, g.cy, g.ux, g.uy, g.hl, g.hw};
        best = fminf(best, ray_rect(ex, ey, dx, dy, r));
      }
    }
    if (side) obs[i] = best;
    else obs[n_first + 6 + i] = best;
  }
}

template <int V, int R>
PGS_HD void phase_f(Smem<V, R>& sm, Thr<V, R>& th, const Tables& T, const State& S, const PgdConfig& cfg, int mode,
                   int obs_dim, float* obs_rows, VisPtr vis, float* reward, uint8_t* done, PgdInfo* info) {
  if (!th.valid) return;
  float* obs = obs_rows + (size_t)th.lane * obs_dim;
  // role 0 holds what it looked up during phase X (the neighbours and the lidar windows were done at its
  // end, phase_x_tail); nothing here touches the lidar part of the rows, which phase L fills without a barrier in between
  if (th.role == 0) {
    task_reward(sm, th, T, S, cfg, mode, obs, reward, done, info);
    task_navi(th, cfg, obs);
  }
  if (cfg.n_side > 0 || cfg.n_lane_line > 0) task_detectors(sm, th, T, cfg, obs);
}

// ---- phase L: lidar as a scatter (cutils.pyx:60-142 restated per chassis instead of per beam) ---------------------
/* Thread (role, lane) of the CTA: the warp `role` takes environments role, role + R, ...; its lanes are the beams
 * of the window of each visible chassis.  The row was pre-filled with 1.0 (no hit); hits are min-ed in. */
PGS_HD void lidar_min(float* cell, float t) {
#ifdef __CUDA_ARCH__
  atomicMin(reinterpret_cast<int*>(cell), __float_as_int(t));  // non-negative floats order like their bit patterns
#else
  if (t < *cell) *cell = t;
#endif
}

template <int V, int R>
PGS_HD void lidar_scatter(const Smem<V, R>& sm, const Tables& T, const State& S, const Pub<V>& P, int e, int lane,
                         int env0, float* row_lidar, VisPtr vis) {
  {
    const int nv = sm.n_vis[e];
    const F4 eg = sm.efin[e];
    const float ex = eg.x, ey = eg.y, eh = sm.ego_h[e];
    float* row = row_lidar;
    const PgdSlot* tpl = slots_of(sm, T, e);
#pragma unroll 1
    for (int k = 0; k < nv; ++k) {
      const int w = vis[k][e];
      const int s = w & 0xff, blo = (w >> 8) & 0xff, bn = w >> 16;
      const Rect r = {P.x[s][e], P.y[s][e], P.hc[s][e], P.hs[s][e], ldg(&tpl[s].length) * 0.5f,
                      ldg(&tpl[s].width) * 0.5f};
#pragma unroll 1
      for (int rel = lane; rel <= bn; rel += PGS_LANES) {
        int i = blo + rel;
        if (i >= PGD_LIDAR_BEAMS) i -= PGD_LIDAR_BEAMS;
        const float ang = (float)i * (PGS_TWO_PI / (float)PGD_LIDAR_BEAMS) + eh;
        float sn, cs;
        PGS_SINCOS(ang, sn, cs);
        const float t = ray_rect(ex, ey, cs * PGS_LIDAR_RANGE, sn * PGS_LIDAR_RANGE, r);
        if (t < 1.0f) lidar_min(row + i, t);
      }
    }
  }
}

/* The warp that scatters an environment's hits first fills its beams with 1.0 = "no hit" (the storage held IDM's
 * look-up data until phase X ended); a warp-level barrier separates the two. */
template <int V, int R>
PGS_HD void phase_l_fill(const Smem<V, R>& sm, int role, int lane, int obs_dim, float* obs_rows) {
  const int head = obs_dim - PGD_LIDAR_BEAMS;
#pragma unroll 1
  for (int e = role; e < PGS_LANES; e += R) {
    if (!sm.wrote[e]) continue;
    float* row = obs_rows + (size_t)e * obs_dim + head;
    for (int i = lane; i < PGD_LIDAR_BEAMS; i += PGS_LANES) row[i] = 1.0f;
  }
}

template <int V, int R>
PGS_HD void phase_l(const Smem<V, R>& sm, const Tables& T, const State& S, int role, int lane, int env0, int obs_dim,
                   float* obs_rows, VisPtr vis) {
  const int head = obs_dim - PGD_LIDAR_BEAMS;
#pragma unroll 1
  for (int e = role; e < PGS_LANES; e += R) {
    if (!sm.wrote[e] || sm.n_vis[e] == 0) continue;
    lidar_scatter<V, R>(sm, T, S, sm.p, e, lane, env0, obs_rows + (size_t)e * obs_dim + head, vis);
  }
}

// ---- phase N: lidar noise on the finished rows (_add_noise_to_cloud_points, obs/state_obs.py:172-182); the warp that
// scattered an environment's hits also adds its noise, so a warp-level barrier separates the two ------------------
template <int V, int R>
PGS_HD void phase_n(const Smem<V, R>& sm, const PgdConfig& cfg, uint32_t call_index, int role, int lane, int env0,
                   int obs_dim, float* obs_rows) {
  if (!(cfg.lidar_gaussian_noise > 0.0f || cfg.lidar_dropout_prob > 0.0f)) return;
  const int head = obs_dim - PGD_LIDAR_BEAMS;
#pragma unroll 1
  for (int e = role; e < PGS_LANES; e += R) {
    if (!sm.wrote[e]) continue;
    float* row = obs_rows + (size_t)e * obs_dim + head;
    for (int i = lane; i < PGD_LIDAR_BEAMS; i += PGS_LANES)
      row[i] = pgd_lidar_noise(row[i], cfg.lidar_gaussian_noise, cfg.lidar_dropout_prob,
                               pgd_noise_key((uint32_t)cfg.noise_seed, call_index, (uint32_t)(env0 + e), (uint32_t)i));
  }
}

}  // namespace pgdstep
#endif
