/* CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A scalar, one-environment-at-a-time restatement of the reference's per-step path
 * (PGDriveEnv.step: before_step -> 5 physics sub-steps -> after_step -> obs / reward / done), used
 * only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs to check
 * and to time against the CUDA path.  Nothing under pgdrive_b200/ links or calls this file.
 *
 * It is deliberately naive: array-of-structs state, every query is a brute-force loop over ALL
 * primitives of the map (no bucket grid), one env after the other.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   pinned against the reference's own Python  : lane Frenet math, navigation info, IDM + PID, observation
 *       packing, reward / done / cost  (tests/golden/step_*.json.gz made by tools/make_golden.py)
 *   pinned against reference-generated fixtures: maps, lanes, traffic slots, routes (host side, pgdrive_b200/)
 *   PARITY UNPINNED                            : everything Bullet (panda3d~=1.10.8, not vendored, not installable
 *       here) computes -- vehicle dynamics, chassis/line/sidewalk contacts, lane-surface ray test, ray-vs-chassis.
 *       These are restated as a planar friction-limited bicycle + exact 2-D rectangle geometry.
 *
 * Reference (paths under /root/reference/pgdrive) is cited per function.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/pgd_math.h" /* shared float32 sin / cos / atan2 / exp: same bits as the CUDA build */
#include "../include/pgd_tables.h"

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

typedef struct {
  float x, y, h, v, w; /* pose, speed [m/s], yaw rate [rad/s] */
  float steer, throttle;
  float hp, hi, lp, li;
  float target_speed;
  int lane, ck0, ck1, rt_lane, timer, rnd_n, airborne;
  int alive, active, on_lane;
  int crashed; /* traffic object already charged (COST_ONCE, traffic_object.py:22) */
  const PgdSlot* s;
} Veh;

typedef struct {
  int episode, n_slots, next_group, done, ep_len;
  float prev_steer, prev_throttle, ep_reward, energy;
  Veh v[PGD_MAX_SLOTS];
} Env;

typedef struct {
  PgdTables t;
  PgdConfig cfg;
  Env* envs;
  int fast; /* orc_set_fast: queries go through the map's bucket grid instead of over every primitive (same results;
               used when the oracle is TIMED as the CPU baseline, so that the baseline is not a strawman) */
  uint32_t call_index; /* API calls so far (set by the wrapper before each reset / step): lidar-noise key */
} Oracle;

static float clipf(float a, float lo, float hi) { return fminf(fmaxf(a, lo), hi); } /* cutils.pyx:153 */

static float wrap_to_pi(float x) { return pgd_wrap_to_pi(x); } /* utils/math_utils.py:32-33 */

static float cosf_(float a) { float s, c; pgd_sincosf(a, &s, &c); return c; }
static float sinf_(float a) { float s, c; pgd_sincosf(a, &s, &c); return s; }

/* ---- lanes: straight_lane.py:53-67, circular_lane.py:46-67 ------------------------------------ */
static void lane_local(const PgdLane* l, float x, float y, float* lon, float* lat) {
  if (l->kind == PGD_LANE_STRAIGHT) {
    float dx = x - l->sx, dy = y - l->sy;
    *lon = dx * l->ax + dy * l->ay;
    *lat = dx * -l->ay + dy * l->ax;
  } else {
    float dx = x - l->ax, dy = y - l->ay;
    float phi = pgd_atan2f(dy, dx);
    phi = l->ph0 + wrap_to_pi(phi - l->ph0);
    float r = sqrtf(dx * dx + dy * dy);
    *lon = l->dir * (phi - l->ph0) * l->radius;
    *lat = l->dir * (l->radius - r);
  }
}

static void lane_position(const PgdLane* l, float lon, float lat, float* x, float* y) {
  if (l->kind == PGD_LANE_STRAIGHT) {
    *x = l->sx + lon * l->ax + lat * -l->ay;
    *y = l->sy + lon * l->ay + lat * l->ax;
  } else {
    float phi = l->dir * lon / l->radius + l->ph0;
    float r = l->radius - lat * l->dir;
    *x = l->ax + r * cosf_(phi);
    *y = l->ay + r * sinf_(phi);
  }
}

static float lane_heading_at(const PgdLane* l, float lon) {
  if (l->kind == PGD_LANE_STRAIGHT) return l->heading;
  float phi = l->dir * lon / l->radius + l->ph0;
  return phi + PI_F / 2 * l->dir;
}

static int lane_precedes(const PgdLane* a, const PgdLane* b) { /* abs_lane.py:114-119 */
  float dx = a->ex - b->sx, dy = a->ey - b->sy;
  return sqrtf(dx * dx + dy * dy) < 1e-1f;
}

/* ---- rectangles ------------------------------------------------------------------------------ */
typedef struct { float cx, cy, ux, uy, hl, hw; } Rect;

static Rect veh_rect(const Veh* v) {
  float sn, cs;
  pgd_sincosf(v->h, &sn, &cs);
  Rect r = {v->x, v->y, cs, sn, v->s->length * 0.5f, v->s->width * 0.5f};
  return r;
}

static int rect_overlap(const Rect* a, const Rect* b) { /* separating-axis test, 4 axes */
  float dx = b->cx - a->cx, dy = b->cy - a->cy;
  float c = fabsf(a->ux * b->ux + a->uy * b->uy);
  float s = fabsf(a->ux * b->uy - a->uy * b->ux);
  if (fabsf(dx * a->ux + dy * a->uy) > a->hl + b->hl * c + b->hw * s) return 0;
  if (fabsf(-dx * a->uy + dy * a->ux) > a->hw + b->hl * s + b->hw * c) return 0;
  if (fabsf(dx * b->ux + dy * b->uy) > b->hl + a->hl * c + a->hw * s) return 0;
  if (fabsf(-dx * b->uy + dy * b->ux) > b->hw + a->hl * s + a->hw * c) return 0;
  return 1;
}

static int rect_contains(const PgdBox* b, float x, float y) {
  float dx = x - b->cx, dy = y - b->cy;
  return fabsf(dx * b->ux + dy * b->uy) <= b->hl && fabsf(-dx * b->uy + dy * b->ux) <= b->hw;
}

/* Fraction in [0,1] along the segment o -> o + d at which it enters the rectangle, 1 if it misses
 * (the closest-hit ray test of cutils.pyx:103-123 against one chassis). */
static float ray_rect(float ox, float oy, float dx, float dy, const Rect* r) {
  float px = ox - r->cx, py = oy - r->cy;
  float lo[2] = {px * r->ux + py * r->uy, -px * r->uy + py * r->ux};
  float ld[2] = {dx * r->ux + dy * r->uy, -dx * r->uy + dy * r->ux};
  float half[2] = {r->hl, r->hw};
  /* Bullet's convex ray cast reports nothing for a shape that contains the ray origin (the ego centre over a line
   * ghost while it crosses a lane line) */
  if (fabsf(lo[0]) <= half[0] && fabsf(lo[1]) <= half[1]) return 1.0f;
  float t0 = 0.0f, t1 = 1.0f;
  for (int k = 0; k < 2; ++k) {
    if (fabsf(ld[k]) < 1e-12f) {
      if (fabsf(lo[k]) > half[k]) return 1.0f;
    } else {
      float inv = 1.0f / ld[k];
      float ta = (-half[k] - lo[k]) * inv, tb = (half[k] - lo[k]) * inv;
      if (ta > tb) { float tmp = ta; ta = tb; tb = tmp; }
      t0 = fmaxf(t0, ta);
      t1 = fminf(t1, tb);
      if (t0 > t1) return 1.0f;
    }
  }
  return t0;
}

/* ---- table access helpers --------------------------------------------------------------------- */
static const PgdEpisode* ep_of(const Oracle* o, const Env* e) { return &o->t.episodes[e->episode]; }
static const PgdMap* map_of(const Oracle* o, const Env* e) { return &o->t.maps[ep_of(o, e)->map]; }
static const PgdLane* lane_at(const Oracle* o, const PgdMap* m, int lane) { return &o->t.lanes[m->lane_off + lane]; }
static const PgdRoad* road_at(const Oracle* o, const PgdMap* m, int road) { return &o->t.roads[m->road_off + road]; }
static int route_road(const Oracle* o, const Veh* v, int k) { return o->t.route_roads[v->s->route_off + k]; }
static int route_node(const Oracle* o, const Veh* v, int k) { return o->t.route_nodes[v->s->route_off + k]; }

/* ---- localisation: scene_utils.py:138-185 + navigation.py:155-211,262-282,328-344 -------------- */
static void localize(const Oracle* o, const PgdMap* m, Veh* v) {
  float hx = cosf_(v->h), hy = sinf_(v->h);
  int cur_road = route_road(o, v, v->ck0);
  int next_road = (v->ck0 == v->ck1) ? -1 : route_road(o, v, v->ck1);
  int first_any = -1, first_cur = -1, first_next = -1;
  int n_cand = m->n_boxes, e0 = 0;
  const int32_t* ent = NULL;
  if (o->fast) { /* lane boxes of the bucket under the vehicle, ascending box id like the loop over all boxes */
    n_cand = 0;
    int cx = (int)floorf((v->x - m->x0) * m->inv_cell), cy = (int)floorf((v->y - m->y0) * m->inv_cell);
    if (cx >= 0 && cy >= 0 && cx < m->nx && cy < m->ny) {
      int cell = m->cell_off + cy * m->nx + cx;
      e0 = o->t.cell_start[cell];
      n_cand = o->t.cell_start[cell + 1] - e0;
      ent = o->t.cell_entries + m->entry_off;
    }
  }
  for (int k = 0; k < n_cand; ++k) {
    int b = k;
    if (ent) {
      if (ent[e0 + k] >= PGD_ENTRY_NOT_LANE) break;
      b = ent[e0 + k];
    }
    const PgdBox* box = &o->t.boxes[m->box_off + b];
    if (box->kind != PGD_BOX_LANE || !rect_contains(box, v->x, v->y)) continue;
    const PgdLane* l = lane_at(o, m, box->lane);
    /* cos(lane.heading_at(long)) * hx + sin(...) * hy > 0 (scene_utils.py:158-170) with the lane direction written
     * without trigonometry: a straight lane's unit vector, an arc's tangent dir * (-dy, dx) / r at the vehicle's
     * bearing from the centre (heading_at(long) = bearing + dir * pi / 2; r > 0 does not change the sign) */
    float dot;
    if (l->kind == PGD_LANE_STRAIGHT) dot = l->ax * hx + l->ay * hy;
    else dot = l->dir * ((v->x - l->ax) * hy - (v->y - l->ay) * hx);
    if (!(dot > 0.0f)) continue;
    if (first_any < 0) first_any = box->lane;
    if (first_cur < 0 && l->road == cur_road) first_cur = box->lane;
    if (first_next < 0 && l->road == next_road) first_next = box->lane;
  }
  int lane = first_cur >= 0 ? first_cur : (first_next >= 0 ? first_next : first_any);
  v->on_lane = 1;
  if (lane < 0) {
    v->on_lane = 0;
    lane = v->lane;
  }
  v->lane = lane;
  /* _update_target_checkpoints */
  if (v->ck0 != v->ck1) {
    const PgdLane* l = lane_at(o, m, lane);
    float lon, lat;
    lane_local(l, v->x, v->y, &lon, &lat);
    int start = road_at(o, m, l->road)->start_node;
    int n = v->s->route_len;
    if (lon < 5.0f) {
      for (int j = v->ck1; j < n - 1; ++j) {
        if (route_node(o, v, j) == start) {
          v->ck0 = j;
          v->ck1 = (j + 1 == n - 1) ? j : j + 1;
          break;
        }
      }
    }
  }
}

/* ---- IDM traffic policy: policy/idm_policy.py ---------------------------------------------------- */
typedef struct {
  int exists[3];     /* left / current / right lane */
  int front[3], back[3];
  float fdist[3], bdist[3];
} FrontBack;

static float speed_kmh(const Veh* v) { return clipf(v->v * 3.6f, 0.0f, 100000.0f); }

static int near_objects(const Env* e, int self, int* out) { /* lidar.py:109-124, centre distance */
  int n = 0;
  const Veh* me = &e->v[self];
  for (int j = 0; j < e->n_slots; ++j) {
    if (j == self || !e->v[j].alive) continue;
    float dx = e->v[j].x - me->x, dy = e->v[j].y - me->y;
    if (dx * dx + dy * dy < LIDAR_RANGE * LIDAR_RANGE) out[n++] = j;
  }
  return n;
}

static void front_back(const Oracle* o, const PgdMap* m, const Env* e, int self, const int* objs, int n_objs,
                       int lane_id, int with_side, FrontBack* fb) { /* idm_policy.py:83-133 */
  const Veh* me = &e->v[self];
  const PgdLane* lane = lane_at(o, m, lane_id);
  const PgdRoad* road = road_at(o, m, lane->road);
  int lanes[3] = {-1, lane_id, -1};
  if (with_side) {
    if (lane->idx > 0) lanes[0] = road->first_lane + lane->idx - 1;
    if (lane->idx + 1 < road->n_lanes) lanes[2] = road->first_lane + lane->idx + 1;
  }
  for (int i = 0; i < 3; ++i) {
    fb->exists[i] = lanes[i] >= 0;
    fb->front[i] = fb->back[i] = -1;
    fb->fdist[i] = fb->bdist[i] = IDM_MAX_LONG;
    if (lanes[i] < 0) continue;
    const PgdLane* l = lane_at(o, m, lanes[i]);
    float cur_long, lat;
    lane_local(l, me->x, me->y, &cur_long, &lat);
    float left_long = l->length - cur_long;
    int found_front = 0, found_back = 0;
    for (int k = 0; k < n_objs; ++k) {
      const Veh* ob = &e->v[objs[k]];
      const PgdLane* ol = lane_at(o, m, ob->lane);
      if (ob->lane == lanes[i]) {
        float lg;
        lane_local(l, ob->x, ob->y, &lg, &lat);
        lg -= cur_long;
        if (fb->fdist[i] > lg && lg > 0.0f) {
          fb->fdist[i] = lg;
          fb->front[i] = objs[k];
          found_front = 1;
        }
        if (lg < 0.0f && fabsf(lg) < fb->bdist[i]) {
          fb->bdist[i] = fabsf(lg);
          fb->back[i] = objs[k];
          found_back = 1;
        }
      } else if (!found_front && lane_precedes(l, ol)) {
        float lg;
        lane_local(ol, ob->x, ob->y, &lg, &lat);
        lg += left_long;
        if (fb->fdist[i] > lg && lg > 0.0f) {
          fb->fdist[i] = lg;
          fb->front[i] = objs[k];
        }
      } else if (!found_back && lane_precedes(ol, l)) {
        float lg;
        lane_local(ol, ob->x, ob->y, &lg, &lat);
        lg = ol->length - lg + cur_long;
        if (fb->bdist[i] > lg) {
          fb->bdist[i] = lg;
          fb->back[i] = objs[k];
        }
      }
    }
  }
}

static float pid(float* p_err, float* i_err, float kp, float ki, float kd, float err) { /* PID_controller.py */
  *i_err += err;
  float d = err - *p_err;
  *p_err = err;
  return -kp * *p_err - ki * *i_err - kd * d;
}

static void idm_act(const Oracle* o, const PgdMap* m, Env* e, int self, float* steer_out, float* acc_out) {
  Veh* v = &e->v[self];
  int cur_road_id = route_road(o, v, v->ck0);
  const PgdRoad* cur_road = road_at(o, m, cur_road_id);
  /* move_to_next_road (:222-242) */
  int ok;
  if (v->rt_lane < 0) {
    v->rt_lane = v->lane;
    ok = lane_at(o, m, v->rt_lane)->road == cur_road_id;
  } else if (lane_at(o, m, v->rt_lane)->road != cur_road_id) {
    ok = 0;
    for (int k = 0; k < cur_road->n_lanes; ++k) {
      if (lane_precedes(lane_at(o, m, v->rt_lane), lane_at(o, m, cur_road->first_lane + k))) {
        v->rt_lane = cur_road->first_lane + k;
        ok = 1;
        break;
      }
    }
  } else if (lane_at(o, m, v->lane)->road == cur_road_id && v->rt_lane != v->lane) {
    v->rt_lane = v->lane;
    v->timer = v->s->rnd25[v->rnd_n % PGD_N_RND25];
    v->rnd_n++;
    ok = 1;
  } else {
    ok = 1;
  }
  int objs[PGD_MAX_SLOTS];
  int n_objs = near_objects(e, self, objs);
  FrontBack fb;
  int front_obj, steer_lane = v->rt_lane;
  float front_dist;
  if (!ok) {
    front_back(o, m, e, self, objs, n_objs, v->rt_lane, 0, &fb);
    front_obj = fb.front[1];
    front_dist = fb.fdist[1];
  } else { /* lane_change_policy (:281-353) */
    front_back(o, m, e, self, objs, n_objs, v->rt_lane, 1, &fb);
    int n_cur = cur_road->n_lanes;
    int lo = 0, hi = n_cur - 1; /* available_routing_index_range */
    int decided = 0;
    int idx = lane_at(o, m, v->rt_lane)->idx;
    if (v->ck0 != v->ck1) {
      const PgdRoad* nxt = road_at(o, m, route_road(o, v, v->ck1));
      int diff = n_cur - nxt->n_lanes;
      if (diff > 0) {
        if (lane_precedes(lane_at(o, m, cur_road->first_lane), lane_at(o, m, nxt->first_lane))) {
          lo = 0;
          hi = nxt->n_lanes - 1;
        } else {
          lo = diff;
          hi = n_cur - 1;
        }
        if (idx < lo || idx > hi) {
          decided = 1;
          int side = idx > hi ? 0 : 2; /* change to left : right */
          if (fb.bdist[side] < IDM_SAFE_DIST || fb.fdist[side] < 5.0f) {
            v->target_speed = IDM_CREEP_SPEED;
            front_obj = fb.front[1];
            front_dist = fb.fdist[1];
          } else {
            v->target_speed = IDM_NORMAL_SPEED;
            front_obj = fb.front[side];
            front_dist = fb.fdist[side];
            steer_lane = cur_road->first_lane + idx + (side == 0 ? -1 : 1);
          }
        }
      }
    }
    if (!decided) {
      float my_speed = speed_kmh(v);
      if (fabsf(my_speed - IDM_NORMAL_SPEED) > 3.0f && fb.front[1] >= 0 &&
          fabsf(speed_kmh(&e->v[fb.front[1]]) - IDM_NORMAL_SPEED) > 3.0f && v->timer > IDM_LANE_CHANGE_FREQ) {
        float side_speed[3];
        int side_ok[3];
        for (int sd = 0; sd < 3; sd += 2) {
          if (fb.front[sd] >= 0) {
            side_speed[sd] = speed_kmh(&e->v[fb.front[sd]]);
            side_ok[sd] = 1;
          } else if (fb.exists[sd] && fb.fdist[sd] > IDM_SAFE_DIST && fb.bdist[sd] > IDM_SAFE_DIST) {
            side_speed[sd] = IDM_MAX_SPEED;
            side_ok[sd] = 1;
          } else {
            side_ok[sd] = 0;
          }
        }
        float front_speed = speed_kmh(&e->v[fb.front[1]]);
        if (side_ok[0] && side_speed[0] - front_speed > IDM_SPEED_INCREASE && idx - 1 >= lo && idx - 1 <= hi) {
          decided = 1;
          front_obj = fb.front[0];
          front_dist = fb.fdist[0];
          steer_lane = cur_road->first_lane + idx - 1;
        } else if (side_ok[2] && side_speed[2] - front_speed > IDM_SPEED_INCREASE && idx + 1 >= lo && idx + 1 <= hi) {
          decided = 1;
          front_obj = fb.front[2];
          front_dist = fb.fdist[2];
          steer_lane = cur_road->first_lane + idx + 1;
        }
      }
    }
    if (!decided) {
      v->target_speed = IDM_NORMAL_SPEED;
      v->timer += 1;
      front_obj = fb.front[1];
      front_dist = fb.fdist[1];
    }
  }
  /* steering_control (:244-252) */
  const PgdLane* tl = lane_at(o, m, steer_lane);
  float lon, lat;
  lane_local(tl, v->x, v->y, &lon, &lat);
  float lane_heading = lane_heading_at(tl, lon + 1.0f);
  float steering = pid(&v->hp, &v->hi, 1.7f, 0.01f, 3.5f, wrap_to_pi(lane_heading - v->h));
  steering += pid(&v->lp, &v->li, 0.3f, 0.002f, 0.05f, -lat);
  /* acceleration (:254-271), speeds in km/h as in the reference */
  float sp = speed_kmh(v);
  float acc = 1.0f - pgd_pow10f(fmaxf(sp, 0.0f) / v->target_speed);
  if (front_obj >= 0) {
    const Veh* f = &e->v[front_obj];
    float hx = cosf_(v->h), hy = sinf_(v->h);
    float fs = speed_kmh(f);
    float dvx = sp * hx - fs * cosf_(f->h), dvy = sp * hy - fs * sinf_(f->h);
    float dv = dvx * hx + dvy * hy;
    float d_star = 10.0f + sp * 1.5f + sp * dv / (2.0f * sqrtf(5.0f));
    float d = front_dist;
    if (!(fabsf(d) > 1e-2f)) d = d > 0.0f ? 1e-2f : -1e-2f; /* not_zero */
    float ratio = d_star / d;
    acc -= ratio * ratio;
  }
  *steer_out = steering;
  *acc_out = acc;
}

/* ---- vehicle dynamics: planar stand-in for BulletVehicle (base_vehicle.py:343-376,488-575) ------- */
static void physics_substep(Veh* v, float dt, int overspeed) {
  if (v->airborne > 0) { /* placed 1 m above the road; no wheel contact while it drops (:311) */
    v->airborne--;
    return;
  }
  const PgdSlot* s = v->s;
  float mu_g = s->friction * GRAVITY;
  float speed = v->v;
  if (v->throttle > 0.0f && !overspeed) { /* engine force on 4 wheels; Bullet ignores the brake when it is non-zero */
    float a = fminf(4.0f * s->max_engine * v->throttle / s->mass, mu_g);
    speed += a * dt;
  } else { /* per-wheel brake impulse: 2.0 idle, |throttle| * max_brake_force when braking */
    float imp = v->throttle >= 0.0f ? 2.0f : -v->throttle * s->max_brake;
    float dv = fminf(4.0f * imp / s->mass, mu_g * dt);
    speed = fmaxf(speed - dv, 0.0f);
  }
  float delta = clipf(-v->steer * s->max_steer, -1.4f, 1.4f); /* +steering = left = heading decreases */
  float tb = s->lr / (s->lf + s->lr) * pgd_tanf(delta);
  float k_yaw = tb / sqrtf(1.0f + tb * tb) / s->lr; /* kinematic yaw rate per unit speed: sin(slip angle) / lr */
  /* yaw rate relaxes towards the kinematic-bicycle value (tyre relaxation + yaw inertia, tau = 0.1 s) ... */
  float yaw = v->w + (speed * k_yaw - v->w) * (dt / YAW_TAU);
  /* ... and the tyres cannot give more than mu*g of lateral acceleration */
  if (speed * fabsf(yaw) > mu_g) yaw = copysignf(mu_g / speed, yaw);
  float sb = speed > 1e-3f ? clipf(yaw * s->lr / speed, -1.0f, 1.0f) : 0.0f;
  float cb = sqrtf(fmaxf(1.0f - sb * sb, 0.0f));
  float ch = cosf_(v->h), sh = sinf_(v->h);
  v->x += speed * (ch * cb - sh * sb) * dt;
  v->y += speed * (sh * cb + ch * sb) * dt;
  float h = v->h + yaw * dt;
  if (h > PI_F) h -= TWO_PI_F;
  if (h < -PI_F) h += TWO_PI_F;
  v->w = yaw;
  v->h = h;
  v->v = speed;
}

/* ---- observation pieces ------------------------------------------------------------------------ */
static void project(float hx, float hy, float vx, float vy, float* fwd, float* side) { /* base_vehicle.py:460-475 */
  const float n = 1.0f + 1e-6f;
  *fwd = (vx * hx + vy * hy) / n;
  *side = (vx * -hy + vy * hx) / n;
}

static void navi_info(const Oracle* o, const PgdMap* m, const Veh* v, int road_id, int n_ref, float* out) {
  /* navigation.py:213-260 */
  const PgdRoad* road = road_at(o, m, road_id);
  const PgdLane* ref = lane_at(o, m, road->first_lane);
  float later_middle = ((float)n_ref / 2.0f - 0.5f) * m->lane_width;
  float cx, cy;
  lane_position(ref, ref->length, later_middle, &cx, &cy);
  float dx = cx - v->x, dy = cy - v->y;
  float dn = sqrtf(dx * dx + dy * dy);
  if (dn > 50.0f) {
    dx = dx / dn * 50.0f;
    dy = dy / dn * 50.0f;
  }
  float ph, ps;
  project(cosf_(v->h), sinf_(v->h), dx, dy, &ph, &ps);
  float bend = 0.0f, dir = 0.0f, angle = 0.0f;
  if (ref->kind == PGD_LANE_ARC) {
    bend = ref->radius / (60.0f + (float)n_ref * m->lane_width);
    dir = ref->dir;
    angle = ref->length / ref->radius;
  }
  out[0] = clipf((ph / 50.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
  out[1] = clipf((ps / 50.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
  out[2] = clipf(bend, 0.0f, 1.0f);
  out[3] = clipf((dir + 1.0f) / 2.0f, 0.0f, 1.0f);
  out[4] = clipf((angle * (180.0f / PI_F) / 135.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
}

static float heading_diff(const PgdLane* l, const Veh* v) { /* base_vehicle.py:433-458 */
  float lx, ly;
  if (l->kind == PGD_LANE_STRAIGHT) {
    lx = -l->ay;
    ly = l->ax;
  } else if (l->dir < 0.0f) {
    lx = v->x - l->ax;
    ly = v->y - l->ay;
  } else {
    lx = l->ax - v->x;
    ly = l->ay - v->y;
  }
  float ln = sqrtf(lx * lx + ly * ly);
  if (!(ln > 0.0f)) return 0.0f;
  float c = (cosf_(v->h) * lx + sinf_(v->h) * ly) / ln;
  return clipf(c, -1.0f, 1.0f) / 2.0f + 0.5f;
}

/* One beam of a side / lane-line detector (distance_detector.py:65-94,137-152; cutils.pyx:43-54,103-123): from the
 * ego centre towards i * 2pi/n + 90 deg + heading, against the line ghosts of the map -- continuous (white / yellow)
 * always, broken ones only for the lane-line detector.  Brute force over every box of the map. */
static float detector_beam(const Oracle* o, const PgdMap* m, const Veh* ego, int i, int n, float distance,
                           int with_broken) {
  float ang = (float)i * (TWO_PI_F / (float)n) + PI_F / 2 + ego->h;
  float dx = cosf_(ang) * distance, dy = sinf_(ang) * distance;
  float best = 1.0f;
  for (int b = 0; b < m->n_boxes; ++b) {
    const PgdBox* box = &o->t.boxes[m->box_off + b];
    if (!(box->kind == PGD_BOX_WHITE || box->kind == PGD_BOX_YELLOW || (with_broken && box->kind == PGD_BOX_BROKEN)))
      continue;
    Rect r = {box->cx, box->cy, box->ux, box->uy, box->hl, box->hw};
    best = fminf(best, ray_rect(ego->x, ego->y, dx, dy, &r));
  }
  return best;
}

/* After-step bookkeeping + observation + reward + done for the ego.  `fresh` = called from reset. */
static void post_step(Oracle* o, Env* e, float last_x, float last_y, float last_h, int crash_bits, int fresh,
                      float* obs, float* reward, uint8_t* done, PgdInfo* info) {
  const PgdMap* m = map_of(o, e);
  const PgdConfig* c = &o->cfg;
  Veh* ego = &e->v[0];
  const int crash_vehicle = crash_bits & 1, crash_object = (crash_bits >> 1) & 1;
  /* after_step of every moving vehicle (agent_manager.py:201-203, traffic_manager.py:91-109) */
  localize(o, m, ego);
  for (int i = 1; i < e->n_slots; ++i) {
    Veh* v = &e->v[i];
    if (!v->alive || !v->active) continue;
    localize(o, m, v);
    if (!v->on_lane) v->alive = 0;
  }
  /* _state_check (base_vehicle.py:615-644): chassis rectangle against line ghosts and sidewalks */
  Rect er = veh_rect(ego);
  uint32_t flags = 0;
  int n_cand = m->n_boxes, e0 = 0;
  const int32_t* ent = NULL;
  if (o->fast) { /* the bucket under the chassis centre lists every box the chassis can touch (PGD_GRID_MARGIN) */
    n_cand = 0;
    int cx = (int)floorf((ego->x - m->x0) * m->inv_cell), cy = (int)floorf((ego->y - m->y0) * m->inv_cell);
    if (cx >= 0 && cy >= 0 && cx < m->nx && cy < m->ny) {
      int cell = m->cell_off + cy * m->nx + cx;
      e0 = o->t.cell_start[cell];
      n_cand = o->t.cell_start[cell + 1] - e0;
      ent = o->t.cell_entries + m->entry_off;
    }
  }
  for (int k = 0; k < n_cand; ++k) {
    int b = ent ? (ent[e0 + k] & PGD_ENTRY_ID_MASK) : k;
    const PgdBox* box = &o->t.boxes[m->box_off + b];
    if (box->kind == PGD_BOX_LANE) continue;
    Rect r = {box->cx, box->cy, box->ux, box->uy, box->hl, box->hw};
    if (!rect_overlap(&er, &r)) continue;
    if (box->kind == PGD_BOX_WHITE) flags |= PGD_F_ON_WHITE;
    else if (box->kind == PGD_BOX_YELLOW) flags |= PGD_F_ON_YELLOW;
    else if (box->kind == PGD_BOX_BROKEN) flags |= PGD_F_ON_BROKEN;
    else flags |= PGD_F_CRASH_SIDEWALK;
  }
  if (ego->on_lane) flags |= PGD_F_ON_LANE;
  if (crash_vehicle) flags |= PGD_F_CRASH_VEHICLE;
  if (crash_object) flags |= PGD_F_CRASH_OBJECT;
  /* route geometry */
  int cur_road_id = route_road(o, ego, ego->ck0);
  const PgdRoad* cur_road = road_at(o, m, cur_road_id);
  int n_ref = cur_road->n_lanes;
  float lon0, lat0;
  lane_local(lane_at(o, m, cur_road->first_lane), ego->x, ego->y, &lon0, &lat0);
  float to_left = lat0 + m->lane_width / 2.0f; /* base_vehicle.py:383-388 */
  float to_right = m->lane_width * (float)n_ref - to_left;
  if (to_left < 0.0f || to_right < 0.0f) flags |= PGD_F_OUT_OF_ROUTE;
  /* arrive_destination (base_vehicle.py:738-745) */
  {
    const PgdRoad* fr = road_at(o, m, route_road(o, ego, ego->s->route_len - 2));
    const PgdLane* fl = lane_at(o, m, fr->first_lane + fr->n_lanes - 1);
    float lon, lat;
    lane_local(fl, ego->x, ego->y, &lon, &lat);
    if (fl->length - 5.0f < lon && lon < fl->length + 5.0f && m->lane_width / 2.0f >= lat &&
        lat >= (0.5f - (float)n_ref) * m->lane_width)
      flags |= PGD_F_ARRIVE_DEST;
  }
  int out_of_road = (flags & (PGD_F_ON_YELLOW | PGD_F_ON_WHITE | PGD_F_CRASH_SIDEWALK)) || !ego->on_lane;
  if (c->out_of_route_done && (flags & PGD_F_OUT_OF_ROUTE)) out_of_road = 1;
  if (out_of_road) flags |= PGD_F_OUT_OF_ROAD;

  /* ---- observation (obs/state_obs.py:58-170) ----
   * layout: [side detector beams | left, right] + 6 state values + [lane-line detector beams] + 10 navi + 16 + 240 */
  float sp = speed_kmh(ego);
  const int n_first = c->n_side > 0 ? c->n_side : 2;
  float* const obs_row = obs;
  if (c->n_side > 0) {
    for (int i = 0; i < c->n_side; ++i) obs[i] = detector_beam(o, m, ego, i, c->n_side, c->side_distance, 0);
  } else {
    obs[0] = clipf(to_left / 18.0f, 0.0f, 1.0f);
    obs[1] = clipf(to_right / 18.0f, 0.0f, 1.0f);
  }
  for (int i = 0; i < c->n_lane_line; ++i)
    obs[n_first + 6 + i] = detector_beam(o, m, ego, i, c->n_lane_line, c->lane_line_distance, 1);
  const int n_extra = c->random_agent_model ? 2 : 0;
  if (n_extra) { /* obs/state_obs.py:103-105: LENGTH / MAX_LENGTH, WIDTH / MAX_WIDTH (base_vehicle.py:83-84) */
    obs_row[n_first + 6 + c->n_lane_line] = clipf(ego->s->length / 10.0f, 0.0f, 1.0f);
    obs_row[n_first + 6 + c->n_lane_line + 1] = clipf(ego->s->width / 2.5f, 0.0f, 1.0f);
  }
  float* const navi = obs_row + n_first + 6 + c->n_lane_line + n_extra;
  obs = obs_row + n_first - 2; /* obs[2..7] below are the six state values */
  obs[2] = heading_diff(lane_at(o, m, cur_road->first_lane + n_ref - 1), ego);
  obs[3] = clipf((sp + 1.0f) / (MAX_SPEED_KMH + 1.0f), 0.0f, 1.0f);
  obs[4] = clipf((ego->steer / 60.0f + 1.0f) / 2.0f, 0.0f, 1.0f);
  obs[5] = clipf((e->prev_steer + 1.0f) / 2.0f, 0.0f, 1.0f);
  obs[6] = clipf((e->prev_throttle + 1.0f) / 2.0f, 0.0f, 1.0f);
  /* yaw rate: arccos(clip(cos(angle between headings), 0, 1)) / 0.1 (state_obs.py:87-94).  arccos is
   * ill-conditioned near 1 in float32, so the identical quantity min(|wrapped heading change|, pi/2) is used. */
  obs[7] = clipf(fminf(fabsf(wrap_to_pi(ego->h - last_h)), PI_F / 2) / 0.1f, 0.0f, 1.0f);
  navi_info(o, m, ego, cur_road_id, n_ref, navi);
  navi_info(o, m, ego, route_road(o, ego, ego->ck1), n_ref, navi + 5);
  obs = navi - 8; /* from here on obs[18..] = neighbours, obs[34..] = lidar, as in the default layout */
  /* 4 nearest vehicles (lidar.py:55-77) */
  {
    int objs[PGD_MAX_SLOTS];
    int n = near_objects(e, 0, objs);
    { /* get_surrounding_vehicles (lidar.py:46-54): cones and barriers are not vehicles */
      int m = 0;
      for (int k = 0; k < n; ++k)
        if (e->v[objs[k]].s->type < PGD_TYPE_OBJECT) objs[m++] = objs[k];
      n = m;
    }
    float d2[PGD_MAX_SLOTS];
    for (int k = 0; k < n; ++k) {
      float dx = e->v[objs[k]].x - ego->x, dy = e->v[objs[k]].y - ego->y;
      d2[k] = dx * dx + dy * dy;
    }
    float hx = cosf_(ego->h), hy = sinf_(ego->h);
    for (int slot = 0; slot < 4; ++slot) {
      int best = -1;
      for (int k = 0; k < n; ++k)
        if (objs[k] >= 0 && (best < 0 || d2[k] < d2[best])) best = k;
      float* q = obs + 18 + 4 * slot;
      if (best < 0) {
        q[0] = q[1] = q[2] = q[3] = 0.0f;
        continue;
      }
      const Veh* w = &e->v[objs[best]];
      objs[best] = -1;
      float pf, ps, vf, vs;
      project(hx, hy, w->x - ego->x, w->y - ego->y, &pf, &ps);
      float ws = speed_kmh(w);
      project(hx, hy, ws * cosf_(w->h) - sp * hx, ws * sinf_(w->h) - sp * hy, &vf, &vs);
      q[0] = clipf((pf / LIDAR_RANGE + 1.0f) / 2.0f, 0.0f, 1.0f);
      q[1] = clipf((ps / LIDAR_RANGE + 1.0f) / 2.0f, 0.0f, 1.0f);
      q[2] = clipf((vf / MAX_SPEED_KMH + 1.0f) / 2.0f, 0.0f, 1.0f);
      q[3] = clipf((vs / MAX_SPEED_KMH + 1.0f) / 2.0f, 0.0f, 1.0f);
    }
  }
  /* 240-beam lidar against every other chassis (cutils.pyx:60-142) */
  {
    Rect rects[PGD_MAX_SLOTS];
    int n_rects = 0;
    for (int j = 1; j < e->n_slots; ++j) {
      if (!e->v[j].alive) continue;
      if (o->fast) { /* a chassis further away than the range plus its half diagonal cannot be hit */
        float dx = e->v[j].x - ego->x, dy = e->v[j].y - ego->y;
        float hl = e->v[j].s->length * 0.5f, hw = e->v[j].s->width * 0.5f;
        float reach = LIDAR_RANGE + sqrtf(hl * hl + hw * hw) * 1.001f + 1e-3f;
        if (dx * dx + dy * dy > reach * reach) continue;
      }
      rects[n_rects++] = veh_rect(&e->v[j]);
    }
    for (int i = 0; i < PGD_LIDAR_BEAMS; ++i) {
      float ang = (float)i * (TWO_PI_F / (float)PGD_LIDAR_BEAMS) + ego->h;
      float sn, cs;
      pgd_sincosf(ang, &sn, &cs);
      float dx = cs * LIDAR_RANGE, dy = sn * LIDAR_RANGE;
      float best = 1.0f;
      for (int j = 0; j < n_rects; ++j) best = fminf(best, ray_rect(ego->x, ego->y, dx, dy, &rects[j]));
      /* _add_noise_to_cloud_points (obs/state_obs.py:172-182): Gaussian noise, clip, dropout */
      if (c->lidar_gaussian_noise > 0.0f || c->lidar_dropout_prob > 0.0f)
        best = pgd_lidar_noise(best, c->lidar_gaussian_noise, c->lidar_dropout_prob,
                               pgd_noise_key((uint32_t)c->noise_seed, o->call_index, (uint32_t)(e - o->envs), (uint32_t)i));
      obs[34 + i] = best;
    }
  }

  /* ---- reward / cost / done (envs/pgdrive_env.py:162-258) ---- */
  float r = 0.0f, step_reward = 0.0f, cost = 0.0f, step_energy = 0.0f;
  int is_done = 0;
  if (!fresh) {
    const PgdLane* el = lane_at(o, m, ego->lane);
    const PgdLane* rl;
    float sign = 1.0f;
    if (el->road == cur_road_id) {
      rl = el;
    } else {
      rl = lane_at(o, m, cur_road->first_lane);
      sign = road_at(o, m, el->road)->negative ? -1.0f : 1.0f;
    }
    float long_last, long_now, lat_last, lat_now;
    lane_local(rl, last_x, last_y, &long_last, &lat_last);
    lane_local(rl, ego->x, ego->y, &long_now, &lat_now);
    float lateral_factor = 1.0f;
    if (c->use_lateral) lateral_factor = clipf(1.0f - 2.0f * fabsf(lat_now) / m->lane_width, 0.0f, 1.0f);
    r += c->driving_reward * (long_now - long_last) * lateral_factor * sign;
    r += c->speed_reward * (sp / MAX_SPEED_KMH) * sign;
    step_reward = r;
    if (flags & PGD_F_ARRIVE_DEST) r = c->success_reward;
    else if (out_of_road) r = -c->out_of_road_penalty;
    else if (crash_vehicle) r = -c->crash_vehicle_penalty;
    else if (crash_object) r = -c->crash_object_penalty;
    if (out_of_road) cost = c->out_of_road_cost;
    else if (crash_vehicle) cost = c->crash_vehicle_cost;
    else if (crash_object) cost = c->crash_object_cost;
    is_done = (flags & PGD_F_ARRIVE_DEST) || out_of_road || crash_vehicle || crash_object;
    /* SafePGDriveEnv.done_function (safe_pgdrive_env.py:44-51), literally: a step with a crash never ends the episode */
    if (c->safe_rl_env && (crash_vehicle || crash_object)) is_done = 0;
    float ddx = last_x - ego->x, ddy = last_y - ego->y; /* base_vehicle.py:278-290 */
    step_energy = 3.25f * pgd_expf(0.01f * sp) * (sqrtf(ddx * ddx + ddy * ddy) / 1000.0f) / 100.0f * 1000.0f;
    e->energy += step_energy;
    e->ep_reward += r;
    e->ep_len += 1;
    if (c->horizon > 0 && e->ep_len >= c->horizon) {
      is_done = 1;
      flags |= PGD_F_MAX_STEP;
    }
    if (e->done) is_done = 1; /* done is sticky (base_env.py:315-316) */
    e->done = is_done;
  } else {
    flags |= PGD_F_WAS_RESET;
  }
  if (reward) *reward = r;
  if (done) *done = (uint8_t)is_done;
  if (info) {
    info->velocity = sp;
    info->steering = ego->steer;
    info->acceleration = ego->throttle;
    info->step_energy = step_energy;
    info->episode_energy = e->energy;
    info->step_reward = step_reward;
    info->episode_reward = e->ep_reward;
    info->cost = cost;
    info->episode_length = e->ep_len;
    info->flags = flags;
  }
}

/* ---- public entry points (ctypes) ---------------------------------------------------------------- */
void* orc_create(const PgdTables* t, const PgdConfig* cfg) {
  for (int i = 0; i < t->n_episodes; ++i) /* same limits as pgd_load_tables */
    if (t->episodes[i].n_slots > PGD_MAX_SLOTS || t->episodes[i].n_slots > cfg->num_slots ||
        t->episodes[i].n_groups > PGD_MAX_GROUPS)
      return NULL;
  Oracle* o = (Oracle*)calloc(1, sizeof(Oracle));
  o->t = *t;
  o->cfg = *cfg;
  o->envs = (Env*)calloc((size_t)cfg->num_envs, sizeof(Env));
  return o;
}

void orc_set_call_index(void* h, uint32_t idx) { ((Oracle*)h)->call_index = idx; }
void orc_set_fast(void* h, int on) { ((Oracle*)h)->fast = on; }

void orc_destroy(void* h) {
  Oracle* o = (Oracle*)h;
  free(o->envs);
  free(o);
}

static void load_template(Oracle* o, Env* e, int episode) {
  const PgdEpisode* ep = &o->t.episodes[episode];
  memset(e, 0, sizeof(Env));
  e->episode = episode;
  e->n_slots = ep->n_slots;
  for (int i = 0; i < ep->n_slots; ++i) {
    const PgdSlot* s = &o->t.slots[ep->slot_off + i];
    Veh* v = &e->v[i];
    v->s = s;
    v->x = s->x;
    v->y = s->y;
    v->h = s->heading;
    v->lane = s->lane;
    v->ck0 = 0;
    v->ck1 = s->route_len > 2 ? 1 : 0;
    v->rt_lane = -1;
    v->timer = s->overtake_timer;
    v->airborne = s->drop_substeps;
    v->target_speed = IDM_NORMAL_SPEED;
    v->alive = 1;
    v->active = i == 0 || s->group == PGD_GROUP_AWAKE; /* respawn mode: traffic drives from the first step */
    v->on_lane = 1;
  }
}

/* reset(force_seed): base_env.py:269-301 (template copy, then after_step + observe) */
void orc_reset(void* h, int env, int episode, float* obs, PgdInfo* info) {
  Oracle* o = (Oracle*)h;
  Env* e = &o->envs[env];
  load_template(o, e, episode);
  /* Spawned traffic keeps its spawn lane (update_map_info, base_vehicle.py:589-613) and gets no
   * after_step until triggered; only the ego is localised by _get_reset_return's engine.after_step(). */
  post_step(o, e, e->v[0].x, e->v[0].y, e->v[0].h, 0, 1, obs, NULL, NULL, info);
}

void orc_step(void* h, int env, const float* action, float* obs, float* reward, uint8_t* done, PgdInfo* info) {
  Oracle* o = (Oracle*)h;
  Env* e = &o->envs[env];
  const PgdConfig* c = &o->cfg;
  if (c->auto_reset && e->done) {
    orc_reset(h, env, e->episode, obs, info);
    *reward = 0.0f;
    *done = 0;
    return;
  }
  const PgdMap* m = map_of(o, e);
  const PgdEpisode* ep = ep_of(o, e);
  Veh* ego = &e->v[0];
  /* EnvInputPolicy.act: clip, NaN -> -1 (env_input_policy.py:17-26, cutils.pyx:153) */
  float steer = clipf(action[0], -1.0f, 1.0f), throttle = clipf(action[1], -1.0f, 1.0f);
  float last_x = ego->x, last_y = ego->y, last_h = ego->h;
  e->prev_throttle = ego->throttle; /* last_current_action[0] after the push (base_vehicle.py:248) */
  if (c->increment_steering) { /* _set_incremental_action (base_vehicle.py:351-358); ego->hp keeps the raw action */
    e->prev_steer = ego->hp;
    ego->hp = steer;
    ego->steer = clipf(ego->steer + steer * 0.05f, -1.0f, 1.0f);
  } else {
    e->prev_steer = ego->steer;
    ego->steer = steer;
  }
  ego->throttle = throttle;
  /* TrafficManager.before_step (traffic_manager.py:71-89): wake the next block's vehicles */
  if (e->next_group < ep->n_groups && lane_at(o, m, ego->lane)->road == ep->trigger_road[e->next_group]) {
    for (int i = 1; i < e->n_slots; ++i)
      if (e->v[i].s->group == e->next_group) e->v[i].active = 1;
    e->next_group++;
  }
  for (int i = 1; i < e->n_slots; ++i) {
    Veh* v = &e->v[i];
    if (!v->alive || !v->active) continue;
    idm_act(o, m, e, i, &v->steer, &v->throttle);
  }
  int overspeed[PGD_MAX_SLOTS];
  for (int i = 0; i < e->n_slots; ++i) overspeed[i] = speed_kmh(&e->v[i]) > MAX_SPEED_KMH;
  /* 5 x doPhysics(0.02) (base_engine.py:206-232); chassis contacts are sticky within the step */
  int crash = 0; /* bit 0: a vehicle chassis, bit 1: a traffic object (collision_callback.py:13-30) */
  int touched[PGD_MAX_SLOTS] = {0};
  for (int k = 0; k < c->decision_repeat; ++k) {
    for (int i = 0; i < e->n_slots; ++i)
      if (e->v[i].alive && e->v[i].s->group != PGD_GROUP_STATIC) physics_substep(&e->v[i], c->dt, overspeed[i]);
      else if (e->v[i].alive && e->v[i].airborne > 0) e->v[i].airborne--; /* a broken-down vehicle still drops */
    Rect er = veh_rect(ego);
    for (int i = 1; i < e->n_slots; ++i) {
      if (!e->v[i].alive) continue;
      Rect r = veh_rect(&e->v[i]);
      if (rect_overlap(&er, &r)) touched[i] = 1;
    }
  }
  for (int i = 1; i < e->n_slots; ++i) {
    if (!touched[i]) continue;
    if (e->v[i].s->type >= PGD_TYPE_OBJECT) {
      if (!e->v[i].crashed) { /* COST_ONCE */
        crash |= 2;
        e->v[i].crashed = 1;
      }
    } else {
      crash |= 1;
    }
  }
  post_step(o, e, last_x, last_y, last_h, crash, 0, obs, reward, done, info);
}

void orc_step_range(void* h, int env0, int env1, const float* actions, float* obs, float* reward, uint8_t* done,
                    PgdInfo* info) {
  for (int e = env0; e < env1; ++e)
    orc_step(h, e, actions + 2 * e, obs + (size_t)pgd_obs_dim(&((Oracle*)h)->cfg) * e, reward + e, done + e, info + e);
}

void orc_get_state(void* h, int env, PgdEnvState* out) {
  Oracle* o = (Oracle*)h;
  const Env* e = &o->envs[env];
  memset(out, 0, sizeof(*out));
  out->episode = e->episode;
  out->next_group = e->next_group;
  out->done = e->done;
  out->ep_len = e->ep_len;
  out->prev_steer = e->prev_steer;
  out->prev_throttle = e->prev_throttle;
  out->ep_reward = e->ep_reward;
  out->energy = e->energy;
  for (int i = 0; i < e->n_slots; ++i) {
    const Veh* v = &e->v[i];
    PgdVehState* s = &out->veh[i];
    s->x = v->x; s->y = v->y; s->heading = v->h; s->speed = v->v;
    s->steer = v->steer; s->throttle = v->throttle;
    s->pid_hp = v->hp; s->pid_hi = v->hi; s->pid_lp = v->lp; s->pid_li = v->li;
    s->target_speed = v->target_speed;
    s->lane = v->lane; s->ck0 = v->ck0; s->ck1 = v->ck1; s->rt_lane = v->rt_lane;
    s->timer = v->timer; s->rnd_n = v->rnd_n; s->airborne = v->airborne; s->yaw_rate = v->w;
    s->flags = (v->alive ? PGD_V_ALIVE : 0) | (v->active ? PGD_V_ACTIVE : 0) | (v->on_lane ? PGD_V_ON_LANE : 0) |
               (v->crashed ? PGD_V_CRASHED : 0);
  }
}

void orc_set_state(void* h, int env, const PgdEnvState* in) {
  Oracle* o = (Oracle*)h;
  Env* e = &o->envs[env];
  load_template(o, e, in->episode);
  e->next_group = in->next_group;
  e->done = in->done;
  e->ep_len = in->ep_len;
  e->prev_steer = in->prev_steer;
  e->prev_throttle = in->prev_throttle;
  e->ep_reward = in->ep_reward;
  e->energy = in->energy;
  for (int i = 0; i < e->n_slots; ++i) {
    Veh* v = &e->v[i];
    const PgdVehState* s = &in->veh[i];
    v->x = s->x; v->y = s->y; v->h = s->heading; v->v = s->speed;
    v->steer = s->steer; v->throttle = s->throttle;
    v->hp = s->pid_hp; v->hi = s->pid_hi; v->lp = s->pid_lp; v->li = s->pid_li;
    v->target_speed = s->target_speed;
    v->lane = s->lane; v->ck0 = s->ck0; v->ck1 = s->ck1; v->rt_lane = s->rt_lane;
    v->timer = s->timer; v->rnd_n = s->rnd_n; v->airborne = s->airborne; v->w = s->yaw_rate;
    v->alive = !!(s->flags & PGD_V_ALIVE); v->active = !!(s->flags & PGD_V_ACTIVE);
    v->on_lane = !!(s->flags & PGD_V_ON_LANE);
    v->crashed = !!(s->flags & PGD_V_CRASHED);
  }
}

/* observation of the CURRENT state without stepping or mutating it (golden-vector tests) */
void orc_observe(void* h, int env, float* obs, PgdInfo* info) {
  Oracle* o = (Oracle*)h;
  Env tmp = o->envs[env];
  post_step(o, &tmp, tmp.v[0].x, tmp.v[0].y, tmp.v[0].h, 0, 1, obs, NULL, NULL, info);
}

/* small probes used by the golden-vector tests */
void orc_lane_local(const PgdLane* l, float x, float y, float* out) { lane_local(l, x, y, out, out + 1); }
void orc_lane_position(const PgdLane* l, float lon, float lat, float* out) { lane_position(l, lon, lat, out, out + 1); }
float orc_lane_heading_at(const PgdLane* l, float lon) { return lane_heading_at(l, lon); }
float orc_ray_rect(float ox, float oy, float dx, float dy, float cx, float cy, float h, float len, float wid) {
  Rect r = {cx, cy, cosf_(h), sinf_(h), len * 0.5f, wid * 0.5f};
  return ray_rect(ox, oy, dx, dy, &r);
}
