/* Procedural map generation + reset decisions, one map per call, for host AND device (PGD_HD).
 *
 * What it computes (reference, paths under /root/reference/pgdrive):
 *   block search with retry / back-tracking            component/algorithm/BIG.py:67-151
 *   the eight PGDrive-v0 block types                   component/blocks/{first_block,straight,curve,intersection,
 *                                                      t_intersection,roundabout,ramp}.py, create_block_utils.py
 *   overlap test of a new road against the map so far  utils/scene_utils.py:40-136
 *   rebuild from the stored block sequence             component/map/pg_map.py:48-71
 *   traffic slots, vehicle parameters, routes          manager/traffic_manager.py:239-290, navigation.py:99-153,
 *                                                      component/road/road_network.py:241-269
 *   static collision primitives + bucket grid          component/blocks/base_block.py:181-463 (as pgdrive_b200/tables.py)
 * and writes the flat tables of include/pgd_tables.h directly.
 *
 * Organisation (not the reference's): no objects, no strings, no dicts.  Lanes and roads live in two stack-like
 * pools (a block owns a contiguous range; undoing a block pops it), nodes are small integers whose sign is the
 * reference's "-" prefix, and the insertion-ordered-dict behaviour that fixes lane numbering is reproduced by
 * grouping a block's roads by first appearance of their start node.  All arithmetic is float64 in the reference's
 * evaluation order; trigonometry goes through pgd_dd.cuh so that host and device builds give identical bits.
 *
 * The host Python path (pgdrive_b200/mapgen.py, episode.py, tables.py) is pinned bit-exactly against the reference's
 * own output; this file is pinned against that path (tests/test_device_mapgen.py: host build here, sm_100a build on
 * the GPU box, same source).
 */
#ifndef PGD_MAPGEN_CUH
#define PGD_MAPGEN_CUH
#include "../../include/pgd_tables.h"
#include "pgd_rng.cuh"

namespace pgdgen {

enum { LT_NONE = 0, LT_BROKEN = 1, LT_CONT = 2, LT_SIDE = 3 };
enum { COL_GREY = 0, COL_YELLOW = 1 };
/* block types in the order of BLOCK_TYPE_DISTRIBUTION_V2 (component/algorithm/blocks_prob_dist.py:31-49); ids >= 7
 * have probability 0 */
enum { BK_C = 0, BK_S = 1, BK_r = 2, BK_R = 3, BK_X = 4, BK_T = 5, BK_O = 6, BK_I = 100 };
enum { ND_START = 1, ND_START2 = 2, ND_START3 = 3, ND_DECO0 = 4, ND_DECO1 = 5, ND_BLOCK_BASE = 16 };
enum {
  GEN_OK = 0, GEN_ERR_LANES = 1, GEN_ERR_ROADS = 2, GEN_ERR_BOXES = 3, GEN_ERR_CELLS = 4, GEN_ERR_ENTRIES = 5,
  GEN_ERR_QUEUE = 6, GEN_ERR_ROUTE = 7, GEN_ERR_CAND = 8, GEN_ERR_SLOTS = 9, GEN_ERR_BACKTRACK = 10,
  GEN_ERR_LOOKUP = 11, GEN_ERR_DEST = 12, GEN_ERR_GROUPS = 13, GEN_ERR_BLOCKS = 14, GEN_ERR_CONFIG = 15
};
enum { MAX_ROAD_LANES = 10, MAX_BLOCK_ROADS = 64, MAX_BLOCK_SOCKETS = 5, MAX_RESPAWN = 8 };

#define PGD_PI 3.141592653589793
#define PGD_DEG2RAD(x) ((x) * (PGD_PI / 180.0))

struct GLane {
  double sx, sy, ex, ey, dx, dy, length, heading, cx, cy, radius, ph0, ph1, width;
  int8_t kind, dir, lt[2], col[2];
};

struct GRoad {
  int32_t from, to;
  int16_t lanes[MAX_ROAD_LANES];
  int16_t n_lanes;
  int8_t removed, bbox_valid;
  double bbox[4];  // x_max, x_min, y_max, y_min
};

struct GSocket {
  int32_t from, to;  // positive road; the negative road is (-to, -from)
  int32_t owner, k;  // "<owner block>-socket<k>"
};

struct GBlock {
  int32_t type, idx;
  GSocket pre;
  int32_t road_begin, road_end, lane_begin;
  GSocket sockets[MAX_BLOCK_SOCKETS];
  int32_t n_sockets;
  int32_t respawn[MAX_RESPAWN][2];
  int32_t n_respawn;
  int32_t ring[4];
  int32_t n_ring;
  int32_t trials, part, road_no;
  uint32_t q[2];
  double length, radius, angle, exit_radius, inner_radius;
  int32_t dir, change_lane_num, decrease_increase, t_type;
  int32_t n_pos, basic, n_cross, nocheck;
  double lane_width;
};

typedef PgdGenConfig GenConfig;  // include/pgd_tables.h
typedef PgdGenCaps GenCaps;

struct GBox {
  double cx, cy, ux, uy, hl, hw;
  int32_t kind, lane;
};

struct GenScratch {  // per map, caller-provided
  GLane* lanes;
  GRoad* roads;
  GBlock* blocks;
  GBox* boxes;
  int32_t* queue;  // BFS entries: node, parent
  int32_t* cand;   // spawn candidates (road, lane, k) x caps.cand, then 4 * caps.roads ints of work arrays
  MT* mt;          // 3 generators
};

struct GenOut {  // this map's slices of the output tables + where they sit in the concatenated arrays
  PgdMap* map;
  PgdLane* lanes;
  PgdRoad* roads;
  PgdBox* boxes;
  int32_t* cell_start;
  int32_t* cell_entries;
  PgdEpisode* episode;
  PgdSlot* slots;
  int32_t* route_nodes;
  int32_t* route_roads;
  int32_t lane_off, road_off, box_off, cell_off, entry_off, slot_off, route_off, map_id;
  int32_t* counts;  // 8 ints: lanes, roads, boxes, cells(+1), entries, slots, route entries, blocks
  int32_t* sequence;  // optional [blocks][12] record of the block sequence (id, socket k, params as doubles bits)
};

struct Gen {
  GenConfig cfg;
  GenCaps caps;
  GenScratch s;
  int32_t n_lanes, n_roads, n_blocks;
  int32_t status;
};

PGD_HD inline void gen_fail(Gen& g, int code) {
  if (g.status == GEN_OK) g.status = code;
}

// ------------------------------------------------------------------------------------------------ lanes
PGD_HD inline double norm2(double x, double y) { return sqrt(x * x + y * y); }
PGD_HD inline double wrap_to_pi(double x) { return py_mod(x + PGD_PI, 2 * PGD_PI) - PGD_PI; }

PGD_HD inline void lane_position(const GLane& l, double lon, double lat, double* x, double* y) {
  if (l.kind == 0) {
    *x = l.sx + lon * l.dx + lat * -l.dy;
    *y = l.sy + lon * l.dy + lat * l.dx;
    return;
  }
  double phi = l.dir * lon / l.radius + l.ph0;
  double r = l.radius - lat * l.dir;
  dd s, c;
  dd_sincos(phi, &s, &c);
  *x = l.cx + r * c.hi;
  *y = l.cy + r * s.hi;
}

PGD_HD inline void lane_local(const GLane& l, double x, double y, double* lon, double* lat) {
  if (l.kind == 0) {
    double ax = x - l.sx, ay = y - l.sy;
    *lon = ax * l.dx + ay * l.dy;
    *lat = ax * -l.dy + ay * l.dx;
    return;
  }
  double ax = x - l.cx, ay = y - l.cy;
  double phi = cr_atan2(ay, ax);
  phi = l.ph0 + wrap_to_pi(phi - l.ph0);
  double r = norm2(ax, ay);
  *lon = l.dir * (phi - l.ph0) * l.radius;
  *lat = l.dir * (l.radius - r);
}

/* same with the platform's atan2 (<= 2 ulp): only for comparisons that are re-done exactly when close */
PGD_HD inline void lane_local_fast(const GLane& l, double x, double y, double* lon, double* lat) {
  if (l.kind == 0) {
    lane_local(l, x, y, lon, lat);
    return;
  }
  double ax = x - l.cx, ay = y - l.cy;
  double phi = atan2(ay, ax);
  phi = l.ph0 + wrap_to_pi(phi - l.ph0);
  double r = norm2(ax, ay);
  *lon = l.dir * (phi - l.ph0) * l.radius;
  *lat = l.dir * (l.radius - r);
}

PGD_HD inline double lane_heading_at(const GLane& l, double lon) {
  if (l.kind == 0) return l.heading;
  double phi = l.dir * lon / l.radius + l.ph0;
  return phi + PGD_PI / 2 * l.dir;
}

PGD_HD inline void lane_refresh(GLane& l) {
  if (l.kind == 0) {
    double vx = l.ex - l.sx, vy = l.ey - l.sy;
    l.length = norm2(vx, vy);
    l.heading = cr_atan2(vy, vx);
    l.dx = vx / l.length;
    l.dy = vy / l.length;
  } else {
    l.length = l.radius * (l.ph1 - l.ph0) * l.dir;
    lane_position(l, 0, 0, &l.sx, &l.sy);
    lane_position(l, l.length, 0, &l.ex, &l.ey);
  }
}

PGD_HD inline GLane lane_straight(double sx, double sy, double ex, double ey, double width, int lt0, int lt1) {
  GLane l;
  l.kind = 0;
  l.sx = sx; l.sy = sy; l.ex = ex; l.ey = ey;
  l.width = width;
  l.lt[0] = (int8_t)lt0; l.lt[1] = (int8_t)lt1;
  l.col[0] = l.col[1] = COL_GREY;
  l.cx = l.cy = l.radius = l.ph0 = l.ph1 = 0.0;
  l.dir = 0;
  lane_refresh(l);
  return l;
}

PGD_HD inline GLane lane_arc(double cx, double cy, double radius, double ph0, double ph1, bool clockwise, double width,
                             int lt0, int lt1) {
  GLane l;
  l.kind = 1;
  l.cx = cx; l.cy = cy;
  l.radius = radius;
  l.ph0 = ph0; l.ph1 = ph1;
  l.dir = clockwise ? 1 : -1;
  l.width = width;
  l.lt[0] = (int8_t)lt0; l.lt[1] = (int8_t)lt1;
  l.col[0] = l.col[1] = COL_GREY;
  l.dx = l.dy = l.heading = 0.0;
  lane_refresh(l);
  return l;
}

/* create_block_utils.py:162-170 */
PGD_HD inline GLane extend_straight(const GLane& lane, double extend_length, int lt0, int lt1) {
  GLane n = lane;
  n.sx = lane.ex; n.sy = lane.ey;
  lane_position(lane, lane.length + extend_length, 0, &n.ex, &n.ey);
  n.lt[0] = (int8_t)lt0; n.lt[1] = (int8_t)lt1;
  lane_refresh(n);
  return n;
}

/* create_block_utils.py:16-59; angle in radians */
PGD_HD inline void bend_then_straight(const GLane& prev, double follow_len, double radius, double angle, bool clockwise,
                                      double width, int lt0, int lt1, GLane* bend, GLane* follow) {
  int bd = clockwise ? 1 : -1;
  double cx, cy;
  lane_position(prev, prev.length, bd * radius, &cx, &cy);
  double x = -prev.dy, y = prev.dx;
  double ph0 = 0;
  if (y == 0) {
    ph0 = (x < 0) ? 0 : -PGD_PI;
  } else if (x == 0) {
    ph0 = (y < 0) ? PGD_PI / 2 : -PGD_PI / 2;
  } else {
    double base = cr_atan(y / x);
    if (x < 0) ph0 = base;
    else if (y < 0) ph0 = PGD_PI + base;
    else if (y > 0) ph0 = -PGD_PI + base;
  }
  double ph1 = ph0 + angle;
  if (!clockwise) {
    ph0 = ph0 - PGD_PI;
    ph1 = ph0 - angle;
  }
  *bend = lane_arc(cx, cy, radius, ph0, ph1, clockwise, width, lt0, lt1);
  double bx, by;
  lane_position(*bend, 2 * radius * angle / 2, 0, &bx, &by);
  double vx = bx - cx, vy = by - cy;
  double vl = norm2(vx, vy);
  double nx, ny;
  if (!clockwise) { nx = vy / vl; ny = -vx / vl; }
  else { nx = -vy / vl; ny = vx / vl; }
  *follow = lane_straight(bx, by, nx * follow_len + bx, ny * follow_len + by, width, lt0, lt1);
}

// ------------------------------------------------------------------------------------------------ road pools
PGD_HD inline int node_of(int block, int part, int road) { return ND_BLOCK_BASE + ((block * 4 + part) * 8 + road); }

PGD_HD inline int push_lane(Gen& g, const GLane& l) {
  if (g.n_lanes >= g.caps.lanes) {
    gen_fail(g, GEN_ERR_LANES);
    return g.caps.lanes - 1;
  }
  g.s.lanes[g.n_lanes] = l;
  return g.n_lanes++;
}

/* road (from, to) among roads [lo, hi) */
PGD_HD inline int find_road(const Gen& g, int from, int to, int lo, int hi) {
  for (int r = lo; r < hi; ++r)
    if (!g.s.roads[r].removed && g.s.roads[r].from == from && g.s.roads[r].to == to) return r;
  return -1;
}

PGD_HD inline int block_road(Gen& g, const GBlock& b, int from, int to) {
  int r = find_road(g, from, to, b.road_begin, g.n_roads);
  if (r < 0) {
    gen_fail(g, GEN_ERR_LOOKUP);
    return b.road_begin < g.n_roads ? b.road_begin : 0;
  }
  return r;
}
PGD_HD inline int world_road(Gen& g, const GBlock& b, int from, int to) {  // the map before this block
  int r = find_road(g, from, to, 0, b.road_begin);
  if (r < 0) {
    gen_fail(g, GEN_ERR_LOOKUP);
    return 0;
  }
  return r;
}

/* RoadNet.add_lane on the block being built; returns the lane's pool index */
PGD_HD inline int add_lane(Gen& g, GBlock& b, int from, int to, const GLane& l) {
  int li = push_lane(g, l);
  int r = find_road(g, from, to, b.road_begin, g.n_roads);
  if (r < 0) {
    if (g.n_roads >= g.caps.roads || g.n_roads - b.road_begin >= MAX_BLOCK_ROADS) {
      gen_fail(g, GEN_ERR_ROADS);
      return li;
    }
    r = g.n_roads++;
    GRoad& rd = g.s.roads[r];
    rd.from = from; rd.to = to;
    rd.n_lanes = 0;
    rd.removed = 0; rd.bbox_valid = 0;
  }
  GRoad& rd = g.s.roads[r];
  if (rd.n_lanes >= MAX_ROAD_LANES) {
    gen_fail(g, GEN_ERR_ROADS);
    return li;
  }
  rd.lanes[rd.n_lanes++] = (int16_t)li;
  return li;
}

/* iteration order of a block's own RoadNet (insertion-ordered dict of dicts): roads grouped by start node in order
 * of the node's first road; decoration last.  Returns the count. */
PGD_HD inline int block_road_order(const Gen& g, int rb, int re, int* out) {
  int n = 0;
  uint64_t done = 0;
  for (int r = rb; r < re; ++r) {
    if (g.s.roads[r].removed || ((done >> (r - rb)) & 1) || g.s.roads[r].from == ND_DECO0) continue;
    int from = g.s.roads[r].from;
    for (int r2 = r; r2 < re; ++r2) {
      if (!g.s.roads[r2].removed && g.s.roads[r2].from == from) {
        out[n++] = r2;
        done |= (uint64_t)1 << (r2 - rb);
      }
    }
  }
  for (int r = rb; r < re; ++r)
    if (!g.s.roads[r].removed && g.s.roads[r].from == ND_DECO0) out[n++] = r;
  return n;
}

// ------------------------------------------------------------------------------------------------ overlap test
/* scene_utils.py:86-128 over the first and last lane of a road */
PGD_HD inline void contour_bbox(const GLane& first, const GLane& last, double* bb) {
  const double extra = 3;
  double xmax = -1e300, xmin = 1e300, ymax = -1e300, ymin = 1e300;
  for (int e = 0; e < 2; ++e) {
    const GLane& lane = e == 0 ? first : last;
    int d = e == 0 ? -1 : 1;
    double px[6], py[6];
    int np = 0;
    lane_position(lane, 0.1, d * (lane.width / 2.0 + extra), &px[np], &py[np]); ++np;
    lane_position(lane, lane.length - 0.1, d * (lane.width / 2.0 + extra), &px[np], &py[np]); ++np;
    if (first.kind != 0) {
      const double pi_2 = PGD_PI / 2.0;
      double ph = py_floordiv(lane.ph0, pi_2) * pi_2;
      ph += (lane.dir == 1) ? pi_2 : 0;
      for (int k = 0; k < 4; ++k) {
        double phi = ph + k * pi_2 * lane.dir;
        if (lane.dir * phi > lane.dir * lane.ph1) break;
        double r = lane.radius - d * (lane.width / 2.0 + extra) * lane.dir;
        dd s, c;
        dd_sincos(phi, &s, &c);
        px[np] = lane.cx + r * c.hi;
        py[np] = lane.cy + r * s.hi;
        ++np;
      }
    }
    for (int i = 0; i < np; ++i) {
      if (px[i] > xmax) xmax = px[i];
      if (px[i] < xmin) xmin = px[i];
      if (py[i] > ymax) ymax = py[i];
      if (py[i] < ymin) ymin = py[i];
    }
  }
  bb[0] = xmax; bb[1] = xmin; bb[2] = ymax; bb[3] = ymin;
}

/* scene_utils.py:40-71: does `lane` (sampled every metre at lateral offset positive * width / 2) enter a lane of the
 * map built so far (roads [0, world_end))?  `ign_from/ign_to` = road to skip (0,0 = none). */
PGD_HD inline bool lane_crosses_world(Gen& g, int world_end, const GLane& lane, double positive, int ign_from,
                                      int ign_to) {
  double bb2[4];
  contour_bbox(lane, lane, bb2);
  const int n_samples = (int)lane.length - 1;  // range(1, int(length), 1)
  const double lat_off = positive * lane.width / 2.0;
  bool have_samples = false;
  for (int r = 0; r < world_end; ++r) {
    GRoad& rd = g.s.roads[r];
    if (rd.removed || rd.n_lanes == 0) continue;
    if (ign_from != 0 && rd.from == ign_from && rd.to == ign_to) continue;
    if (rd.from == ND_DECO0) continue;
    if (!rd.bbox_valid) {
      contour_bbox(g.s.lanes[rd.lanes[0]], g.s.lanes[rd.lanes[rd.n_lanes - 1]], rd.bbox);
      rd.bbox_valid = 1;
    }
    const double* bb1 = rd.bbox;
    if (bb1[1] > bb2[0] || bb2[1] > bb1[0] || bb1[3] > bb2[2] || bb2[3] > bb1[2]) continue;
    if (!have_samples) {  // the box scratch is idle during the search: park the sample points there
      if (n_samples > g.caps.boxes) { gen_fail(g, GEN_ERR_BOXES); return true; }
      for (int i = 1; i <= n_samples; ++i) lane_position(lane, (double)i, lat_off, &g.s.boxes[i - 1].cx, &g.s.boxes[i - 1].cy);
      have_samples = true;
    }
    for (int i = 1; i <= n_samples; ++i) {
      const double px = g.s.boxes[i - 1].cx, py = g.s.boxes[i - 1].cy;
      for (int k = 0; k < rd.n_lanes; ++k) {
        const GLane& other = g.s.lanes[rd.lanes[k]];
        double half = other.width / 2.0;
        double lon, lat;
        lane_local_fast(other, px, py, &lon, &lat);
        const double guard = 1e-7;
        if (fabs(fabs(lat) - half) < guard || fabs(lon) < guard || fabs(lon - other.length) < guard)
          lane_local(other, px, py, &lon, &lat);
        if (fabs(lat) <= half && 0 <= lon && lon <= other.length) return true;
      }
    }
  }
  return false;
}

// ------------------------------------------------------------------------------------------------ road builders
PGD_HD inline bool crosses(Gen& g, const GBlock& b, const GLane& lane, double positive, int ign_from = 0,
                           int ign_to = 0) {
  if (b.nocheck) return true;  // scene_utils.py:49-50: with checking off the reference reports "crossing"
  return lane_crosses_world(g, b.road_begin, lane, positive, ign_from, ign_to);
}

/* create_block_utils.py:62-159.  `lane` becomes the outermost (toward_smaller) or innermost lane of the road.
 * Returns the pool index of that lane through *origin_idx (callers that keep using the lane read it back). */
PGD_HD inline bool road_from(Gen& g, GBlock& b, const GLane& lane_in, int lane_num, int from, int to,
                             bool toward_smaller = true, int ign_from = 0, int ign_to = 0, int center = LT_CONT,
                             int side = LT_SIDE, int inner = LT_BROKEN, int center_color = COL_YELLOW,
                             int* origin_idx = nullptr) {
  const int extra = lane_num - 1;
  GLane built[MAX_ROAD_LANES];  // creation order: origin's neighbour first
  GLane origin = lane_in;
  GLane cur = lane_in;
  const double w = lane_in.width;
  int nb = 0;
  for (int i = extra; i > 0; --i) {
    GLane s = cur;
    if (cur.kind == 0) {
      double off = toward_smaller ? -w : w;
      lane_position(cur, 0, off, &s.sx, &s.sy);
      lane_position(cur, cur.length, off, &s.ex, &s.ey);
    } else {
      bool cw = cur.dir == 1;
      if (!toward_smaller) s.radius = cw ? cur.radius - w : cur.radius + w;
      else s.radius = cw ? cur.radius + w : cur.radius - w;
      lane_refresh(s);
    }
    if (i == 1) {
      if (toward_smaller) { s.lt[0] = (int8_t)center; s.lt[1] = (int8_t)inner; }
      else { s.lt[0] = (int8_t)inner; s.lt[1] = (int8_t)side; }
    } else {
      s.lt[0] = s.lt[1] = (int8_t)inner;
    }
    if (nb < MAX_ROAD_LANES - 1) built[nb++] = s;
    else gen_fail(g, GEN_ERR_ROADS);
    cur = s;
  }
  const int total = nb + 1;
  if (toward_smaller) {
    origin.lt[0] = (int8_t)(total > 1 ? inner : center);
    origin.lt[1] = (int8_t)side;
  } else if (total > 1) {
    origin.lt[1] = built[nb - 1].lt[0];
  }
  double factor = (3.0 + 0.6 + w / 2.0) * 2.0 / w;
  bool ok = !crosses(g, b, origin, factor, ign_from, ign_to);
  // final order of the road's lanes
  int first_idx = -1, last_idx = -1, org = -1;
  if (toward_smaller) {
    for (int k = nb - 1; k >= 0; --k) {
      int li = add_lane(g, b, from, to, built[k]);
      if (first_idx < 0) first_idx = li;
      last_idx = li;
    }
    org = add_lane(g, b, from, to, origin);
    if (first_idx < 0) first_idx = org;
    last_idx = org;
  } else {
    org = add_lane(g, b, from, to, origin);
    first_idx = last_idx = org;
    for (int k = 0; k < nb; ++k) last_idx = add_lane(g, b, from, to, built[k]);
  }
  if (extra == 0) {
    g.s.lanes[last_idx].lt[0] = (int8_t)center;
    g.s.lanes[last_idx].lt[1] = (int8_t)side;
  }
  g.s.lanes[first_idx].col[0] = (int8_t)center_color;
  g.s.lanes[first_idx].col[1] = COL_GREY;
  if (origin_idx) *origin_idx = org;
  return ok;
}

/* create_block_utils.py:177-230 */
PGD_HD inline bool adverse_road(Gen& g, GBlock& b, int from, int to, int ign_from = 0, int ign_to = 0,
                                int center = LT_CONT, int side = LT_SIDE, int inner = LT_BROKEN,
                                int center_color = COL_YELLOW) {
  int r = block_road(g, b, from, to);
  const GRoad& rd = g.s.roads[r];
  const GLane ref = g.s.lanes[rd.lanes[rd.n_lanes - 1]];
  const int num = rd.n_lanes * 2;
  const double w = ref.width;
  GLane sym;
  if (ref.kind == 0) {
    double sx, sy, ex, ey;
    lane_position(ref, ref.length, -(num - 1) * w, &sx, &sy);
    lane_position(ref, 0, -(num - 1) * w, &ex, &ey);
    sym = lane_straight(sx, sy, ex, ey, w, ref.lt[0], ref.lt[1]);
  } else {
    bool cw = ref.dir != 1;
    double radius = !cw ? ref.radius + (num - 1) * w : ref.radius - (num - 1) * w;
    sym = lane_arc(ref.cx, ref.cy, radius, ref.ph1, ref.ph0, cw, w, ref.lt[0], ref.lt[1]);
  }
  bool ok = road_from(g, b, sym, num / 2, -to, -from, true, ign_from, ign_to, center, side, inner, center_color);
  int r0 = block_road(g, b, from, to);
  GLane& l0 = g.s.lanes[g.s.roads[r0].lanes[0]];
  l0.col[0] = (int8_t)center_color;
  l0.col[1] = COL_GREY;
  return ok;
}

// ------------------------------------------------------------------------------------------------ blocks
PGD_HD inline int new_node(GBlock& b) {
  b.road_no += 1;
  return node_of(b.idx, b.part, b.road_no - 1);
}
PGD_HD inline void set_part(GBlock& b, int part) {
  b.part = part;
  b.road_no = 0;
}
PGD_HD inline void add_socket(Gen& g, GBlock& b, int from, int to) {
  if (b.n_sockets >= MAX_BLOCK_SOCKETS) {
    gen_fail(g, GEN_ERR_LOOKUP);
    return;
  }
  GSocket& s = b.sockets[b.n_sockets];
  s.from = from; s.to = to;
  s.owner = b.idx; s.k = b.n_sockets;
  b.n_sockets++;
}
PGD_HD inline void respawn_add(Gen& g, GBlock& b, int from, int to) {
  if (b.n_respawn >= MAX_RESPAWN) {
    gen_fail(g, GEN_ERR_LOOKUP);
    return;
  }
  b.respawn[b.n_respawn][0] = from;
  b.respawn[b.n_respawn][1] = to;
  b.n_respawn++;
}
PGD_HD inline void respawn_remove(GBlock& b, int from, int to) {
  for (int i = 0; i < b.n_respawn; ++i) {
    if (b.respawn[i][0] == from && b.respawn[i][1] == to) {
      for (int j = i; j + 1 < b.n_respawn; ++j) {
        b.respawn[j][0] = b.respawn[j + 1][0];
        b.respawn[j][1] = b.respawn[j + 1][1];
      }
      b.n_respawn--;
      return;
    }
  }
}

/* Block.socket(): intersections / roundabouts stop spawning traffic on the arm the next block plugs into */
PGD_HD inline GSocket take_socket(GBlock& b, int pos) {
  GSocket s = b.sockets[pos];
  if (b.type == BK_X || b.type == BK_T || b.type == BK_O) respawn_remove(b, -s.to, -s.from);
  return s;
}

PGD_HD inline void block_clear(Gen& g, GBlock& b) {  // also undoes the merge into the world (stack discipline)
  g.n_roads = b.road_begin;
  g.n_lanes = b.lane_begin;
  b.road_end = b.road_begin;
  b.part = 0;
  b.road_no = 0;
  b.n_respawn = 0;
  b.n_sockets = 0;
}

PGD_HD inline const GLane& basic_lane(const Gen& g, const GBlock& b) { return g.s.lanes[b.basic]; }

PGD_HD inline bool build_straight(Gen& g, GBlock& b) {
  set_part(b, 0);
  GLane nw = extend_straight(basic_lane(g, b), b.length, LT_BROKEN, LT_SIDE);
  int from = b.pre.to, to = new_node(b);
  bool ok = road_from(g, b, nw, b.n_pos, from, to);
  ok = adverse_road(g, b, from, to) && ok;
  add_socket(g, b, from, to);
  return ok;
}

PGD_HD inline bool build_curve(Gen& g, GBlock& b) {
  int from = b.pre.to, to = new_node(b);
  GLane bend, straight;
  bend_then_straight(basic_lane(g, b), b.length, b.radius, PGD_DEG2RAD(b.angle), b.dir != 0, basic_lane(g, b).width,
                     LT_BROKEN, LT_SIDE, &bend, &straight);
  bool ok = road_from(g, b, bend, b.n_pos, from, to);
  ok = adverse_road(g, b, from, to) && ok;
  from = to;
  to = new_node(b);
  ok = road_from(g, b, straight, b.n_pos, from, to) && ok;
  ok = adverse_road(g, b, from, to) && ok;
  add_socket(g, b, from, to);
  return ok;
}

/* intersection.py:151-206 (one arm: left turn, straight through, right turn); attach = road entering the junction */
PGD_HD inline bool intersection_part(Gen& g, GBlock& b, int attach_road, bool attach_in_world, const int* nodes,
                                     int part, GLane* right_exit) {
  const double radius = b.radius;
  const int n = (part == 0 || part == 2) ? b.n_cross : b.n_pos;
  GRoad att = g.s.roads[attach_road];  // copy: the pool may grow
  (void)attach_in_world;
  const GLane left = g.s.lanes[att.lanes[0]];
  const double w = left.width;
  const int n_turn = b.n_pos < b.n_cross ? b.n_pos : b.n_cross;
  const double left_r = radius + n * w;
  GLane bend, tmp;
  // change_lane_num is forced to 0 (std_intersection.py:6-9), so the lane-count-changing branch never runs
  bend_then_straight(left, 30, left_r, PGD_DEG2RAD(90), false, w, LT_NONE, LT_NONE, &bend, &tmp);
  road_from(g, b, bend, n_turn, att.to, nodes[2], false, 0, 0, LT_NONE, LT_NONE, LT_NONE);
  // straight through
  GLane src[MAX_ROAD_LANES];
  for (int k = 0; k < att.n_lanes; ++k) src[k] = g.s.lanes[att.lanes[k]];
  const double through = 2 * radius + (2 * n - 1) * src[0].width;
  for (int k = 0; k < att.n_lanes; ++k)
    add_lane(g, b, att.to, nodes[1], extend_straight(src[k], through, LT_NONE, LT_NONE));
  // right turn
  const GLane right = src[att.n_lanes - 1];
  GLane rbend, rstraight;
  bend_then_straight(right, 30, radius, PGD_DEG2RAD(90), true, right.width, LT_NONE, LT_SIDE, &rbend, &rstraight);
  bool ok = !crosses(g, b, rbend, 1);
  road_from(g, b, rbend, n_turn, att.to, nodes[0], true, 0, 0, LT_NONE, LT_SIDE, LT_NONE);
  rstraight.lt[0] = LT_BROKEN;
  rstraight.lt[1] = LT_SIDE;
  *right_exit = rstraight;
  return ok;
}

/* intersection.py:45-96 */
PGD_HD inline bool build_intersection(Gen& g, GBlock& b) {
  b.change_lane_num = 0;
  int di = b.decrease_increase == 0 ? -1 : 1;
  if (b.n_pos <= 1) di = 1;
  else if (b.n_pos >= 4) di = -1;
  b.n_cross = b.n_pos + di * b.change_lane_num;
  bool ok = true;
  int attach = world_road(g, b, b.pre.from, b.pre.to);
  int nodes[4] = {node_of(b.idx, 0, 0), node_of(b.idx, 1, 0), node_of(b.idx, 2, 0), -b.pre.to};
  for (int i = 0; i < 4; ++i) {
    GLane right_lane;
    bool good = intersection_part(g, b, attach, i == 0, nodes, i, &right_lane);
    int n0 = nodes[0];
    nodes[0] = nodes[1]; nodes[1] = nodes[2]; nodes[2] = nodes[3]; nodes[3] = n0;
    ok = ok && good;
    if (i != 3) {
      int n = (i == 1) ? b.n_pos : b.n_cross;
      int ef = node_of(b.idx, i, 0), et = node_of(b.idx, i, 1);
      ok = road_from(g, b, right_lane, n, ef, et) && ok;
      ok = adverse_road(g, b, ef, et) && ok;
      respawn_add(g, b, -et, -ef);
      add_socket(g, b, ef, et);
      attach = block_road(g, b, -et, -ef);
    }
  }
  return ok;
}

/* RoadNetwork.remove_all_roads (road_network.py:120-133) on the block's own net, with the lazy breadth-first
 * enumeration of road_network.py:241-256: roads of a found path are removed before the search continues. */
PGD_HD inline void remove_all_roads(Gen& g, GBlock& b, int start, int goal) {
  int32_t* q = g.s.queue;  // entries (node, parent)
  int head = 0, tail = 0;
  q[0] = start; q[1] = -1;
  tail = 1;
  const int rb = b.road_begin;
  while (head < tail) {
    int cur = head++;
    int node = q[2 * cur];
    // children snapshot: to-nodes of live roads from `node`, in insertion order, not already on the path
    int kids[MAX_BLOCK_ROADS];
    int nk = 0;
    for (int r = rb; r < g.n_roads; ++r) {
      if (g.s.roads[r].removed || g.s.roads[r].from != node) continue;
      int to = g.s.roads[r].to;
      bool on_path = false;
      for (int p = cur; p >= 0; p = q[2 * p + 1])
        if (q[2 * p] == to) { on_path = true; break; }
      if (!on_path && nk < MAX_BLOCK_ROADS) kids[nk++] = to;
    }
    for (int k = 0; k < nk; ++k) {
      int nxt = kids[k];
      if (nxt == goal) {
        // remove every road of path + [goal]
        int b_node = goal;
        for (int p = cur; p >= 0; p = q[2 * p + 1]) {
          int a_node = q[2 * p];
          int r = find_road(g, a_node, b_node, rb, g.n_roads);
          if (r >= 0) g.s.roads[r].removed = 1;
          b_node = a_node;
        }
      } else {
        bool has_out = false;
        for (int r = rb; r < g.n_roads; ++r)
          if (!g.s.roads[r].removed && g.s.roads[r].from == nxt) { has_out = true; break; }
        if (has_out) {
          if (tail >= g.caps.queue) { gen_fail(g, GEN_ERR_QUEUE); return; }
          q[2 * tail] = nxt;
          q[2 * tail + 1] = cur;
          ++tail;
        }
      }
    }
  }
}

PGD_HD inline void road_of_socket(const GSocket& s, bool neg, int* from, int* to) {
  if (!neg) { *from = s.from; *to = s.to; }
  else { *from = -s.to; *to = -s.from; }
}

/* t_intersection.py:17-86: build the four-arm crossing, delete one arm, re-label the through road */
PGD_HD inline bool build_t_intersection(Gen& g, GBlock& b) {
  bool ok = build_intersection(g, b);
  const int t = b.t_type;
  // sockets 0..2 are this block's exits; "socket 3" is the socket we are plugged into
  GSocket all[4] = {b.sockets[0], b.sockets[1], b.sockets[2], b.pre};
  const GSocket gone = all[t];
  const int start_node = -gone.from, end_node = gone.from;  // gone.neg[1], gone.pos[0]
  for (int i = 0; i < 4; ++i) {
    if (i == t) continue;
    const GSocket& s = all[i];
    int exit_node = (i != 3) ? s.from : -s.to;
    remove_all_roads(g, b, start_node, exit_node);
    int entry_node = (i != 3) ? -s.from : s.to;
    remove_all_roads(g, b, entry_node, end_node);
  }
  {  // _change_vis (t_intersection.py:22-51)
    const GSocket& nxt = all[(t + 1) % 4];
    const GSocket& last = all[(t + 3) % 4];
    int np_f, np_t, nn_f, nn_t, lp_f, lp_t, ln_f, ln_t;
    road_of_socket(nxt, false, &np_f, &np_t);
    road_of_socket(nxt, true, &nn_f, &nn_t);
    road_of_socket(last, false, &lp_f, &lp_t);
    road_of_socket(last, true, &ln_f, &ln_t);
    if (t == 2) {
      road_of_socket(nxt, true, &np_f, &np_t);
      road_of_socket(nxt, false, &nn_f, &nn_t);
    }
    if (t == 0) {
      road_of_socket(last, true, &lp_f, &lp_t);
      road_of_socket(last, false, &ln_f, &ln_t);
    }
    const int rf[2] = {ln_t, nn_t}, rt[2] = {np_f, lp_f};
    for (int i = 0; i < 2; ++i) {
      int r = block_road(g, b, rf[i], rt[i]);
      GRoad& rd = g.s.roads[r];
      int outside = (i == 0) ? LT_SIDE : LT_NONE;
      for (int k = 0; k < rd.n_lanes; ++k) {
        GLane& lane = g.s.lanes[rd.lanes[k]];
        lane.lt[0] = LT_BROKEN;
        lane.lt[1] = (int8_t)((k != rd.n_lanes - 1) ? LT_BROKEN : outside);
        if (k == 0) {
          lane.col[0] = COL_YELLOW;
          lane.col[1] = COL_GREY;
          if (i == 1) lane.lt[0] = LT_NONE;
        }
      }
    }
  }
  // drop the arm's socket (the others keep their numbers) and its two roads
  int rr = find_road(g, gone.from, gone.to, b.road_begin, g.n_roads);
  if (rr >= 0) g.s.roads[rr].removed = 1;  // remove_all_roads(pos[0], pos[1]): the direct road
  rr = find_road(g, -gone.to, -gone.from, b.road_begin, g.n_roads);
  if (rr >= 0) g.s.roads[rr].removed = 1;
  respawn_remove(b, -gone.to, -gone.from);
  int n = 0;
  for (int i = 0; i < 3; ++i)
    if (i != t) b.sockets[n++] = all[i];
  b.n_sockets = n;
  return ok;
}

PGD_HD inline GLane tool_lane(const GLane& straight, double back) {
  double sx, sy, ex, ey;
  lane_position(straight, -back, 0, &sx, &sy);
  lane_position(straight, 0, 0, &ex, &ey);
  return lane_straight(sx, sy, ex, ey, 4, LT_BROKEN, LT_BROKEN);
}

/* roundabout.py:49-191: one quarter of the ring; road = the road entering this quarter */
PGD_HD inline bool roundabout_part(Gen& g, GBlock& b, int road_from_node, int road_to_node, int part, int* exit_from,
                                   int* exit_to) {
  bool ok = true;
  set_part(b, part);
  const int n = b.n_pos;
  const double w = b.lane_width;
  const double r_exit = b.exit_radius, r_inner = b.inner_radius, angle = b.angle;
  const double r_big = (n * 2 - 1) * w + r_inner;
  // entry arc
  int seg_f = road_to_node, seg_t = new_node(b);
  int r_in = (part == 0) ? world_road(g, b, road_from_node, road_to_node) : block_road(g, b, road_from_node, road_to_node);
  const GLane last_in = g.s.lanes[g.s.roads[r_in].lanes[g.s.roads[r_in].n_lanes - 1]];
  GLane bend, straight, to_next;
  bend_then_straight(last_in, 10, r_exit, PGD_DEG2RAD(angle), true, w, LT_BROKEN, LT_SIDE, &bend, &straight);
  int skip = node_of(b.idx, (part + 3) % 4, 0);
  ok = road_from(g, b, bend, n, seg_f, seg_t, true, skip, skip) && ok;
  {
    GRoad& rd = g.s.roads[block_road(g, b, seg_f, seg_t)];
    for (int k = 0; k < rd.n_lanes; ++k) {
      g.s.lanes[rd.lanes[k]].lt[0] = LT_NONE;
      g.s.lanes[rd.lanes[k]].lt[1] = (int8_t)((k == n - 1) ? LT_SIDE : LT_NONE);
    }
  }
  // ring arc
  bend_then_straight(tool_lane(straight, 5), 10, r_big, PGD_DEG2RAD(2 * angle - 90), false, w, LT_BROKEN, LT_SIDE,
                     &bend, &to_next);
  seg_f = seg_t;
  seg_t = new_node(b);
  ok = road_from(g, b, bend, n, seg_f, seg_t) && ok;
  if (b.n_ring < 4) b.ring[b.n_ring++] = block_road(g, b, seg_f, seg_t);
  // exit arc + exit straight
  bend_then_straight(tool_lane(to_next, 5), 30, r_exit, PGD_DEG2RAD(angle), true, w, LT_BROKEN, LT_SIDE, &bend,
                     &straight);
  seg_f = seg_t;
  seg_t = (part < 3) ? new_node(b) : -b.pre.to;
  ok = road_from(g, b, bend, n, seg_f, seg_t) && ok;
  {
    GRoad& rd = g.s.roads[block_road(g, b, seg_f, seg_t)];
    for (int k = 0; k < rd.n_lanes; ++k) {
      g.s.lanes[rd.lanes[k]].lt[0] = LT_NONE;
      g.s.lanes[rd.lanes[k]].lt[1] = (int8_t)((k == n - 1) ? LT_SIDE : LT_NONE);
    }
  }
  *exit_from = seg_t;
  *exit_to = new_node(b);
  if (part < 3) {
    ok = road_from(g, b, straight, n, *exit_from, *exit_to) && ok;
    add_socket(g, b, *exit_from, *exit_to);
  }
  // inner connector to the next quarter
  seg_f = node_of(b.idx, part, 1);
  seg_t = node_of(b.idx, (part + 1) % 4, 0);
  double beneath = (n * 2 - 1) * w / 2 + r_exit;
  double r_seg = beneath / cr_cos(PGD_DEG2RAD(angle)) - r_exit;
  GLane tmp;
  bend_then_straight(tool_lane(to_next, 6), 5, r_seg, PGD_DEG2RAD(180 - 2 * angle), false, w, LT_BROKEN, LT_SIDE,
                     &bend, &tmp);
  road_from(g, b, bend, n, seg_f, seg_t);
  {
    GRoad& rd = g.s.roads[block_road(g, b, seg_f, seg_t)];
    for (int k = 0; k < rd.n_lanes; ++k) {
      GLane& ln = g.s.lanes[rd.lanes[k]];
      if (k == 0) {
        ln.lt[0] = LT_CONT;
        ln.lt[1] = (int8_t)(n > 1 ? LT_BROKEN : LT_NONE);
      } else {
        ln.lt[0] = ln.lt[1] = LT_BROKEN;
      }
    }
  }
  return ok;
}

PGD_HD inline bool build_roundabout(Gen& g, GBlock& b) {
  b.n_ring = 0;
  bool ok = true;
  int af = b.pre.from, at = b.pre.to;
  for (int i = 0; i < 4; ++i) {
    int ef, et;
    bool good = roundabout_part(g, b, af, at, i, &ef, &et);
    ok = ok && good;
    if (i < 3) {
      ok = adverse_road(g, b, ef, et) && ok;
      af = -et;
      at = -ef;
    }
  }
  for (int i = 0; i < b.n_sockets; ++i) respawn_add(g, b, -b.sockets[i].to, -b.sockets[i].from);
  return ok;
}

#define RAMP_RADIUS 40
#define RAMP_ANGLE 10
#define RAMP_CONNECT 20
#define RAMP_LEN 15

/* ramp.py:43-204 */
PGD_HD inline bool build_in_ramp(Gen& g, GBlock& b) {
  const double acc_len = b.length;
  const int n = b.n_pos;
  const double w = b.lane_width;
  const double extra_part = 10, socket_len = 20;
  bool ok = true;
  set_part(b, 0);
  const double sin_a = cr_sin(PGD_DEG2RAD(RAMP_ANGLE)), cos_a = cr_cos(PGD_DEG2RAD(RAMP_ANGLE));
  const double lon_len = sin_a * RAMP_RADIUS * 2 + cos_a * RAMP_CONNECT + RAMP_LEN;
  GLane extend = extend_straight(basic_lane(g, b), lon_len + extra_part, LT_BROKEN, LT_CONT);
  int ext_f = b.pre.to, ext_t = new_node(b);
  int org;
  ok = road_from(g, b, extend, n, ext_f, ext_t, true, 0, 0, LT_CONT, LT_CONT, LT_BROKEN, COL_YELLOW, &org) && ok;
  g.s.lanes[org].lt[0] = (int8_t)(n != 1 ? LT_BROKEN : LT_CONT);
  g.s.lanes[org].lt[1] = LT_CONT;
  ok = adverse_road(g, b, ext_f, ext_t) && ok;
  extend = g.s.lanes[org];  // the reference keeps using the live object
  {
    GRoad& rd = g.s.roads[block_road(g, b, -ext_t, -ext_f)];
    GLane& ln = g.s.lanes[rd.lanes[rd.n_lanes - 1]];
    ln.lt[0] = (int8_t)(n == 1 ? LT_NONE : LT_BROKEN);
    ln.lt[1] = LT_SIDE;
  }
  // acceleration part
  GLane acc_side = extend_straight(extend, acc_len + w, extend.lt[0], LT_SIDE);
  int acc_f = ext_t, acc_t = new_node(b);
  ok = road_from(g, b, acc_side, n, acc_f, acc_t, true, 0, 0, LT_CONT, LT_CONT, LT_BROKEN, COL_YELLOW, &org) && ok;
  ok = adverse_road(g, b, acc_f, acc_t) && ok;
  {
    GRoad& rd = g.s.roads[block_road(g, b, acc_f, acc_t)];
    GLane& ln = g.s.lanes[rd.lanes[rd.n_lanes - 1]];
    ln.lt[0] = (int8_t)(n == 1 ? LT_CONT : LT_BROKEN);
    ln.lt[1] = LT_BROKEN;
  }
  acc_side = g.s.lanes[org];
  // socket part
  GLane sock_side = extend_straight(acc_side, socket_len, acc_side.lt[0], acc_side.lt[1]);
  int sock_f = acc_t, sock_t = new_node(b);
  ok = road_from(g, b, sock_side, n, sock_f, sock_t, true, 0, 0, LT_CONT, LT_CONT) && ok;
  ok = adverse_road(g, b, sock_f, sock_t) && ok;
  add_socket(g, b, sock_f, sock_t);
  // the ramp itself
  set_part(b, 1);
  const double lat = (1 - cos_a) * RAMP_RADIUS * 2 + sin_a * RAMP_CONNECT;
  double ex, ey, sx, sy;
  lane_position(extend, extra_part + RAMP_LEN, lat + w, &ex, &ey);
  lane_position(extend, extra_part, lat + w, &sx, &sy);
  GLane entry = lane_straight(sx, sy, ex, ey, w, LT_CONT, LT_CONT);
  int en_f = new_node(b), en_t = new_node(b);
  add_lane(g, b, en_f, en_t, entry);
  ok = (!crosses(g, b, entry, 0.95)) && ok;
  respawn_add(g, b, en_f, en_t);
  GLane bend1, connect;
  bend_then_straight(entry, RAMP_CONNECT, RAMP_RADIUS, PGD_DEG2RAD(RAMP_ANGLE), false, w, LT_CONT, LT_CONT, &bend1,
                     &connect);
  int b1_f = en_t, b1_t = new_node(b);
  int co_f = b1_t, co_t = new_node(b);
  add_lane(g, b, b1_f, b1_t, bend1);
  add_lane(g, b, co_f, co_t, connect);
  ok = (!crosses(g, b, bend1, 0.95)) && ok;
  ok = (!crosses(g, b, connect, 0.95)) && ok;
  GLane bend2, acc_lane;
  bend_then_straight(connect, acc_len, RAMP_RADIUS, PGD_DEG2RAD(RAMP_ANGLE), true, w, LT_CONT, LT_CONT, &bend2,
                     &acc_lane);
  acc_lane.lt[0] = LT_BROKEN;
  acc_lane.lt[1] = LT_CONT;
  add_lane(g, b, co_t, node_of(b.idx, 0, 0), bend2);
  add_lane(g, b, acc_f, acc_t, acc_lane);
  ok = (!crosses(g, b, bend2, 0.95)) && ok;
  ok = (!crosses(g, b, acc_lane, 0.95)) && ok;
  GLane merge, tmp;
  bend_then_straight(acc_lane, 10, w / 2, PGD_PI / 2, false, w, LT_BROKEN, LT_CONT, &merge, &tmp);
  add_lane(g, b, ND_DECO0, ND_DECO1, merge);
  return ok;
}

/* ramp.py:235-365 */
PGD_HD inline bool build_out_ramp(Gen& g, GBlock& b) {
  const int n = b.n_pos;
  const double w = b.lane_width;
  bool ok = true;
  const double sin_a = cr_sin(PGD_DEG2RAD(RAMP_ANGLE)), cos_a = cr_cos(PGD_DEG2RAD(RAMP_ANGLE));
  const double lon_len = sin_a * RAMP_RADIUS * 2 + cos_a * RAMP_CONNECT + RAMP_LEN + 15;
  set_part(b, 0);
  const double dec_len = b.length;
  const GLane basic = basic_lane(g, b);
  GLane dec_lane = extend_straight(basic, dec_len + w, basic.lt[0], LT_SIDE);
  int dec_f = b.pre.to, dec_t = new_node(b);
  ok = road_from(g, b, dec_lane, n, dec_f, dec_t, true, 0, 0, LT_CONT, LT_CONT) && ok;
  ok = adverse_road(g, b, dec_f, dec_t) && ok;
  GLane dec_right;
  {
    GRoad& rd = g.s.roads[block_road(g, b, dec_f, dec_t)];
    GLane& ln = g.s.lanes[rd.lanes[rd.n_lanes - 1]];
    ln.lt[0] = (int8_t)(n == 1 ? LT_CONT : LT_BROKEN);
    ln.lt[1] = LT_NONE;
    dec_right = ln;
  }
  GLane extend = extend_straight(dec_right, lon_len, dec_right.lt[0], LT_CONT);
  int ext_f = dec_t, ext_t = new_node(b);
  ok = road_from(g, b, extend, n, ext_f, ext_t, true, 0, 0, LT_CONT, LT_CONT) && ok;
  ok = adverse_road(g, b, ext_f, ext_t) && ok;
  {
    GRoad& rd = g.s.roads[block_road(g, b, -ext_t, -ext_f)];
    GLane& ln = g.s.lanes[rd.lanes[rd.n_lanes - 1]];
    ln.lt[0] = (int8_t)(n == 1 ? LT_NONE : LT_BROKEN);
    ln.lt[1] = LT_SIDE;
  }
  add_socket(g, b, ext_f, ext_t);
  // deceleration lane + ramp
  set_part(b, 1);
  double sx, sy, ex, ey;
  lane_position(dec_right, w, w, &sx, &sy);
  lane_position(dec_right, dec_right.length, w, &ex, &ey);
  GLane side_lane = lane_straight(sx, sy, ex, ey, w, LT_BROKEN, LT_CONT);
  add_lane(g, b, dec_f, dec_t, side_lane);
  ok = (!crosses(g, b, side_lane, 0.95)) && ok;
  GLane bend1, connect;
  bend_then_straight(side_lane, RAMP_CONNECT, RAMP_RADIUS, PGD_DEG2RAD(RAMP_ANGLE), true, w, LT_CONT, LT_CONT, &bend1,
                     &connect);
  int b1_f = dec_t, b1_t = new_node(b);
  int co_f = b1_t, co_t = new_node(b);
  add_lane(g, b, b1_f, b1_t, bend1);
  add_lane(g, b, co_f, co_t, connect);
  ok = (!crosses(g, b, bend1, 0.95)) && ok;
  ok = (!crosses(g, b, connect, 0.95)) && ok;
  GLane bend2, exit_lane;
  bend_then_straight(connect, RAMP_LEN, RAMP_RADIUS, PGD_DEG2RAD(RAMP_ANGLE), false, w, LT_CONT, LT_CONT, &bend2,
                     &exit_lane);
  int b2_f = co_t, b2_t = new_node(b);
  int ex_f = b2_t, ex_t = new_node(b);
  add_lane(g, b, b2_f, b2_t, bend2);
  add_lane(g, b, ex_f, ex_t, exit_lane);
  ok = (!crosses(g, b, bend2, 0.95)) && ok;
  ok = (!crosses(g, b, exit_lane, 0.95)) && ok;
  GLane tool = lane_straight(side_lane.ex, side_lane.ey, side_lane.sx, side_lane.sy, side_lane.width, LT_BROKEN,
                             LT_BROKEN);
  GLane deco, tmp;
  bend_then_straight(tool, 10, w / 2, PGD_PI / 2, true, w, LT_CONT, LT_BROKEN, &deco, &tmp);
  add_lane(g, b, ND_DECO0, ND_DECO1, deco);
  return ok;
}

/* FirstPGBlock (first_block.py:25-89) */
PGD_HD inline void build_first(Gen& g, bool nocheck) {
  GBlock& b = g.s.blocks[0];
  b.type = BK_I;
  b.idx = 0;
  b.pre.from = ND_DECO0; b.pre.to = ND_DECO1; b.pre.owner = -1; b.pre.k = 0;
  b.road_begin = b.road_end = 0;
  b.lane_begin = 0;
  b.n_sockets = b.n_respawn = b.n_ring = 0;
  b.trials = b.part = b.road_no = 0;
  b.nocheck = nocheck;
  g.n_lanes = g.n_roads = 0;
  const double lw = g.cfg.lane_width;
  const int ln = g.cfg.lane_num;
  GLane basic = lane_straight(0, lw * (ln - 1), 10, lw * (ln - 1), lw, LT_BROKEN, LT_SIDE);
  road_from(g, b, basic, ln, ND_START, ND_START2);
  adverse_road(g, b, ND_START, ND_START2);
  GLane nxt = extend_straight(basic, g.cfg.exit_length - 10, LT_BROKEN, LT_SIDE);
  road_from(g, b, nxt, ln, ND_START2, ND_START3);
  adverse_road(g, b, ND_START2, ND_START3);
  add_socket(g, b, ND_START2, ND_START3);
  respawn_add(g, b, ND_START2, ND_START3);
  b.road_end = g.n_roads;
  g.n_blocks = 1;
}

/* parameters of a block from the uniform sample u of its parameter stream (utils/space.py:263-306) */
PGD_HD inline void sample_params(GBlock& b, double u) {
  switch (b.type) {
    case BK_S: b.length = box_f32(40.0, 80.0, u); break;
    case BK_C:
      b.angle = box_f32(45, 135, u);
      b.dir = box_int(0, 1, u);
      b.length = box_f32(40.0, 80.0, u);
      b.radius = box_f32(25.0, 60.0, u);
      break;
    case BK_X:
    case BK_T:
      b.radius = box_f32(10, 10, u);
      b.change_lane_num = box_int(0, 1, u);
      b.decrease_increase = box_int(0, 1, u);
      if (b.type == BK_T) b.t_type = box_int(0, 2, u);
      break;
    case BK_O:
      b.exit_radius = box_f32(5, 15, u);
      b.inner_radius = box_f32(15, 45, u);
      b.angle = box_f32(60, 60, u);
      break;
    case BK_r:
    case BK_R: b.length = box_f32(20, 40, u); break;
    default: break;
  }
}

PGD_HD inline bool run_builder(Gen& g, GBlock& b) {
  switch (b.type) {
    case BK_S: return build_straight(g, b);
    case BK_C: return build_curve(g, b);
    case BK_X: return build_intersection(g, b);
    case BK_T: return build_t_intersection(g, b);
    case BK_O: return build_roundabout(g, b);
    case BK_r: return build_in_ramp(g, b);
    case BK_R: return build_out_ramp(g, b);
    default: gen_fail(g, GEN_ERR_BACKTRACK); return true;
  }
}

/* Block.__init__ minus the RNG: attach to `sock` of the map so far */
PGD_HD inline void block_attach(Gen& g, GBlock& b, int type, int idx, const GSocket& sock, bool nocheck) {
  b.type = type;
  b.idx = idx;
  b.pre = sock;
  b.road_begin = b.road_end = g.n_roads;
  b.lane_begin = g.n_lanes;
  b.n_sockets = b.n_respawn = b.n_ring = 0;
  b.trials = b.part = b.road_no = 0;
  b.nocheck = nocheck;
  b.length = b.radius = b.angle = b.exit_radius = b.inner_radius = 0;
  b.dir = b.change_lane_num = b.decrease_increase = b.t_type = 0;
  int r = find_road(g, sock.from, sock.to, 0, g.n_roads);
  if (r < 0) { gen_fail(g, GEN_ERR_LOOKUP); r = 0; }
  const GRoad& rd = g.s.roads[r];
  b.n_pos = rd.n_lanes;
  b.basic = rd.lanes[rd.n_lanes - 1];
  b.lane_width = g.s.lanes[b.basic].width;
  b.n_cross = b.n_pos;
}

/* construct_block (base_block.py:72-96) with the parameters already in b */
PGD_HD inline bool block_build(Gen& g, GBlock& b) {
  block_clear(g, b);
  b.trials += 1;
  bool ok = run_builder(g, b);
  b.road_end = g.n_roads;
  return ok;
}

/* BIG.py:67-151: returns the number of blocks (including the first) */
PGD_HD inline void search_blocks(Gen& g, uint64_t seed) {
  MT* rs = &g.s.mt[0];
  MT* tmp = &g.s.mt[1];
  mt_seeded(rs, seed);
  build_first(g, false);
  const double block_prob[13] = {0.3, 0.1, 0.1, 0.1, 0.15, 0.15, 0.1, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  const int target = (g.cfg.n_fixed > 0 ? g.cfg.n_fixed : g.cfg.block_num) + 1;
  if (target > g.caps.blocks) { gen_fail(g, GEN_ERR_BLOCKS); return; }
  enum { FORWARD, DESTRUCT, SIBLING, BACK };
  int step = FORWARD;
  int guard = 0;
  while (!(g.n_blocks >= target && step == FORWARD)) {
    if (++guard > 100000 || g.status != GEN_OK) { gen_fail(g, GEN_ERR_BACKTRACK); return; }
    if (step == FORWARD) {
      int type;
      if (g.cfg.n_fixed > 0) type = g.cfg.fixed_types[g.n_blocks - 1];
      else type = mt_choice_p(rs, block_prob, 13);
      GBlock& prev = g.s.blocks[g.n_blocks - 1];
      int pick = (int)mt_randint(rs, (uint32_t)prev.n_sockets);
      GSocket sock = take_socket(prev, pick);
      uint32_t bseed = mt_randint(rs, 10000);
      GBlock& blk = g.s.blocks[g.n_blocks];
      block_attach(g, blk, type, g.n_blocks, sock, false);
      // the block's own stream: __init__ samples once (unused), every build() samples again
      mt_seeded(tmp, bseed);
      mt_randint(tmp, 1000000);
      blk.q[0] = mt_randint(tmp, 1000000);
      blk.q[1] = mt_randint(tmp, 1000000);
      g.n_blocks++;
      sample_params(blk, first_sample_of(tmp, blk.q[0]));
      step = block_build(g, blk) ? FORWARD : DESTRUCT;
    } else if (step == DESTRUCT) {
      GBlock& blk = g.s.blocks[g.n_blocks - 1];
      block_clear(g, blk);
      step = blk.trials < 2 ? SIBLING : BACK;
    } else if (step == SIBLING) {
      GBlock& blk = g.s.blocks[g.n_blocks - 1];
      if (blk.trials < 2) {
        if (blk.type == BK_I) { gen_fail(g, GEN_ERR_BACKTRACK); return; }
        sample_params(blk, first_sample_of(tmp, blk.q[blk.trials]));
        step = block_build(g, blk) ? FORWARD : DESTRUCT;
      } else {
        step = BACK;
      }
    } else {
      g.n_blocks--;
      if (g.n_blocks < 1) { gen_fail(g, GEN_ERR_BACKTRACK); return; }
      block_clear(g, g.s.blocks[g.n_blocks - 1]);
      step = SIBLING;
    }
  }
}

/* pg_map.py:48-71: rebuild every block from its stored parameters with overlap checking off */
PGD_HD inline void rebuild_from_sequence(Gen& g) {
  const int nb = g.n_blocks;
  // the records survive in g.s.blocks[i] (type, socket, parameters); only the pools are rebuilt
  build_first(g, true);
  for (int i = 1; i < nb; ++i) {
    GBlock& blk = g.s.blocks[i];
    GBlock rec = blk;
    GBlock& prev = g.s.blocks[i - 1];
    int pick = -1;
    for (int k = 0; k < prev.n_sockets; ++k)
      if (prev.sockets[k].owner == rec.pre.owner && prev.sockets[k].k == rec.pre.k) pick = k;
    if (pick < 0) { gen_fail(g, GEN_ERR_LOOKUP); return; }
    GSocket sock = take_socket(prev, pick);
    block_attach(g, blk, rec.type, i, sock, true);
    blk.length = rec.length; blk.radius = rec.radius; blk.angle = rec.angle;
    blk.exit_radius = rec.exit_radius; blk.inner_radius = rec.inner_radius;
    blk.dir = rec.dir; blk.change_lane_num = rec.change_lane_num;
    blk.decrease_increase = rec.decrease_increase; blk.t_type = rec.t_type;
    block_build(g, blk);
    g.n_blocks = i + 1;
  }
}

// ------------------------------------------------------------------------------------------------ tables
struct MapIndex {  // numbering used by the tables (tables.py MapIndex)
  int32_t n_roads, n_lanes, n_nodes;
};

PGD_HD inline void emit_box(Gen& g, int& nb, double p0x, double p0y, double p1x, double p1y, double mx, double my,
                            double hl, double hw, int kind, int lane) {
  if (nb >= g.caps.boxes) { gen_fail(g, GEN_ERR_BOXES); return; }
  double dx = p1x - p0x, dy = p1y - p0y;
  double d = norm2(dx, dy);
  GBox& bx = g.s.boxes[nb++];
  bx.cx = mx; bx.cy = my;
  bx.ux = dx / d; bx.uy = dy / d;
  bx.hl = hl; bx.hw = hw;
  bx.kind = kind; bx.lane = lane;
}

/* lane-surface boxes used for localisation (base_block.py:396-463) */
PGD_HD inline void surface_boxes(Gen& g, int& nb, const GLane& lane, int lane_id) {
  const double width = lane.width + 0.6 * 2;
  if (lane.kind == 0) {
    double mx, my, ex, ey;
    lane_position(lane, lane.length / 2, 0, &mx, &my);
    lane_position(lane, lane.length, 0, &ex, &ey);
    emit_box(g, nb, mx, my, ex, ey, mx, my, (lane.length + 0.1) / 2, width / 2, PGD_BOX_LANE, lane_id);
  } else {
    int n = (int)(lane.length / 4.0);
    for (int i = 0; i < n; ++i) {
      double mx, my, ex, ey;
      lane_position(lane, lane.length * (i + .5) / n, 0, &mx, &my);
      lane_position(lane, lane.length * (i + 1) / n, 0, &ex, &ey);
      emit_box(g, nb, mx, my, ex, ey, mx, my, (lane.length * 1.3 / n + 0.1) / 2, width / 2, PGD_BOX_LANE, lane_id);
    }
  }
}

/* lane-line ghosts and sidewalks (base_block.py:181-394) */
PGD_HD inline void line_boxes(Gen& g, int& nb, const GLane& lane, int lane_in_road, int lane_id) {
  const double w = lane.width;
  const bool straight = lane.kind == 0;
  const double LINE_HALF = 0.15 / 2, SEG = 4.0, STRIPE = 1.5, SW_SEG = 3.0, SW_W = 3.0, SW_GAP = 0.6;
  for (int k = 0; k < 2; ++k) {
    const int side = k == 0 ? -1 : 1;
    const int lt = lane.lt[k];
    if (lt == LT_NONE || (lane_in_road != 0 && k == 0)) {
      if (straight || lane.radius != w / 2) continue;
    }
    const double lat = side * w / 2;
    const int colour = lane.col[k];
    if (lt == LT_CONT || lt == LT_SIDE) {
      const int kind = colour == COL_YELLOW ? PGD_BOX_YELLOW : PGD_BOX_WHITE;
      if (straight) {
        double ax, ay, bx, by, mx, my;
        lane_position(lane, 0, lat, &ax, &ay);
        lane_position(lane, lane.length, lat, &bx, &by);
        lane_position(lane, lane.length / 2, lat, &mx, &my);
        emit_box(g, nb, ax, ay, bx, by, mx, my, norm2(bx - ax, by - ay) / 2, LINE_HALF, kind, lane_id);
      } else {
        int n = (int)(lane.length / SEG);
        for (int s = 0; s <= n; ++s) {
          double s0 = s * SEG, s1 = (s < n) ? (s + 1) * SEG : lane.length;
          double ax, ay, bx, by;
          lane_position(lane, s0, lat, &ax, &ay);
          lane_position(lane, s1, lat, &bx, &by);
          double ln = norm2(bx - ax, by - ay);
          if (ln <= 0) continue;
          emit_box(g, nb, ax, ay, bx, by, (ax + bx) / 2, (ay + by) / 2, ln / 2, LINE_HALF, kind, lane_id);
        }
      }
      if (lt == LT_SIDE) {
        const double radius = straight ? 0.0 : lane.radius;
        int n = (int)(lane.length / SW_SEG);
        for (int j = 0; j <= n; ++j) {
          double s0 = j * SW_SEG, s1 = (j < n) ? (j + 1) * SW_SEG : lane.length;
          double ax, ay, bx, by;
          lane_position(lane, s0, lat, &ax, &ay);
          lane_position(lane, s1, lat, &bx, &by);
          double ln = norm2(bx - ax, by - ay);
          if (j == n && !(ln > 1e-1)) continue;
          double factor;
          if (radius == 0) factor = 1.0;
          else if (lane.dir == 1) factor = 1 - SW_GAP / radius;
          else factor = (1 + SW_W / radius) * (1 + SW_GAP / radius);
          double mx = (ax + bx) / 2, my = (ay + by) / 2;
          double vx = -(by - ay) / ln, vy = (bx - ax) / ln;
          double off = SW_W / 2 + SW_GAP;
          emit_box(g, nb, ax, ay, bx, by, mx + vx * off, my + vy * off, ln * factor / 2, SW_W / 2, PGD_BOX_SIDEWALK,
                   lane_id);
        }
      }
    } else if (lt == LT_BROKEN) {
      if (straight) {
        double ax, ay, bx, by, mx, my;
        lane_position(lane, 0, lat, &ax, &ay);
        lane_position(lane, lane.length, lat, &bx, &by);
        lane_position(lane, lane.length / 2, lat, &mx, &my);
        emit_box(g, nb, ax, ay, bx, by, mx, my, norm2(bx - ax, by - ay) / 2, LINE_HALF, PGD_BOX_BROKEN, lane_id);
      } else {
        int n = (int)(lane.length / (2 * STRIPE));
        for (int s = 0; s < n; ++s) {
          double ax, ay, bx, by, mx, my;
          lane_position(lane, s * STRIPE * 2, lat, &ax, &ay);
          lane_position(lane, s * STRIPE * 2 + STRIPE, lat, &bx, &by);
          double ln = norm2(bx - ax, by - ay);
          if (ln <= 0) continue;
          lane_position(lane, s * STRIPE * 2 + STRIPE / 2, lat, &mx, &my);
          emit_box(g, nb, ax, ay, bx, by, mx, my, ln, LINE_HALF, PGD_BOX_BROKEN, lane_id);
        }
        double ax, ay, bx, by;
        lane_position(lane, n * STRIPE * 2, lat, &ax, &ay);
        lane_position(lane, lane.length + STRIPE, lat, &bx, &by);
        double ln = norm2(bx - ax, by - ay);
        if (ln > 0) emit_box(g, nb, ax, ay, bx, by, (ax + bx) / 2, (ay + by) / 2, ln, LINE_HALF, PGD_BOX_BROKEN, lane_id);
      }
    }
  }
}

/* Breadth-first first simple path start -> goal over the whole map (road_network.py:241-269); children in the
 * map's insertion order.  Writes node codes to path[], returns the length (0 = none). */
PGD_HD inline int shortest_path(Gen& g, const int* order, int n_order, int start, int goal, int* path, int max_path) {
  int32_t* q = g.s.queue;
  int head = 0, tail = 1;
  q[0] = start; q[1] = -1;
  while (head < tail) {
    int cur = head++;
    int node = q[2 * cur];
    for (int oi = 0; oi < n_order; ++oi) {
      const GRoad& rd = g.s.roads[order[oi]];
      if (rd.from != node) continue;
      int nxt = rd.to;
      bool on_path = false;
      for (int p = cur; p >= 0; p = q[2 * p + 1])
        if (q[2 * p] == nxt) { on_path = true; break; }
      if (on_path) continue;
      if (nxt == goal) {
        int len = 1;
        for (int p = cur; p >= 0; p = q[2 * p + 1]) ++len;
        if (len > max_path) { gen_fail(g, GEN_ERR_ROUTE); return 0; }
        path[len - 1] = goal;
        int k = len - 2;
        for (int p = cur; p >= 0; p = q[2 * p + 1]) path[k--] = q[2 * p];
        return len;
      }
      bool has_out = false;
      for (int oj = 0; oj < n_order; ++oj)
        if (g.s.roads[order[oj]].from == nxt) { has_out = true; break; }
      if (has_out) {
        if (tail >= g.caps.queue) { gen_fail(g, GEN_ERR_QUEUE); return 0; }
        q[2 * tail] = nxt;
        q[2 * tail + 1] = cur;
        ++tail;
      }
    }
  }
  return 0;
}

struct VehicleBody {
  double length, width, height, mass, lf, lr, tyre, track;
};
/* component/vehicle/vehicle_type.py:7-78; index = TYPE_ID of tables.py: s, m, l, xl, default */
PGD_HD inline VehicleBody vehicle_body(int type) {
  switch (type) {
    case 0: return VehicleBody{4.25, 1.7, 1.7, 800.0, 1.4126, 1.07, 0.376, 0.7};
    case 1: return VehicleBody{4.4, 1.85, 1.37, 1200.0, 1.285, 1.203, 0.39, 0.803};
    case 2: return VehicleBody{4.5, 1.86, 1.85, 1300.0, 1.391, 1.10751, 0.39, 0.75};
    case 3: return VehicleBody{5.8, 2.3, 2.8, 1600.0, 1.726, 1.075, 0.37, 0.831};
    default: return VehicleBody{4.51, 1.852, 1.19, 1100.0, 1.05234, 1.4166, 0.313, 0.815};
  }
}
/* utils/space.py:219-255 in the literal (low, high) order of the reference */
PGD_HD inline void vehicle_params(int type, double u, double* engine, double* brake, double* steer_deg,
                                  double* friction) {
  switch (type) {
    case 0: *friction = box_f32(0.9, 0.9, u); *engine = box_f32(550, 350, u); *brake = box_f32(80, 35, u); *steer_deg = box_f32(50, 50, u); break;
    case 1: *friction = box_f32(0.75, 0.75, u); *engine = box_f32(850, 650, u); *brake = box_f32(150, 60, u); *steer_deg = box_f32(45, 45, u); break;
    case 2: *friction = box_f32(0.8, 0.8, u); *engine = box_f32(650, 450, u); *brake = box_f32(120, 60, u); *steer_deg = box_f32(40, 40, u); break;
    case 3: *friction = box_f32(0.7, 0.7, u); *engine = box_f32(700, 500, u); *brake = box_f32(100, 50, u); *steer_deg = box_f32(35, 35, u); break;
    default: *friction = box_f32(0.9, 0.9, u); *engine = box_f32(850, 750, u); *brake = box_f32(180, 80, u); *steer_deg = box_f32(40, 40, u); break;
  }
}
PGD_HD inline int drop_substeps(int type) {
  VehicleBody vb = vehicle_body(type);
  double axis = type == 3 ? 0.3 : 0.2;
  double fall = vb.height / 2 + 1 - (vb.tyre + axis);
  return (int)ceil(sqrt(2 * fall / 9.81) / 0.02);
}

/* The whole reset path of one seed: search, rebuild, tables, episode template. */
PGD_HD inline int generate_one(uint64_t seed, const GenConfig& cfg, const GenCaps& caps, const GenScratch& scratch,
                               GenOut& out) {
  Gen g;
  g.cfg = cfg;
  if (cfg.random_lane_width || cfg.random_lane_num) {
    // MapManager.add_random_to_map (manager/map_manager.py:157-169) on the stream MapManager.seed(seed) sets up
    MT* rs = &scratch.mt[0];
    mt_seeded(rs, seed);
    if (cfg.random_lane_width) g.cfg.lane_width = mt_double(rs) * (4.5 - 3.0) + 3.0;
    if (cfg.random_lane_num) g.cfg.lane_num = 2 + (int)mt_randint(rs, 1);  // randint(2, 3): always 2, no draw
  }
  g.caps = caps;
  g.s = scratch;
  g.n_lanes = g.n_roads = g.n_blocks = 0;
  g.status = GEN_OK;
  for (int i = 0; i < 8; ++i) out.counts[i] = 0;
  if (g.cfg.lane_num < 1 || g.cfg.lane_num > MAX_ROAD_LANES - 2 || caps.blocks < 2) {
    return GEN_ERR_CONFIG;
  }
  search_blocks(g, seed);
  if (g.status != GEN_OK) return g.status;
  if (out.sequence) {
    for (int i = 0; i < g.n_blocks; ++i) {
      const GBlock& b = g.s.blocks[i];
      int32_t* rec = out.sequence + 16 * i;
      rec[0] = b.type;
      rec[1] = i == 0 ? -1 : b.pre.owner;
      rec[2] = i == 0 ? -1 : b.pre.k;
      rec[3] = b.dir; rec[4] = b.change_lane_num; rec[5] = b.decrease_increase; rec[6] = b.t_type;
      float* f = (float*)(rec + 8);
      f[0] = (float)b.length; f[1] = (float)b.radius; f[2] = (float)b.angle;
      f[3] = (float)b.exit_radius; f[4] = (float)b.inner_radius;
    }
  }
  rebuild_from_sequence(g);
  if (g.status != GEN_OK) return g.status;

  // ---- numbering: roads in the map's iteration order, decoration last
  // three work arrays live behind the candidate list (the caller reserves 4 * caps.roads ints there): the emission
  // order of the roads, the first flat lane id of every pool road, and the node codes in order of first appearance
  int* order = g.s.cand + 3 * g.caps.cand;
  int n_order = 0;
  for (int bi = 0; bi < g.n_blocks; ++bi) {
    const GBlock& b = g.s.blocks[bi];
    int tmp[MAX_BLOCK_ROADS + 1];
    int n = block_road_order(g, b.road_begin, b.road_end, tmp);
    for (int i = 0; i < n; ++i)
      if (g.s.roads[tmp[i]].from != ND_DECO0) order[n_order++] = tmp[i];
  }
  // pooled decoration road
  int deco_lanes = 0;
  for (int r = 0; r < g.n_roads; ++r)
    if (!g.s.roads[r].removed && g.s.roads[r].from == ND_DECO0) deco_lanes += g.s.roads[r].n_lanes;
  const int n_roads = n_order + (deco_lanes > 0 ? 1 : 0);
  if (n_roads > g.caps.roads) return GEN_ERR_ROADS;
  // flat lane ids: first lane of every pool road (-1 when not emitted); node ids by first appearance
  int32_t* first_lane = order + g.caps.roads;       // [caps.roads] indexed by pool road
  int32_t* node_codes = first_lane + g.caps.roads;  // [2 * caps.roads] node id -> code
  int n_nodes = 0, n_lanes = 0;
  for (int r = 0; r < g.n_roads; ++r) first_lane[r] = -1;
  auto node_id = [&](int code) -> int {
    for (int i = 0; i < n_nodes; ++i)
      if (node_codes[i] == code) return i;
    node_codes[n_nodes] = code;
    return n_nodes++;
  };
  for (int oi = 0; oi < n_roads; ++oi) {
    const bool deco = oi >= n_order;
    int from = deco ? ND_DECO0 : g.s.roads[order[oi]].from;
    int to = deco ? ND_DECO1 : g.s.roads[order[oi]].to;
    int nl = deco ? deco_lanes : g.s.roads[order[oi]].n_lanes;
    if (n_lanes + nl > g.caps.lanes) return GEN_ERR_LANES;
    int fid = node_id(from), tid = node_id(to);
    PgdRoad& pr = out.roads[oi];
    pr.first_lane = n_lanes;
    pr.n_lanes = nl;
    pr.start_node = fid;
    pr.end_node = tid;
    pr.negative = (to < 0 && !deco) ? 1 : 0;
    pr.pad[0] = pr.pad[1] = pr.pad[2] = 0;
    if (!deco) first_lane[order[oi]] = n_lanes;
    int i = 0;
    for (int r = deco ? 0 : order[oi]; r < (deco ? g.n_roads : order[oi] + 1); ++r) {
      const GRoad& rd = g.s.roads[r];
      if (deco && (rd.removed || rd.from != ND_DECO0)) continue;
      for (int k = 0; k < rd.n_lanes; ++k, ++i) {
        const GLane& ln = g.s.lanes[rd.lanes[k]];
        PgdLane& pl = out.lanes[n_lanes + i];
        pl.sx = (float)ln.sx; pl.sy = (float)ln.sy; pl.ex = (float)ln.ex; pl.ey = (float)ln.ey;
        pl.length = (float)ln.length; pl.width = (float)ln.width;
        if (ln.kind == 0) {
          pl.ax = (float)ln.dx; pl.ay = (float)ln.dy;
          pl.radius = 0.f; pl.ph0 = 0.f; pl.dir = 0.f;
          pl.heading = (float)ln.heading;
        } else {
          pl.ax = (float)ln.cx; pl.ay = (float)ln.cy;
          pl.radius = (float)ln.radius; pl.ph0 = (float)ln.ph0; pl.dir = (float)ln.dir;
          pl.heading = 0.f;
        }
        pl.road = oi; pl.idx = i; pl.kind = ln.kind; pl.pad = 0;
      }
    }
    n_lanes += nl;
  }
  // ---- static primitives, road by road: surfaces of all lanes, then lines of all lanes
  int n_boxes = 0;
  for (int oi = 0; oi < n_roads; ++oi) {
    const bool deco = oi >= n_order;
    for (int pass = 0; pass < 2; ++pass) {
      int i = 0;
      for (int r = deco ? 0 : order[oi]; r < (deco ? g.n_roads : order[oi] + 1); ++r) {
        const GRoad& rd = g.s.roads[r];
        if (deco && (rd.removed || rd.from != ND_DECO0)) continue;
        for (int k = 0; k < rd.n_lanes; ++k, ++i) {
          const GLane& ln = g.s.lanes[rd.lanes[k]];
          int lane_id = out.roads[oi].first_lane + i;
          if (pass == 0) surface_boxes(g, n_boxes, ln, lane_id);
          else line_boxes(g, n_boxes, ln, i, lane_id);
        }
      }
    }
  }
  if (g.status != GEN_OK) return g.status;
  // ---- bucket grid (tables.py add_map): PGD_GRID_CELL cells over the boxes grown by PGD_GRID_MARGIN
  const double CELL = PGD_GRID_CELL, MARGIN = PGD_GRID_MARGIN;
  double minx = 1e300, maxx = -1e300, miny = 1e300, maxy = -1e300;
  for (int i = 0; i < n_boxes; ++i) {
    const GBox& b = g.s.boxes[i];
    double ex = fabs(b.ux) * b.hl + fabs(b.uy) * b.hw + MARGIN;
    double ey = fabs(b.uy) * b.hl + fabs(b.ux) * b.hw + MARGIN;
    if (b.cx - ex < minx) minx = b.cx - ex;
    if (b.cx + ex > maxx) maxx = b.cx + ex;
    if (b.cy - ey < miny) miny = b.cy - ey;
    if (b.cy + ey > maxy) maxy = b.cy + ey;
    PgdBox& pb = out.boxes[i];
    pb.cx = (float)b.cx; pb.cy = (float)b.cy; pb.ux = (float)b.ux; pb.uy = (float)b.uy;
    pb.hl = (float)b.hl; pb.hw = (float)b.hw; pb.kind = b.kind; pb.lane = b.lane;
  }
  const double x0 = floor(minx), y0 = floor(miny);
  const int nx = (int)ceil((maxx - x0) / CELL) + 1, ny = (int)ceil((maxy - y0) / CELL) + 1;
  const int n_cells = nx * ny;
  if (n_cells + 1 > g.caps.cells) return GEN_ERR_CELLS;
  for (int c = 0; c <= n_cells; ++c) out.cell_start[c] = 0;
  // pass 0 counts; passes 1 and 2 fill: lane-surface boxes first, then the rest with PGD_ENTRY_NOT_LANE set
  for (int pass = 0; pass < 3; ++pass) {
    for (int i = 0; i < n_boxes; ++i) {
      const GBox& b = g.s.boxes[i];
      if (pass == 1 && b.kind != PGD_BOX_LANE) continue;
      if (pass == 2 && b.kind == PGD_BOX_LANE) continue;
      double ex = fabs(b.ux) * b.hl + fabs(b.uy) * b.hw + MARGIN;
      double ey = fabs(b.uy) * b.hl + fabs(b.ux) * b.hw + MARGIN;
      int ix0 = (int)floor((b.cx - ex - x0) / CELL), ix1 = (int)floor((b.cx + ex - x0) / CELL);
      int iy0 = (int)floor((b.cy - ey - y0) / CELL), iy1 = (int)floor((b.cy + ey - y0) / CELL);
      for (int iy = iy0; iy <= iy1; ++iy)
        for (int ix = ix0; ix <= ix1; ++ix) {
          int c = iy * nx + ix;
          if (pass == 0) out.cell_start[c + 1]++;
          else out.cell_entries[out.cell_start[c]++] = pass == 1 ? i : (i | PGD_ENTRY_NOT_LANE);
        }
    }
    if (pass == 1) continue;
    if (pass == 0) {
      int total = 0;
      for (int c = 0; c < n_cells; ++c) {
        int cnt = out.cell_start[c + 1];
        out.cell_start[c + 1] = total;  // start of cell c, parked one slot up
        total += cnt;
      }
      if (total > g.caps.entries) return GEN_ERR_ENTRIES;
      // shift down: cell_start[c] = start of c ; the fill pass advances it to the start of c + 1
      for (int c = 0; c < n_cells; ++c) out.cell_start[c] = out.cell_start[c + 1];
      out.cell_start[n_cells] = total;
      out.counts[4] = total;
    } else {
      // after filling, cell_start[c] holds the END of cell c: shift back up
      for (int c = n_cells; c > 0; --c) out.cell_start[c] = out.cell_start[c - 1];
      out.cell_start[0] = 0;
    }
  }
  PgdMap& pm = *out.map;
  pm.lane_off = out.lane_off; pm.n_lanes = n_lanes;
  pm.road_off = out.road_off; pm.n_roads = n_roads;
  pm.box_off = out.box_off; pm.n_boxes = n_boxes;
  pm.cell_off = out.cell_off; pm.entry_off = out.entry_off;
  pm.nx = nx; pm.ny = ny;
  pm.x0 = (float)x0; pm.y0 = (float)y0; pm.inv_cell = (float)(1.0 / CELL);
  pm.lane_width = (float)g.cfg.lane_width; pm.lane_num = g.cfg.lane_num; pm.pad = 0;
  out.counts[0] = n_lanes; out.counts[1] = n_roads; out.counts[2] = n_boxes; out.counts[3] = n_cells + 1;
  out.counts[7] = g.n_blocks;

  // ---- episode template (episode.py make_episode + tables.py add_episode)
  MT* engine_rs = &g.s.mt[0];
  MT* traffic_rs = &g.s.mt[1];
  MT* tmp = &g.s.mt[2];
  mt_seeded(engine_rs, seed);
  mt_seeded(traffic_rs, seed);
  PgdEpisode& ep = *out.episode;
  ep.map = out.map_id; ep.seed = (int32_t)seed; ep.slot_off = out.slot_off; ep.n_slots = 0; ep.n_groups = 0;
  for (int i = 0; i < PGD_MAX_GROUPS; ++i) ep.trigger_road[i] = -1;
  int n_slots = 0, n_route = 0;
  // destination sockets: one draw per polarity from a fresh stream of the map seed (navigation.py:99-121)
  const GBlock& first_blk = g.s.blocks[0];
  const GBlock& last_blk = g.s.blocks[g.n_blocks - 1];
  mt_seeded(tmp, seed);
  const GSocket dest_pos = last_blk.sockets[mt_randint(tmp, (uint32_t)last_blk.n_sockets)];
  mt_seeded(tmp, seed);
  const GSocket dest_neg = first_blk.sockets[mt_randint(tmp, (uint32_t)first_blk.n_sockets)];

  auto emit_slot = [&](int type, double u_params, int road, int lane_i, double lon, double lat, int group, int timer,
                       int64_t idm_seed) {
    if (n_slots >= PGD_MAX_SLOTS) { gen_fail(g, GEN_ERR_SLOTS); return; }
    const GRoad& rd = g.s.roads[road];
    const GLane& ln = g.s.lanes[rd.lanes[lane_i]];
    PgdSlot& ps = out.slots[n_slots];
    double x, y;
    lane_position(ln, lon, lat, &x, &y);
    VehicleBody vb = vehicle_body(type);
    double engine, brake, steer_deg, friction;
    vehicle_params(type, u_params, &engine, &brake, &steer_deg, &friction);
    double heading = py_mod(lane_heading_at(ln, lon) + PGD_PI, 2 * PGD_PI) - PGD_PI;
    ps.x = (float)x; ps.y = (float)y; ps.heading = (float)heading;
    ps.length = (float)vb.length; ps.width = (float)vb.width; ps.mass = (float)vb.mass;
    ps.lf = (float)vb.lf; ps.lr = (float)vb.lr;
    ps.max_engine = (float)engine; ps.max_brake = (float)brake;
    ps.max_steer = (float)(steer_deg * (PGD_PI / 180.0)); ps.friction = (float)friction;
    ps.lane = first_lane[road] + lane_i;
    ps.type = type; ps.group = group; ps.drop_substeps = drop_substeps(type); ps.overtake_timer = timer;
    ps.pad = 0;
    for (int i = 0; i < PGD_N_RND25; ++i) ps.rnd25[i] = 0;
    if (idm_seed >= 0) {
      mt_seeded(tmp, (uint64_t)idm_seed);
      mt_randint(tmp, 50);
      for (int i = 0; i < PGD_N_RND25; ++i) ps.rnd25[i] = (uint8_t)mt_randint(tmp, 25);
    }
    // route (episode.py route_for)
    const bool negative = rd.to < 0;
    const GSocket& sock = negative ? dest_neg : dest_pos;
    const int n_sock = negative ? first_blk.n_sockets : last_blk.n_sockets;
    const int start = rd.from;
    if (n_sock > 1 && (start == sock.from || start == sock.to || start == -sock.to || start == -sock.from))
      gen_fail(g, GEN_ERR_DEST);
    const int final_node = negative ? -sock.from : sock.to;
    int path[64];
    int len = shortest_path(g, order, n_order, start, final_node, path, 64);
    if (len <= 2) {
      path[0] = rd.from; path[1] = rd.to;
      len = 2;
    }
    if (n_route + len > g.caps.route) { gen_fail(g, GEN_ERR_ROUTE); return; }
    ps.route_off = out.route_off + n_route;
    ps.route_len = len;
    for (int i = 0; i < len; ++i) {
      out.route_nodes[n_route + i] = node_id(path[i]);
      int rid = -1;
      if (i + 1 < len) {
        for (int oi = 0; oi < n_order; ++oi)
          if (g.s.roads[order[oi]].from == path[i] && g.s.roads[order[oi]].to == path[i + 1]) { rid = oi; break; }
        if (rid < 0) gen_fail(g, GEN_ERR_LOOKUP);
      }
      out.route_roads[n_route + i] = rid;
    }
    n_route += len;
    ++n_slots;
  };

  // ego
  {
    uint32_t ego_seed = mt_randint(engine_rs, 65536);
    mt_seeded(tmp, ego_seed);
    uint32_t q = mt_randint(tmp, 1000000);
    double u = first_sample_of(tmp, q);
    int r0 = find_road(g, ND_START, ND_START2, 0, g.n_roads);
    if (r0 < 0 || cfg.spawn_lane < 0 || cfg.spawn_lane >= g.s.roads[r0].n_lanes) return GEN_ERR_CONFIG;
    emit_slot(4, u, r0, cfg.spawn_lane, cfg.spawn_long, cfg.spawn_lat, -1, 0, -1);
  }
  if (!(fabs(cfg.density) < 1e-2)) {
    if (g.n_blocks - 1 > PGD_MAX_GROUPS) return GEN_ERR_GROUPS;
    const double type_prob[5] = {0.2, 0.3, 0.3, 0.2, 0};
    for (int bi = 1; bi < g.n_blocks && g.status == GEN_OK; ++bi) {
      const GBlock& b = g.s.blocks[bi];
      // spawn roads of the block (get_intermediate_spawn_lanes of each block type)
      int spawn[MAX_BLOCK_ROADS + MAX_RESPAWN + 4];
      int ns = 0;
      if (b.type == BK_X || b.type == BK_T || b.type == BK_O) {
        for (int i = 0; i < b.n_respawn; ++i) spawn[ns++] = block_road(g, b, b.respawn[i][0], b.respawn[i][1]);
        if (b.type == BK_O)
          for (int i = 0; i < b.n_ring; ++i) spawn[ns++] = b.ring[i];
      } else {
        int tmpo[MAX_BLOCK_ROADS + 1];
        int n = block_road_order(g, b.road_begin, b.road_end, tmpo);
        for (int i = 0; i < n; ++i) {
          const GRoad& rd = g.s.roads[tmpo[i]];
          if (rd.to >= 0 && rd.from != ND_DECO0) spawn[ns++] = tmpo[i];
        }
        for (int i = 0; i < b.n_respawn; ++i) {
          int r = block_road(g, b, b.respawn[i][0], b.respawn[i][1]);
          bool have = false;
          for (int k = 0; k < ns; ++k) have = have || spawn[k] == r;
          if (!have) spawn[ns++] = r;
        }
      }
      int32_t* cand = g.s.cand;
      int nc = 0;
      double total_length = 0;
      bool first_term = true;
      for (int si = 0; si < ns; ++si) {
        const GRoad& rd = g.s.roads[spawn[si]];
        for (int k = 0; k < rd.n_lanes; ++k) {
          const GLane& ln = g.s.lanes[rd.lanes[k]];
          int cnt = (int)(ln.length / 10);
          for (int j = 0; j < cnt; ++j) {
            if (nc >= g.caps.cand) { gen_fail(g, GEN_ERR_CAND); break; }
            cand[3 * nc] = spawn[si]; cand[3 * nc + 1] = k; cand[3 * nc + 2] = j * 10;
            ++nc;
          }
          total_length = first_term ? ln.length : total_length + ln.length;
          first_term = false;
        }
      }
      int total = (int)floor((double)(int)floor(total_length / 10) * cfg.density);
      // RandomState.shuffle of a list
      for (int i = nc - 1; i >= 1; --i) {
        int j = (int)mt_interval(traffic_rs, (uint32_t)i);
        for (int c = 0; c < 3; ++c) {
          int32_t t = cand[3 * i + c];
          cand[3 * i + c] = cand[3 * j + c];
          cand[3 * j + c] = t;
        }
      }
      int take = total < nc ? total : nc;
      const int group = bi - 1;
      ep.trigger_road[group] = -1;
      for (int oi = 0; oi < n_order; ++oi)
        if (g.s.roads[order[oi]].from == b.pre.from && g.s.roads[order[oi]].to == b.pre.to) ep.trigger_road[group] = oi;
      if (ep.trigger_road[group] < 0) gen_fail(g, GEN_ERR_LOOKUP);
      for (int v = 0; v < take && g.status == GEN_OK; ++v) {
        int type = mt_choice_p(traffic_rs, type_prob, 5);
        uint32_t vseed = mt_randint(engine_rs, 65536);
        mt_seeded(tmp, vseed);
        uint32_t q = mt_randint(tmp, 1000000);
        double u = first_sample_of(tmp, q);
        uint32_t idm_seed = mt_randint(traffic_rs, 65536);
        mt_seeded(tmp, idm_seed);
        int timer = (int)mt_randint(tmp, 50);
        emit_slot(type, u, cand[3 * v], cand[3 * v + 1], (double)cand[3 * v + 2], 0.0, group, timer, (int64_t)idm_seed);
      }
      ep.n_groups = group + 1;
    }
  }
  ep.n_slots = n_slots;
  out.counts[5] = n_slots;
  out.counts[6] = n_route;
  return g.status;
}

}  // namespace pgdgen
#endif
