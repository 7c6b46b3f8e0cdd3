/* Plain-old-data tables read by the step kernel (and, for checking, by oracle/).
 *
 * DATA FORMATS ONLY -- no algorithm lives in this header.  pgdrive_b200/tables.py writes these
 * records (numpy structured dtypes with the same field order); every id stored in a record is
 * map-local and PgdMap carries the offsets into the concatenated arrays.
 *
 * Reference objects each record flattens (paths under /root/reference/pgdrive):
 *   PgdLane     StraightLane / CircularLane            component/lane/straight_lane.py:13-67, circular_lane.py:9-67
 *   PgdRoad     Road + RoadNetwork.graph[from][to]     component/road/road.py:11-48, road_network.py:18-56
 *   PgdBox      Bullet static primitives of a block    component/blocks/base_block.py:181-463
 *   PgdSlot     one vehicle as reset() creates it      component/vehicle/base_vehicle.py:292-339, vehicle_type.py:7-78,
 *                                                      manager/traffic_manager.py:239-290, policy/idm_policy.py:180-188
 *   PgdEpisode  what reset(force_seed=s) decides       envs/base_env.py:269-301
 */
#ifndef PGD_TABLES_H
#define PGD_TABLES_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { PGD_LANE_STRAIGHT = 0, PGD_LANE_ARC = 1 };
enum { PGD_BOX_LANE = 0, PGD_BOX_WHITE = 1, PGD_BOX_YELLOW = 2, PGD_BOX_BROKEN = 3, PGD_BOX_SIDEWALK = 4 };
/* Bucket-grid entries (cell_entries) are box ids; inside a cell the lane-surface boxes come first (ascending id), then
 * everything else (ascending id) with this bit set, so that a scan for lanes stops at the first flagged entry and a
 * scan for line ghosts skips the others without touching the box records. */
enum { PGD_ENTRY_NOT_LANE = 1 << 30, PGD_ENTRY_ID_MASK = (1 << 30) - 1 };
/* Bucket grid geometry: square cells of PGD_GRID_CELL metres; a box is listed in every cell that its bounding rectangle
 * grown by PGD_GRID_MARGIN touches (>= the half diagonal of the largest chassis, 5.8 x 2.3 m -> 3.12 m, so the cell
 * under a chassis centre lists everything the chassis can overlap; a detector ray samples a point every 2 margins). */
#define PGD_GRID_CELL 4.0
#define PGD_GRID_MARGIN 3.2
/* PGD_OBS_DIM is the PGDrive-v0 observation (no side / lane-line detector); with detectors see pgd_obs_dim(). */
enum { PGD_MAX_SLOTS = 32, PGD_MAX_GROUPS = 11, PGD_N_RND25 = 16, PGD_OBS_DIM = 274, PGD_LIDAR_BEAMS = 240,
       PGD_MAX_DETECTOR_BEAMS = 240 };

typedef struct {          /* 64 B */
  float sx, sy, ex, ey;   /* start / end point of the centre line */
  float ax, ay;           /* straight: unit direction; arc: centre */
  float length, width;
  float radius, ph0;      /* arc: radius, start phase */
  float dir;              /* arc: +1 clockwise, -1 counter-clockwise; straight: 0 */
  float heading;          /* straight: atan2(direction) */
  int32_t road, idx, kind, pad;
} PgdLane;

typedef struct {          /* 32 B */
  int32_t first_lane, n_lanes, start_node, end_node, negative, pad[3];
} PgdRoad;

typedef struct {          /* 32 B: rectangle centre, unit axis, half extents */
  float cx, cy, ux, uy, hl, hw;
  int32_t kind, lane;
} PgdBox;

typedef struct {          /* 64 B */
  int32_t lane_off, n_lanes, road_off, n_roads, box_off, n_boxes;
  int32_t cell_off, entry_off, nx, ny;   /* bucket grid: cell_start[cell_off + c], cell_entries[entry_off + k] */
  float x0, y0, inv_cell;
  float lane_width;       /* map_config lane_width */
  int32_t lane_num, pad;
} PgdMap;

/* PgdSlot.group: >= 0 trigger group (woken when the ego reaches the group's road), -1 the ego, PGD_GROUP_AWAKE a
 * traffic vehicle that drives from the first step (traffic_mode "respawn", manager/traffic_manager.py:63-66,224-237),
 * PGD_GROUP_STATIC an object or broken-down vehicle of an accident scene (manager/object_manager.py:40-124) */
enum { PGD_GROUP_AWAKE = -2, PGD_GROUP_STATIC = -3 };
/* PgdSlot.type: 0..4 vehicle types s, m, l, xl, default (vehicle_type.py); >= PGD_TYPE_OBJECT a traffic cone / warning
 * tripod / barrier of an accident scene (component/static_object/traffic_object.py:37-103), always PGD_GROUP_STATIC like a
 * broken-down vehicle: present, an obstacle for IDM and the lidar, never driving.  Touching one sets crash_object (once
 * per object: COST_ONCE), touching a vehicle crash_vehicle (engine/core/collision_callback.py:7-35). */
enum { PGD_TYPE_OBJECT = 5, PGD_TYPE_CONE = 5, PGD_TYPE_WARNING = 6, PGD_TYPE_BARRIER = 7 };

typedef struct {          /* 96 B */
  float x, y, heading;    /* spawn pose */
  float length, width, mass, lf, lr;
  float max_engine, max_brake, max_steer, friction;   /* max_steer in radians */
  int32_t lane, type, group, drop_substeps, overtake_timer, route_off, route_len, pad;
  uint8_t rnd25[PGD_N_RND25];   /* successive randint(0, 25) draws of the slot's IDM stream */
} PgdSlot;

typedef struct {          /* 64 B */
  int32_t map, seed, slot_off, n_slots, n_groups;
  int32_t trigger_road[PGD_MAX_GROUPS];  /* group g wakes when the ego is on this road (in order) */
} PgdEpisode;

/* Host-side view of a whole table set (pointers into caller-owned memory). */
typedef struct {
  const PgdMap* maps;            int32_t n_maps;
  const PgdLane* lanes;          int32_t n_lanes;
  const PgdRoad* roads;          int32_t n_roads;
  const PgdBox* boxes;           int32_t n_boxes;
  const int32_t* cell_start;     int32_t n_cell_start;
  const int32_t* cell_entries;   int32_t n_cell_entries;
  const PgdEpisode* episodes;    int32_t n_episodes;
  const PgdSlot* slots;          int32_t n_slots;
  const int32_t* route_nodes;    /* route_len node ids per slot */
  const int32_t* route_roads;    /* road id of (node[k], node[k+1]); -1 after the last node */
  int32_t n_route;
} PgdTables;

/* Settings of the on-device reset path (pgd_generate_tables): what the reference keeps in map_config
 * (component/map/base_map.py:16-35: block_num | block_sequence, lane_num, lane_width, exit_length), the traffic
 * density (manager/traffic_manager.py:263-270) and the ego spawn (base_vehicle.py:299-311). */
typedef struct {
  int32_t block_num;       /* number of searched blocks */
  int32_t lane_num;
  int32_t n_fixed;         /* > 0: the block types are given (map_config type "block_sequence") */
  int32_t spawn_lane;      /* lane index on the first road (">", ">>") */
  double lane_width, exit_length, density, spawn_long, spawn_lat;
  int8_t fixed_types[32];  /* 0..6 = C S r R X T O (order of BLOCK_TYPE_DISTRIBUTION_V2) */
  /* manager/map_manager.py:157-169: per-seed lane width = rand() * 1.5 + 3.0 and lane number = randint(2, 3) from
   * the map manager's stream of the seed, replacing lane_width / lane_num above */
  int32_t random_lane_width, random_lane_num;
} PgdGenConfig;

/* Per-map capacities of the generated tables (map m owns the fixed-stride slice [m * cap, (m + 1) * cap) of every
 * table; PgdMap / PgdEpisode / PgdSlot carry the offsets as usual).  A map that does not fit is reported, never
 * truncated. */
typedef struct {
  int32_t blocks, lanes, roads, boxes, cells, entries, queue, route, cand;
} PgdGenCaps;

/* Reward / termination scheme (envs/pgdrive_env.py:93-108) and stepping (envs/base_env.py:33,70). */
typedef struct {
  int32_t num_envs;
  int32_t num_slots;        /* vehicle slots per env: 16, 24 or 32 */
  int32_t decision_repeat;  /* physics sub-steps per env step (5) */
  int32_t horizon;          /* 0 = none */
  float dt;                 /* physics_world_step_size (0.02) */
  float success_reward, out_of_road_penalty, crash_vehicle_penalty;
  float driving_reward, speed_reward;
  float out_of_road_cost, crash_vehicle_cost;
  int32_t use_lateral, out_of_route_done;
  int32_t auto_reset;       /* 1: an env that reported done is reset at its next step (action ignored) */
  /* ray fans against lane-line ghosts (vehicle_module/distance_detector.py:137-152; off in PGDrive-v0):
   * side detector = continuous lines only, lane-line detector = continuous + broken; beam i points at
   * i * 2pi / n + 90 degrees from the heading */
  int32_t n_side, n_lane_line;
  float side_distance, lane_line_distance;
  /* envs/base_env.py:29, obs/state_obs.py:18-23,103-105: the ego is one of the five vehicle types (chosen per seed by
   * the host) and the observation gains LENGTH / 10 and WIDTH / 2.5 after the lane-line beams */
  int32_t random_agent_model;
  /* obs/state_obs.py:165-182: clip(p + N(0, sigma), 0, 1) on the 240 lidar values, then each set to 0 with probability
   * dropout.  The reference draws from numpy's process-global generator; here a counter-based generator keyed by
   * (noise_seed, API call, environment, beam) -- include/pgd_math.h -- so only the distribution matches. */
  float lidar_gaussian_noise, lidar_dropout_prob;
  int32_t noise_seed;
  /* base_vehicle.py:249,351-358 (vehicle_config.increment_steering): steering += action[0] * 0.05, clipped to [-1, 1];
   * the raw steering action of the last step (observation value 5) is kept in the ego's otherwise unused pid_hp */
  int32_t increment_steering;
  /* accident scenes (envs/safe_pgdrive_env.py:7-63, envs/pgdrive_env.py:197-258): reward / cost of touching a traffic
   * object; safe_rl_env: an episode does not end on a crash (vehicle or object), it only costs */
  float crash_object_penalty, crash_object_cost;
  int32_t safe_rl_env;
} PgdConfig;

/* Observation length (obs/state_obs.py:18-23,108-115,125-130): (n_side or 2) + 6 + n_lane_line [+ 2 vehicle
 * dimensions] + 10 + 16 + 240. */
static inline int32_t pgd_obs_dim(const PgdConfig* c) {
  return (c->n_side > 0 ? c->n_side : 2) + 6 + c->n_lane_line + (c->random_agent_model ? 2 : 0) + 10 + 16 +
         PGD_LIDAR_BEAMS;
}

/* per-step info (base_vehicle.py:262-272, pgdrive_env.py:165-207, base_env.py:335-339) */
enum {
  PGD_F_CRASH_VEHICLE = 1 << 0, PGD_F_OUT_OF_ROAD = 1 << 1, PGD_F_ARRIVE_DEST = 1 << 2, PGD_F_MAX_STEP = 1 << 3,
  PGD_F_ON_YELLOW = 1 << 4, PGD_F_ON_WHITE = 1 << 5, PGD_F_ON_BROKEN = 1 << 6, PGD_F_CRASH_SIDEWALK = 1 << 7,
  PGD_F_ON_LANE = 1 << 8, PGD_F_OUT_OF_ROUTE = 1 << 9, PGD_F_WAS_RESET = 1 << 10, PGD_F_CRASH_OBJECT = 1 << 11
};
typedef struct {          /* 40 B */
  float velocity, steering, acceleration, step_energy, episode_energy;
  float step_reward, episode_reward, cost;
  int32_t episode_length;
  uint32_t flags;
} PgdInfo;

/* Exchange format for pgd_get_state / pgd_set_state (parity debugging; not the device layout). */
enum { PGD_V_ALIVE = 1, PGD_V_ACTIVE = 2, PGD_V_ON_LANE = 4, PGD_V_CRASHED = 16 /* object already charged (COST_ONCE) */ };
typedef struct {          /* 80 B */
  float x, y, heading, speed;
  float steer, throttle;
  float pid_hp, pid_hi, pid_lp, pid_li;   /* heading / lateral PID: last error, summed error */
  float target_speed;
  int32_t lane, ck0, ck1, rt_lane, timer, rnd_n, airborne, flags;
  float yaw_rate;
} PgdVehState;
typedef struct {
  int32_t episode, next_group, done, ep_len;
  float prev_steer, prev_throttle, ep_reward, energy;
  PgdVehState veh[PGD_MAX_SLOTS];
} PgdEnvState;

#ifdef __cplusplus
}
#endif
#endif
