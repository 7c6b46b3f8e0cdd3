"""Flatten generated maps and episode templates into the plain-old-data tables the step kernel reads.

Layouts are the C structs of ``include/pgd_tables.h`` (numpy structured dtypes here, same field
order).  One table set holds many maps / seeds; every id stored in a record is map-local and the
per-map header carries the offsets.

What is tabulated and where the reference defines it:
  lanes / roads      component/lane/{straight,circular}_lane.py, component/road/road_network.py
  boxes              the static collision primitives the reference hands to Bullet:
                     lane surfaces (base_block.py:396-463), lane-line ghosts (:181-365),
                     sidewalks (:219-234,367-394); constants.py:228-256 for the sizes
  grid               a uniform bucket grid over the boxes (ours; the reference relies on Bullet's
                     broad phase)
  slots / episodes   pgdrive_b200/episode.py (reset decisions) + vehicle_type.py dimensions
"""
import math

import numpy as np

from . import rng
from .episode import OBJECT_BODY, VEHICLE_BODY, make_episode
from .roadnet import BROKEN, CONTINUOUS, DECO, NONE, SIDE, YELLOW, is_negative, norm2

LANE_DT = np.dtype([
    ("sx", "f4"), ("sy", "f4"), ("ex", "f4"), ("ey", "f4"), ("ax", "f4"), ("ay", "f4"), ("length", "f4"),
    ("width", "f4"), ("radius", "f4"), ("ph0", "f4"), ("dir", "f4"), ("heading", "f4"), ("road", "i4"), ("idx", "i4"),
    ("kind", "i4"), ("pad", "i4")
])
ROAD_DT = np.dtype([
    ("first_lane", "i4"), ("n_lanes", "i4"), ("start_node", "i4"), ("end_node", "i4"), ("negative", "i4"),
    ("pad", "i4", (3, ))
])
BOX_DT = np.dtype([("cx", "f4"), ("cy", "f4"), ("ux", "f4"), ("uy", "f4"), ("hl", "f4"), ("hw", "f4"), ("kind", "i4"),
                   ("lane", "i4")])
MAP_DT = np.dtype([
    ("lane_off", "i4"), ("n_lanes", "i4"), ("road_off", "i4"), ("n_roads", "i4"), ("box_off", "i4"), ("n_boxes", "i4"),
    ("cell_off", "i4"), ("entry_off", "i4"), ("nx", "i4"), ("ny", "i4"), ("x0", "f4"), ("y0", "f4"),
    ("inv_cell", "f4"), ("lane_width", "f4"), ("lane_num", "i4"), ("pad", "i4")
])
SLOT_DT = np.dtype([
    ("x", "f4"), ("y", "f4"), ("heading", "f4"), ("length", "f4"), ("width", "f4"), ("mass", "f4"), ("lf", "f4"),
    ("lr", "f4"), ("max_engine", "f4"), ("max_brake", "f4"), ("max_steer", "f4"), ("friction", "f4"), ("lane", "i4"),
    ("type", "i4"), ("group", "i4"), ("drop_substeps", "i4"), ("overtake_timer", "i4"), ("route_off", "i4"),
    ("route_len", "i4"), ("pad", "i4"), ("rnd25", "u1", (16, ))
])
EPISODE_DT = np.dtype([
    ("map", "i4"), ("seed", "i4"), ("slot_off", "i4"), ("n_slots", "i4"), ("n_groups", "i4"),
    ("trigger_road", "i4", (11, ))
])
assert LANE_DT.itemsize == 64 and ROAD_DT.itemsize == 32 and BOX_DT.itemsize == 32
assert MAP_DT.itemsize == 64 and SLOT_DT.itemsize == 96 and EPISODE_DT.itemsize == 64

BOX_LANE, BOX_WHITE, BOX_YELLOW, BOX_BROKEN, BOX_SIDEWALK = 0, 1, 2, 3, 4
TYPE_ID = {"s": 0, "m": 1, "l": 2, "xl": 3, "default": 4, "TrafficCone": 5, "TrafficWarning": 6, "TrafficBarrier": 7}
MAX_GROUPS = 11
N_RND25 = 16

# constants.py:228-256
LINE_HALF_WIDTH = 0.15 / 2
CIRCULAR_SEGMENT = 4.0
STRIPE = 1.5
SIDEWALK_SEG = 3.0
SIDEWALK_WIDTH = 3.0
SIDEWALK_GAP = 0.6
CELL = 4.0  # bucket size [m] (PGD_GRID_CELL, include/pgd_tables.h)
GRID_MARGIN = 3.2  # PGD_GRID_MARGIN: >= half diagonal of the largest chassis (5.8 x 2.3 -> 3.12 m)
ENTRY_NOT_LANE = 1 << 30  # PGD_ENTRY_NOT_LANE (include/pgd_tables.h)
GROUP_STATIC = -3  # PGD_GROUP_STATIC: objects and broken-down vehicles of accident scenes
GROUP_AWAKE = -2  # PGD_GROUP_AWAKE: PgdSlot.group of a vehicle that drives from the first step (traffic_mode respawn)

GRAVITY = 9.81


def _seg_box(p0, p1, mid, half_len, half_w, kind, lane):
    dx, dy = p1[0] - p0[0], p1[1] - p0[1]
    d = norm2(dx, dy)
    return (mid[0], mid[1], dx / d, dy / d, half_len, half_w, kind, lane)


def _mid(a, b):
    return ((a[0] + b[0]) / 2, (a[1] + b[1]) / 2)


def lane_boxes(lane, lane_in_road, lane_id):
    """Line ghosts and sidewalks of one lane (_add_pgdrive_lanes, base_block.py:181-265)."""
    out = []
    w = lane.width
    straight = lane.kind == "S"
    for k, side in enumerate((-1, 1)):
        lt = lane.line_types[k]
        if lt == NONE or (lane_in_road != 0 and k == 0):
            if straight or lane.radius != w / 2:
                continue
        lat = side * w / 2
        colour = lane.line_color[k]
        if lt in (CONTINUOUS, SIDE):
            kind = BOX_YELLOW if colour == YELLOW else BOX_WHITE
            if straight:
                a, b = lane.position(0, lat), lane.position(lane.length, lat)
                out.append(_seg_box(a, b, lane.position(lane.length / 2, lat), norm2(b[0] - a[0], b[1] - a[1]) / 2,
                                    LINE_HALF_WIDTH, kind, lane_id))
            else:
                n = int(lane.length / CIRCULAR_SEGMENT)
                cuts = [s * CIRCULAR_SEGMENT for s in range(n + 1)] + [lane.length]
                for s0, s1 in zip(cuts[:-1], cuts[1:]):
                    a, b = lane.position(s0, lat), lane.position(s1, lat)
                    ln = norm2(b[0] - a[0], b[1] - a[1])
                    if ln <= 0:
                        continue
                    out.append(_seg_box(a, b, _mid(a, b), ln / 2, LINE_HALF_WIDTH, kind, lane_id))
            if lt == SIDE:
                radius = 0.0 if straight else lane.radius
                n = int(lane.length / SIDEWALK_SEG)
                cuts = [s * SIDEWALK_SEG for s in range(n + 1)] + [lane.length]
                for j, (s0, s1) in enumerate(zip(cuts[:-1], cuts[1:])):
                    a, b = lane.position(s0, lat), lane.position(s1, lat)
                    ln = norm2(b[0] - a[0], b[1] - a[1])
                    if j == n and not ln > 1e-1:
                        continue
                    if radius == 0:
                        factor = 1.0
                    elif lane.dir == 1:
                        factor = 1 - SIDEWALK_GAP / radius
                    else:
                        factor = (1 + SIDEWALK_WIDTH / radius) * (1 + SIDEWALK_GAP / radius)
                    m = _mid(a, b)
                    vx, vy = -(b[1] - a[1]) / ln, (b[0] - a[0]) / ln
                    off = SIDEWALK_WIDTH / 2 + SIDEWALK_GAP
                    out.append(_seg_box(a, b, (m[0] + vx * off, m[1] + vy * off), ln * factor / 2, SIDEWALK_WIDTH / 2,
                                        BOX_SIDEWALK, lane_id))
        elif lt == BROKEN:
            if straight:
                a, b = lane.position(0, lat), lane.position(lane.length, lat)
                out.append(_seg_box(a, b, lane.position(lane.length / 2, lat), norm2(b[0] - a[0], b[1] - a[1]) / 2,
                                    LINE_HALF_WIDTH, BOX_BROKEN, lane_id))
            else:
                n = int(lane.length / (2 * STRIPE))
                for s in range(n):
                    a = lane.position(s * STRIPE * 2, lat)
                    b = lane.position(s * STRIPE * 2 + STRIPE, lat)
                    ln = norm2(b[0] - a[0], b[1] - a[1])
                    if ln <= 0:
                        continue
                    # ghost half-extent is the full stripe length (base_block.py:339-345)
                    out.append(_seg_box(a, b, lane.position(s * STRIPE * 2 + STRIPE / 2, lat), ln, LINE_HALF_WIDTH,
                                        BOX_BROKEN, lane_id))
                a = lane.position(n * STRIPE * 2, lat)
                b = lane.position(lane.length + STRIPE, lat)
                ln = norm2(b[0] - a[0], b[1] - a[1])
                if ln > 0:
                    out.append(_seg_box(a, b, _mid(a, b), ln, LINE_HALF_WIDTH, BOX_BROKEN, lane_id))
    return out


def surface_boxes(lane, lane_id):
    """Lane-surface boxes used for localisation (_add_lane_surface / _add_lane2bullet)."""
    out = []
    width = lane.width + SIDEWALK_GAP * 2
    if lane.kind == "S":
        mid, end = lane.position(lane.length / 2, 0), lane.position(lane.length, 0)
        out.append(_seg_box(mid, end, mid, (lane.length + 0.1) / 2, width / 2, BOX_LANE, lane_id))
    else:
        n = int(lane.length / CIRCULAR_SEGMENT)
        for i in range(n):
            mid = lane.position(lane.length * (i + .5) / n, 0)
            end = lane.position(lane.length * (i + 1) / n, 0)
            out.append(_seg_box(mid, end, mid, (lane.length * 1.3 / n + 0.1) / 2, width / 2, BOX_LANE, lane_id))
    return out


def drop_substeps(vtype):
    """Sub-steps a freshly placed vehicle spends falling onto its wheels: it is placed with its
    origin HEIGHT/2 + 1 m above the road (base_vehicle.py:311) and rests at about tyre radius +
    wheel-axis offset (base_vehicle.py:543-546)."""
    length, width, height, mass, lf, lr, tyre, track = VEHICLE_BODY[vtype]
    axis = 0.3 if vtype == "xl" else 0.2
    fall = height / 2 + 1 - (tyre + axis)
    return int(math.ceil(math.sqrt(2 * fall / GRAVITY) / 0.02))


class MapIndex:
    """Name <-> id maps of one map (node / road / lane numbering used by the tables)."""
    def __init__(self, pgmap):
        self.nodes = {}
        self.roads = {}
        self.lane_of = {}
        self.lanes = []
        self.road_list = []
        for (frm, to), lanes in pgmap.net.roads():
            for n in (frm, to):
                self.nodes.setdefault(n, len(self.nodes))
            self.roads[(frm, to)] = len(self.road_list)
            self.road_list.append((frm, to, len(self.lanes), len(lanes)))
            for i, ln in enumerate(lanes):
                self.lane_of[(frm, to, i)] = len(self.lanes)
                self.lanes.append(ln)


class TableSet:
    """Concatenated tables for a list of seeds (one map + one episode template per seed)."""
    def __init__(self):
        self.maps, self.lanes, self.roads, self.boxes = [], [], [], []
        self.cell_start, self.cell_entries = [], []
        self.episodes, self.slots, self.route_nodes, self.route_roads = [], [], [], []
        self.seeds = []
        self.index = []  # MapIndex per map (host-side debugging / tests)

    # -- maps ------------------------------------------------------------------------------------
    def add_map(self, pgmap):
        mi = MapIndex(pgmap)
        lane_off, road_off, box_off = len(self.lanes), len(self.roads), len(self.boxes)
        for rid, (frm, to, first, n) in enumerate(mi.road_list):
            neg = 1 if (is_negative((frm, to)) and (frm, to) != DECO) else 0
            self.roads.append((first, n, mi.nodes[frm], mi.nodes[to], neg, (0, 0, 0)))
            for i in range(n):
                ln = mi.lanes[first + i]
                if ln.kind == "S":
                    rec = (ln.sx, ln.sy, ln.ex, ln.ey, ln.dx, ln.dy, ln.length, ln.width, 0.0, 0.0, 0.0, ln.heading,
                           rid, i, 0, 0)
                else:
                    rec = (ln.sx, ln.sy, ln.ex, ln.ey, ln.cx, ln.cy, ln.length, ln.width, ln.radius, ln.ph0,
                           float(ln.dir), 0.0, rid, i, 1, 0)
                self.lanes.append(rec)
        boxes = []
        for rid, (frm, to, first, n) in enumerate(mi.road_list):
            for i in range(n):
                boxes += surface_boxes(mi.lanes[first + i], first + i)
            for i in range(n):
                boxes += lane_boxes(mi.lanes[first + i], i, first + i)
        self.boxes += boxes
        # bucket grid over box bounding rectangles grown by GRID_MARGIN
        arr = np.array([b[:6] for b in boxes], dtype=np.float64)
        ex = np.abs(arr[:, 2]) * arr[:, 4] + np.abs(arr[:, 3]) * arr[:, 5] + GRID_MARGIN
        ey = np.abs(arr[:, 3]) * arr[:, 4] + np.abs(arr[:, 2]) * arr[:, 5] + GRID_MARGIN
        x0, y0 = float(np.floor((arr[:, 0] - ex).min())), float(np.floor((arr[:, 1] - ey).min()))
        nx = int(math.ceil(((arr[:, 0] + ex).max() - x0) / CELL)) + 1
        ny = int(math.ceil(((arr[:, 1] + ey).max() - y0) / CELL)) + 1
        cells = [[] for _ in range(nx * ny)]
        ix0 = np.floor((arr[:, 0] - ex - x0) / CELL).astype(int)
        ix1 = np.floor((arr[:, 0] + ex - x0) / CELL).astype(int)
        iy0 = np.floor((arr[:, 1] - ey - y0) / CELL).astype(int)
        iy1 = np.floor((arr[:, 1] + ey - y0) / CELL).astype(int)
        for b in range(len(boxes)):
            for iy in range(iy0[b], iy1[b] + 1):
                for ix in range(ix0[b], ix1[b] + 1):
                    cells[iy * nx + ix].append(b)
        cell_off, entry_off = len(self.cell_start), len(self.cell_entries)
        pos = 0
        for c in cells:  # lane-surface boxes first, the rest flagged (PGD_ENTRY_NOT_LANE, include/pgd_tables.h)
            self.cell_start.append(pos)
            self.cell_entries += [b for b in c if boxes[b][6] == 0] + [b | ENTRY_NOT_LANE for b in c if boxes[b][6] != 0]
            pos += len(c)
        self.cell_start.append(pos)
        self.maps.append((lane_off, len(mi.lanes), road_off, len(mi.road_list), box_off, len(boxes), cell_off,
                          entry_off, nx, ny, x0, y0, 1.0 / CELL, pgmap.lane_width, pgmap.lane_num, 0))
        self.index.append(mi)
        return len(self.maps) - 1

    # -- episodes --------------------------------------------------------------------------------
    def _route(self, mi, checkpoints):
        off = len(self.route_nodes)
        self.route_nodes += [mi.nodes[c] for c in checkpoints]
        self.route_roads += [mi.roads[(a, b)] for a, b in zip(checkpoints[:-1], checkpoints[1:])] + [-1]
        return off, len(checkpoints)

    def _slot(self, pgmap, mi, vtype, params, lane_index, lon, lat, group, timer, idm_seed, checkpoints):
        ln = pgmap.net.lanes((lane_index[0], lane_index[1]))[lane_index[2]]
        x, y = ln.position(lon, lat)
        length, width, height, mass, lf, lr, tyre, track = VEHICLE_BODY[vtype]
        rnd = np.zeros(N_RND25, dtype=np.uint8)
        if idm_seed is not None:
            rs = rng.seeded(idm_seed)
            rs.randint(0, 50)  # the constructor's overtake_timer draw (idm_policy.py:185)
            rnd[:] = [rs.randint(0, 25) for _ in range(N_RND25)]  # move_to_next_road draws (idm_policy.py:239)
        off, n = self._route(mi, checkpoints)
        # spawn heading wrapped into [-pi, pi): the simulator keeps headings wrapped (Panda's getH() does too)
        heading = (ln.heading_at(lon) + math.pi) % (2 * math.pi) - math.pi
        self.slots.append((x, y, heading, length, width, mass, lf, lr, params["max_engine_force"],
                           params["max_brake_force"], math.radians(params["max_steering"]), params["wheel_friction"],
                           mi.lane_of[tuple(lane_index)], TYPE_ID[vtype], group, drop_substeps(vtype), timer, off, n, 0,
                           rnd))

    def _static(self, pgmap, mi, obj):
        """A cone / tripod / barrier or broken-down vehicle (episode.StaticObject) as a slot that never wakes."""
        ln = pgmap.net.lanes((obj.lane[0], obj.lane[1]))[obj.lane[2]]
        x, y = ln.position(obj.long, obj.lat)
        heading = (ln.heading_at(obj.long) + math.pi) % (2 * math.pi) - math.pi
        if obj.kind == "vehicle":
            length, width, height, mass, lf, lr, tyre, track = VEHICLE_BODY[obj.type]
            p, tid, drop = obj.params, TYPE_ID[obj.type], drop_substeps(obj.type)
            eng, brk, steer, fric = p["max_engine_force"], p["max_brake_force"], math.radians(p["max_steering"]), \
                p["wheel_friction"]
        else:
            length, width, mass = OBJECT_BODY[obj.kind]
            lf = lr = length / 2
            tid, drop, eng, brk, steer, fric = TYPE_ID[obj.kind], 0, 0.0, 0.0, 0.0, 0.9
        self.slots.append((x, y, heading, length, width, mass, lf, lr, eng, brk, steer, fric,
                           mi.lane_of[tuple(obj.lane)], tid, GROUP_STATIC, drop, 0, 0, 0, 0,
                           np.zeros(N_RND25, dtype=np.uint8)))

    def add_episode(self, pgmap, map_id, ep, spawn_lane=(">", ">>", 0), spawn_long=5.0, spawn_lat=0.0):
        mi = self.index[map_id]
        slot_off = len(self.slots)
        self._slot(pgmap, mi, getattr(ep, "ego_type", "default"), ep.ego_params, spawn_lane, spawn_long, spawn_lat, -1, 0,
                   None, ep.ego_checkpoints)
        groups = list(reversed(ep.block_vehicles))  # trigger order: block 1 first
        awake = [vs for road, vs in groups if road is None]  # traffic_mode "respawn": no trigger, awake from step 0
        groups = [(road, vs) for road, vs in groups if road is not None]
        if len(groups) > MAX_GROUPS:
            raise ValueError("more than %d traffic trigger groups" % MAX_GROUPS)
        trig = [-1] * MAX_GROUPS
        for vehicles in awake:
            for v in vehicles:
                self._slot(pgmap, mi, v.type, v.params, v.lane, v.long, 0.0, GROUP_AWAKE, v.overtake_timer, v.idm_seed,
                           v.checkpoints)
        for g, (road, vehicles) in enumerate(groups):
            trig[g] = mi.roads[tuple(road)]
            for v in vehicles:
                self._slot(pgmap, mi, v.type, v.params, v.lane, v.long, 0.0, g, v.overtake_timer, v.idm_seed,
                           v.checkpoints)
        for obj in getattr(ep, "objects", []):  # accident scenes last: traffic slots keep their numbers
            self._static(pgmap, mi, obj)
        self.episodes.append((map_id, ep.seed, slot_off, len(self.slots) - slot_off, len(groups), trig))
        self.seeds.append(ep.seed)
        return len(self.episodes) - 1

    def finish(self):
        def arr(rows, dt):
            return np.array(rows, dtype=dt) if rows else np.zeros(0, dtype=dt)

        out = dict(
            maps=arr(self.maps, MAP_DT), lanes=arr(self.lanes, LANE_DT), roads=arr(self.roads, ROAD_DT),
            boxes=arr(self.boxes, BOX_DT), cell_start=np.array(self.cell_start, dtype=np.int32),
            cell_entries=np.array(self.cell_entries, dtype=np.int32), episodes=arr(self.episodes, EPISODE_DT),
            slots=arr(self.slots, SLOT_DT), route_nodes=np.array(self.route_nodes, dtype=np.int32),
            route_roads=np.array(self.route_roads, dtype=np.int32)
        )
        out["max_slots"] = int(out["episodes"]["n_slots"].max()) if len(self.episodes) else 0
        return out


def build_tables(seeds, density=0.1, map_kwargs=None, generate=None):
    """Tables for ``seeds`` (one map + episode per seed).  ``generate(seed) -> PGMapData``."""
    from . import mapgen
    ts = TableSet()
    for s in seeds:
        pgmap = generate(s) if generate else mapgen.generate_map(s, **(map_kwargs or {}))
        mid = ts.add_map(pgmap)
        ts.add_episode(pgmap, mid, make_episode(pgmap, s, density))
    return ts
