"""Per-seed episode template (host side): everything ``reset(force_seed=s)`` decides before the
first physics step -- ego parameters and route, traffic slots (type, lane, longitude, sampled
parameters, IDM seed / overtake timer, route) and the trigger road of every block.

The draw order over the reference's RNG streams is SURVEY.md appendix A; sources:
engine/base_engine.py:92-112,300-304 (object seeds), base_class/base_runnable.py:81-88 (parameters),
manager/traffic_manager.py:239-290,311-314 (traffic), component/vehicle_module/navigation.py:99-153
(destination + route), policy/idm_policy.py:180-188 (IDM seed, overtake timer).

Checked against tests/golden/reset_*.json.gz, produced by the unmodified reference.
"""
import math

from . import rng
from .roadnet import is_negative

VEHICLE_GAP = 10  # traffic_manager.py:30

# utils/space.py:219-255 -- bounds in the literal (positional) order of the reference: low, high
VEHICLE_SPACE = {
    "default": dict(wheel_friction=("c", 0.9), max_engine_force=("f", 850, 750), max_brake_force=("f", 180, 80),
                    max_steering=("c", 40), max_speed=("c", 80)),
    "s": dict(wheel_friction=("c", 0.9), max_engine_force=("f", 550, 350), max_brake_force=("f", 80, 35),
              max_steering=("c", 50), max_speed=("c", 80)),
    "m": dict(wheel_friction=("c", 0.75), max_engine_force=("f", 850, 650), max_brake_force=("f", 150, 60),
              max_steering=("c", 45), max_speed=("c", 80)),
    "l": dict(wheel_friction=("c", 0.8), max_engine_force=("f", 650, 450), max_brake_force=("f", 120, 60),
              max_steering=("c", 40), max_speed=("c", 80)),
    "xl": dict(wheel_friction=("c", 0.7), max_engine_force=("f", 700, 500), max_brake_force=("f", 100, 50),
               max_steering=("c", 35), max_speed=("c", 80)),
}
# component/vehicle/vehicle_type.py:7-78: LENGTH, WIDTH, HEIGHT, MASS, front / rear wheelbase, tyre radius, track/2
VEHICLE_BODY = {
    "default": (4.51, 1.852, 1.19, 1100.0, 1.05234, 1.4166, 0.313, 0.815),
    "xl": (5.8, 2.3, 2.8, 1600.0, 1.726, 1.075, 0.37, 0.831),
    "l": (4.5, 1.86, 1.85, 1300.0, 1.391, 1.10751, 0.39, 0.75),
    "m": (4.4, 1.85, 1.37, 1200.0, 1.285, 1.203, 0.39, 0.803),
    "s": (4.25, 1.7, 1.7, 800.0, 1.4126, 1.07, 0.376, 0.7),
}
TYPE_KEYS = ["s", "m", "l", "xl", "default"]
TYPE_PROB = [0.2, 0.3, 0.3, 0.2, 0]


def sample_vehicle(vtype, seed):
    """Randomizable(seed) -> sample_parameters()."""
    return rng.sample_space(VEHICLE_SPACE[vtype], rng.seeded(seed))


def route_for(pgmap, lane_index, seed, final_node=None):
    """Navigation.update + set_route: destination = random socket of the last block (first block for
    vehicles born on a negative road), route = first breadth-first path."""
    start = lane_index[0]
    if final_node is None:
        negative = is_negative((lane_index[0], lane_index[1]))
        block = pgmap.blocks[0] if negative else pgmap.blocks[-1]
        sockets = list(block.sockets.values())
        sock = sockets[int(rng.seeded(seed).choice(len(sockets)))]
        if len(sockets) > 1 and start in (sock.pos[0], sock.pos[1], sock.neg[0], sock.neg[1]):
            # navigation.py:114-121 loops forever / raises in this case; PG maps never reach it
            raise ValueError("Can not set a destination!")
        final_node = sock.neg[1] if negative else sock.pos[1]
    path = pgmap.net.shortest_path(start, final_node)
    if len(path) <= 2:
        path = [lane_index[0], lane_index[1]]
    return path


class VehicleSlot:
    __slots__ = ("type", "lane", "long", "seed", "params", "idm_seed", "overtake_timer", "checkpoints")


class StaticObject:
    """A traffic cone / warning tripod / barrier or a broken-down vehicle of an accident scene."""
    __slots__ = ("kind", "type", "lane", "long", "lat", "seed", "params")


# component/static_object/traffic_object.py:37-103: footprint (extent along the lane, across the lane) and mass.  Cones
# and tripods are cylinders (radius 0.25 / 0.5): their footprint is the enclosing square.  A barrier's LENGTH (2.0) lies
# ACROSS the lane: objects are oriented with setH(panda_heading(h)), vehicles with an extra -90 degrees
# (base_vehicle.py:681).
OBJECT_BODY = {"TrafficCone": (0.5, 0.5, 1.0), "TrafficWarning": (1.0, 1.0, 1.0), "TrafficBarrier": (0.3, 2.0, 10.0)}
ACCIDENT_BLOCKS = ("S", "C", "r", "R")  # object_manager.py:51-53: Straight, Curve, InRampOnStraight, OutRampOnStraight
ALERT_DIST, ACCIDENT_AREA_LEN, CONE_LONGITUDE, CONE_LATERAL, PROHIBIT_SCENE_PROB = 10, 10, 2, 1, 0.67


def make_accidents(pgmap, seed, accident_prob, engine_rs, traffic_rs):
    """TrafficObjectManager.reset (manager/object_manager.py:40-124) on its own stream of the seed; every spawned object
    draws its seed from the ENGINE's stream (before the ego: the manager's PRIORITY is 9) and a broken-down vehicle's
    type comes from the TRAFFIC manager's stream.  Returns (objects, accident_lanes)."""
    objects, accident_lanes = [], []
    if abs(accident_prob) < 1e-2:
        return objects, accident_lanes
    rs = rng.seeded(seed)
    lane_width = pgmap.lane_width

    def spawn(kind, lane, lon, lat, vtype=None):
        o = StaticObject()
        o.kind, o.type, o.lane, o.long, o.lat = kind, vtype, lane, float(lon), float(lat)
        o.seed = rng.draw_seed(engine_rs)
        o.params = sample_vehicle(vtype, o.seed) if vtype else None
        objects.append(o)

    for block in pgmap.blocks:
        if block.id not in ACCIDENT_BLOCKS:
            continue
        if rs.rand() > accident_prob:
            continue
        road_1 = (block.pre_socket.pos[1], block.node(0, 0))
        road_2 = (block.node(0, 0), block.node(0, 1)) if block.id != "S" else None
        is_ramp = block.id in ("r", "R")
        if rs.rand() > PROHIBIT_SCENE_PROB:  # a coned-off lane end
            road = (road_1, road_2)[int(rs.choice(2))] if block.id != "C" else road_2
            road = road_1 if road is None else road
            on_left = bool(rs.rand() > 0.5 or (road is road_2 and is_ramp))
            lanes = pgmap.net.lanes(road)
            idx = 0 if on_left else len(lanes) - 1
            lane = lanes[idx]
            longitude = lane.length - ACCIDENT_AREA_LEN
            accident_lanes += [(road[0], road[1], i) for i in range(len(lanes))]
            lat_num = int(lane_width / CONE_LATERAL)
            longitude_num = int(ACCIDENT_AREA_LEN / CONE_LONGITUDE)
            lats = [k * CONE_LATERAL for k in range(lat_num)] + [lat_num * CONE_LATERAL] * (longitude_num + 1) + \
                [(lat_num - k - 1) * CONE_LATERAL for k in range(lat_num)]
            total = lat_num * 2 + longitude_num + 1
            left = 1 if on_left else -1
            for k, lat in zip(range(-int(total / 2), int(total / 2)), lats):
                spawn("TrafficCone", (road[0], road[1], idx), k * CONE_LONGITUDE + longitude, left * (lat - lane.width / 2))
        else:  # a broken-down vehicle with its warning tripod, or a barrier
            road = (road_1, road_2)[int(rs.choice(2))]
            road = road_1 if road is None else road
            on_left = bool(rs.rand() > 0.5 or (road is road_2 and is_ramp))
            lanes = pgmap.net.lanes(road)
            idx = int(rs.randint(0, len(lanes) - 1)) if on_left else len(lanes) - 1
            lane = lanes[idx]
            longitude = rs.rand() * lane.length / 2 + lane.length / 2
            if rs.rand() > 0.5:
                vtype = TYPE_KEYS[int(traffic_rs.choice(len(TYPE_KEYS), p=TYPE_PROB))]
                spawn("vehicle", (road[0], road[1], idx), longitude, 0.0, vtype)
                spawn("TrafficWarning", (road[0], road[1], idx), longitude - ALERT_DIST, 0.0)
            else:
                spawn("TrafficBarrier", (road[0], road[1], idx), longitude, 0.0)
    return objects, accident_lanes



class EpisodeTemplate:
    def __init__(self, seed, density):
        self.seed = seed
        self.density = density
        self.ego_seed = None
        self.ego_type = "default"
        self.ego_params = None
        self.ego_checkpoints = None
        self.block_vehicles = []  # [(trigger_road, [VehicleSlot])], LAST element triggers first
        self.objects = []         # [StaticObject]: accident scenes (SafePGDriveEnv)


def respawn_lanes(pgmap):
    """TrafficManager._get_available_respawn_lanes (traffic_manager.py:292-309): a road listed by two blocks drops out."""
    roads = []
    for block in pgmap.blocks:
        for road in block.respawn:
            if road in roads:
                roads.remove(road)
            else:
                roads.append(road)
    lanes = []
    for road in roads:
        lanes += [(road, i, ln) for i, ln in enumerate(pgmap.net.lanes(road))]
    return lanes


def make_episode(pgmap, seed, density=0.1, spawn_lane=(">", ">>", 0), random_agent_model=False, traffic_mode="trigger",
                 accident_prob=0.0, traffic_rs=None):
    """``traffic_mode`` (traffic_manager.py:21-27): "trigger" and "hybrid" create every block's vehicles once and wake
    them when the ego reaches the block (in this version of the reference the two are the same code path, :63-69,
    :76-85); "respawn" fills every respawn lane with one vehicle per 10 m -- the density only switches traffic on --
    and all of them drive from the first step (:224-237; the re-spawn itself is commented out at :100-105).

    ``traffic_rs``: the traffic manager's random stream.  None = a fresh stream of ``seed`` (the reference re-seeds every
    manager at reset, base_engine.py:300-305); with ``random_traffic`` the traffic manager skips that re-seeding
    (traffic_manager.py:348-350), i.e. the caller passes ONE generator that lives across resets."""
    if traffic_mode not in ("trigger", "hybrid", "respawn"):
        raise ValueError("No such mode named {}".format(traffic_mode))
    engine_rs = rng.seeded(seed)
    if traffic_rs is None:
        traffic_rs = rng.seeded(seed)
    ep = EpisodeTemplate(seed, density)
    # random_agent_model (manager/agent_manager.py:63-71, component/vehicle/vehicle_type.py:84-86): the agent manager's
    # own stream of the seed picks one of the five types with equal probability
    ep.ego_type = "default"
    if random_agent_model:
        ep.ego_type = TYPE_KEYS[int(rng.seeded(seed).choice(len(TYPE_KEYS), p=[1 / len(TYPE_KEYS)] * len(TYPE_KEYS)))]
    ep.objects, accident_lanes = make_accidents(pgmap, seed, accident_prob, engine_rs, traffic_rs)
    ep.ego_seed = rng.draw_seed(engine_rs)
    ep.ego_params = sample_vehicle(ep.ego_type, ep.ego_seed)
    ep.ego_checkpoints = route_for(pgmap, spawn_lane, seed)
    if abs(density) < 1e-2:
        return ep
    lane_index = {}
    for (frm, to), lanes in pgmap.net.roads():
        for i, ln in enumerate(lanes):
            lane_index[id(ln)] = (frm, to, i)

    def new_slot(lane, lon):
        v = VehicleSlot()
        v.type = TYPE_KEYS[int(traffic_rs.choice(len(TYPE_KEYS), p=TYPE_PROB))]
        v.lane, v.long = lane, float(lon)
        v.seed = rng.draw_seed(engine_rs)
        v.params = sample_vehicle(v.type, v.seed)
        v.checkpoints = route_for(pgmap, lane, seed)
        v.idm_seed = rng.draw_seed(traffic_rs)
        v.overtake_timer = int(rng.seeded(v.idm_seed).randint(0, 50))
        return v

    if traffic_mode == "respawn":  # _create_respawn_vehicles -> _create_vehicles_on_lane (traffic_manager.py:188-237)
        slots = []
        for road, i, ln in respawn_lanes(pgmap):
            longs = [k * VEHICLE_GAP for k in range(int(ln.length / VEHICLE_GAP))]
            traffic_rs.shuffle(longs)
            slots += [new_slot((road[0], road[1], i), lon) for lon in longs]
        ep.block_vehicles.append((None, slots))  # no trigger road: awake from the first step
        return ep
    for block in pgmap.blocks[1:]:
        spawn = block.spawn_lanes()
        cand = []
        for lanes in spawn:
            for ln in lanes:
                if lane_index[id(ln)] in accident_lanes:  # traffic_manager.py:251-252
                    continue
                for k in range(int(ln.length / VEHICLE_GAP)):
                    cand.append((lane_index[id(ln)], k * VEHICLE_GAP))
        total_length = sum(ln.length for lanes in spawn for ln in lanes)
        total = int(math.floor(int(math.floor(total_length / VEHICLE_GAP)) * density))
        traffic_rs.shuffle(cand)
        slots = [new_slot(lane, lon) for lane, lon in cand[:min(total, len(cand))]]
        ep.block_vehicles.append((block.pre_socket.pos, slots))
    ep.block_vehicles.reverse()
    return ep
