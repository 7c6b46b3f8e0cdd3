"""Generate golden fixtures by running the UNMODIFIED reference (read-only /root/reference) under
tools/ref_stub.py.  Run only inside the build container:

    python tools/make_golden.py maps      # tests/golden/maps_*.json   (seed -> lanes / sockets / spawn lanes)
    python tools/make_golden.py reset     # tests/golden/reset_*.json  (seed -> ego params, route, traffic slots)

What the reference itself computes here (no restatement involved):
  * BIG block search + block classes -> road network  (pgdrive/component/algorithm/BIG.py, component/blocks/*)
  * save_map() block sequence                        (pgdrive/component/map/base_map.py:103-118)
  * ego / traffic RNG chain, _create_vehicles_once   (pgdrive/manager/traffic_manager.py:239-290)
  * Navigation.update -> checkpoints                 (pgdrive/component/vehicle_module/navigation.py:99-153)
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_stub  # noqa: E402

ref_stub.install()

from pgdrive.utils.config import Config  # noqa: E402
from pgdrive.engine.base_engine import BaseEngine  # noqa: E402
from pgdrive.component.map.pg_map import PGMap  # noqa: E402
from pgdrive.component.lane.straight_lane import StraightLane  # noqa: E402
from pgdrive.component.lane.circular_lane import CircularLane  # noqa: E402
from pgdrive.base_class.base_runnable import BaseRunnable  # noqa: E402
from pgdrive.utils.random_utils import get_np_random  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
MAP_CONFIG = dict(type="block_num", config=3, lane_width=3.5, lane_num=3, exit_length=50)


def make_engine(seed, density=0.1):
    eng = ref_stub.FakeEngine(
        seed,
        Config(
            dict(
                draw_map_resolution=1024,
                vehicle_config={},
                traffic_mode="trigger",
                random_traffic=False,
                traffic_density=density,
                start_seed=seed,
                environment_num=1
            )
        )
    )
    BaseEngine.singleton = eng
    return eng


_SHIPPED = None


def shipped_sequence(seed):
    """Block sequence of `seed` in the reference's shipped 30 000-map JSON (pgdrive_env.py:18-21)."""
    global _SHIPPED
    if _SHIPPED is None:
        path = os.path.join(
            ref_stub.REF_ROOT, "assets", "maps", "20210814_generated_maps_start_seed_0_environment_num_30000.json"
        )
        with open(path) as f:
            _SHIPPED = json.load(f)["map_data"]
    return _SHIPPED.get(str(seed))


def build_map(seed):
    """Live BIG first (pg_map.py:34-46), then rebuild through the reference's DEFAULT path
    (load_map_from_json=True -> _config_generate, pg_map.py:48-71, map_manager.py:84-91).  Socket
    side effects (InterSection.get_socket removes respawn roads) only happen for the FINAL socket
    choices on that path, so it -- not the search-time map -- is what the simulator drives on."""
    eng = make_engine(seed)
    cfg = dict(MAP_CONFIG)
    cfg["seed"] = seed
    big = PGMap(map_config=cfg, random_seed=None)
    saved = big.save_map()
    ship = shipped_sequence(seed)
    if ship is not None:
        a = json.loads(json.dumps(saved["block_sequence"]))
        assert a == ship["block_sequence"], ("live BIG != shipped JSON", seed)
    eng = make_engine(seed)
    cfg2 = dict(MAP_CONFIG)
    cfg2["type"] = "pg_map_file"
    cfg2["config"] = dict(seed=seed, block_sequence=json.loads(json.dumps(saved["block_sequence"])))
    m = PGMap(map_config=cfg2, random_seed=None)
    m.big_block_sequence = saved["block_sequence"]
    eng.current_map = m
    return eng, m


def lane_record(frm, to, idx, lane):
    rec = dict(
        frm=frm,
        to=to,
        idx=idx,
        width=float(lane.width),
        length=float(lane.length),
        line_types=[str(t) for t in lane.line_types],
        line_color=[[float(c) for c in col] for col in lane.line_color],
        speed_limit=float(lane.speed_limit),
        start=[float(lane.start[0]), float(lane.start[1])],
        end=[float(lane.end[0]), float(lane.end[1])],
    )
    if isinstance(lane, StraightLane):
        rec["kind"] = "S"
    elif isinstance(lane, CircularLane):
        rec["kind"] = "C"
        rec.update(
            center=[float(lane.center[0]), float(lane.center[1])],
            radius=float(lane.radius),
            start_phase=float(lane.start_phase),
            end_phase=float(lane.end_phase),
            direction=int(lane.direction)
        )
    else:
        raise ValueError(type(lane))
    return rec


def dump_map(seed):
    eng, m = build_map(seed)
    lanes = []
    for frm, td in m.road_network.graph.items():
        for to, ls in td.items():
            for i, l in enumerate(ls):
                lanes.append(lane_record(frm, to, i, l))
    blocks = []
    for b in m.blocks:
        socks = []
        for s in b.get_socket_list():
            socks.append(
                dict(
                    index=s.index,
                    pos=[s.positive_road.start_node, s.positive_road.end_node],
                    neg=[s.negative_road.start_node, s.negative_road.end_node]
                )
            )
        spawn = []
        if b.block_index != 0:
            for ls in b.get_intermediate_spawn_lanes():
                # lane.index is only assigned by _add_lane2bullet (Bullet); recover it from the graph
                spawn.append([find_index(m, l) for l in ls])
        blocks.append(
            dict(
                id=b.ID,
                name=b.name,
                sockets=socks,
                respawn_roads=[[r.start_node, r.end_node] for r in b.get_respawn_roads()],
                spawn_lanes=spawn,
                trigger_road=[b.pre_block_socket.positive_road.start_node, b.pre_block_socket.positive_road.end_node]
            )
        )
    return dict(seed=seed, block_sequence=m.big_block_sequence, lanes=lanes, blocks=blocks)


def find_index(m, lane):
    for frm, td in m.road_network.graph.items():
        for to, ls in td.items():
            for i, l in enumerate(ls):
                if l is lane:
                    return [frm, to, i]
    raise KeyError("lane not in graph")


def cmd_maps(seeds, tag):
    out = {}
    for s in seeds:
        out[str(s)] = dump_map(s)
    path = os.path.join(GOLD, "maps_%s.json.gz" % tag)
    with gzip.open(path, "wt") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path))


# ---------------------------------------------------------------------------------------------
class _ParamSampler(BaseRunnable):
    """BaseRunnable's own __init__ -> Randomizable(seed) -> sample_parameters() (base_runnable.py:19-29,81-88)."""
    pass


def sample_vehicle_params(vclass, seed):
    cls = type("S_" + vclass.__name__, (_ParamSampler, ), dict(PARAMETER_SPACE=vclass.PARAMETER_SPACE))
    obj = cls(random_seed=seed)
    return {k: float(v) for k, v in obj.get_config().items()}


def dump_reset(seed, density=0.1):
    from pgdrive.component.vehicle_module.navigation import Navigation
    from pgdrive.component.vehicle.vehicle_type import vehicle_type, DefaultVehicle
    from pgdrive.manager.traffic_manager import TrafficManager
    import pgdrive.policy.idm_policy as idm_mod

    eng, m = build_map(seed)
    eng.global_config["traffic_density"] = density
    # make every lane know its index, as _add_lane2bullet would have done (base_block.py:447)
    for frm, td in m.road_network.graph.items():
        for to, ls in td.items():
            for i, l in enumerate(ls):
                l.index = (frm, to, i)

    rec = dict(seed=seed, density=density)
    # --- ego (agent manager runs before traffic manager: PRIORITY tie, registration order) ---
    ego_seed = eng.generate_seed()
    rec["ego_seed"] = int(ego_seed)
    rec["ego_params"] = sample_vehicle_params(DefaultVehicle, ego_seed)
    nav = Navigation(eng)
    nav.update(m, current_lane_index=(">", ">>", 0), final_road_node=None, random_seed=seed)
    rec["ego_checkpoints"] = list(nav.checkpoints)

    # --- traffic: run the reference's own _create_vehicles_once against a recording engine ---
    tm = TrafficManager.__new__(TrafficManager)
    tm.engine = eng
    tm.random_seed = seed
    tm.np_random = get_np_random(seed)
    tm.spawned_objects = {}
    tm._traffic_vehicles = []
    tm.block_triggered_vehicles = []
    tm.mode = "trigger"
    tm.random_traffic = False
    tm.density = density
    vehicles = []

    class RecVehicle:
        def __init__(self, vclass, cfg, vseed):
            self.vclass = vclass
            self.cfg = dict(cfg)
            self.seed = vseed
            self.id = "v%d" % len(vehicles)
            self.idm_seed = None
            self.timer = None

    def spawn_object(vclass, vehicle_config=None, **kw):
        vseed = eng.generate_seed()  # base_engine.py:102-103
        v = RecVehicle(vclass, vehicle_config, vseed)
        vehicles.append(v)
        return v

    tm.spawn_object = spawn_object

    def add_policy(vid, policy):
        v = [x for x in vehicles if x.id == vid][0]
        v.idm_seed = int(policy.random_seed)
        v.timer = int(policy.overtake_timer)

    eng.add_policy = add_policy
    eng.object_manager = type("OM", (), dict(accident_lanes=[]))()
    type(eng).map_manager = property(lambda self: type("MM", (), dict(current_map=m))())
    if abs(density) >= 1e-2:
        tm._create_vehicles_once(m, density)
    name_of = {v: k for k, v in vehicle_type.items()}
    out_blocks = []
    for bv in tm.block_triggered_vehicles:  # already reversed: last element triggers first
        vs = []
        for v in bv.vehicles:
            nv = Navigation(eng)
            nv.update(m, current_lane_index=tuple(v.cfg["spawn_lane_index"]), final_road_node=None, random_seed=seed)
            vs.append(
                dict(
                    type=name_of[v.vclass],
                    lane=list(v.cfg["spawn_lane_index"]),
                    long=float(v.cfg["spawn_longitude"]),
                    seed=int(v.seed),
                    params=sample_vehicle_params(v.vclass, v.seed),
                    idm_seed=v.idm_seed,
                    overtake_timer=v.timer,
                    checkpoints=list(nv.checkpoints),
                )
            )
        out_blocks.append(dict(trigger_road=[bv.trigger_road.start_node, bv.trigger_road.end_node], vehicles=vs))
    rec["block_vehicles"] = out_blocks
    return rec


def cmd_reset(seeds, tag, density=0.1):
    out = {}
    for s in seeds:
        out[str(s)] = dump_reset(s, density)
    path = os.path.join(GOLD, "reset_%s.json.gz" % tag)
    with gzip.open(path, "wt") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "maps":
        cmd_maps(list(range(1000, 1100)), "v0_1000_1099")
        cmd_maps([0, 1, 2, 99, 1500, 1999, 2999, 12345, 29999], "misc")
    elif what == "reset":
        cmd_reset(list(range(1000, 1100)), "v0_1000_1099")
        cmd_reset([0, 1, 2, 99, 1500, 1999, 2999, 12345, 29999], "misc")
    elif what == "probe":
        print(json.dumps(dump_reset(int(sys.argv[2])), indent=1)[:6000])
