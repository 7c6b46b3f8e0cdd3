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


def dump_reset(seed, density=0.1, mode="trigger", accident_prob=0.0):
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
    tm = TrafficManager.__new__(TrafficManager)  # created early: the object manager borrows its random_vehicle_type
    tm.engine = eng
    tm.random_seed = seed
    tm.np_random = get_np_random(seed)
    accident_lanes = []
    if accident_prob > 0:
        # --- accident scenes (manager/object_manager.py:40-124; PRIORITY 9: its reset runs before the agent manager's) ---
        from pgdrive.manager.object_manager import TrafficObjectManager
        om = TrafficObjectManager.__new__(TrafficObjectManager)
        om.engine = eng
        om.random_seed = seed
        om.np_random = get_np_random(seed)
        om.accident_prob = accident_prob
        om.accident_lanes = []
        om.spawned_objects = {}
        objs = []

        def om_spawn(cls, **kw):
            oseed = eng.generate_seed()  # base_engine.py:102-103
            name = cls.__name__
            if "vehicle_config" in kw:
                vc = kw["vehicle_config"]
                r = dict(kind="vehicle", type={v: k for k, v in vehicle_type.items()}[cls], lane=list(vc["spawn_lane_index"]),
                         long=float(vc["spawn_longitude"]), lat=0.0, seed=int(oseed),
                         params=sample_vehicle_params(cls, oseed))
            else:
                r = dict(kind=name, lane=list(kw["lane"].index), long=float(kw["longitude"]), lat=float(kw["lateral"]),
                         seed=int(oseed))
            objs.append(r)
            return type("Obj", (), dict(set_break_down=lambda self, *a: None))()

        om.spawn_object = om_spawn
        eng.traffic_manager = tm
        om.reset()
        rec["accident_prob"] = accident_prob
        rec["objects"] = objs
        rec["accident_lanes"] = [list(l.index) for l in om.accident_lanes]
        accident_lanes = om.accident_lanes
    # --- ego (agent manager runs before traffic manager: PRIORITY tie, registration order) ---
    ego_seed = eng.generate_seed()
    rec["ego_seed"] = int(ego_seed)
    rec["ego_params"] = sample_vehicle_params(DefaultVehicle, ego_seed)
    nav = Navigation(eng)
    nav.update(m, current_lane_index=(">", ">>", 0), final_road_node=None, random_seed=seed)
    rec["ego_checkpoints"] = list(nav.checkpoints)

    # --- traffic: run the reference's own _create_vehicles_once against a recording engine ---
    tm.spawned_objects = {}
    tm._traffic_vehicles = []
    tm.block_triggered_vehicles = []
    tm.mode = mode
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
    eng.object_manager = type("OM", (), dict(accident_lanes=accident_lanes))()
    type(eng).map_manager = property(lambda self: type("MM", (), dict(current_map=m))())
    if abs(density) >= 1e-2:
        if mode == "respawn":  # traffic_manager.py:63-66: every respawn lane is filled, all vehicles act from step 0
            tm.respawn_lanes = tm._get_available_respawn_lanes(m)
            tm._create_respawn_vehicles(m, density)
        else:
            tm._create_vehicles_once(m, density)
    name_of = {v: k for k, v in vehicle_type.items()}
    out_blocks = []
    groups = list(tm.block_triggered_vehicles)  # already reversed: last element triggers first
    if mode == "respawn":
        groups = [type("BV", (), dict(trigger_road=None, vehicles=list(tm._traffic_vehicles)))()]
    for bv in groups:
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
        out_blocks.append(dict(trigger_road=None if bv.trigger_road is None else
                               [bv.trigger_road.start_node, bv.trigger_road.end_node], vehicles=vs))
    rec["block_vehicles"] = out_blocks
    return rec


def cmd_reset(seeds, tag, density=0.1, mode="trigger", accident_prob=0.0):
    out = {}
    for s in seeds:
        out[str(s)] = dump_reset(s, density, mode, accident_prob)
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
    elif what == "reset_respawn":  # traffic_mode="respawn" (traffic_manager.py:21-27,63-66,224-237)
        cmd_reset(list(range(1000, 1030)) + [0, 1, 2, 99], "respawn", mode="respawn")
    elif what == "reset_accidents":  # SafePGDriveEnv: accident_prob 0.8, traffic_density 0.05 (safe_pgdrive_env.py:8-25)
        cmd_reset(list(range(100, 140)) + [0, 1, 2, 1000, 1001, 1002, 1003, 1005], "accidents", density=0.05,
                  accident_prob=0.8)
    elif what == "probe":
        print(json.dumps(dump_reset(int(sys.argv[2])), indent=1)[:6000])


# =================================================================================================
# step-level vectors: the reference's OWN Python (IDM + PID, navigation info, checkpoint update, state
# observation, neighbour features, reward, arrive_destination) evaluated on simulator states.
# States come from a roll-out of this repo's CPU oracle (they are just inputs); every expected value
# in the fixture is computed by unmodified reference code running under tools/ref_stub.py.
# =================================================================================================
def cmd_step(seeds, tag, steps=260, every=4, density=0.1, only_interesting=False, max_records=100000):
    import base64
    from collections import deque
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle.oracle import Oracle
    from pgdrive_b200 import cabi, tables as ptables, mapgen, episode as pepisode
    from pgdrive.component.vehicle.base_vehicle import BaseVehicle
    from pgdrive.component.vehicle_module.navigation import Navigation
    from pgdrive.component.vehicle_module.lidar import Lidar
    from pgdrive.component.road.road import Road
    from pgdrive.obs.state_obs import StateObservation
    from pgdrive.policy.idm_policy import IDMPolicy
    from pgdrive.policy.base_policy import BasePolicy
    from pgdrive.envs.pgdrive_env import PGDriveEnv
    from pgdrive.utils.math_utils import Vector
    import pgdrive.policy.idm_policy as idm_mod

    class FakeVehicle:
        """Exactly the attributes the reference's pure-Python step code reads from a BaseVehicle."""
        MAX_STEERING = 60
        max_speed = 80
        heading_diff = BaseVehicle.heading_diff
        projection = BaseVehicle.projection
        _dist_to_route_left_right = BaseVehicle._dist_to_route_left_right
        arrive_destination = BaseVehicle.arrive_destination

        def __init__(self, st, slot_rec, ref_lanes, world):
            self.position = np.array([float(st["x"]), float(st["y"])])
            self.heading_theta = float(st["heading"])
            self.heading = Vector((math.cos(self.heading_theta), math.sin(self.heading_theta)))
            self.speed = float(np.clip(st["speed"] * 3.6, 0.0, 100000.0))
            self.velocity = self.speed * np.asarray([math.cos(self.heading_theta), math.sin(self.heading_theta)])
            self.lane = ref_lanes[int(st["lane"])]
            self.lane_index = self.lane.index
            self.LENGTH, self.WIDTH = float(slot_rec["length"]), float(slot_rec["width"])
            self.world = world
            self.lidar = self
            self.engine = None

        @property
        def current_road(self):
            return Road(*self.lane_index[0:-1])

        def get_surrounding_objects(self, vehicle):  # Lidar.get_surrounding_objects stand-in (50 m, centre distance)
            return [o for o in self.world if o is not vehicle and
                    (o.position[0] - vehicle.position[0])**2 + (o.position[1] - vehicle.position[1])**2 < 2500.0]

    import math
    out = []
    for seed in seeds:
        eng, m = build_map(seed)
        eng.global_config["traffic_density"] = density
        pgmap = mapgen.generate_map(seed)
        ts = ptables.TableSet()
        mid = ts.add_map(pgmap)
        ts.add_episode(pgmap, mid, pepisode.make_episode(pgmap, seed, density))
        T = ts.finish()
        mi = ts.index[0]
        ref_lanes = []
        for (frm, to, first, n) in mi.road_list:
            for i in range(n):
                ln = m.road_network.graph[frm][to][i]
                ln.index = (frm, to, i)
                ref_lanes.append(ln)
        node_name = {v: k for k, v in mi.nodes.items()}
        slots = T["slots"]
        n_slots = int(T["episodes"][0]["n_slots"])
        if n_slots > 32:
            print("seed", seed, "needs", n_slots, "vehicle slots: skipped")
            continue
        orc = Oracle(T, 1, auto_reset=False, num_slots=32)
        orc.reset([0], [0])
        rs = np.random.RandomState(seed)
        policies = {}

        def make_nav(st, slot):
            nav = Navigation.__new__(Navigation)
            nav.map = m
            off, ln = int(slots[slot]["route_off"]), int(slots[slot]["route_len"])
            nav.checkpoints = [node_name[int(x)] for x in T["route_nodes"][off:off + ln]]
            nav._target_checkpoints_index = [int(st["ck0"]), int(st["ck1"])]
            c = nav.checkpoints
            i0, i1 = nav._target_checkpoints_index
            nav.current_ref_lanes = m.road_network.graph[c[i0]][c[i0 + 1]]
            nav.current_road = Road(c[i0], c[i0 + 1])
            if i0 == i1:
                nav.next_ref_lanes, nav.next_road = None, None
            else:
                nav.next_ref_lanes, nav.next_road = m.road_network.graph[c[i1]][c[i1 + 1]], Road(c[i1], c[i1 + 1])
            nav.final_road = Road(c[-2], c[-1])
            nav.final_lane = nav.final_road.get_lanes(m.road_network)[-1]
            nav._navi_info = np.zeros((10, ))
            nav._show_navi_info = False
            return nav

        lat_err = 0.0
        for t in range(steps):
            s0 = orc.get_state(0)
            v0 = s0["veh"][0]
            # a lane-following ego (keeps the episode alive so that traffic wakes up); any policy would do
            ego_lane = ref_lanes[int(v0[0]["lane"])]
            lon, lat = ego_lane.local_coordinates((float(v0[0]["x"]), float(v0[0]["y"])))
            herr = ((ego_lane.heading_at(lon + 2.0) - float(v0[0]["heading"]) + np.pi) % (2 * np.pi)) - np.pi
            a = np.array([[np.clip(-1.5 * herr + 0.25 * lat + rs.uniform(-0.03, 0.03), -1, 1),
                           np.clip(0.6 - float(v0[0]["speed"]) / 15.0 + rs.uniform(-0.1, 0.1), -1, 1)]], np.float32)
            obs, rew, done, info = orc.step(a)
            s1 = orc.get_state(0)
            v1 = s1["veh"][0]
            if t % every == 0 and t > 0:
                rec = dict(seed=seed, t=t, density=density, s0=base64.b64encode(s0.tobytes()).decode(),
                           action=[float(a[0, 0]), float(a[0, 1])])
                # ---- world before the step (IDM inputs) ----
                world0 = {}
                for i in range(n_slots):
                    if int(v0[i]["flags"]) & cabi.V_ALIVE:
                        world0[i] = FakeVehicle(v0[i], slots[i], ref_lanes, None)
                for fv in world0.values():
                    fv.world = list(world0.values())
                idm = []
                for i in range(1, n_slots):
                    ran = (int(v1[i]["flags"]) & cabi.V_ACTIVE) and (int(v0[i]["flags"]) & cabi.V_ALIVE)
                    if not ran:
                        continue
                    fv = world0[i]
                    fv.navigation = make_nav(v0[i], i)
                    pol = IDMPolicy.__new__(IDMPolicy)
                    BasePolicy.__init__(pol, control_object=fv, random_seed=0)
                    # the policy's own stream, advanced to where this vehicle's stream stands
                    from pgdrive.utils.random_utils import get_np_random
                    idm_seed = None
                    pol.np_random = None
                    pol.target_speed = float(v0[i]["target_speed"])
                    pol.routing_target_lane = None if int(v0[i]["rt_lane"]) < 0 else ref_lanes[int(v0[i]["rt_lane"])]
                    pol.available_routing_index_range = None
                    pol.overtake_timer = int(v0[i]["timer"])
                    from pgdrive.component.vehicle_module.PID_controller import PIDController
                    pol.heading_pid = PIDController(1.7, 0.01, 3.5)
                    pol.lateral_pid = PIDController(0.3, .002, 0.05)
                    pol.heading_pid.p_error, pol.heading_pid.i_error = float(v0[i]["pid_hp"]), float(v0[i]["pid_hi"])
                    pol.lateral_pid.p_error, pol.lateral_pid.i_error = float(v0[i]["pid_lp"]), float(v0[i]["pid_li"])

                    class _Stream:  # replays the tabulated randint(0, 25) draws of this vehicle's IDM stream
                        def __init__(self, draws, n):
                            self.draws, self.n = draws, n

                        def randint(self, lo, hi):
                            assert (lo, hi) == (0, 25)
                            v = int(self.draws[self.n % len(self.draws)])
                            self.n += 1
                            return v

                    pol.np_random = _Stream(slots[i]["rnd25"], int(v0[i]["rnd_n"]))
                    # A decision that flips when every other vehicle is nudged by 0.3 mm sits on one of IDM's exact
                    # thresholds (traffic spawns on a 10 m grid and MAX_LONG_DIST is 30 m, SAFE_LANE_CHANGE_DISTANCE
                    # 15 m): the reference's own outcome then depends on float32 position round-off inside Bullet.
                    # Such samples are marked as ties and not used as known answers.
                    import copy as _copy
                    outcomes = []
                    for nudge in (0.0, 3e-4, -3e-4):
                        trial = _copy.copy(pol)
                        trial.heading_pid, trial.lateral_pid = _copy.copy(pol.heading_pid), _copy.copy(pol.lateral_pid)
                        trial.np_random = _Stream(slots[i]["rnd25"], int(v0[i]["rnd_n"]))
                        saved_pos = {}
                        for j, other in world0.items():
                            if j != i:
                                saved_pos[j] = other.position
                                other.position = other.position + nudge * np.asarray([other.heading[0], other.heading[1]])
                        st_, acc_ = trial.act()
                        for j, ppos in saved_pos.items():
                            world0[j].position = ppos
                        outcomes.append((float(st_), float(acc_), float(trial.target_speed), int(trial.overtake_timer),
                                         ref_lanes.index(trial.routing_target_lane), trial))
                    steering, acc = outcomes[0][0], outcomes[0][1]
                    pol = outcomes[0][5]
                    tie = any(abs(o_[0] - steering) > 2e-3 or abs(o_[1] - acc) > 2e-3 or o_[2:5] != outcomes[0][2:5]
                              for o_ in outcomes[1:])
                    idm.append(dict(
                        slot=i, steering=float(steering), acc=float(acc), target_speed=float(pol.target_speed),
                        timer=int(pol.overtake_timer), rt_lane=ref_lanes.index(pol.routing_target_lane),
                        pid=[float(pol.heading_pid.p_error), float(pol.heading_pid.i_error),
                             float(pol.lateral_pid.p_error), float(pol.lateral_pid.i_error)], tie=bool(tie)
                    ))
                rec["idm"] = idm
                if only_interesting:
                    # keep the step only if some vehicle crept, braked, finished a lane change (timer re-drawn) or
                    # moved its routing lane: the branches a sparse roll-out rarely reaches
                    keep = False
                    for g in idm:
                        v = v0[g["slot"]]
                        keep |= g["target_speed"] == 5.0 or g["acc"] < 0 or g["timer"] < int(v["timer"]) or \
                            (int(v["rt_lane"]) >= 0 and g["rt_lane"] != int(v["rt_lane"]))
                    if not keep:
                        continue
                # ---- world after the step (observation / reward inputs) ----
                world1 = {}
                for i in range(n_slots):
                    if int(v1[i]["flags"]) & cabi.V_ALIVE:
                        world1[i] = FakeVehicle(v1[i], slots[i], ref_lanes, None)
                for fv in world1.values():
                    fv.world = list(world1.values())
                ego = world1[0]
                ego.navigation = make_nav(v1[0], 0)
                # checkpoint advance: reference rule applied to the pre-step indices and the post-step lane
                nav0 = make_nav(v0[0], 0)
                lon1, _ = ego.lane.local_coordinates(ego.position)
                nav0._update_target_checkpoints(ego.lane_index, lon1)
                rec["ck"] = [int(x) for x in nav0._target_checkpoints_index]
                navi = ego.navigation._get_info_for_checkpoint(0, ego.navigation.current_ref_lanes, ego)[0]
                c = ego.navigation.checkpoints
                i1 = ego.navigation._target_checkpoints_index[1]
                navi += ego.navigation._get_info_for_checkpoint(1, m.road_network.graph[c[i1]][c[i1 + 1]], ego)[0]
                rec["navi"] = [float(x) for x in navi]
                ego.dist_to_left_side, ego.dist_to_right_side = ego._dist_to_route_left_right()
                ego.steering = float(v1[0]["steer"])
                ego.throttle_brake = float(v1[0]["throttle"])
                ego.last_current_action = deque([(float(s1["prev_steer"][0]), float(s1["prev_throttle"][0])),
                                                 (ego.steering, ego.throttle_brake)], maxlen=2)
                ego.last_position = np.array([float(v0[0]["x"]), float(v0[0]["y"])])
                h0 = float(v0[0]["heading"])
                ego.last_heading_dir = Vector((math.cos(h0), math.sin(h0)))

                class _Off:
                    available = False

                ego.side_detector = ego.lane_line_detector = _Off()
                ego.navigation.map = type("M", (), dict(MAX_LANE_NUM=3, MAX_LANE_WIDTH=4.5, _config=m._config,
                                                        LANE_WIDTH="lane_width", road_network=m.road_network))()
                so = StateObservation.__new__(StateObservation)
                so.config = dict(random_agent_model=False)
                rec["state"] = [float(x) for x in so.vehicle_state(ego)]
                lid = type("L", (), dict(perceive_distance=50, get_surrounding_vehicles=staticmethod(lambda objs: set(objs))))()
                rec["neighbours"] = [float(x) for x in Lidar.get_surrounding_vehicles_info(
                    lid, ego, ego.get_surrounding_objects(ego), 4)]
                ego.on_yellow_continuous_line = ego.on_white_continuous_line = ego.crash_sidewalk = False
                ego.on_lane = True
                ego.crash_vehicle = ego.crash_object = ego.out_of_route = False
                fenv = type("E", (), dict(vehicles={"default_agent": ego}, config=dict(
                    use_lateral=False, driving_reward=1.0, speed_reward=0.1, success_reward=10.0,
                    out_of_road_penalty=5.0, crash_vehicle_penalty=5.0, crash_object_penalty=5.0,
                    out_of_route_done=False), _is_out_of_road=lambda self, v: False))()
                r, rinfo = PGDriveEnv.reward_function(fenv, "default_agent")
                rec["step_reward"] = float(rinfo["step_reward"])
                rec["arrive_dest"] = bool(ego.arrive_destination.fget(ego) if isinstance(
                    ego.arrive_destination, property) else FakeVehicle.arrive_destination.fget(ego))
                out.append(rec)
            if done[0] or len(out) >= max_records:
                break
        orc.close()
        print("seed", seed, "steps", t + 1, "records so far", len(out), "idm samples", sum(len(r["idm"]) for r in out))
    path = os.path.join(GOLD, "step_%s.json.gz" % tag)
    with gzip.open(path, "wt") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__" and sys.argv[1] == "step":
    cmd_step([1000, 1003, 1008, 1015, 1021, 1042, 1055, 1077], "v0")
if __name__ == "__main__" and sys.argv[1] == "step_dense":
    # ramps (lane count drops -> forced lane changes / creeping) and 3x the default traffic, every step sampled,
    # only the steps that exercise IDM's rarer branches kept
    cmd_step([1000, 1002, 1006, 1011, 1013, 1016, 1017, 1024, 1028, 1033, 1036, 1039], "dense", steps=400, every=1,
             density=0.3, only_interesting=True, max_records=100000)


# =================================================================================================
# lidar vectors: the reference's own per-beam loop (the Python twin of cutils_perceive that ships in
# pgdrive/utils/cutils.py:36-97) driven by a 2-D stand-in for Bullet's rayTestClosest that intersects the
# beam with the four edges of every chassis rectangle analytically (segment / segment, the construction
# the reference's own mask test uses, tests/test_component/test_detector_mask.py:132-154).
# =================================================================================================
def cmd_lidar(tag, n_scenes=40):
    import base64
    import math
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle.oracle import Oracle
    from pgdrive_b200 import cabi, tables as ptables
    from pgdrive.utils.cutils import _get_fake_cutils
    fake = _get_fake_cutils()

    class Hit:
        def __init__(self, node, frac, pos):
            self.node, self.frac, self.pos = node, frac, pos

        def getNode(self):
            return self.node

        def getHitFraction(self):
            return self.frac

        def hasHit(self):
            return self.node is not None

        def getHitPos(self):
            return self.pos

    class World2D:
        """rayTestClosest over chassis rectangles given in PGDrive coordinates; arguments arrive in Panda
        coordinates (x, -y, z) exactly as cutils_perceive passes them."""
        def __init__(self, boxes):
            self.edges = []
            self.rects = [(x, y, math.cos(h), math.sin(h), length / 2, width / 2) for x, y, h, length, width in boxes]
            for k, (x, y, h, length, width) in enumerate(boxes):
                c, s = math.cos(h), math.sin(h)
                pts = [(x + c * a * length / 2 - s * b * width / 2, y + s * a * length / 2 + c * b * width / 2)
                       for a, b in ((1, 1), (1, -1), (-1, -1), (-1, 1))]
                for i in range(4):
                    self.edges.append((k, pts[i], pts[(i + 1) % 4]))

        def rayTestClosest(self, start, end, mask):
            p = (start[0], -start[1])
            r = (end[0] - start[0], -(end[1] - start[1]))
            best, node = 1.0, None
            # Bullet's convex ray cast reports no hit for a shape that contains the ray origin
            inside = {k for k, (x, y, c, s, hl, hw) in enumerate(self.rects)
                      if abs((p[0] - x) * c + (p[1] - y) * s) <= hl and abs(-(p[0] - x) * s + (p[1] - y) * c) <= hw}
            for k, a, b in self.edges:
                if k in inside:
                    continue
                sx, sy = b[0] - a[0], b[1] - a[1]
                den = r[0] * sy - r[1] * sx
                if den == 0:
                    continue
                qx, qy = a[0] - p[0], a[1] - p[1]
                t = (qx * sy - qy * sx) / den
                u = (qx * r[1] - qy * r[0]) / den
                if 0.0 <= t <= 1.0 and 0.0 <= u <= 1.0 and t < best:
                    best, node = t, k
            return Hit(node, best, (p[0] + best * r[0], -(p[1] + best * r[1]), start[2]))

    T = ptables.build_tables([1003]).finish()
    n_slots = int(T["episodes"][0]["n_slots"])
    slots = T["slots"]
    orc = Oracle(T, 1, auto_reset=False)
    rs = np.random.RandomState(7)
    lidar_range = np.arange(0, 240) * (2 * np.pi / 240)
    out = []
    for scene in range(n_scenes):
        orc.reset([0], [0])
        s = orc.get_state(0)
        v = s["veh"][0]
        ex, ey = float(v[0]["x"]) + rs.uniform(-3, 3), float(v[0]["y"]) + rs.uniform(-1, 1)
        eh = rs.uniform(-np.pi, np.pi)
        v[0]["x"], v[0]["y"], v[0]["heading"] = ex, ey, eh
        boxes = []
        for i in range(1, n_slots):
            d, ang = rs.uniform(4.5, 58.0), rs.uniform(-np.pi, np.pi)
            v[i]["x"], v[i]["y"] = ex + d * math.cos(ang), ey + d * math.sin(ang)
            v[i]["heading"] = rs.uniform(-np.pi, np.pi)
            alive = rs.rand() > 0.15
            v[i]["flags"] = (cabi.V_ALIVE | cabi.V_ON_LANE) if alive else 0
            if alive:  # float32 poses, as the simulator stores them
                boxes.append((float(v[i]["x"]), float(v[i]["y"]), float(v[i]["heading"]), float(slots[i]["length"]),
                              float(slots[i]["width"])))
        s["veh"][0] = v
        cloud = np.ones(240)
        cloud, _, _ = fake.cutils_perceive(
            cloud, None, None, lidar_range, 50.0, float(v[0]["heading"]), float(v[0]["x"]), float(v[0]["y"]), 240, 1.2,
            World2D(boxes), set(), False, True, 0, 0, 0)
        out.append(dict(state=base64.b64encode(s.tobytes()).decode(), cloud=[float(x) for x in cloud],
                        hits=int((cloud < 1.0).sum())))
    orc.close()
    path = os.path.join(GOLD, "lidar_%s.json.gz" % tag)
    with gzip.open(path, "wt") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "beams hitting:", sum(r["hits"] for r in out))


if __name__ == "__main__" and sys.argv[1] == "lidar":
    cmd_lidar("v0")


def cmd_digests(seeds, tag):
    """Compact fixtures for many seeds: one sha1 per seed over the reference's lanes / sockets and one over its
    reset decisions (tools/golden_hash.py defines what is hashed)."""
    import golden_hash
    out = {}
    for s in seeds:
        d = dump_map(s)
        lanes = []
        for l in d["lanes"]:
            l = dict(l)
            l["colours"] = ["Y" if abs(c[0] - 1) > 1e-6 else "G" for c in l.pop("line_color")]
            lanes.append(l)
        blocks = [[b["id"], [[x["index"], x["pos"], x["neg"]] for x in b["sockets"]], b["respawn_roads"],
                   b["trigger_road"], b["spawn_lanes"]] for b in d["blocks"]]
        r = dump_reset(s)
        rec = dict(ego_seed=r["ego_seed"], ego_params=r["ego_params"], ego_checkpoints=r["ego_checkpoints"],
                   block_vehicles=[(g["trigger_road"], g["vehicles"]) for g in r["block_vehicles"]])
        out[str(s)] = [golden_hash.map_digest(lanes, blocks), golden_hash.episode_digest(rec),
                       "".join(b["id"] for b in d["block_sequence"])]
        if s % 100 == 0:
            print(s, out[str(s)])
    path = os.path.join(GOLD, "digests_%s.json.gz" % tag)
    with gzip.open(path, "wt") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__" and sys.argv[1] == "digests":
    cmd_digests(list(range(1000, 2000)), "1000_1999")


# =================================================================================================
# side / lane-line detector vectors: the reference's beam loop (cutils.py twin) with the SideDetector /
# LaneLineDetector conventions (phase offset 90 deg, distance_detector.py:137-152) over an analytic 2-D world made
# of the line-ghost rectangles that the reference's own block code (component/blocks/base_block.py) builds for the
# map.  The rectangles are captured from the UNMODIFIED _add_lane_line2bullet / _add_box_body by recording the
# arguments they pass to the (stubbed) Bullet shape / node calls.
# =================================================================================================
def capture_line_boxes(seed, all_kinds=False):
    """Run the reference's own primitive builders (BaseBlock._add_pgdrive_lanes / _add_lane_surface, unmodified) for
    every block of the map with recording stand-ins for the Bullet shape / node classes, and return each primitive
    as (cx, cy, theta, half_length, half_width, name) in PGDrive coordinates."""
    import math
    from pgdrive.component.blocks import base_block as bb
    from pgdrive.constants import BodyName
    rec = []
    line_names = (BodyName.White_continuous_line, BodyName.Yellow_continuous_line, BodyName.Broken_line)

    class Shape:
        def __init__(self, half):
            self.half = half

    class Node:
        def __init__(self, name=None, *a, **k):
            self.name = a[0] if a else name  # BaseRigidBodyNode(lane, BodyName.Lane)
            self.shape = None

        def addShape(self, shape, *a):
            self.shape = shape

        def __getattr__(self, k):
            return lambda *a, **kw: None

    class NP:
        def __init__(self, node=None):
            self._node = node
            self.pos = None
            self.theta = None

        def node(self):
            return self._node

        def attachNewNode(self, node):
            return NP(node if isinstance(node, Node) else Node(node))

        def setPos(self, p):
            self.pos = p

        def setQuat(self, q):
            node = self._node
            self.theta = -2 * math.atan2(q[3], q[0])
            if isinstance(node, Node) and node.shape is not None:
                if node.name in line_names:
                    rec.append((self.pos[0], -self.pos[1], self.theta, node.shape.half[0], node.shape.half[1], node.name))
                elif node.name == BodyName.Lane and all_kinds:  # BulletBoxShape(length / 2, 0.1, width / 2)
                    rec.append((self.pos[0], -self.pos[1], self.theta, node.shape.half[0], node.shape.half[2], node.name))

        def setScale(self, sx, sy, sz):
            if isinstance(self._node, Node) and self._node.name == BodyName.Sidewalk and all_kinds:
                rec.append((self.pos[0], -self.pos[1], self.theta, sx / 2, sy / 2, self._node.name))

        def __getattr__(self, k):
            return lambda *a, **kw: None

    names = ("BulletBoxShape", "BulletGhostNode", "BulletRigidBodyNode", "BaseRigidBodyNode", "NodePath", "Vec3",
             "LQuaternionf", "panda_position")
    saved = {k: getattr(bb, k) for k in names}
    bb.BulletBoxShape = lambda v: Shape(v)
    bb.BulletGhostNode = Node
    bb.BulletRigidBodyNode = Node
    bb.BaseRigidBodyNode = Node
    bb.NodePath = NP
    bb.Vec3 = lambda *a: tuple(a)
    bb.LQuaternionf = lambda *a: tuple(a)
    bb.panda_position = lambda p, z=0.0: (p[0], -p[1], z)
    try:
        eng, m = build_map(seed)
        rec.clear()  # build_map itself ran the patched builders (block search + rebuild); keep one clean pass only
        for blk in m.blocks:
            parent = NP(Node("root"))
            blk.sidewalk_node_path = NP(Node("sidewalks"))
            blk.lane_node_path = NP(Node("lanes"))
            blk.lane_vis_node_path = NP(Node("vis"))
            for _from, to_dict in blk.block_network.graph.items():
                for _to, lanes in to_dict.items():
                    if all_kinds:
                        blk._add_lane_surface(_from, _to, lanes)
                    for _id, l in enumerate(lanes):
                        blk._add_pgdrive_lanes(l, _id, l.width_at(0), l.line_color, parent)
    finally:
        for k, v in saved.items():
            setattr(bb, k, v)
    return rec


def cmd_primitives(tag, seeds=(1000, 1003, 1008, 1015, 1021, 1042, 1055, 1077, 1096)):
    """Every static primitive (lane surfaces, line ghosts, sidewalks) of a few maps as the reference's block code
    builds them -> tests/golden/primitives_*.json.gz (pins pgdrive_b200/tables.py lane_boxes / surface_boxes)."""
    out = {}
    for s in seeds:
        out[str(s)] = [[float(x) for x in b[:5]] + [str(b[5])] for b in capture_line_boxes(s, all_kinds=True)]
        print("seed", s, "primitives", len(out[str(s)]))
    path = os.path.join(GOLD, "primitives_%s.json.gz" % tag)
    with gzip.open(path, "wt") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path))


def cmd_detectors(tag, seeds=(1000, 1003, 1015, 1042), n_poses=24):
    import base64
    import math
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle.oracle import Oracle
    from pgdrive_b200 import cabi, tables as ptables
    from pgdrive.utils.cutils import _get_fake_cutils
    from pgdrive.constants import BodyName
    fake = _get_fake_cutils()

    class Hit:
        def __init__(self, node, frac):
            self.node, self.frac = node, frac

        def getNode(self):
            return self.node

        def getHitFraction(self):
            return self.frac

        def hasHit(self):
            return self.node is not None

        def getHitPos(self):
            return (0, 0, 0)

    class World2D:
        def __init__(self, boxes):
            self.edges = []
            self.rects = [(x, y, math.cos(h), math.sin(h), hl, hw) for x, y, h, hl, hw in boxes]
            for k, (x, y, h, hl, hw) in enumerate(boxes):
                c, s = math.cos(h), math.sin(h)
                pts = [(x + c * a * hl - s * b * hw, y + s * a * hl + c * b * hw)
                       for a, b in ((1, 1), (1, -1), (-1, -1), (-1, 1))]
                for i in range(4):
                    self.edges.append((k, pts[i], pts[(i + 1) % 4]))

        def rayTestClosest(self, start, end, mask):
            p = (start[0], -start[1])
            r = (end[0] - start[0], -(end[1] - start[1]))
            best, node = 1.0, None
            # Bullet's convex ray cast reports no hit for a shape that contains the ray origin
            inside = {k for k, (x, y, c, s, hl, hw) in enumerate(self.rects)
                      if abs((p[0] - x) * c + (p[1] - y) * s) <= hl and abs(-(p[0] - x) * s + (p[1] - y) * c) <= hw}
            for k, a, b in self.edges:
                if k in inside:
                    continue
                sx, sy = b[0] - a[0], b[1] - a[1]
                den = r[0] * sy - r[1] * sx
                if den == 0:
                    continue
                qx, qy = a[0] - p[0], a[1] - p[1]
                t = (qx * sy - qy * sx) / den
                u = (qx * r[1] - qy * r[0]) / den
                if 0.0 <= t <= 1.0 and 0.0 <= u <= 1.0 and t < best:
                    best, node = t, k
            return Hit(node, best)

    out = []
    for seed in seeds:
        boxes = capture_line_boxes(seed)
        cont = [b[:5] for b in boxes if b[5] != BodyName.Broken_line]
        every = [b[:5] for b in boxes]
        T = ptables.build_tables([seed]).finish()
        # the table must hold the same ghosts as the reference built (multiset of rounded rectangles)
        mine = T["boxes"][T["boxes"]["kind"] > 0]
        mine = mine[mine["kind"] < 4]
        assert len(mine) == len(boxes), (seed, len(mine), len(boxes))
        orc = Oracle(T, 1, auto_reset=False)
        orc.reset([0], [0])
        rs = np.random.RandomState(seed)
        lanes = T["lanes"]
        for k in range(n_poses):
            ln = lanes[rs.randint(len(lanes))]
            s = orc.get_state(0)
            v = s["veh"][0]
            t = rs.uniform(0.1, 0.9)
            if ln["kind"] == 0:
                x = ln["sx"] + t * (ln["ex"] - ln["sx"]) + rs.uniform(-1, 1)
                y = ln["sy"] + t * (ln["ey"] - ln["sy"]) + rs.uniform(-1, 1)
            else:
                x, y = ln["sx"] + rs.uniform(-2, 2), ln["sy"] + rs.uniform(-2, 2)
            v[0]["x"], v[0]["y"], v[0]["heading"] = x, y, rs.uniform(-np.pi, np.pi)
            s["veh"][0] = v
            ex, ey, eh = float(v[0]["x"]), float(v[0]["y"]), float(v[0]["heading"])
            rec = dict(seed=seed, state=base64.b64encode(s.tobytes()).decode())
            for name, n, dist, world in (("side", 120, 50.0, World2D(cont)), ("lane_line", 40, 20.0, World2D(every))):
                rng_ = np.arange(0, n) * (2 * np.pi / n) + np.deg2rad(90)  # set_start_phase_offset(90)
                cloud, _, _ = fake.cutils_perceive(np.ones(n), None, None, rng_, dist, eh, ex, ey, n, 0.2, world, set(),
                                                   False, False, 0, 0, 0)
                rec[name] = [float(c) for c in cloud]
            out.append(rec)
        orc.close()
        print("seed", seed, "ghosts", len(boxes), "continuous", len(cont))
    path = os.path.join(GOLD, "detectors_%s.json.gz" % tag)
    with gzip.open(path, "wt") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__" and sys.argv[1] == "detectors":
    cmd_detectors("v0")


if __name__ == "__main__" and sys.argv[1] == "primitives":
    cmd_primitives("v0")


# ---------------------------------------------------------------------------------------------
def cmd_random_lane(seeds, tag):
    """random_lane_width / random_lane_num (manager/map_manager.py:157-169): the reference's own add_random_to_map on
    a stream seeded like MapManager.seed(current_seed) (engine/base_engine.py:300-304, base_class/randomizable.py:16-18),
    then the LIVE block search with that lane configuration (load_map_from_json must be off with these options, so
    the simulator drives on the search-time map, pg_map.py:34-46)."""
    from pgdrive.manager.map_manager import MapManager

    class _Eng:
        def __init__(self, flags):
            self.global_config = flags

    class _Self:
        pass

    out = {}
    for s in seeds:
        rec = {}
        for name, flags in (("both", dict(random_lane_width=True, random_lane_num=True)),
                            ("width", dict(random_lane_width=True, random_lane_num=False)),
                            ("num", dict(random_lane_width=False, random_lane_num=True))):
            me = _Self()
            me.engine = _Eng(dict(flags, load_map_from_json=False))
            me.np_random = get_np_random(s)
            cfg = dict(MAP_CONFIG)
            cfg["seed"] = s
            cfg = MapManager.add_random_to_map(me, cfg)
            rec[name] = dict(lane_width=float(cfg["lane_width"]), lane_num=int(cfg["lane_num"]))
            if name != "both":
                continue
            make_engine(s)
            m = PGMap(map_config=dict(cfg), random_seed=None)
            lanes = []
            for frm, td in m.road_network.graph.items():
                for to, ls in td.items():
                    for i, l in enumerate(ls):
                        lanes.append(lane_record(frm, to, i, l))
            blocks = []
            for b in m.blocks:
                blocks.append(dict(
                    id=b.ID, name=b.name,
                    sockets=[dict(index=k.index, pos=[k.positive_road.start_node, k.positive_road.end_node],
                                  neg=[k.negative_road.start_node, k.negative_road.end_node])
                             for k in b.get_socket_list()],
                    respawn_roads=[[r.start_node, r.end_node] for r in b.get_respawn_roads()],
                    spawn_lanes=[[find_index(m, l) for l in ls] for ls in b.get_intermediate_spawn_lanes()]
                    if b.block_index != 0 else [],
                ))
            rec["lanes"], rec["blocks"] = lanes, blocks
            rec["block_sequence"] = json.loads(json.dumps(m.save_map()["block_sequence"]))
        out[str(s)] = rec
    path = os.path.join(GOLD, "maps_%s.json.gz" % tag)
    with gzip.open(path, "wt") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__" and sys.argv[1] == "random_lane":
    cmd_random_lane(list(range(1000, 1012)) + [5, 77, 2500], "random_lane")


def cmd_random_agent(seeds, tag):
    """random_agent_model: the ego's vehicle type per seed, from the reference's own random_vehicle_type on the stream
    AgentManager.seed(current_seed) sets up (manager/agent_manager.py:63-71, vehicle_type.py:84-86), and the type's
    dimensions / sampled parameters."""
    from pgdrive.component.vehicle.vehicle_type import random_vehicle_type, vehicle_type
    names = {v: k for k, v in vehicle_type.items()}
    out = {}
    for s in seeds:
        cls = random_vehicle_type(get_np_random(s))
        ego_seed = int(get_np_random(s).randint(0, 65536))  # BaseEngine.spawn_object -> generate_seed (first draw)
        out[str(s)] = dict(type=names[cls], length=float(cls.LENGTH), width=float(cls.WIDTH),
                           max_length=float(cls.MAX_LENGTH), max_width=float(cls.MAX_WIDTH),
                           params=sample_vehicle_params(cls, ego_seed))
    path = os.path.join(GOLD, "reset_%s.json.gz" % tag)
    with gzip.open(path, "wt") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__" and sys.argv[1] == "random_agent":
    cmd_random_agent(list(range(1000, 1040)) + [0, 5, 77, 2500], "random_agent")


def cmd_vehicle_types():
    """The reference's vehicle-type table (component/vehicle/vehicle_type.py:7-86) and parameter spaces
    (utils/space.py:219-255) as data, for tests/test_dynamics_pins.py."""
    from pgdrive.component.vehicle.vehicle_type import vehicle_type
    from pgdrive.component.vehicle.base_vehicle import BaseVehicle
    from pgdrive.component.static_object.traffic_object import TrafficCone, TrafficWarning, TrafficBarrier
    out = {}
    for key, cls in vehicle_type.items():
        from pgdrive.utils.space import VehicleParameterSpace
        raw = dict(default=VehicleParameterSpace.DEFAULT_VEHICLE, s=VehicleParameterSpace.S_VEHICLE,
                   m=VehicleParameterSpace.M_VEHICLE, l=VehicleParameterSpace.L_VEHICLE,
                   xl=VehicleParameterSpace.XL_VEHICLE)[key]
        space = {name: dict(type=type(sp).__name__, fields=[float(x) for x in sp]) for name, sp in raw.items()}
        out[key] = dict(LENGTH=cls.LENGTH, WIDTH=cls.WIDTH, HEIGHT=cls.HEIGHT, MASS=cls.MASS, TIRE_RADIUS=cls.TIRE_RADIUS,
                        FRONT_WHEELBASE=cls.FRONT_WHEELBASE, REAR_WHEELBASE=cls.REAR_WHEELBASE,
                        LATERAL_TIRE_TO_CENTER=cls.LATERAL_TIRE_TO_CENTER,
                        CHASSIS_TO_WHEEL_AXIS=getattr(cls, "CHASSIS_TO_WHEEL_AXIS", BaseVehicle.CHASSIS_TO_WHEEL_AXIS),
                        space=space)
    out["_base"] = dict(MAX_LENGTH=BaseVehicle.MAX_LENGTH, MAX_WIDTH=BaseVehicle.MAX_WIDTH,
                        MAX_STEERING=BaseVehicle.MAX_STEERING, STEERING_INCREMENT=BaseVehicle.STEERING_INCREMENT)
    out["_objects"] = dict(TrafficCone=dict(RADIUS=TrafficCone.RADIUS, MASS=TrafficCone.MASS),
                           TrafficWarning=dict(RADIUS=TrafficWarning.RADIUS, MASS=TrafficWarning.MASS),
                           TrafficBarrier=dict(LENGTH=TrafficBarrier.LENGTH, WIDTH=TrafficBarrier.WIDTH,
                                               MASS=TrafficBarrier.MASS))
    path = os.path.join(GOLD, "vehicle_types.json.gz")
    with gzip.open(path, "wt") as f:
        json.dump(out, f)
    print("wrote", path, json.dumps(out["s"])[:300])


if __name__ == "__main__" and sys.argv[1] == "vehicle_types":
    cmd_vehicle_types()
