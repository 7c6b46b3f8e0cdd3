"""The gym surface of the reference over the batched CUDA step.

``PGDriveEnv``     drop-in for /root/reference/pgdrive/envs/pgdrive_env.py:112-336 (one environment;
                   ``reset(force_seed=)``, ``step(a) -> (obs, reward, done, info)``, spaces, ``seed``,
                   ``current_seed``, ``close``; same config dict and KeyError on unknown keys).
``VecPGDriveEnv``  the same environment batched: ``num_envs`` copies advanced by ONE kernel launch per
                   step, observations / rewards / dones as ``[N, ...]`` arrays.

Both call the C-ABI of include/pgdrive_b200.h through ctypes; torch tensors are only the device-buffer
container.  There is no CPU path: constructing either class without the CUDA library or without a
CUDA device raises.
"""
import ctypes as C
import hashlib
import os
import pickle

import numpy as np

from . import cabi, devgen, episode, mapgen, tables
from .config import ENGINE_CONFIG, Config, check_supported, default_config, post_process_config
from .spaces import Box, MultiDiscrete

ENVIRONMENTS = {  # /root/reference/pgdrive/register.py:7-40
    "PGDrive-test-v0": dict(start_seed=0, environment_num=200),
    "PGDrive-validation-v0": dict(start_seed=200, environment_num=800),
    "PGDrive-v0": dict(start_seed=1000, environment_num=100),
    "PGDrive-10envs-v0": dict(start_seed=1000, environment_num=10),
    "PGDrive-1000envs-v0": dict(start_seed=1000, environment_num=1000),
    "PGDrive-training0-v0": dict(start_seed=3000, environment_num=1000),
    "PGDrive-training1-v0": dict(start_seed=5000, environment_num=1000),
    "PGDrive-training2-v0": dict(start_seed=7000, environment_num=1000),
}

INFO_FLAGS = dict(
    crash_vehicle=cabi.F_CRASH_VEHICLE, out_of_road=cabi.F_OUT_OF_ROAD, arrive_dest=cabi.F_ARRIVE_DEST,
    max_step=cabi.F_MAX_STEP
)


def parse_map_config(cfg):
    """component/map/base_map.py:16-35: ``map`` shorthand (int = block count, str = block ids) unless the
    user overrode ``map_config``."""
    mc = cfg["map_config"].get_dict()
    default_mc = default_config()["map_config"].get_dict()
    if mc != default_mc:
        out = dict(default_mc)
        out.update(mc)
        return out
    easy = cfg["map"]
    if isinstance(easy, bool) or not isinstance(easy, (int, str)):
        raise ValueError("Unkown easy map config: {} and original map config: {}".format(easy, mc))
    mc["type"] = "block_num" if isinstance(easy, int) else "block_sequence"
    mc["config"] = easy
    return mc


def effective_horizon(cfg, map_config):
    """Step limit handed to the kernel: ``horizon`` (base_env.py:190-192) and, with ``auto_termination``, the
    reference's 250 steps per block including the first one (base_env.py:318-326); both end the episode with
    ``max_step`` set, so the smaller one decides.  Every map of an environment has the same number of blocks."""
    limits = []
    if cfg["horizon"]:
        limits.append(int(cfg["horizon"]))
    if cfg["auto_termination"]:
        if map_config["type"] == "block_num":
            n_blocks = int(map_config["config"]) + 1
        elif map_config["type"] == "block_sequence":
            n_blocks = len(str(map_config["config"])) + 1
        else:
            raise ValueError("Map can not be created by {}".format(map_config["type"]))
        limits.append(250 * n_blocks)
    return min(limits) if limits else 0


def make_action_space(cfg):
    """base_vehicle.py:721-727."""
    if cfg["discrete_action"]:
        return MultiDiscrete([cfg["discrete_steering_dim"], cfg["discrete_throttle_dim"]])
    # extra_action_dim only widens the space; the vehicle reads action[0] and action[1] (base_vehicle.py:343-358)
    return Box(-1.0, 1.0, shape=(2 + int(cfg["vehicle_config"]["extra_action_dim"]), ), dtype=np.float32)


def discrete_to_continuous(actions, cfg):
    """EnvInputPolicy.act + convert_to_continuous_action (policy/env_input_policy.py:17-31), literally: the index is
    clipped to [-1, 1] BEFORE it is scaled, so every index >= 1 maps to ``unit - 1``.  Works on numpy arrays and
    torch tensors alike."""
    su = 2.0 / (cfg["discrete_steering_dim"] - 1)
    tu = 2.0 / (cfg["discrete_throttle_dim"] - 1)
    if isinstance(actions, np.ndarray) or not hasattr(actions, "clamp"):
        a = np.clip(np.asarray(actions, dtype=np.float32), -1.0, 1.0)
        return np.stack([a[..., 0] * su - 1.0, a[..., 1] * tu - 1.0], axis=-1).astype(np.float32)
    a = actions.float().clamp(-1.0, 1.0)
    return (a * a.new_tensor([su, tu]) - 1.0).contiguous()


def seed_map_config(mc, seed, random_lane_width=False, random_lane_num=False):
    """MapManager.add_random_to_map (manager/map_manager.py:157-169) on the stream MapManager.seed(current_seed) sets
    up (engine/base_engine.py:300-304): lane width = rand() * (4.5 - 3.0) + 3.0, then lane number =
    randint(MIN_LANE_NUM=2, MAX_LANE_NUM=3) -- literally, i.e. always 2 (component/map/pg_map.py:13-16)."""
    if not (random_lane_width or random_lane_num):
        return mc
    from . import rng
    rs = rng.seeded(int(seed))
    out = dict(mc)
    if random_lane_width:
        out["lane_width"] = float(rs.rand() * (4.5 - 3.0) + 3.0)
    if random_lane_num:
        out["lane_num"] = int(rs.randint(2, 3))
    return out


def _seed_tables(args):
    seed, mc, density, spawn = args[:4]
    stored = args[4] if len(args) > 4 else None
    random_agent = bool(args[5]) if len(args) > 5 else False
    traffic_mode = args[6] if len(args) > 6 else "trigger"
    accident_prob = float(args[7]) if len(args) > 7 else 0.0
    traffic_rs = args[8] if len(args) > 8 else None  # random_traffic: the traffic manager's stream lives across resets
    kw = dict(lane_num=mc["lane_num"], lane_width=mc["lane_width"], exit_length=mc["exit_length"])
    if stored is not None:  # restored from a map file: no block search (pg_map.py:48-71)
        pgmap = mapgen.build_from_sequence(seed, stored, **kw)
    elif mc["type"] == "block_num":
        pgmap = mapgen.generate_map(seed, block_num=mc["config"], **kw)
    elif mc["type"] == "block_sequence":
        pgmap = mapgen.generate_map(seed, sequence=mc["config"], **kw)
    else:
        raise ValueError("Map can not be created by {}".format(mc["type"]))
    ts = tables.TableSet()
    mid = ts.add_map(pgmap)
    lane, lon, lat = spawn
    ts.add_episode(pgmap, mid, episode.make_episode(pgmap, seed, density, tuple(lane), random_agent, traffic_mode,
                                                      accident_prob, traffic_rs), tuple(lane), lon, lat)
    return ts.finish()


def merge_tables(parts):
    """Concatenate per-seed table sets, rebasing the offsets stored in the records."""
    keys = ["maps", "lanes", "roads", "boxes", "cell_start", "cell_entries", "episodes", "slots", "route_nodes",
            "route_roads"]
    off = {k: 0 for k in keys}
    out = {k: [] for k in keys}
    for p in parts:
        maps, eps, slots = p["maps"].copy(), p["episodes"].copy(), p["slots"].copy()
        maps["lane_off"] += off["lanes"]
        maps["road_off"] += off["roads"]
        maps["box_off"] += off["boxes"]
        maps["cell_off"] += off["cell_start"]
        maps["entry_off"] += off["cell_entries"]
        eps["map"] += off["maps"]
        eps["slot_off"] += off["slots"]
        slots["route_off"] += off["route_nodes"]
        for k, a in (("maps", maps), ("episodes", eps), ("slots", slots)):
            out[k].append(a)
        for k in keys:
            if k not in ("maps", "episodes", "slots"):
                out[k].append(p[k])
            off[k] += len(p[k])
    T = {k: np.concatenate(v) if v else None for k, v in out.items()}
    T["max_slots"] = int(T["episodes"]["n_slots"].max())
    return T


def load_map_file(path_or_dict, map_config, seeds):
    """The reference's map-collection format (``PGDriveEnv.dump_all_maps``, envs/pgdrive_env.py:260-288; read back by
    manager/map_manager.py:43-91): ``{"map_config": {...}, "map_data": {seed: {"block_sequence": [...]}}}``.
    Returns {seed: block_sequence} when the file's map_config equals ours and it covers ``seeds``, else None
    (the reference then falls back to generating the maps, map_manager.py:55-64)."""
    import json
    data = path_or_dict
    if not isinstance(data, dict):
        with open(path_or_dict) as f:
            data = json.load(f)
    if set(data.keys()) != {"map_config", "map_data"}:
        raise ValueError("a map file holds exactly the keys map_config and map_data")
    have = {int(k): v for k, v in data["map_data"].items()}
    if dict(data["map_config"]) != dict(map_config) or not set(int(s) for s in seeds).issubset(have):
        return None
    return {s: have[s]["block_sequence"] for s in have}


def dump_maps(seeds, map_config):
    """``dump_all_maps``: block sequences of ``seeds`` in the reference's JSON-serialisable format."""
    kw = dict(lane_num=map_config["lane_num"], lane_width=map_config["lane_width"],
              exit_length=map_config["exit_length"])
    out = {}
    for s in seeds:
        if map_config["type"] == "block_num":
            seq = mapgen.search_sequence(int(s), block_num=map_config["config"], **kw)
        else:
            seq = mapgen.search_sequence(int(s), sequence=map_config["config"], **kw)
        out[int(s)] = {"block_sequence": seq}
    return dict(map_config=dict(map_config), map_data=out)


def build_seed_tables(seeds, map_config, density, spawn, workers=None, stored=None, random_lane=(False, False),
                      random_agent_model=False, traffic_mode="trigger", accident_prob=0.0):
    """Tables for a list of seeds, built in worker processes when there are many, with an optional
    on-disk cache ($PGDRIVE_B200_CACHE) because map search costs ~50 ms per seed.  ``stored`` = {seed: block
    sequence} restored from a map file."""
    seeds = [int(s) for s in seeds]
    jobs = [(s, seed_map_config(map_config, s, *random_lane), density, spawn, (stored or {}).get(s),
             bool(random_agent_model), traffic_mode, float(accident_prob)) for s in seeds]
    cache_dir = os.environ.get("PGDRIVE_B200_CACHE")
    path = None
    if cache_dir:
        src = b"".join(open(os.path.join(os.path.dirname(__file__), f), "rb").read()
                       for f in ("mapgen.py", "roadnet.py", "episode.py", "tables.py", "rng.py"))
        key = hashlib.sha1(repr(jobs).encode() + src).hexdigest()
        path = os.path.join(cache_dir, "tables_%s.pkl" % key)
        if os.path.exists(path):
            with open(path, "rb") as f:
                return pickle.load(f)
    if workers is None:
        workers = min(len(os.sched_getaffinity(0)), 32) if len(seeds) >= 256 else 1
    if workers > 1:
        # fork (not spawn): children only run numpy code, and spawn would re-import the caller's __main__
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            parts = pool.map(_seed_tables, jobs, chunksize=max(1, len(jobs) // (workers * 4)))
    else:
        parts = [_seed_tables(j) for j in jobs]
    T = merge_tables(parts)
    if path:
        os.makedirs(cache_dir, exist_ok=True)
        with open(path, "wb") as f:
            pickle.dump(T, f)
    return T


def accident_prob_of(cfg):
    """Accident scenes exist only when the TrafficObjectManager is registered (SafePGDriveEnv.setup_engine)."""
    return float(cfg["accident_prob"]) if cfg.get("object_manager", False) else 0.0


def pick_slots(need):
    """Smallest vehicle-slot count the kernel is instantiated for (16 / 24 / 32) that holds ``need`` vehicles."""
    for v in (16, 24, 32):
        if need <= v:
            return v
    raise ValueError("an episode needs %d vehicle slots; at most 32 are supported" % need)


class _Engine:
    """One C-ABI handle + its loaded tables."""
    def __init__(self, cfg, num_envs, num_slots, device, auto_reset, horizon=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("pgdrive_b200 needs a CUDA device (there is no CPU fallback)")
        self.torch = torch
        self.lib = cabi.load_library()
        self.device = torch.device("cuda", device)
        horizon = (cfg["horizon"] or 0) if horizon is None else horizon
        self.pcfg = cabi.make_config(
            num_envs, num_slots, cfg["decision_repeat"], horizon, cfg["physics_world_step_size"],
            cfg["success_reward"], cfg["out_of_road_penalty"], cfg["crash_vehicle_penalty"], cfg["driving_reward"],
            cfg["speed_reward"], cfg["out_of_road_cost"], cfg["crash_vehicle_cost"], cfg["use_lateral"],
            cfg["out_of_route_done"], auto_reset,
            n_side=cfg["vehicle_config"]["side_detector"]["num_lasers"],
            side_distance=cfg["vehicle_config"]["side_detector"]["distance"],
            n_lane_line=cfg["vehicle_config"]["lane_line_detector"]["num_lasers"],
            lane_line_distance=cfg["vehicle_config"]["lane_line_detector"]["distance"],
            random_agent_model=bool(cfg["random_agent_model"]),
            lidar_gaussian_noise=cfg["vehicle_config"]["lidar"]["gaussian_noise"],
            lidar_dropout_prob=cfg["vehicle_config"]["lidar"]["dropout_prob"],
            noise_seed=int(cfg.get("noise_seed", 0) or 0) + 7919 * int(device),
            increment_steering=bool(cfg["vehicle_config"]["increment_steering"]),
            crash_object_penalty=cfg["crash_object_penalty"], crash_object_cost=cfg["crash_object_cost"],
            safe_rl_env=bool(cfg.get("safe_rl_env", False))
        )
        self.obs_dim = cabi.obs_dim(self.pcfg)
        self.h = C.c_void_p()
        cabi.check(self.lib, self.lib.pgd_create(C.byref(self.pcfg), device, C.byref(self.h)))
        self.num_envs, self.num_slots = num_envs, num_slots

    def load(self, T):
        t, keep = cabi.pack_tables(T)
        cabi.check(self.lib, self.lib.pgd_load_tables(self.h, C.byref(t)))

    def stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def close(self):
        if self.h:
            self.lib.pgd_destroy(self.h)
            self.h = C.c_void_p()


class VecPGDriveEnv:
    """``num_envs`` PGDrive environments advanced together.

    Environment ``i`` plays seed ``start_seed + i % environment_num`` unless ``reset(seeds=...)`` says
    otherwise.  With ``auto_reset`` (default) an environment that reported ``done`` restarts on the
    same seed at its next ``step`` (that step ignores the action and returns the reset observation
    with reward 0 -- the "next-step" convention of vector environments); the reference itself never
    resets on its own (README.md:96-101).
    """
    def __init__(self, config=None, tables_dict=None, obs_out=None):
        merged = default_config()
        merged.update(ENGINE_CONFIG)
        self.config = post_process_config(merged.update(config or {}, allow_add_new_key=False))
        check_supported(self.config)
        cfg = self.config
        if cfg["random_traffic"]:
            # the environments of a seed share one traffic template in HBM; PGDriveEnv (one environment) redraws it at
            # every reset like the reference
            raise NotImplementedError("random_traffic needs per-reset traffic templates: use PGDriveEnv")
        self.num_envs = int(cfg["num_envs"])
        self.start_seed, self.env_num = int(cfg["start_seed"]), int(cfg["environment_num"])
        self.map_config = parse_map_config(cfg)
        vc = cfg["vehicle_config"]
        self._spawn = (tuple(vc["spawn_lane_index"]), float(vc["spawn_longitude"]), float(vc["spawn_lateral"]))
        seeds = list(range(self.start_seed, self.start_seed + self.env_num))
        stored = None
        if cfg["load_map_from_json"] and cfg["_load_map_from_json"] is not None:
            stored = load_map_file(cfg["_load_map_from_json"], self.map_config, seeds)
        self._T = None
        random_lane = (bool(cfg["random_lane_width"]), bool(cfg["random_lane_num"]))
        why_not = None  # what the device generator (pgd_mapgen.cuh) does not cover
        if cfg["random_agent_model"]:
            why_not = "device_mapgen spawns the default ego vehicle"
        elif cfg["traffic_mode"] == "respawn":
            why_not = "device_mapgen builds trigger-mode traffic"
        elif accident_prob_of(cfg) > 0:
            why_not = "device_mapgen builds no accident scenes"
        elif tuple(self._spawn[0][:2]) != (">", ">>"):
            why_not = "device_mapgen spawns the ego on the first road"
        elif tables_dict is not None or stored is not None:
            why_not = "tables / a map file were given"
        if cfg["device_mapgen"] and why_not and tables_dict is None and stored is None:
            raise NotImplementedError(why_not)
        use_device = (cfg["device_mapgen"] is None or cfg["device_mapgen"]) and why_not is None
        self.reset_path = "host"
        if use_device:
            try:
                self._build_on_device(cfg, seeds, random_lane)
                self.reset_path = "device"
            except (RuntimeError, ValueError):
                if cfg["device_mapgen"]:  # required explicitly
                    raise
                # auto: a configuration outside the generator's table capacities (many lanes / blocks): host path
        if self.reset_path == "host":
            self._T = tables_dict if tables_dict is not None else build_seed_tables(
                seeds, self.map_config, cfg["traffic_density"], self._spawn, stored=stored, random_lane=random_lane,
                random_agent_model=cfg["random_agent_model"], traffic_mode=cfg["traffic_mode"],
                accident_prob=accident_prob_of(cfg)
            )
            self.episode_of_seed = {int(s): i for i, s in enumerate(self._T["episodes"]["seed"])}
            need = int(self._T["max_slots"])
            slots = cfg["num_slots"] or pick_slots(need)
            if need > slots:
                raise ValueError("the loaded seeds need %d vehicle slots; num_slots=%d" % (need, slots))
            self.engine = _Engine(cfg, self.num_envs, slots, int(cfg["device"]), bool(cfg["auto_reset"]),
                                  effective_horizon(cfg, self.map_config))
            self.engine.load(self._T)
        torch = self.engine.torch
        dev = self.engine.device
        n = self.num_envs
        self.obs_dim = self.engine.obs_dim  # 274 unless side / lane-line detectors are on (state_obs.py:108-130)
        self.obs = obs_out if obs_out is not None else torch.empty((n, self.obs_dim), dtype=torch.float32, device=dev)
        self.reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self.done = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.info = torch.zeros((n, cabi.INFO_DT.itemsize // 4), dtype=torch.int32, device=dev)
        # host-path results live in page-locked memory so that pgd_step_host can DMA straight into them
        self._pinned = [
            torch.empty((n, self.obs_dim), dtype=torch.float32, pin_memory=True),
            torch.empty(n, dtype=torch.float32, pin_memory=True),
            torch.empty(n, dtype=torch.uint8, pin_memory=True),
            torch.empty((n, cabi.INFO_DT.itemsize // 4), dtype=torch.int32, pin_memory=True),
        ]
        self._h_obs, self._h_reward, self._h_done = [t.numpy() for t in self._pinned[:3]]
        self._h_info = self._pinned[3].numpy().view(cabi.INFO_DT).reshape(n)
        self.observation_space = Box(-0.0, 1.0, shape=(self.obs_dim, ), dtype=np.float32)
        self.action_space = make_action_space(cfg)
        self.env_seeds = np.array([self.start_seed + i % self.env_num for i in range(n)], dtype=np.int64)

    def _build_on_device(self, cfg, seeds, random_lane):
        """The whole reset path on the GPU (pgd_generate_tables: one warp per seed writes maps, collision primitives,
        bucket grid, traffic slots and routes straight into the tables the step kernel reads; nothing visits the host).
        The few seeds on which the reference's own result hangs on the last bit of a glibc call (devgen.tie_seeds) are
        built by the reference-pinned host path and patched in, so device tables equal host tables on every seed."""
        gc = devgen.make_gen_config(self.map_config, cfg["traffic_density"], self._spawn, random_lane)
        slots = cfg["num_slots"] or 16
        while True:
            self.engine = _Engine(cfg, self.num_envs, slots, int(cfg["device"]), bool(cfg["auto_reset"]),
                                  effective_horizon(cfg, self.map_config))
            try:
                devgen.generate(self.engine, seeds, gc)
                break
            except RuntimeError:
                counts, status = getattr(self.engine, "gen_counts", None), getattr(self.engine, "gen_status", None)
                self.engine.close()
                need = int(counts[:, 5].max()) if counts is not None else 0
                if status is None or not (status == 0).all() or cfg["num_slots"] or need <= slots or need > 32:
                    raise
                slots = pick_slots(need)  # some seed needs more vehicle slots than tried
        ties = sorted(devgen.tie_seeds(self.map_config, cfg["traffic_density"], self._spawn, random_lane) & set(seeds))
        for s in ties:
            part = _seed_tables((s, seed_map_config(self.map_config, s, *random_lane), cfg["traffic_density"], self._spawn))
            devgen.patch(self.engine, seeds.index(s), part, self.engine.gen_caps)
        self.device_mapgen_patched = ties
        self.episode_of_seed = {int(s): i for i, s in enumerate(seeds)}

    @property
    def T(self):
        """Host copy of the tables (downloaded from the device when they were generated there)."""
        if self._T is None:
            self._T = devgen.download(self.engine)
        return self._T

    # -- reset / step ------------------------------------------------------------------------------
    def reset(self, seeds=None, env_ids=None):
        """Restart ``env_ids`` (default: all) on ``seeds`` (default: each env's current seed).  Returns the
        ``[N, 274]`` observation tensor (rows of untouched environments keep their last observation)."""
        if env_ids is None:
            env_ids = np.arange(self.num_envs, dtype=np.int32)
        env_ids = np.ascontiguousarray(env_ids, dtype=np.int32)
        new_seeds = self.env_seeds[env_ids] if seeds is None else \
            np.broadcast_to(np.asarray(seeds, dtype=np.int64), env_ids.shape)
        try:  # resolve first, assign after: a bad seed must not stick to the environment (base_env.py:451-458 asserts)
            eps = np.array([self.episode_of_seed[int(s)] for s in new_seeds], dtype=np.int32)
        except KeyError as e:
            raise KeyError("seed %s is outside [start_seed, start_seed + environment_num)" % e)
        self.env_seeds[env_ids] = new_seeds
        e = self.engine
        cabi.check(
            e.lib,
            e.lib.pgd_reset(e.h, env_ids.ctypes.data, eps.ctypes.data, len(env_ids), self.obs.data_ptr(),
                            self.info.data_ptr(), e.stream())
        )
        return self.obs

    def step_into(self, actions, obs_ptr, reward_ptr, done_ptr):
        """Device step writing results to raw device pointers (e.g. this rank's rows of a peer-mapped gather buffer,
        pgdrive_b200.sharding.PeerGather).  ``actions``: float32 CUDA tensor [N, 2]."""
        e = self.engine
        a = actions.contiguous()
        cabi.check(
            e.lib,
            e.lib.pgd_step(e.h, a.data_ptr(), int(obs_ptr), int(reward_ptr), int(done_ptr), self.info.data_ptr(),
                           e.stream())
        )

    def step(self, actions, out=None, copy=True):
        """``actions``: ``[N, 2]`` float32, either a CUDA tensor (device path: returns CUDA tensors, no
        synchronisation) or a numpy array (host path: pinned staging, returns numpy arrays).  ``out`` (device path
        only) = ``(obs, reward, done)`` tensors to write into instead of the environment's own buffers, e.g. this
        rank's rows of an all-gather buffer.

        Host path: the results land in page-locked staging arrays that the NEXT step overwrites.  ``copy=True``
        (default) returns fresh arrays, like the reference; ``copy=False`` returns the staging arrays themselves
        (no 75 MB copy per step at 65 536 environments), read-only -- valid only until the next ``step`` / ``reset``.

        Observation rows cross PCIe packed (head + lidar hit mask + the beams that are not 1.0) and are expanded by a
        pool of host threads inside the library; the staging rows keep the previous step's content and only what
        changed is rewritten (include/pgdrive_b200.h: pgd_step_host)."""
        e = self.engine
        torch = e.torch
        if self.config["discrete_action"]:
            actions = discrete_to_continuous(actions, self.config)
        extra = 0 if self.config["discrete_action"] else int(self.config["vehicle_config"]["extra_action_dim"])
        if isinstance(actions, torch.Tensor):
            if actions.device != e.device or actions.dtype != torch.float32 or \
                    tuple(actions.shape) != (self.num_envs, 2 + extra):
                raise ValueError("actions must be a float32 [num_envs, %d] tensor on %s" % (2 + extra, e.device))
            a = (actions[:, :2] if extra else actions).contiguous()
            obs, reward, done = out if out is not None else (self.obs, self.reward, self.done)
            if out is not None:
                n = self.num_envs
                ok = (tuple(obs.shape) == (n, self.obs_dim) and obs.dtype == torch.float32 and obs.is_contiguous()
                      and tuple(reward.shape) == (n, ) and reward.dtype == torch.float32 and reward.is_contiguous()
                      and tuple(done.shape) == (n, ) and done.dtype == torch.uint8 and done.is_contiguous()
                      and obs.device == reward.device == done.device == e.device)
                if not ok:
                    raise ValueError("out must be contiguous (f32 [N,274], f32 [N], u8 [N]) tensors on %s" % e.device)
            cabi.check(
                e.lib,
                e.lib.pgd_step(e.h, a.data_ptr(), obs.data_ptr(), reward.data_ptr(), done.data_ptr(),
                               self.info.data_ptr(), e.stream())
            )
            return obs, reward, done, self.info
        if out is not None:
            raise ValueError("out= is only supported with device actions")
        a = np.asarray(actions, dtype=np.float32)
        if a.shape != (self.num_envs, 2 + extra):
            raise ValueError("actions must have shape [num_envs, %d]" % (2 + extra))
        a = np.ascontiguousarray(a[:, :2])
        cabi.check(
            e.lib,
            e.lib.pgd_step_host(e.h, a.ctypes.data, self._h_obs.ctypes.data, self._h_reward.ctypes.data,
                                self._h_done.ctypes.data, self._h_info.ctypes.data)
        )
        if copy:
            return self._h_obs.copy(), self._h_reward.copy(), self._h_done.copy(), self._h_info.copy()
        # read-only views: the next step writes only what changed in the observation rows (pgd_step_host)
        views = [a.view() for a in (self._h_obs, self._h_reward, self._h_done, self._h_info)]
        for v in views:
            v.flags.writeable = False
        return tuple(views)

    def rows_to_host(self, obs, reward, done, obs_out, reward_out=None, done_out=None):
        """Device rows -> host arrays through the packed PCIe path (pgd_rows_to_host): ``obs`` [R, obs_dim] float32 CUDA
        tensor (any batch in this GPU's memory, e.g. the gathered batch on rank 0), ``reward`` / ``done`` CUDA tensors or
        None; ``obs_out`` [R, obs_dim] float32 numpy array (C-contiguous; pass the same one every step: only what
        changed is rewritten), ``reward_out`` / ``done_out`` numpy arrays.  Returns when the host arrays are valid."""
        e = self.engine
        rows, od = int(obs.shape[0]), int(obs.shape[1])
        if not (obs.is_cuda and obs.is_contiguous() and obs.dtype == e.torch.float32):
            raise ValueError("obs must be a contiguous float32 CUDA tensor")
        if obs_out.shape != (rows, od) or obs_out.dtype != np.float32 or not obs_out.flags.c_contiguous:
            raise ValueError("obs_out must be a C-contiguous float32 [%d, %d] array" % (rows, od))
        for dev, host, dt, name in ((reward, reward_out, np.float32, "reward"), (done, done_out, np.uint8, "done")):
            if (dev is None) != (host is None):
                raise ValueError("%s needs both its device tensor and its host array" % name)
            if dev is not None and (host.shape != (rows, ) or host.dtype != dt or not dev.is_contiguous()):
                raise ValueError("%s: [%d] %s on both sides" % (name, rows, np.dtype(dt).name))
        cabi.check(e.lib, e.lib.pgd_rows_to_host(
            e.h, obs.data_ptr(), None if reward is None else reward.data_ptr(), None if done is None else done.data_ptr(),
            rows, od, obs_out.ctypes.data, None if reward_out is None else reward_out.ctypes.data,
            None if done_out is None else done_out.ctypes.data, e.stream()))

    def host_transfer_bytes(self):
        """(host-to-device, device-to-host) bytes the last host-path ``step`` moved over PCIe."""
        import ctypes
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        e = self.engine
        cabi.check(e.lib, e.lib.pgd_host_transfer_bytes(e.h, ctypes.byref(a), ctypes.byref(b)))
        return int(a.value), int(b.value)

    def info_numpy(self):
        """Device info of the last device-path step/reset as a structured array (synchronises)."""
        return self.info.cpu().numpy().view(cabi.INFO_DT).reshape(self.num_envs)

    # -- state exchange (parity debugging) -----------------------------------------------------------
    def get_state(self, env):
        s = np.zeros(1, cabi.ENV_STATE_DT)
        e = self.engine
        cabi.check(e.lib, e.lib.pgd_get_state(e.h, int(env), s.ctypes.data))
        return s

    def set_state(self, env, state):
        s = np.ascontiguousarray(state)
        e = self.engine
        cabi.check(e.lib, e.lib.pgd_set_state(e.h, int(env), s.ctypes.data))

    @property
    def launch_count(self):
        return int(self.engine.lib.pgd_launch_count(self.engine.h))

    def close(self):
        self.engine.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class EgoView:
    """Read-only stand-in for ``env.vehicle`` (component/vehicle/base_vehicle.py:390-425,683-698): the ego's pose and
    flags read back from the simulator.  ``state`` / ``flags`` are callables so that every access is current."""
    LENGTH, WIDTH, HEIGHT, MASS = 4.51, 1.852, 1.19, 1100.0  # DefaultVehicle (vehicle_type.py:7-20)
    MAX_LENGTH, MAX_WIDTH, MAX_STEERING = 10, 2.5, 60  # base_vehicle.py:83-85

    def __init__(self, state, flags, lane_name, destination, spawn_road):
        self._state, self._flags, self._lane_name = state, flags, lane_name
        self._destination, self._spawn_road = destination, spawn_road

    def _veh(self):
        return self._state()["veh"][0][0]

    @property
    def position(self):
        v = self._veh()
        return np.array([float(v["x"]), float(v["y"])])

    @property
    def heading_theta(self):
        return float(self._veh()["heading"])

    @property
    def heading(self):
        h = self.heading_theta
        return np.array([np.cos(h), np.sin(h)])

    @property
    def speed(self):
        """km/h, clipped at 0 like base_vehicle.py:400-408."""
        return float(np.clip(self._veh()["speed"] * 3.6, 0.0, 100000.0))

    @property
    def velocity(self):
        return self.heading * self.speed

    @property
    def steering(self):
        return float(self._veh()["steer"])

    @property
    def throttle_brake(self):
        return float(self._veh()["throttle"])

    @property
    def lane_index(self):
        return self._lane_name(int(self._veh()["lane"]))

    @property
    def on_lane(self):
        return bool(self._veh()["flags"] & cabi.V_ON_LANE)

    crash_vehicle = property(lambda self: bool(self._flags() & cabi.F_CRASH_VEHICLE))
    crash_sidewalk = property(lambda self: bool(self._flags() & cabi.F_CRASH_SIDEWALK))
    out_of_route = property(lambda self: bool(self._flags() & cabi.F_OUT_OF_ROUTE))
    on_yellow_continuous_line = property(lambda self: bool(self._flags() & cabi.F_ON_YELLOW))
    on_white_continuous_line = property(lambda self: bool(self._flags() & cabi.F_ON_WHITE))
    on_broken_line = property(lambda self: bool(self._flags() & cabi.F_ON_BROKEN))
    arrive_destination = property(lambda self: bool(self._flags() & cabi.F_ARRIVE_DEST))

    def get_state(self):
        return {
            "heading": self.heading_theta, "position": self.position.tolist(),
            "done": self.crash_vehicle or self.out_of_route or self.crash_sidewalk or not self.on_lane,
            "speed": self.speed, "spawn_road": self._spawn_road, "destination": self._destination(),
        }


class PGDriveEnv:
    """Single-environment drop-in (a ``num_envs=1`` view of the batched engine).  Maps are built lazily,
    one seed at a time, like the reference's map manager (manager/map_manager.py:98-155)."""
    DEFAULT_AGENT = "default_agent"

    @classmethod
    def default_config(cls):
        return default_config()

    def __init__(self, config=None):
        self.config = post_process_config(self.default_config().update(config or {}, allow_add_new_key=False))
        check_supported(self.config)
        self.start_seed, self.env_num = int(self.config["start_seed"]), int(self.config["environment_num"])
        self.map_config = parse_map_config(self.config)
        vc = self.config["vehicle_config"]
        self._spawn = (tuple(vc["spawn_lane_index"]), float(vc["spawn_longitude"]), float(vc["spawn_lateral"]))
        self.obs_dim = ((vc["side_detector"]["num_lasers"] or 2) + 6 + vc["lane_line_detector"]["num_lasers"] + 266 +
                        (2 if self.config["random_agent_model"] else 0))
        self.observation_space = Box(-0.0, 1.0, shape=(self.obs_dim, ), dtype=np.float32)
        self.action_space = make_action_space(self.config)
        self._parts, self._episode_of_seed, self._loaded_seed = {}, {}, None
        self._maps = {}
        self._stored = None
        if self.config["load_map_from_json"] and self.config["_load_map_from_json"] is not None:
            self._stored = load_map_file(self.config["_load_map_from_json"], self.map_config,
                                         range(self.start_seed, self.start_seed + self.env_num))
        self._engine = None
        self._seed = None
        self._rs = np.random.RandomState()
        self._traffic_rs = np.random.RandomState() if self.config["random_traffic"] else None
        self.episode_steps = 0
        self._obs = self._reward = self._done = self._info = None

    # lazily create the engine the first time reset() runs (base_env.py:166-178)
    def _ensure_seed(self, seed):
        """Tables of ``seed`` (built once, kept on the host like the reference's map cache, map_manager.py:98-155) and an
        engine that holds exactly the tables of the CURRENT seed: a visit uploads one map (tens of KB), not every map
        seen so far."""
        if seed not in self._parts or self._traffic_rs is not None:
            # random_traffic (traffic_manager.py:348-350): the traffic manager is not re-seeded at reset, every visit of a
            # map draws new traffic from one stream -- the seed's tables are rebuilt and uploaded again
            mc = seed_map_config(self.map_config, seed, self.config["random_lane_width"], self.config["random_lane_num"])
            self._parts[seed] = _seed_tables((seed, mc, self.config["traffic_density"], self._spawn,
                                              (self._stored or {}).get(seed), bool(self.config["random_agent_model"]),
                                              self.config["traffic_mode"], accident_prob_of(self.config),
                                              self._traffic_rs))
            self._episode_of_seed[seed] = 0
            if self._traffic_rs is not None:
                self._loaded_seed = None
        part = self._parts[seed]
        slots = pick_slots(int(part["max_slots"]))
        if self._engine is not None and self._engine.num_slots < slots:
            self._engine.close()
            self._engine = None
            self._loaded_seed = None
        if self._engine is None:
            self._engine = _Engine(self.config, 1, slots, int(os.environ.get("PGDRIVE_B200_DEVICE", 0)), False,
                                   effective_horizon(self.config, self.map_config))
            torch = self._engine.torch
            dev = self._engine.device
            self._obs = torch.empty((1, self.obs_dim), dtype=torch.float32, device=dev)
            self._reward = torch.zeros(1, dtype=torch.float32, device=dev)
            self._done = torch.zeros(1, dtype=torch.uint8, device=dev)
            self._info = torch.zeros((1, cabi.INFO_DT.itemsize // 4), dtype=torch.int32, device=dev)
            self._act = torch.zeros((1, 2), dtype=torch.float32, device=dev)
        if self._loaded_seed != seed:
            self._engine.load(part)
            self._loaded_seed = seed

    # -- read-only views of the reference's object graph (envs/base_env.py:371-462) ----------------------------------
    def _map_of(self, seed):
        if seed not in self._maps:
            mc = seed_map_config(self.map_config, seed, self.config["random_lane_width"], self.config["random_lane_num"])
            kw = dict(lane_num=mc["lane_num"], lane_width=mc["lane_width"], exit_length=mc["exit_length"])
            stored = (self._stored or {}).get(seed)
            if stored is not None:
                self._maps[seed] = mapgen.build_from_sequence(seed, stored, **kw)
            elif mc["type"] == "block_num":
                self._maps[seed] = mapgen.generate_map(seed, block_num=mc["config"], **kw)
            else:
                self._maps[seed] = mapgen.generate_map(seed, sequence=mc["config"], **kw)
        return self._maps[seed]

    @property
    def current_map(self):
        """The map of the current seed as the host generator's object (blocks, road network, block sequence)."""
        return None if self._seed is None else self._map_of(self._seed)

    @property
    def maps(self):
        """{seed: map or None}: which maps of [start_seed, start_seed + environment_num) have been visited."""
        return {s: (self._map_of(s) if s in self._episode_of_seed else None)
                for s in range(self.start_seed, self.start_seed + self.env_num)}

    @property
    def vehicle(self):
        assert self._engine is not None, "Please initialize the environment first!"
        flat = [(f, t, i) for (f, t), lanes in self.current_map.net.roads() for i in range(len(lanes))]
        route = episode.route_for(self.current_map, tuple(self._spawn[0]), self._seed)
        return EgoView(self.get_state, lambda: int(self._info.cpu().numpy().view(cabi.INFO_DT).reshape(-1)[0]["flags"]),
                       lambda k: flat[k], lambda: (route[-2], route[-1]), tuple(self._spawn[0][:-1]))

    @property
    def vehicles(self):
        return {self.DEFAULT_AGENT: self.vehicle}

    def dump_all_maps(self):
        """envs/pgdrive_env.py:260-288: every map of [start_seed, start_seed + environment_num) as block sequences."""
        return dump_maps(range(self.start_seed, self.start_seed + self.env_num), self.map_config)

    def seed(self, seed=None):
        if seed is not None:
            self._seed = int(seed)

    @property
    def current_seed(self):
        return self._seed

    def reset(self, episode_data=None, force_seed=None):
        if episode_data is not None:
            raise NotImplementedError("episode replay is not supported")
        if force_seed is not None:
            seed = int(force_seed)
        else:  # base_env.py:451-458: an unseeded generator picks the map
            seed = int(self._rs.randint(self.start_seed, self.start_seed + self.env_num))
        self._seed = seed
        self._ensure_seed(seed)
        e = self._engine
        ids = np.zeros(1, np.int32)
        eps = np.array([self._episode_of_seed[seed]], np.int32)
        cabi.check(
            e.lib,
            e.lib.pgd_reset(e.h, ids.ctypes.data, eps.ctypes.data, 1, self._obs.data_ptr(), self._info.data_ptr(),
                            e.stream())
        )
        self.episode_steps = 0
        # the reference returns float64 although the space says float32 (state_obs.py:146,150)
        return self._obs[0].cpu().numpy().astype(np.float64)

    def step(self, action):
        if self._engine is None:
            raise RuntimeError("call reset() before step()")
        e = self._engine
        self.episode_steps += 1
        raw = np.asarray(action, dtype=np.float32).reshape(-1)[:2]
        a = discrete_to_continuous(raw, self.config) if self.config["discrete_action"] else raw
        self._act.copy_(e.torch.from_numpy(a).reshape(1, 2))
        cabi.check(
            e.lib,
            e.lib.pgd_step(e.h, self._act.data_ptr(), self._obs.data_ptr(), self._reward.data_ptr(),
                           self._done.data_ptr(), self._info.data_ptr(), e.stream())
        )
        obs = self._obs[0].cpu().numpy().astype(np.float64)
        rec = self._info.cpu().numpy().view(cabi.INFO_DT).reshape(-1)[0]
        flags = int(rec["flags"])
        info = {k: bool(flags & bit) for k, bit in INFO_FLAGS.items()}
        info.update(
            crash_object=bool(flags & cabi.F_CRASH_OBJECT), crash_building=False,
            crash=bool(flags & (cabi.F_CRASH_VEHICLE | cabi.F_CRASH_OBJECT)),
            cost=float(rec["cost"]), velocity=float(rec["velocity"]), steering=float(rec["steering"]),
            acceleration=float(rec["acceleration"]), step_energy=float(rec["step_energy"]),
            episode_energy=float(rec["episode_energy"]), step_reward=float(rec["step_reward"]),
            episode_reward=float(rec["episode_reward"]), episode_length=int(rec["episode_length"]),
            raw_action=(float(raw[0]), float(raw[1])), overtake_vehicle_num=0,
            on_yellow_continuous_line=bool(flags & cabi.F_ON_YELLOW), on_white_continuous_line=bool(flags & cabi.F_ON_WHITE),
            on_broken_line=bool(flags & cabi.F_ON_BROKEN), crash_sidewalk=bool(flags & cabi.F_CRASH_SIDEWALK),
            on_lane=bool(flags & cabi.F_ON_LANE), out_of_route=bool(flags & cabi.F_OUT_OF_ROUTE)
        )
        return obs, float(self._reward.item()), bool(self._done.item()), info

    def get_state(self):
        s = np.zeros(1, cabi.ENV_STATE_DT)
        cabi.check(self._engine.lib, self._engine.lib.pgd_get_state(self._engine.h, 0, s.ctypes.data))
        return s

    def render(self, *a, **k):
        raise NotImplementedError("the batched simulator is headless")

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None


class SafePGDriveEnv(PGDriveEnv):
    """envs/safe_pgdrive_env.py:7-63: accident scenes (cones, warning tripods, barriers, broken-down vehicles) on 80 % of
    the eligible blocks, sparse traffic, and crashes that cost instead of ending the episode (``safe_rl_env``); ``info``
    gains ``total_cost``."""
    @classmethod
    def default_config(cls):
        config = default_config()
        config.update(dict(environment_num=100, accident_prob=0.8, traffic_density=0.05, crash_vehicle_cost=1,
                           crash_object_cost=1, out_of_road_cost=1., use_lateral=False))
        config.update(dict(safe_rl_env=True, cost_to_reward=False, object_manager=True), allow_add_new_key=True)
        return config

    def __init__(self, config=None):
        super(SafePGDriveEnv, self).__init__(config)
        self.episode_cost = 0

    def reset(self, *args, **kwargs):
        self.episode_cost = 0
        return super(SafePGDriveEnv, self).reset(*args, **kwargs)

    def step(self, action):
        obs, reward, done, info = super(SafePGDriveEnv, self).step(action)
        self.episode_cost += info["cost"]
        info["total_cost"] = self.episode_cost
        return obs, reward, done, info


def make(env_id, **overrides):
    """``gym.make(id)`` for the ids the reference registers (register.py:7-43)."""
    if env_id not in ENVIRONMENTS:
        raise KeyError("unknown environment id %r; known: %s" % (env_id, sorted(ENVIRONMENTS)))
    cfg = dict(ENVIRONMENTS[env_id])
    cfg.update(overrides)
    return PGDriveEnv(cfg)
