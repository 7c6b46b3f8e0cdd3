"""Configuration: the reference's default dicts and its merge / key-check semantics.

Defaults restate BASE_DEFAULT_CONFIG (/root/reference/pgdrive/envs/base_env.py:19-90) and
PGDriveEnv_DEFAULT_CONFIG (envs/pgdrive_env.py:22-109); ``Config.update(..., allow_add_new_key=False)``
raises ``KeyError`` on an unknown key exactly where the reference does (utils/config.py:115-125).
Keys that only drive rendering are accepted and ignored (headless simulator).
"""
import copy


class Config:
    def __init__(self, config=None):
        if isinstance(config, Config):
            config = config.get_dict()
        self._config = {}
        for k, v in copy.deepcopy(config or {}).items():
            self._config[k] = Config(v) if isinstance(v, dict) else v

    def get_dict(self):
        return {k: (v.get_dict() if isinstance(v, Config) else copy.deepcopy(v)) for k, v in self._config.items()}

    def copy(self):
        return Config(self)

    def update(self, new_dict=None, allow_add_new_key=True):
        new_dict = copy.deepcopy(new_dict.get_dict() if isinstance(new_dict, Config) else (new_dict or {}))
        if not allow_add_new_key:
            diff = set(new_dict).difference(self._config)
            if diff:
                raise KeyError(
                    "'{}' does not exist in existing config. Please use config.update(...) to update the config. "
                    "Existing keys: {}.".format(diff, self._config.keys())
                )
        for k, v in new_dict.items():
            cur = self._config.get(k)
            if isinstance(cur, Config):
                if not isinstance(v, dict):
                    if allow_add_new_key:
                        self._config[k] = v
                        continue
                    raise TypeError(
                        "Type error! The item {} has original type {} and updating type {}.".format(k, type(cur), type(v))
                    )
                cur.update(v, allow_add_new_key=allow_add_new_key)
            else:
                self._config[k] = Config(v) if isinstance(v, dict) else v
        return self

    def __getitem__(self, k):
        if k not in self._config:
            raise KeyError(
                "'{}' does not exist in existing config. Please use config.update(...) to update the config. "
                "Existing keys: {}.".format(k, self._config.keys())
            )
        return self._config[k]

    def __setitem__(self, k, v):
        self._config[k] = Config(v) if isinstance(v, dict) else v

    def __contains__(self, k):
        return k in self._config

    def get(self, k, default=None):
        return self._config.get(k, default)

    def keys(self):
        return self._config.keys()

    def items(self):
        return self._config.items()

    def __repr__(self):
        return "Config(%r)" % (self.get_dict(), )


BASE_DEFAULT_CONFIG = dict(
    start_seed=0,
    environment_num=1,
    num_agents=1,
    is_multi_agent=False,
    allow_respawn=False,
    delay_done=0,
    random_agent_model=False,
    IDM_agent=False,
    decision_repeat=5,
    discrete_action=False,
    discrete_steering_dim=5,
    discrete_throttle_dim=5,
    use_render=False,
    debug=False,
    fast=False,
    cull_scene=True,
    manual_control=False,
    controller="keyboard",
    use_chase_camera_follow_lane=False,
    camera_height=1.8,
    camera_dist=7,
    prefer_track_agent=None,
    draw_map_resolution=1024,
    top_down_camera_initial_x=0,
    top_down_camera_initial_y=0,
    top_down_camera_initial_z=200,
    vehicle_config=dict(
        increment_steering=False,
        vehicle_model="default",
        show_navi_mark=True,
        extra_action_dim=0,
        enable_reverse=False,
        random_navi_mark_color=False,
        show_dest_mark=False,
        show_line_to_dest=False,
        am_i_the_special_one=False
    ),
    window_size=(1200, 900),
    physics_world_step_size=2e-2,
    show_fps=True,
    global_light=False,
    onscreen_message=True,
    debug_physics_world=False,
    debug_static_world=False,
    headless_machine_render=False,
    pstats=False,
    max_distance=None,
    _debug_crash_object=False,
    record_episode=False,
    horizon=None,
)

PGDRIVE_DEFAULT_CONFIG = dict(
    start_seed=0,
    environment_num=1,
    map=3,
    random_lane_width=False,
    random_lane_num=False,
    map_config={"type": "block_num", "config": None, "lane_width": 3.5, "lane_num": 3, "exit_length": 50},
    load_map_from_json=True,
    _load_map_from_json=None,
    use_topdown=False,
    offscreen_render=False,
    _disable_detector_mask=False,
    traffic_density=0.1,
    traffic_mode="trigger",
    random_traffic=False,
    accident_prob=0.,
    auto_termination=False,
    use_saver=False,
    save_level=0.5,
    vehicle_config=dict(
        lidar=dict(num_lasers=240, distance=50, num_others=4, gaussian_noise=0.0, dropout_prob=0.0),
        side_detector=dict(num_lasers=0, distance=50, gaussian_noise=0.0, dropout_prob=0.0),
        lane_line_detector=dict(num_lasers=0, distance=20, gaussian_noise=0.0, dropout_prob=0.0),
        show_lidar=False,
        mini_map=(84, 84, 250),
        rgb_camera=(84, 84),
        depth_camera=(84, 84, True),
        show_side_detector=False,
        show_lane_line_detector=False,
        image_source="rgb_camera",
        spawn_lane_index=(">", ">>", 0),
        spawn_longitude=5.0,
        spawn_lateral=0.0,
        destination_node=None,
        overtake_stat=False,
        action_check=False,
        random_color=False,
    ),
    rgb_clip=True,
    gaussian_noise=0.0,
    dropout_prob=0.0,
    success_reward=10.0,
    out_of_road_penalty=5.0,
    crash_vehicle_penalty=5.0,
    crash_object_penalty=5.0,
    acceleration_penalty=0.0,
    low_speed_penalty=0.0,
    driving_reward=1.0,
    general_penalty=0.0,
    speed_reward=0.1,
    use_lateral=False,
    crash_vehicle_cost=1,
    crash_object_cost=1,
    out_of_road_cost=1.,
    out_of_route_done=False,
)

# knobs of the batched engine itself (not in the reference)
ENGINE_CONFIG = dict(
    num_envs=1,          # environments stepped per launch
    num_slots=None,      # vehicle slots per env (16 / 24 / 32); None = smallest that fits the loaded seeds
    device=0,            # CUDA device ordinal
    auto_reset=True,     # VecPGDriveEnv only: a finished env restarts at its next step (action ignored)
    # SafePGDriveEnv (envs/safe_pgdrive_env.py:7-63).  object_manager = its setup_engine registered the
    # TrafficObjectManager: accident scenes are built with probability accident_prob per eligible block (without it
    # accident_prob has no effect, exactly as in the reference's plain PGDriveEnv); safe_rl_env = a crash costs but does
    # not end the episode; cost_to_reward = the costs are added to the penalties.
    object_manager=False,
    safe_rl_env=False,
    cost_to_reward=False,
    noise_seed=0,        # key of the counter-based generator behind lidar gaussian_noise / dropout_prob
    # VecPGDriveEnv: run the reset path (map search, tables, episode templates) on the GPU.  None (default) = whenever
    # the device generator covers the configuration (default ego type, trigger / hybrid traffic, no accident scenes, no
    # map file), else the host path; True = require it (raises otherwise); False = host path.
    device_mapgen=None,
)


def default_config():
    c = Config(BASE_DEFAULT_CONFIG)
    c.update(PGDRIVE_DEFAULT_CONFIG)
    return c


# Options of the reference that this simulator does not implement.  They are accepted at their default
# value and rejected otherwise, so that a config that silently changed behaviour cannot slip through.
UNSUPPORTED_IF_CHANGED = {
    "num_agents": 1, "is_multi_agent": False, "IDM_agent": False,
    "use_render": False, "manual_control": False, "use_topdown": False, "offscreen_render": False,
    "record_episode": False,
}


def post_process_config(cfg):
    """PGDriveEnv._post_process_config (envs/pgdrive_env.py:131-157): the top-level gaussian_noise / dropout_prob fan out
    to the three sensors' configs (asserting that those were left at 0)."""
    if cfg.get("cost_to_reward", False):  # safe_pgdrive_env.py:36-42
        cfg["crash_vehicle_penalty"] += cfg["crash_vehicle_cost"]
        cfg["crash_object_penalty"] += cfg["crash_object_cost"]
        cfg["out_of_road_penalty"] += cfg["out_of_road_cost"]
    vc = cfg["vehicle_config"]
    for key in ("gaussian_noise", "dropout_prob"):
        if cfg[key] > 0:
            for sensor in ("lidar", "side_detector", "lane_line_detector"):
                assert vc[sensor][key] == 0, "You already provide config!"
                vc[sensor][key] = cfg[key]
    return cfg


def check_supported(cfg):
    for k, v in UNSUPPORTED_IF_CHANGED.items():
        if cfg[k] != v:
            raise NotImplementedError("config[%r]=%r is not supported by the batched simulator (only %r)" % (k, cfg[k], v))
    if cfg["traffic_mode"] not in ("trigger", "hybrid", "respawn"):
        raise ValueError("No such mode named {}".format(cfg["traffic_mode"]))  # traffic_manager.py:69
    if (cfg["random_lane_width"] or cfg["random_lane_num"]) and cfg["load_map_from_json"]:
        # manager/map_manager.py:158-166
        raise AssertionError("You are supposed to turn off the load_map_from_json")
    vc = cfg["vehicle_config"]
    lid = vc["lidar"]
    if (lid["num_lasers"], lid["distance"], lid["num_others"]) != (240, 50, 4):
        raise NotImplementedError("lidar must be 240 beams x 50 m with 4 neighbours")
    if lid["gaussian_noise"] < 0 or not 0 <= lid["dropout_prob"] <= 1:
        raise ValueError("lidar gaussian_noise must be >= 0 and dropout_prob in [0, 1]")
    for name in ("side_detector", "lane_line_detector"):
        det = vc[name]
        if not 0 <= det["num_lasers"] <= 240:
            raise NotImplementedError("%s supports 0..240 lasers" % name)
        # gaussian_noise / dropout_prob of the two detectors are accepted and unused, as in the reference: only the lidar's
        # cloud points go through _add_noise_to_cloud_points (obs/state_obs.py:64-66,96-97,165-170)
        if det["num_lasers"] and det["distance"] <= 0:
            raise ValueError("%s distance must be positive" % name)
    if vc["enable_reverse"]:
        raise NotImplementedError("enable_reverse is not supported (the planar vehicle model has no reverse gear)")
    if vc["overtake_stat"]:
        # base_vehicle.py:702 calls the static Lidar.get_surrounding_vehicles() without its argument: the reference
        # itself raises TypeError on the first step with this option
        raise NotImplementedError("overtake_stat raises in the reference at this version (base_vehicle.py:702)")
    if vc["extra_action_dim"] < 0:
        raise ValueError("extra_action_dim must be >= 0")
    if cfg["decision_repeat"] < 1:
        raise ValueError("decision_repeat must be >= 1")
