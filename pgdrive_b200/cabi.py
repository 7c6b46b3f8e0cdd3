"""ctypes mirror of include/pgd_tables.h and include/pgdrive_b200.h, and the loader of the CUDA library.

There is no CPU fallback: if ``pgdrive_b200/csrc/libpgdrive_b200.so`` is missing or fails to load,
:func:`load_library` raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PGDRIVE_B200_LIB") or os.path.join(HERE, "csrc", "libpgdrive_b200.so")

OBS_DIM = 274  # PGDrive-v0 (no side / lane-line detector)
MAX_SLOTS = 32
MAX_DETECTOR_BEAMS = 240

INFO_DT = np.dtype([
    ("velocity", "f4"), ("steering", "f4"), ("acceleration", "f4"), ("step_energy", "f4"), ("episode_energy", "f4"),
    ("step_reward", "f4"), ("episode_reward", "f4"), ("cost", "f4"), ("episode_length", "i4"), ("flags", "u4")
])
VEH_STATE_DT = np.dtype([
    ("x", "f4"), ("y", "f4"), ("heading", "f4"), ("speed", "f4"), ("steer", "f4"), ("throttle", "f4"), ("pid_hp", "f4"),
    ("pid_hi", "f4"), ("pid_lp", "f4"), ("pid_li", "f4"), ("target_speed", "f4"), ("lane", "i4"), ("ck0", "i4"),
    ("ck1", "i4"), ("rt_lane", "i4"), ("timer", "i4"), ("rnd_n", "i4"), ("airborne", "i4"), ("flags", "i4"),
    ("yaw_rate", "f4")
])
ENV_STATE_DT = np.dtype([
    ("episode", "i4"), ("next_group", "i4"), ("done", "i4"), ("ep_len", "i4"), ("prev_steer", "f4"),
    ("prev_throttle", "f4"), ("ep_reward", "f4"), ("energy", "f4"), ("veh", VEH_STATE_DT, (MAX_SLOTS, ))
])
assert INFO_DT.itemsize == 40 and VEH_STATE_DT.itemsize == 80 and ENV_STATE_DT.itemsize == 32 + 80 * MAX_SLOTS

F_CRASH_VEHICLE, F_OUT_OF_ROAD, F_ARRIVE_DEST, F_MAX_STEP = 1, 2, 4, 8
F_ON_YELLOW, F_ON_WHITE, F_ON_BROKEN, F_CRASH_SIDEWALK = 16, 32, 64, 128
F_ON_LANE, F_OUT_OF_ROUTE, F_WAS_RESET, F_CRASH_OBJECT = 256, 512, 1024, 2048
V_ALIVE, V_ACTIVE, V_ON_LANE, V_CRASHED = 1, 2, 4, 16


class PgdTables(C.Structure):
    _fields_ = [
        ("maps", C.c_void_p), ("n_maps", C.c_int32), ("lanes", C.c_void_p), ("n_lanes", C.c_int32),
        ("roads", C.c_void_p), ("n_roads", C.c_int32), ("boxes", C.c_void_p), ("n_boxes", C.c_int32),
        ("cell_start", C.c_void_p), ("n_cell_start", C.c_int32), ("cell_entries", C.c_void_p),
        ("n_cell_entries", C.c_int32), ("episodes", C.c_void_p), ("n_episodes", C.c_int32), ("slots", C.c_void_p),
        ("n_slots", C.c_int32), ("route_nodes", C.c_void_p), ("route_roads", C.c_void_p), ("n_route", C.c_int32)
    ]


class PgdConfig(C.Structure):
    _fields_ = [
        ("num_envs", C.c_int32), ("num_slots", C.c_int32), ("decision_repeat", C.c_int32), ("horizon", C.c_int32),
        ("dt", C.c_float), ("success_reward", C.c_float), ("out_of_road_penalty", C.c_float),
        ("crash_vehicle_penalty", C.c_float), ("driving_reward", C.c_float), ("speed_reward", C.c_float),
        ("out_of_road_cost", C.c_float), ("crash_vehicle_cost", C.c_float), ("use_lateral", C.c_int32),
        ("out_of_route_done", C.c_int32), ("auto_reset", C.c_int32), ("n_side", C.c_int32),
        ("n_lane_line", C.c_int32), ("side_distance", C.c_float), ("lane_line_distance", C.c_float),
        ("random_agent_model", C.c_int32), ("lidar_gaussian_noise", C.c_float), ("lidar_dropout_prob", C.c_float),
        ("noise_seed", C.c_int32), ("increment_steering", C.c_int32), ("crash_object_penalty", C.c_float),
        ("crash_object_cost", C.c_float), ("safe_rl_env", C.c_int32)
    ]


def make_config(num_envs, num_slots=16, decision_repeat=5, horizon=0, dt=0.02, success_reward=10.0,
                out_of_road_penalty=5.0, crash_vehicle_penalty=5.0, driving_reward=1.0, speed_reward=0.1,
                out_of_road_cost=1.0, crash_vehicle_cost=1.0, use_lateral=False, out_of_route_done=False,
                auto_reset=True, n_side=0, side_distance=50.0, n_lane_line=0, lane_line_distance=20.0,
                random_agent_model=False, lidar_gaussian_noise=0.0, lidar_dropout_prob=0.0, noise_seed=0,
                increment_steering=False, crash_object_penalty=5.0, crash_object_cost=1.0, safe_rl_env=False):
    if not (0 <= n_side <= MAX_DETECTOR_BEAMS and 0 <= n_lane_line <= MAX_DETECTOR_BEAMS):
        raise ValueError("side / lane-line detectors support 0..%d lasers" % MAX_DETECTOR_BEAMS)
    return PgdConfig(num_envs, num_slots, decision_repeat, int(horizon or 0), dt, success_reward, out_of_road_penalty,
                     crash_vehicle_penalty, driving_reward, speed_reward, out_of_road_cost, crash_vehicle_cost,
                     int(use_lateral), int(out_of_route_done), int(auto_reset), int(n_side), int(n_lane_line),
                     float(side_distance), float(lane_line_distance), int(bool(random_agent_model)),
                     float(lidar_gaussian_noise), float(lidar_dropout_prob), int(noise_seed),
                     int(bool(increment_steering)), float(crash_object_penalty), float(crash_object_cost),
                     int(bool(safe_rl_env)))


def obs_dim(cfg):
    """pgd_obs_dim of include/pgd_tables.h."""
    return (cfg.n_side if cfg.n_side > 0 else 2) + 6 + cfg.n_lane_line + (2 if cfg.random_agent_model else 0) + 10 + 16 + 240


def pack_tables(T):
    """``T``: dict of numpy arrays from TableSet.finish().  Returns (PgdTables, keepalive list)."""
    keep = []

    def ptr(name):
        a = np.ascontiguousarray(T[name])
        keep.append(a)
        return a.ctypes.data, len(a)

    t = PgdTables()
    t.maps, t.n_maps = ptr("maps")
    t.lanes, t.n_lanes = ptr("lanes")
    t.roads, t.n_roads = ptr("roads")
    t.boxes, t.n_boxes = ptr("boxes")
    t.cell_start, t.n_cell_start = ptr("cell_start")
    t.cell_entries, t.n_cell_entries = ptr("cell_entries")
    t.episodes, t.n_episodes = ptr("episodes")
    t.slots, t.n_slots = ptr("slots")
    t.route_nodes, t.n_route = ptr("route_nodes")
    t.route_roads, _ = ptr("route_roads")
    return t, keep


_LIB = None


def load_library():
    """Load the CUDA library and declare the prototypes of include/pgdrive_b200.h."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "CUDA library %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % LIB_PATH
        )
    lib = C.CDLL(LIB_PATH)
    vp, i32 = C.c_void_p, C.c_int32
    lib.pgd_last_error.restype = C.c_char_p
    lib.pgd_create.argtypes = [C.POINTER(PgdConfig), i32, C.POINTER(vp)]
    lib.pgd_destroy.argtypes = [vp]
    lib.pgd_load_tables.argtypes = [vp, C.POINTER(PgdTables)]
    lib.pgd_reset.argtypes = [vp, vp, vp, i32, vp, vp, vp]
    lib.pgd_step.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.pgd_step_host.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.pgd_host_invalidate.argtypes = [vp]
    lib.pgd_rows_to_host.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp, vp]
    lib.pgd_host_transfer_bytes.argtypes = [vp, vp, vp]
    lib.pgd_host_expand_rows.argtypes = [vp, vp, i32, i32, i32, vp, vp, i32]
    lib.pgd_host_pool_selftest.argtypes = [i32, i32, i32]
    lib.pgd_get_state.argtypes = [vp, i32, vp]
    lib.pgd_set_state.argtypes = [vp, i32, vp]
    lib.pgd_state_bytes_per_env.argtypes = [vp]
    lib.pgd_state_bytes_per_env.restype = C.c_int64
    lib.pgd_launch_count.argtypes = [vp]
    lib.pgd_launch_count.restype = C.c_int64
    lib.pgd_last_kernel_ms.argtypes = [vp]
    lib.pgd_last_kernel_ms.restype = C.c_float
    lib.pgd_set_timing.argtypes = [vp, i32]
    lib.pgd_words_checksum.argtypes = [vp, vp, C.c_uint64, vp, vp]
    lib.pgd_pack_rows.argtypes = [vp, vp, i32, i32, vp]
    lib.pgd_expand_rows.argtypes = [vp, vp, i32, i32, vp]
    lib.pgd_expand_rows_delta.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    lib.pgd_packed_row_words.argtypes = [i32]
    lib.pgd_packed_row_words.restype = i32
    lib.pgd_peer_alloc.argtypes = [vp, C.c_uint64, C.POINTER(vp), C.c_char_p]
    lib.pgd_peer_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    lib.pgd_peer_release.argtypes = [vp, vp, i32]
    lib.pgd_generate_tables.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp]
    lib.pgd_table_sizes.argtypes = [vp, vp]
    lib.pgd_patch_tables.argtypes = [vp, i32, C.POINTER(PgdTables), vp]
    lib.pgd_download_tables.argtypes = [vp, C.POINTER(PgdTables)]
    for name in ("pgd_create", "pgd_destroy", "pgd_load_tables", "pgd_reset", "pgd_step", "pgd_step_host",
                 "pgd_host_invalidate", "pgd_rows_to_host", "pgd_host_transfer_bytes", "pgd_host_expand_rows",
                 "pgd_host_pool_selftest", "pgd_get_state", "pgd_set_state", "pgd_set_timing", "pgd_words_checksum",
                 "pgd_pack_rows", "pgd_expand_rows", "pgd_expand_rows_delta", "pgd_packed_row_words",
                 "pgd_patch_tables", "pgd_peer_alloc", "pgd_peer_open", "pgd_peer_release", "pgd_generate_tables",
                 "pgd_table_sizes", "pgd_download_tables"):
        getattr(lib, name).restype = i32
    _LIB = lib
    return lib


def check(lib, rc):
    if rc != 0:
        raise RuntimeError("pgdrive_b200: %s (code %d)" % (lib.pgd_last_error().decode(), rc))
