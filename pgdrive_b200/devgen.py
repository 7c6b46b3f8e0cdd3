"""Device-side reset path: seeds -> maps + episode templates generated ON the GPU, straight into the tables the step
kernel reads (no host map search, no table upload).

What runs on the device is the reference's whole reset decision chain -- BIG block search with retry / back-tracking
(component/algorithm/BIG.py:67-151), the block builders, the rebuild from the block sequence (pg_map.py:48-71), the
static collision primitives, traffic slots / vehicle parameters / routes (traffic_manager.py:239-290,
navigation.py:99-153) -- with the reference's own random streams (sha512-seeded MT19937, numpy legacy draws).
Source: pgdrive_b200/csrc/pgd_mapgen.cuh (+ pgd_rng.cuh, pgd_dd.cuh), kernel and C-ABI in pgd_mapgen.cu.

This module only holds the ctypes mirrors of the generator's config structs and the table-capacity rule.
"""
import ctypes as C

BLOCK_CODE = {"C": 0, "S": 1, "r": 2, "R": 3, "X": 4, "T": 5, "O": 6}  # order of BLOCK_TYPE_DISTRIBUTION_V2
CODE_BLOCK = {v: k for k, v in BLOCK_CODE.items()}
CODE_BLOCK[100] = "I"

GEN_ERRORS = {
    1: "lane pool full", 2: "road pool full", 3: "box table full", 4: "grid cell table full",
    5: "grid entry table full", 6: "search queue full", 7: "route table full", 8: "spawn candidate list full",
    9: "more than 32 vehicle slots", 10: "block search could not finish", 11: "road lookup failed",
    12: "can not set a destination", 13: "more than 11 traffic trigger groups", 14: "too many blocks",
    15: "unsupported generator config"
}


class GenConfig(C.Structure):
    _fields_ = [
        ("block_num", C.c_int32), ("lane_num", C.c_int32), ("n_fixed", C.c_int32), ("spawn_lane", C.c_int32),
        ("lane_width", C.c_double), ("exit_length", C.c_double), ("density", C.c_double), ("spawn_long", C.c_double),
        ("spawn_lat", C.c_double), ("fixed_types", C.c_int8 * 32), ("random_lane_width", C.c_int32),
        ("random_lane_num", C.c_int32)
    ]


class GenCaps(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("blocks", "lanes", "roads", "boxes", "cells", "entries", "queue", "route",
                                         "cand")]


def make_gen_config(map_config, density, spawn=((">", ">>", 0), 5.0, 0.0), random_lane=(False, False)):
    """``map_config``: the reference's map_config dict (type block_num | block_sequence)."""
    lane, lon, lat = spawn
    if tuple(lane[:2]) != (">", ">>"):
        raise ValueError("device map generation spawns the ego on the first road ('>', '>>')")
    gc = GenConfig()
    gc.lane_num = int(map_config["lane_num"])
    gc.lane_width = float(map_config["lane_width"])
    gc.exit_length = float(map_config["exit_length"])
    gc.density = float(density)
    gc.spawn_lane, gc.spawn_long, gc.spawn_lat = int(lane[2]), float(lon), float(lat)
    gc.random_lane_width, gc.random_lane_num = int(bool(random_lane[0])), int(bool(random_lane[1]))
    if map_config["type"] == "block_num":
        gc.block_num, gc.n_fixed = int(map_config["config"]), 0
    elif map_config["type"] == "block_sequence":
        seq = str(map_config["config"])
        if len(seq) > 32 or any(ch not in BLOCK_CODE for ch in seq):
            raise ValueError("block_sequence may hold up to 32 of %s" % sorted(BLOCK_CODE))
        gc.block_num, gc.n_fixed = len(seq), len(seq)
        for i, ch in enumerate(seq):
            gc.fixed_types[i] = BLOCK_CODE[ch]
    else:
        raise ValueError("Map can not be created by {}".format(map_config["type"]))
    return gc


def caps_for(gen_config):
    """Per-map table capacities (fixed stride in the device tables).  Measured over 1 300 maps: a 3-block map needs
    <= 198 lanes, 66 roads, 1 891 boxes, 1 441 cells, 14 236 grid entries, 178 route entries; the generator reports an
    error (never truncates) when a map does not fit."""
    b = int(gen_config.block_num) + 1
    c = GenCaps()
    c.blocks = b
    c.lanes = 88 * b
    c.roads = 32 * b
    c.boxes = 900 * b
    c.cells = 4800 * b + 1
    c.entries = 12000 * b
    c.queue = 4096
    c.route = 32 * (3 * b + 10)
    c.cand = 160 * b
    return c


def generate(engine, seeds, gen_config, caps=None):
    """Generate maps + episode templates of ``seeds`` on ``engine``'s device and install them as its tables
    (pgd_generate_tables).  Returns (status [n], counts [n, 8]); raises when a seed cannot be generated."""
    import numpy as np
    from . import cabi
    caps = caps or caps_for(gen_config)
    s = np.ascontiguousarray(seeds, dtype=np.int32)
    status = np.zeros(len(s), np.int32)
    counts = np.zeros((len(s), 8), np.int32)
    rc = engine.lib.pgd_generate_tables(engine.h, s.ctypes.data, len(s), C.addressof(gen_config), C.addressof(caps),
                                        status.ctypes.data, counts.ctypes.data, engine.stream())
    engine.gen_status, engine.gen_counts, engine.gen_caps = status, counts, caps
    cabi.check(engine.lib, rc)
    return status, counts


def tie_seeds(map_config, density, spawn=((">", ">>", 0), 5.0, 0.0), random_lane=(False, False)):
    """Seeds whose reference result hangs on the last bit of a glibc sin / cos / atan2 call (tools/find_tie_seeds.py):
    known for the default map config at density 0.1 (all 30 000 shipped seeds checked); for any other config the set is
    unknown and empty."""
    import json
    import os
    path = os.path.join(os.path.dirname(__file__), "devgen_ties.json")
    if not os.path.exists(path):
        return set()
    rec = json.load(open(path))
    same = all(map_config.get(k) == v for k, v in rec["map_config"].items()) and abs(density - rec["traffic_density"]) < 1e-12
    same = same and tuple(spawn[0]) == (">", ">>", 0) and (spawn[1], spawn[2]) == (5.0, 0.0) and not any(random_lane)
    return set(rec["tie_seeds"]) if same else set()


def patch(engine, index, T_one, caps):
    """Replace map / episode ``index`` of device-generated tables by a host-built single-seed table set."""
    from . import cabi
    t, keep = cabi.pack_tables(T_one)
    cabi.check(engine.lib, engine.lib.pgd_patch_tables(engine.h, int(index), C.byref(t), C.addressof(caps)))


def download(engine):
    """Device tables -> dict of numpy arrays in the layout of tables.TableSet.finish() (fixed stride per map when
    they were generated on the device)."""
    import numpy as np
    from . import cabi, tables as tb
    sizes = np.zeros(9, np.int64)
    cabi.check(engine.lib, engine.lib.pgd_table_sizes(engine.h, sizes.ctypes.data))
    dts = [tb.MAP_DT, tb.LANE_DT, tb.ROAD_DT, tb.BOX_DT, np.int32, np.int32, tb.EPISODE_DT, tb.SLOT_DT, np.int32]
    names = ["maps", "lanes", "roads", "boxes", "cell_start", "cell_entries", "episodes", "slots", "route_nodes"]
    T = {k: np.zeros(int(n), dt) for k, n, dt in zip(names, sizes, dts)}
    T["route_roads"] = np.zeros(int(sizes[8]), np.int32)
    t, keep = cabi.pack_tables(T)
    cabi.check(engine.lib, engine.lib.pgd_download_tables(engine.h, C.byref(t)))
    T["max_slots"] = int(T["episodes"]["n_slots"].max()) if len(T["episodes"]) else 0
    return T


def compact(T, m):
    """Map / episode ``m`` of fixed-stride tables as a stand-alone table set with offsets rebased to 0 (the form
    tables.TableSet.finish() gives for a single seed)."""
    import numpy as np
    mp, ep = T["maps"][m:m + 1].copy(), T["episodes"][m:m + 1].copy()
    r = mp[0]
    nc = int(r["nx"]) * int(r["ny"]) + 1
    cs = T["cell_start"][r["cell_off"]:r["cell_off"] + nc]
    slots = T["slots"][ep[0]["slot_off"]:ep[0]["slot_off"] + ep[0]["n_slots"]].copy()
    r0 = int(slots["route_off"].min()) if len(slots) else 0
    nr = int((slots["route_off"] + slots["route_len"]).max()) - r0 if len(slots) else 0
    out = dict(
        lanes=T["lanes"][r["lane_off"]:r["lane_off"] + r["n_lanes"]],
        roads=T["roads"][r["road_off"]:r["road_off"] + r["n_roads"]],
        boxes=T["boxes"][r["box_off"]:r["box_off"] + r["n_boxes"]],
        cell_start=cs, cell_entries=T["cell_entries"][r["entry_off"]:r["entry_off"] + int(cs[-1])],
        route_nodes=T["route_nodes"][r0:r0 + nr], route_roads=T["route_roads"][r0:r0 + nr],
    )
    slots["route_off"] -= r0
    for k in ("lane_off", "road_off", "box_off", "cell_off", "entry_off"):
        mp[k] = 0
    ep["map"] = 0
    ep["slot_off"] = 0
    out.update(maps=mp, episodes=ep, slots=slots, max_slots=int(ep[0]["n_slots"]))
    return out
