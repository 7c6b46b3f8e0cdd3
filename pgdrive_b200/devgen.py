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
        ("spawn_lat", C.c_double), ("fixed_types", C.c_int8 * 32)
    ]


class GenCaps(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("blocks", "lanes", "roads", "boxes", "cells", "entries", "queue", "route",
                                         "cand")]


def make_gen_config(map_config, density, spawn=((">", ">>", 0), 5.0, 0.0)):
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
    c.cells = 1200 * b + 1
    c.entries = 7000 * b
    c.queue = 4096
    c.route = 32 * (3 * b + 10)
    c.cand = 160 * b
    return c
