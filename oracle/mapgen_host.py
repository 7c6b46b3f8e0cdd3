"""TEST INFRASTRUCTURE ONLY: ctypes loader of oracle/_build/libpgd_mapgen_host.so, the host (g++) build of the device
map generator's source.  Used by tests to check the generator's logic without a GPU and as the bit-exact expectation
for the sm_100a build.  Nothing under pgdrive_b200/ imports this."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libpgd_mapgen_host.so")
_lib = None


def build(force=False):
    src = [os.path.join(HERE, "mapgen_host.cpp")] + [
        os.path.join(HERE, "..", "pgdrive_b200", "csrc", f) for f in ("pgd_mapgen.cuh", "pgd_rng.cuh", "pgd_dd.cuh")
    ] + [os.path.join(HERE, "..", "include", "pgd_tables.h")]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(s) for s in src):
        return LIB
    subprocess.check_call(["make", "-C", HERE, "_build/libpgd_mapgen_host.so"], stdout=subprocess.DEVNULL)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.pgd_host_hash_seed.restype = C.c_uint64
        _lib.pgd_host_hash_seed.argtypes = [C.c_uint64]
        for f in ("pgd_host_sin", "pgd_host_cos", "pgd_host_atan"):
            getattr(_lib, f).restype = C.c_double
            getattr(_lib, f).argtypes = [C.c_double]
        _lib.pgd_host_atan2.restype = C.c_double
        _lib.pgd_host_atan2.argtypes = [C.c_double, C.c_double]
        _lib.pgd_host_rng_script.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib.pgd_hostgen.argtypes = [C.c_uint64] + [C.c_void_p] * 14
    return _lib


def rng_script(seed, ops, probs=None):
    """ops: list of (op, arg); returns the list of doubles the generator's RNG produced."""
    o = np.array([a for a, _ in ops], np.int32)
    a = np.array([b for _, b in ops], np.int32)
    n_out = int(sum(b if op == 3 else 1 for op, b in ops))
    out = np.zeros(max(n_out, 1), np.float64)
    p = np.ascontiguousarray(probs if probs is not None else [1.0], np.float64)
    k = lib().pgd_host_rng_script(int(seed), o.ctypes.data, a.ctypes.data, len(ops), p.ctypes.data, out.ctypes.data)
    return out[:k]


def generate(seed, gen_config, caps):
    """Run the host build for one seed.  ``gen_config`` / ``caps``: ctypes structs of pgdrive_b200.devgen.
    Returns (status, tables dict trimmed to the real sizes, sequence array)."""
    from pgdrive_b200 import tables as tb
    maps = np.zeros(1, tb.MAP_DT)
    lanes = np.zeros(caps.lanes, tb.LANE_DT)
    roads = np.zeros(caps.roads, tb.ROAD_DT)
    boxes = np.zeros(caps.boxes, tb.BOX_DT)
    cell_start = np.zeros(caps.cells, np.int32)
    cell_entries = np.zeros(caps.entries, np.int32)
    eps = np.zeros(1, tb.EPISODE_DT)
    slots = np.zeros(32, tb.SLOT_DT)
    rn = np.zeros(caps.route, np.int32)
    rr = np.zeros(caps.route, np.int32)
    counts = np.zeros(8, np.int32)
    seq = np.zeros((caps.blocks, 16), np.int32)
    rc = lib().pgd_hostgen(
        int(seed), C.addressof(gen_config), C.addressof(caps), maps.ctypes.data, lanes.ctypes.data, roads.ctypes.data,
        boxes.ctypes.data, cell_start.ctypes.data, cell_entries.ctypes.data, eps.ctypes.data, slots.ctypes.data,
        rn.ctypes.data, rr.ctypes.data, counts.ctypes.data, seq.ctypes.data
    )
    nl, nr, nb, nc, ne, ns, nrt, nblk = [int(v) for v in counts]
    T = dict(maps=maps, lanes=lanes[:nl], roads=roads[:nr], boxes=boxes[:nb], cell_start=cell_start[:nc],
             cell_entries=cell_entries[:ne], episodes=eps, slots=slots[:ns], route_nodes=rn[:nrt],
             route_roads=rr[:nrt])
    T["max_slots"] = ns
    return rc, T, seq[:nblk]
