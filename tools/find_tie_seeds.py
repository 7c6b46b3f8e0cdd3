"""Seeds on which the device reset path (host build of pgd_mapgen.cuh, correctly rounded trigonometry) and the
reference-pinned Python path (glibc trigonometry, like the reference itself) produce different tables: the reference
decides discrete things on the last bit of libm results (DESIGN.md "The reset path on the device").  Writes
pgdrive_b200/devgen_ties.json; VecPGDriveEnv(device_mapgen=True) builds exactly these seeds on the host and patches them
into the device tables.
    python tools/find_tie_seeds.py [first] [last] [workers]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

V0 = dict(type="block_num", config=3, lane_num=3, lane_width=3.5, exit_length=50)
SPAWN = ((">", ">>", 0), 5.0, 0.0)
KEYS = ("maps", "roads", "lanes", "boxes", "cell_start", "cell_entries", "episodes", "slots", "route_nodes", "route_roads")


def check(chunk):
    from oracle import mapgen_host
    from pgdrive_b200 import devgen, env
    gc = devgen.make_gen_config(V0, 0.1, SPAWN)
    caps = devgen.caps_for(gc)
    bad = []
    for seed in chunk:
        rc, Th, _ = mapgen_host.generate(seed, gc, caps)
        Tp = env._seed_tables((seed, V0, 0.1, SPAWN))
        if rc != 0 or not all(len(Th[k]) == len(Tp[k]) and Th[k].tobytes() == Tp[k].tobytes() for k in KEYS):
            bad.append(int(seed))
    return bad


if __name__ == "__main__":
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    last = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
    workers = int(sys.argv[3]) if len(sys.argv) > 3 else len(os.sched_getaffinity(0))
    import multiprocessing as mp
    seeds = list(range(first, last))
    chunks = [seeds[i:i + 50] for i in range(0, len(seeds), 50)]
    with mp.get_context("fork").Pool(workers) as pool:
        bad = sorted(sum(pool.map(check, chunks), []))
    out = dict(map_config=V0, traffic_density=0.1, seeds_checked=[first, last], tie_seeds=bad,
               note="host build of the device generator vs the reference-pinned Python path; any byte of any table")
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pgdrive_b200", "devgen_ties.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print(len(bad), "tie seeds of", len(seeds), ":", bad)
