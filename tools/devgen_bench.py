"""Time the reset path: maps + episode templates of N seeds generated on the GPU (pgd_generate_tables) against the
host Python path (one process).  Usage: python tools/devgen_bench.py [n_seeds ...]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from pgdrive_b200 import devgen, env
    from pgdrive_b200.config import ENGINE_CONFIG, default_config
    sizes = [int(a) for a in sys.argv[1:]] or [100, 1000]
    mc = dict(type="block_num", config=3, lane_num=3, lane_width=3.5, exit_length=50)
    cfg = default_config()
    cfg.update(ENGINE_CONFIG)
    eng = env._Engine(cfg, 1, 32, 0, True)
    gc = devgen.make_gen_config(mc, 0.1)
    out = []
    for n in sizes:
        seeds = list(range(1000, 1000 + n))
        devgen.generate(eng, seeds[:4], gc)  # warm-up (module load, local-memory allocation)
        torch.cuda.synchronize()
        t = time.perf_counter()
        devgen.generate(eng, seeds, gc)
        torch.cuda.synchronize()
        dev_s = time.perf_counter() - t
        k = min(n, 40)
        t = time.perf_counter()
        for s in seeds[:k]:
            env._seed_tables((s, mc, 0.1, ((">", ">>", 0), 5.0, 0.0)))
        host_s = (time.perf_counter() - t) / k * n
        out.append(dict(seeds=n, device_s=round(dev_s, 4), host_python_s_est=round(host_s, 2),
                        maps_per_s_device=round(n / dev_s, 1), speedup=round(host_s / dev_s, 1)))
        print(json.dumps(out[-1]), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
