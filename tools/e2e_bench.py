"""One GPU: the host-buffer step (VecPGDriveEnv.step(numpy, copy=False) -> pgd_step_host) at 65 536 environments in the
steady state, for a few settings of the host pool / chunking and against dense rows.  Usage: python tools/e2e_bench.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(tag, envvars, n=65536, preroll=1024, steps=48):
    import numpy as np
    import torch
    from pgdrive_b200 import VecPGDriveEnv
    for k in ("PGDRIVE_B200_HOST_DENSE", "PGDRIVE_B200_HOST_THREADS", "PGDRIVE_B200_HOST_CHUNKS"):
        os.environ.pop(k, None)
    os.environ.update(envvars)
    env = VecPGDriveEnv(dict(num_envs=n, start_seed=1000, environment_num=100, traffic_density=0.1))
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    acts = torch.rand((256, n, 2), generator=g, device="cuda") * 2 - 1
    for t in range(preroll):
        env.step(acts[t % 256])
    h = acts[:steps].cpu().numpy()
    for t in range(4):
        env.step(h[t], copy=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(steps):
        o, r, d, i = env.step(h[t], copy=False)
    dt = time.perf_counter() - t0
    h2d, d2h = env.host_transfer_bytes()
    rec = dict(tag=tag, env=envvars, ms_per_step=dt / steps * 1e3, env_steps_per_s=n * steps / dt, h2d=h2d, d2h=d2h,
               checksum=float(o[:, :8].sum()))
    print(json.dumps(rec), flush=True)
    env.close()


if __name__ == "__main__":
    run("default", {})
    run("dense rows", {"PGDRIVE_B200_HOST_DENSE": "1"})
    for th in sys.argv[1:2] and sys.argv[1].split(","):
        run("threads " + th, {"PGDRIVE_B200_HOST_THREADS": th})
    for ch in sys.argv[2:3] and sys.argv[2].split(","):
        run("chunks " + ch, {"PGDRIVE_B200_HOST_CHUNKS": ch})
