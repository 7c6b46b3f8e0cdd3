"""The reference's own throughput harness (pgdrive/examples/profile_pgdrive.py:6-28) on the drop-in classes.

    python examples/profile_pgdrive.py              # one environment, like the reference: steps/s of env.step
    python examples/profile_pgdrive.py --envs 4096  # the same policy on a batch (device tensors in, device tensors out)
    python examples/profile_pgdrive.py --envs 65536 --host  # numpy in, numpy out: what a gym-style caller sees

Same workload as the reference script: environment_num=1000, start_seed=1010, constant action [0, 1], reset on done."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402


def single(steps):
    from pgdrive_b200 import PGDriveEnv
    env = PGDriveEnv(dict(environment_num=1000, start_seed=1010))
    env.reset()
    start = time.time()
    action = [0.0, 1.]
    for s in range(steps):
        o, r, d, i = env.step(action)
        if d:
            env.reset()
        if (s + 1) % 1000 == 0:
            print("Finish {}/{} simulation steps. Time elapse: {:.4f}. Average FPS: {:.4f}".format(
                s + 1, steps, time.time() - start, (s + 1) / (time.time() - start)))
    print(f"Total Time Elapse: {time.time() - start}")
    env.close()


def batched(n, steps, host=False):
    import torch
    from pgdrive_b200 import VecPGDriveEnv
    env = VecPGDriveEnv(dict(environment_num=100, start_seed=1010, num_envs=n))
    env.reset()
    a = torch.tensor([[0.0, 1.0]], device="cuda").repeat(n, 1)
    if host:  # host arrays: observation rows cross PCIe packed and are expanded by the library's host threads;
        # copy=False returns read-only views of the staging arrays, valid until the next step
        a = a.cpu().numpy()
        torch.cuda.synchronize()
        start = time.time()
        for s in range(steps):
            obs, reward, done, info = env.step(a, copy=False)
        dt = time.time() - start
        print("%d envs x %d steps in %.3f s: %.1f env-steps/s (host arrays)" % (n, steps, dt, n * steps / dt))
        env.close()
        return
    torch.cuda.synchronize()
    start = time.time()
    for s in range(steps):
        env.step(a)  # finished environments restart by themselves at their next step
    torch.cuda.synchronize()
    dt = time.time() - start
    print("%d envs x %d steps in %.3f s: %.1f env-steps/s" % (n, steps, dt, n * steps / dt))
    env.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--host", action="store_true", help="numpy actions in, numpy results out (pgd_step_host)")
    args = ap.parse_args()
    if args.envs == 1:
        single(args.steps)
    else:
        batched(args.envs, args.steps, args.host)
