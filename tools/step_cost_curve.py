"""Kernel time per step as a function of the step index since reset (B200): how long does the episode distribution take
to reach its steady state under each policy?  Prints one line per block of 64 steps.
    ACTIONS=uniform|forward STEPS=4096 python tools/step_cost_curve.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pgdrive_b200 import VecPGDriveEnv
n = int(os.environ.get("ENVS", 65536)); K = int(os.environ.get("STEPS", 4096)); B = 64
T = bench.build_tables()
env = VecPGDriveEnv(dict(start_seed=1000, environment_num=100, num_envs=n, traffic_density=0.1, num_slots=16), tables_dict=T)
env.reset()
mode = os.environ.get("ACTIONS", "uniform")
g = torch.Generator(device="cuda"); g.manual_seed(1)
a = torch.rand((256, n, 2), generator=g, device="cuda") * 2 - 1
if mode == "forward":
    a[..., 1] = a[..., 1].abs(); a[..., 0] *= 0.1
out = []
for b in range(K // B):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(B):
        env.step(a[(b * B + t) % 256])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / B
    info = env.info_numpy()
    out.append((b * B, ms, float(env.done.float().mean()), float((info["episode_length"]).mean())))
    print("%s steps %5d-%5d: %.4f ms/step  %.1f M env-steps/s  done rate %.5f  mean episode length %.0f" % (
        mode, b * B, b * B + B - 1, ms, n / ms / 1e3, out[-1][2], out[-1][3]), flush=True)
