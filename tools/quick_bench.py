"""Device-resident timing of the step kernel only (A/B of kernel variants):  PGDRIVE_B200_LIB=... python tools/quick_bench.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pgdrive_b200 import VecPGDriveEnv
n = int(os.environ.get("ENVS", 65536)); K = int(os.environ.get("STEPS", 200)); W = int(os.environ.get('WARM', 2048))  # steady state of the episode distribution (profiles/r02i_cost_curve_*.log)
T = bench.build_tables()
env = VecPGDriveEnv(dict(start_seed=1000, environment_num=100, num_envs=n, traffic_density=0.1, num_slots=16), tables_dict=T)
if os.environ.get("BLOCKED") == "1":  # one map per CTA of 32 environments (map-staging experiment)
    env.reset(seeds=1000 + (np.arange(n) // 32) % 100)
else:
    env.reset()
mode = os.environ.get("ACTIONS", "uniform")
g = torch.Generator(device="cuda"); g.manual_seed(1)
NA = 256
a = torch.rand((NA, n, 2), generator=g, device="cuda") * 2 - 1
if mode == "forward":
    a[..., 1] = a[..., 1].abs(); a[..., 0] *= 0.1
for t in range(W): env.step(a[t % NA])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(K): env.step(a[(W + t) % NA])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print("%s actions=%s: %.4f ms/step, %.1f M env-steps/s" % (os.path.basename(os.environ.get("PGDRIVE_B200_LIB", "") or "default"), mode, ms, n / ms / 1e3))

if hasattr(env.engine.lib, "pgd_debug_phase_clocks"):  # diagnostic build (-DPGS_PHASE_CLOCKS)
    import ctypes
    buf = (ctypes.c_ulonglong * 32)()
    env.engine.lib.pgd_debug_phase_clocks(buf, 1)  # forget warm-up and the timed steps above
    for t in range(K): env.step(a[(W + K + t) % NA])
    torch.cuda.synchronize()
    env.engine.lib.pgd_debug_phase_clocks(buf, 1)
    names = ["init", "A", "wait A", "B + wait", "X (role 0)", "-", "-", "wait X", "F", "wait F", "L", "N", "store"]
    ctas = (n + 31) // 32 * K
    tot = sum(buf[:13])
    print("phase clocks (cycles per CTA, thread 0):", ", ".join("%s %.0f" % (nm, buf[i] / ctas) for i, nm in enumerate(names)),
          "| total %.0f" % (tot / ctas),
          "| role 0 in X: sub-steps %.0f, bucket scan %.0f, checkpoints %.0f, reward look-ups = the rest" % (buf[13] / ctas, buf[14] / ctas, buf[15] / ctas), end=" || ")
    nt, nsc = max(buf[16 + 10], 1), max(buf[16 + 11], 1)
    parts = ["take", "load", "IDM", "wait ego + sub-step setup", "sub-steps", "localise", "store"]
    print("traffic warp 1, cycles per batch of <= 32 vehicles:", ", ".join("%s %.0f" % (nm, buf[16 + i] / nt) for i, nm in enumerate(parts)),
          "| awake vehicles per CTA %.1f, vehicle batches of warp 1 per CTA %.2f" % (buf[24] / ctas, nt / ctas))
