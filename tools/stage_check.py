"""Map-staging experiment (libvar_stage.so): every CTA's 32 environments play ONE seed, so the lane table is staged in
shared memory by TMA; the rollout must stay bit-identical to the oracle.  PGDRIVE_B200_LIB=... python tools/stage_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.oracle import Oracle
from pgdrive_b200 import VecPGDriveEnv
n = 128
env = VecPGDriveEnv(dict(start_seed=1000, environment_num=4, num_envs=n))
seeds = 1000 + (np.arange(n) // 32) % 4
obs = env.reset(seeds=seeds).cpu().numpy()
ref = Oracle(env.T, n, auto_reset=True, num_slots=env.engine.num_slots)
ro = ref.reset(range(n), [env.episode_of_seed[int(s)] for s in seeds])
assert np.array_equal(obs, ro)
rs = np.random.RandomState(3)
for t in range(300):
    a = np.c_[np.clip((ro[:, 2] - 0.5) * 8.0 + rs.uniform(-0.05, 0.05, n), -1, 1), np.where(ro[:, 3] < 0.3, 0.6, 0.0)].astype(np.float32)
    o, r, d, _ = env.step(torch.from_numpy(a).cuda())
    ro, rr, rd, _ = ref.step(a)
    assert np.array_equal(o.cpu().numpy(), ro) and np.array_equal(r.cpu().numpy(), rr) and np.array_equal(d.cpu().numpy(), rd), t
    ro = ro.copy()
print("staged rollout bit-identical to the oracle:", os.path.basename(os.environ.get("PGDRIVE_B200_LIB", "default")))
