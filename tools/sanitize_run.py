"""Run the host builds of the generator, the step and the oracle under ASan + UBSan."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import mapgen_host as mh, oracle as orc, step_host as sh
SAN = os.path.join(ROOT, "oracle", "_san")
mh.LIB = os.path.join(SAN, "libpgd_mapgen_host.so"); mh.build = lambda force=False: mh.LIB
orc.LIB = os.path.join(SAN, "libpgd_oracle.so"); orc.build = lambda force=False: orc.LIB
sh.LIB = os.path.join(SAN, "libpgd_step_host.so")
import subprocess
sh.subprocess = type("S", (), {"check_call": staticmethod(lambda *a, **k: 0), "DEVNULL": None})
from pgdrive_b200 import devgen, env as E
V0 = dict(type="block_num", config=3, lane_num=3, lane_width=3.5, exit_length=50)
SP = ((">", ">>", 0), 5.0, 0.0)
# generator: many seeds / configs
n = 0
for mc, dens, seeds in [(V0, 0.1, range(1000, 1150)), (dict(V0, lane_num=1), 0.2, range(0, 30)), (dict(V0, lane_num=4, lane_width=3.0), 0.1, range(0, 30)),
                        (dict(V0, config=8), 0.05, range(0, 20)), (dict(V0, type="block_sequence", config="SCrRXTO"), 0.1, range(0, 10)), (V0, 0.4, range(5000, 5020))]:
    gc = devgen.make_gen_config(mc, dens, SP); caps = devgen.caps_for(gc)
    for s in seeds:
        rc, T, seq = mh.generate(s, gc, caps); n += 1
# tight caps: overflow paths
gc = devgen.make_gen_config(V0, 0.1, SP)
for field, val in [("lanes", 30), ("roads", 10), ("boxes", 200), ("cells", 100), ("entries", 500), ("queue", 8), ("route", 20), ("cand", 10)]:
    caps = devgen.caps_for(gc); setattr(caps, field, val)
    for s in range(1000, 1010):
        rc, T, seq = mh.generate(s, gc, caps); n += 1
print("generator runs under sanitizers:", n)
# step + oracle rollouts
seeds = list(range(1000, 1030))
T = E.merge_tables([E._seed_tables((s, V0, 0.1, SP)) for s in seeds])
a = orc.Oracle(T, 90, auto_reset=True, num_slots=16); b = sh.HostStep(T, 90, auto_reset=True, num_slots=16)
eps = [i % 30 for i in range(90)]
a.reset(range(90), eps); b.reset(range(90), eps)
rs = np.random.RandomState(0)
for t in range(250):
    act = rs.uniform(-1, 1, (90, 2)).astype(np.float32); act[:, 1] = np.abs(act[:, 1]); act[:, 0] *= 0.1
    r1 = a.step(act); r2 = b.step(act)
    assert np.array_equal(r1[0], r2[0])
cfg = dict(auto_reset=True, n_side=12, side_distance=50.0, n_lane_line=8, lane_line_distance=20.0)
a = orc.Oracle(T, 30, num_slots=16, **cfg); b = sh.HostStep(T, 30, num_slots=16, **cfg)
a.reset(range(30), range(30)); b.reset(range(30), range(30))
for t in range(100):
    act = rs.uniform(-1, 1, (30, 2)).astype(np.float32); act[:, 1] = np.abs(act[:, 1])
    assert np.array_equal(a.step(act)[0], b.step(act)[0])
print("step / oracle rollouts under sanitizers: ok")
