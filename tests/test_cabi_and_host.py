"""CPU-side checks: the C-ABI library loads and exports every declared symbol, table layouts match the C
structs, config semantics mirror the reference, sharding arithmetic, and the world-size-2 gather under gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from pgdrive_b200 import cabi
    hdr = open(os.path.join(ROOT, "include", "pgdrive_b200.h")).read()
    names = sorted(set(re.findall(r"\b(pgd_[a-z_]+)\s*\(", hdr)))
    assert len(names) >= 12
    lib = ctypes.CDLL(cabi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    lib2 = cabi.load_library()  # prototypes declared; no compute call without a GPU
    assert lib2.pgd_last_error() is not None


def test_struct_sizes_match_header():
    from pgdrive_b200 import cabi, tables
    src = r"""
    #include <stdio.h>
    #include "include/pgdrive_b200.h"
    int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(PgdLane), sizeof(PgdRoad), sizeof(PgdBox),
      sizeof(PgdMap), sizeof(PgdSlot), sizeof(PgdEpisode), sizeof(PgdInfo), sizeof(PgdVehState), sizeof(PgdEnvState),
      sizeof(PgdConfig)); return 0; }
    """
    exe = os.path.join(ROOT, "oracle", "_build", "sizes")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["gcc", "-x", "c", "-", "-I", ROOT, "-o", exe], input=src.encode(), check=True, cwd=ROOT)
    got = [int(x) for x in subprocess.check_output([exe]).split()]
    want = [tables.LANE_DT.itemsize, tables.ROAD_DT.itemsize, tables.BOX_DT.itemsize, tables.MAP_DT.itemsize,
            tables.SLOT_DT.itemsize, tables.EPISODE_DT.itemsize, cabi.INFO_DT.itemsize, cabi.VEH_STATE_DT.itemsize,
            cabi.ENV_STATE_DT.itemsize, ctypes.sizeof(cabi.PgdConfig)]
    assert got == want


def test_generator_struct_sizes_match_header():
    from pgdrive_b200 import devgen
    src = r"""
    #include <stdio.h>
    #include <stddef.h>
    #include "include/pgdrive_b200.h"
    int main(void){ printf("%zu %zu %zu %zu\n", sizeof(PgdGenConfig), sizeof(PgdGenCaps),
      offsetof(PgdGenConfig, lane_width), offsetof(PgdGenConfig, fixed_types)); return 0; }
    """
    exe = os.path.join(ROOT, "oracle", "_build", "gen_sizes")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["gcc", "-x", "c", "-", "-I", ROOT, "-o", exe], input=src.encode(), check=True, cwd=ROOT)
    got = [int(x) for x in subprocess.check_output([exe]).split()]
    assert got == [ctypes.sizeof(devgen.GenConfig), ctypes.sizeof(devgen.GenCaps), devgen.GenConfig.lane_width.offset,
                   devgen.GenConfig.fixed_types.offset]


def test_config_rejects_unknown_keys_like_the_reference():
    from pgdrive_b200 import PGDriveEnv
    from pgdrive_b200.config import default_config
    with pytest.raises(KeyError):
        PGDriveEnv(dict(this_key_does_not_exist=1))
    with pytest.raises(KeyError):
        PGDriveEnv(dict(vehicle_config=dict(lidar=dict(num_beams=3))))
    with pytest.raises(NotImplementedError):
        PGDriveEnv(dict(use_render=True))
    c = default_config()
    assert c["decision_repeat"] == 5 and c["physics_world_step_size"] == 2e-2
    assert c["vehicle_config"]["lidar"]["num_lasers"] == 240
    assert c["traffic_density"] == 0.1 and c["map"] == 3
    env = PGDriveEnv(dict(start_seed=1000, environment_num=100))  # PGDrive-v0
    assert env.observation_space.shape == (274, ) and env.action_space.shape == (2, )
    assert float(env.observation_space.low.min()) == 0.0 and float(env.observation_space.high.max()) == 1.0
    assert env.current_seed is None


def test_registered_ids():
    from pgdrive_b200 import ENVIRONMENTS
    assert ENVIRONMENTS["PGDrive-v0"] == dict(start_seed=1000, environment_num=100)
    assert ENVIRONMENTS["PGDrive-1000envs-v0"] == dict(start_seed=1000, environment_num=1000)  # register.py:24-27


def test_gym_registration_with_a_stand_in_registry(monkeypatch):
    """register.py:42-43 of the reference: every id registered with kwargs=dict(config=...).  gym is not installed in
    this image, so a stand-in module records the calls."""
    import sys
    import types
    calls = {}
    reg = types.ModuleType("gym.envs.registration")
    reg.registry = {}
    reg.register = lambda id, entry_point, kwargs: calls.__setitem__(id, (entry_point, kwargs))
    gym = types.ModuleType("gym")
    envs = types.ModuleType("gym.envs")
    for name, mod in (("gym", gym), ("gym.envs", envs), ("gym.envs.registration", reg)):
        monkeypatch.setitem(sys.modules, name, mod)
    gym.envs, envs.registration = envs, reg
    from pgdrive_b200 import register
    monkeypatch.setattr(register, "registered_with", [])
    assert "gym" in register.register_all()
    assert set(calls) == set(register.get_env_list()) and len(calls) == 8
    assert calls["PGDrive-1000envs-v0"] == ("pgdrive_b200.env:PGDriveEnv",
                                            dict(config=dict(start_seed=1000, environment_num=1000)))


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pgdrive_b200 import PGDriveEnv
    env = PGDriveEnv(dict(start_seed=1000, environment_num=1))
    with pytest.raises(RuntimeError):
        env.reset(force_seed=1000)


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle/_build|libpgd_oracle|pgd_oracle\.c", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pgdrive_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def test_shard_ranges_cover_the_batch():
    from pgdrive_b200.sharding import seed_of_env, shard_range
    for total, world in ((524288, 8), (65536, 1), (10, 4), (7, 8)):
        cuts = [shard_range(total, world, r) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        assert max(hi - lo for lo, hi in cuts) - min(hi - lo for lo, hi in cuts) <= 1
    assert shard_range(524288, 8, 3) == (196608, 262144)
    assert seed_of_env(1234, 1000, 100) == 1034
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_balanced_shards_give_rank_0_less_to_simulate():
    """sharding.balanced_sizes: sizes sum to the total, whole CTAs (multiples of 32), ranks 1.. within one granule of
    each other, rank 0 never above an equal share and smaller the more ranks it gathers from."""
    from pgdrive_b200.sharding import ROW_COST_NS, balanced_sizes, sizes_with_rank0
    assert balanced_sizes(65536, 1) == [65536]
    last = None
    for world in (2, 3, 4, 8):
        for mode in ("sparse", "copy", "peer"):
            total = world * 65536
            s = balanced_sizes(total, world, mode)
            assert len(s) == world and sum(s) == total and all(v > 0 and v % 32 == 0 for v in s)
            assert max(s[1:]) - min(s[1:]) <= 32 and s[0] <= 65536 <= min(s[1:])
            c = ROW_COST_NS
            pack, expand = (c["pack"], c["expand"]) if mode == "sparse" else (0.0, 0.0)
            t0 = c["step"] * s[0] + expand * (total - s[0]) + c["consume"] * total + c["fixed"]
            t1 = (c["step"] + pack) * max(s[1:])
            assert t0 <= t1 * 1.01 or s[0] == 4096  # balanced, or rank 0 already at its floor
        now = balanced_sizes(world * 65536, world, "sparse")[0]
        assert last is None or now <= last
        last = now
    assert balanced_sizes(8 * 65536, 8, "sparse", cost=dict(step=1e6)) == [65536] * 8  # gather work negligible: equal
    assert sizes_with_rank0(4 * 4096, 4, 1024) == [1024, 5120, 5120, 5120]
    assert sum(sizes_with_rank0(8 * 65536, 8, 4096)) == 8 * 65536
    for bad in ((100, 2), (64, 4)):
        with pytest.raises(ValueError):
            balanced_sizes(*bad)
    with pytest.raises(ValueError):
        sizes_with_rank0(4096, 2, 100)


GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from pgdrive_b200.sharding import GatherBuffers, shard_range
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"], rank=rank, world_size=world)
n = 6
buf = GatherBuffers(torch, n, world, rank, "cpu", obs_dim=274)
lo, hi = shard_range(world * n, world, rank)
# what the step kernel would write: rows of this rank only (pattern = global env index)
idx = torch.arange(lo, hi, dtype=torch.float32)
buf.local(buf.obs).copy_(idx[:, None].expand(n, 274))
buf.local(buf.reward).copy_(idx * 0.5)
buf.local(buf.done).copy_((idx %% 2).to(torch.uint8))
buf.all_gather(dist)
full = torch.arange(world * n, dtype=torch.float32)
assert torch.equal(buf.obs, full[:, None].expand(world * n, 274)), "obs"
assert torch.equal(buf.reward, full * 0.5), "reward"
assert torch.equal(buf.done, (full %% 2).to(torch.uint8)), "done"
dist.barrier()
dist.destroy_process_group()
print("rank %%d ok" %% rank)
"""


def test_world_size_2_gather_under_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % ROOT)
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o


def test_abi_argument_errors_without_a_gpu():
    """Error behaviour of the C-ABI that does not need a device: negative code + message (include/pgdrive_b200.h)."""
    import __graft_entry__
    __graft_entry__.build()
    from pgdrive_b200 import cabi
    lib = cabi.load_library()
    h = ctypes.c_void_p()
    assert lib.pgd_create(None, 0, ctypes.byref(h)) == -1
    assert b"null" in lib.pgd_last_error()
    bad = cabi.make_config(0)
    assert lib.pgd_create(ctypes.byref(bad), 0, ctypes.byref(h)) == -1
    assert b"num_envs" in lib.pgd_last_error()
    bad = cabi.make_config(4, num_slots=12)
    assert lib.pgd_create(ctypes.byref(bad), 0, ctypes.byref(h)) == -1
    assert b"num_slots" in lib.pgd_last_error()
    assert lib.pgd_step(None, None, None, None, None, None, None) == -1
    assert lib.pgd_reset(None, None, None, 0, None, None, None) == -1
    assert lib.pgd_destroy(None) == 0
    assert lib.pgd_launch_count(None) == 0


def test_vec_env_rejects_unsupported_options():
    from pgdrive_b200 import VecPGDriveEnv
    for cfg in (dict(random_traffic=True),  # (PGDriveEnv supports it: the traffic templates are per reset there)
                dict(num_agents=2), dict(vehicle_config=dict(enable_reverse=True)),
                dict(vehicle_config=dict(overtake_stat=True)),
                dict(vehicle_config=dict(lidar=dict(num_lasers=120))),
                dict(vehicle_config=dict(side_detector=dict(num_lasers=500)))):
        with pytest.raises(NotImplementedError):
            VecPGDriveEnv(cfg)
    with pytest.raises(ValueError):
        VecPGDriveEnv(dict(traffic_mode="no such mode"))  # traffic_manager.py:69
    with pytest.raises(AssertionError):  # pgdrive_env.py:143-146: "You already provide config!"
        VecPGDriveEnv(dict(gaussian_noise=0.1, vehicle_config=dict(lidar=dict(gaussian_noise=0.2))))
    with pytest.raises(KeyError):
        VecPGDriveEnv(dict(no_such_key=True))


def test_discrete_actions_follow_the_reference_conversion():
    """policy/env_input_policy.py:17-31: clip to [-1, 1] first, then index * unit - 1 (so indices >= 1 coincide)."""
    from pgdrive_b200 import PGDriveEnv
    from pgdrive_b200.env import discrete_to_continuous
    env = PGDriveEnv(dict(discrete_action=True))
    assert repr(env.action_space) == "MultiDiscrete([5, 5])"
    assert env.action_space.contains(env.action_space.sample())
    got = discrete_to_continuous(np.array([[0, 0], [1, 4], [3, 2], [4, 1]]), env.config)
    np.testing.assert_allclose(got, [[-1, -1], [-0.5, -0.5], [-0.5, -0.5], [-0.5, -0.5]])
    env7 = PGDriveEnv(dict(discrete_action=True, discrete_steering_dim=3, discrete_throttle_dim=9))
    np.testing.assert_allclose(discrete_to_continuous(np.array([1, 1]), env7.config), [0.0, -0.75])


def test_effective_horizon_combines_horizon_and_auto_termination():
    """base_env.py:190-192 (horizon) and :318-326 (auto_termination: 250 steps per block, first block included)."""
    from pgdrive_b200.config import check_supported, default_config
    from pgdrive_b200.env import effective_horizon, parse_map_config
    cfg = default_config()
    mc = parse_map_config(cfg)
    assert effective_horizon(cfg, mc) == 0
    cfg.update(dict(auto_termination=True))
    check_supported(cfg)
    assert effective_horizon(cfg, mc) == 1000  # map = 3 blocks + the first block
    cfg.update(dict(horizon=300))
    assert effective_horizon(cfg, mc) == 300
    cfg.update(dict(horizon=5000, map="SCrRX"))
    assert effective_horizon(cfg, parse_map_config(cfg)) == 1500
    cfg.update(dict(auto_termination=False))
    assert effective_horizon(cfg, parse_map_config(cfg)) == 5000


def test_ego_view_and_map_views_mirror_the_reference_properties():
    """env.vehicle / env.current_map stand-ins (envs/base_env.py:371-462, base_vehicle.py:390-425,683-698), driven by
    a synthetic state record (no GPU)."""
    from pgdrive_b200 import cabi, mapgen
    from pgdrive_b200.env import EgoView
    st = np.zeros(1, cabi.ENV_STATE_DT)
    st["veh"][0][0]["x"], st["veh"][0][0]["y"] = 12.5, -3.0
    st["veh"][0][0]["heading"], st["veh"][0][0]["speed"] = 0.5, 10.0
    st["veh"][0][0]["lane"], st["veh"][0][0]["flags"] = 4, cabi.V_ALIVE | cabi.V_ACTIVE | cabi.V_ON_LANE
    m = mapgen.generate_map(1000)
    flat = [(f, t, i) for (f, t), lanes in m.net.roads() for i in range(len(lanes))]
    v = EgoView(lambda: st, lambda: cabi.F_ON_LANE | cabi.F_ON_BROKEN, lambda k: flat[k], lambda: ("a", "b"), (">", ">>"))
    assert v.position.tolist() == [12.5, -3.0] and abs(v.speed - 36.0) < 1e-5 and v.heading_theta == 0.5
    assert np.allclose(v.velocity, 36.0 * np.array([np.cos(0.5), np.sin(0.5)]))
    assert v.lane_index == flat[4] and v.on_lane and v.on_broken_line and not v.crash_vehicle
    s = v.get_state()
    assert s["done"] is False and s["destination"] == ("a", "b") and s["spawn_road"] == (">", ">>")
    assert m.num_blocks == 4 and m.road_network is m.net
    assert [b["id"] for b in m.save_map()["block_sequence"]] == [b.id for b in m.blocks]


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU oracle port on the host cores): one JSON line with the contract's keys, the
    SAME `config` object as the CUDA arm prints for the same arguments, `e2e` equal to the line's own value."""
    import argparse
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--envs", "1024"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "env-steps/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == dict(value=d["value"], unit="env-steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    sys.path.insert(0, ROOT)
    import bench
    _, _, n_slots, desc = bench.WORKLOADS["v0"]
    assert d["config"] == bench.line_config(argparse.Namespace(envs=1024), 1, n_slots, desc)


def test_clock_sampler_counts_the_samples_of_the_loaded_window():
    """bench.ClockSampler: samples that arrive between mark() and stop() count; when none did (nvidia-smi too slow to
    start) it falls back to all samples and says so; throttle reasons are collected."""
    import time
    sys.path.insert(0, ROOT)
    import bench

    class Proc:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            pass

    c = bench.ClockSampler(0)
    assert c.stop()["reasons"] == ["nvidia-smi unavailable"]  # never started
    c.proc = Proc()
    now = time.time()
    c.rows = [(now - 2.0, "1000, 1965, 300, Not Active, Not Active, Not Active, Not Active"),
              (now - 0.5, "1965, 1965, 700, Not Active, Not Active, Not Active, Active"),
              (now - 0.2, "1950, 1965, 700, Not Active, Not Active, Not Active, Not Active"),
              (now - 0.1, "garbage")]
    c.t0 = now - 1.0
    r = c.stop()
    assert r["samples"] == 2 and r["sm_mhz"] == 1957.5 and r["sm_max_mhz"] == 1965.0 and r["reasons"] == ["sw_power_cap"]
    assert r["window"].startswith("pre-roll")
    c.t0 = now + 60
    r = c.stop()
    assert r["samples"] == 3 and r["window"].startswith("all samples")
