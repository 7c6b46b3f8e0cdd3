"""More GPU parity rollouts of the step kernel against the CPU oracle, bit for bit, through the C-ABI: 32 and 24 vehicle
slots, horizon, partial CTAs, and the full-size timing smoke."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rollouts(n, seeds, steps, policy, density=0.1, **cfg):
    import torch
    from test_gpu_parity import _pair, _reset_both, _rollout
    env, ref = _pair(n, seeds, density=density, **cfg)
    obs, ro = _reset_both(env, ref)
    assert np.array_equal(obs, ro)
    rs = np.random.RandomState(2)

    def act(t):
        a = rs.uniform(-1, 1, (n, 2))
        if policy == "forward":
            a[:, 1] = np.abs(a[:, 1])
            a[:, 0] *= 0.1
        elif policy == "lane":
            a[:, 0] = 0.0
            a[:, 1] = 0.6
        return a

    dones = _rollout(env, ref, steps, act)
    env.close()
    return dones


def test_step_random_and_forward_policies_match_oracle():
    seeds = list(range(1000, 1100))
    _rollouts(400, seeds, 100, "uniform")
    assert _rollouts(400, seeds, 250, "forward") > 0


def test_step_lane_following_meets_traffic():
    assert _rollouts(300, list(range(1000, 1100)), 350, "lane") > 0


def test_step_32_and_24_slots_and_horizon():
    _rollouts(96, list(range(1000, 1012)), 200, "lane", density=0.2)
    _rollouts(70, list(range(1000, 1012)), 120, "lane", density=0.1, num_slots=24)  # 70 envs: a partly empty CTA
    _rollouts(64, [1000, 1001], 40, "forward", horizon=9, auto_reset=False)


def test_step_full_size_throughput_smoke():
    """65 536 environments: runs, replicas of a seed stay identical, and the time per step is printed."""
    import torch
    from pgdrive_b200 import VecPGDriveEnv
    n = 65536
    env = VecPGDriveEnv(dict(start_seed=1000, environment_num=100, num_envs=n, traffic_density=0.1,
                             ))
    env.reset()
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    acts = torch.rand((60, n, 2), generator=g, device="cuda") * 2 - 1
    acts[:, 100:200] = acts[:, :100]  # envs 100..199 replay envs 0..99 (same seeds: i % 100)
    for t in range(20):
        env.step(acts[t])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(20, 60):
        obs, r, d, _ = env.step(acts[t])
    e1.record()
    torch.cuda.synchronize()
    print("role per warp: %.4f ms / step" % (e0.elapsed_time(e1) / 40))
    assert torch.equal(obs[:100], obs[100:200])
    env.close()


def _free_run(env, ref, n, steps, act_fn):
    import torch
    from test_gpu_parity import FLAG_MASK
    obs = env.reset().cpu().numpy().copy()
    ro = ref.reset(range(n), [env.episode_of_seed[int(s)] for s in env.env_seeds]).copy()
    assert np.array_equal(obs, ro)
    seen = 0
    for t in range(steps):
        a = act_fn(t, ro).astype(np.float32)
        o, r, d, _ = env.step(torch.from_numpy(a).cuda())
        o, r, d = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
        info = env.info_numpy()
        ro, rr, rd, rinfo = ref.step(a)
        assert np.array_equal(d, rd) and np.array_equal(o, ro) and np.array_equal(r, rr), t
        assert info.tobytes() == rinfo.tobytes(), t
        seen |= int(np.bitwise_or.reduce(info["flags"]))
        ro = ro.copy()
    return seen


def test_step_options_respawn_noise_increment_and_accident_scenes():
    """The options added in round 2, on the GPU against the oracle bit for bit: traffic_mode="respawn" (all traffic awake
    from step 0), lidar gaussian noise + dropout with increment_steering, SafePGDriveEnv's accident scenes with and
    without safe_rl_env."""
    from oracle.oracle import Oracle
    from pgdrive_b200 import VecPGDriveEnv
    rs = np.random.RandomState(5)

    def hold_lane(t, obs):
        a = np.zeros((obs.shape[0], 2))
        a[:, 0] = np.clip((obs[:, 2] - 0.5) * 8.0 + rs.uniform(-0.02, 0.02, obs.shape[0]), -1, 1)
        a[:, 1] = np.where(obs[:, 3] < 0.25, 0.5, 0.0)
        return a

    # respawn mode: seeds whose maps need at most 32 slots
    n = 64
    env = VecPGDriveEnv(dict(start_seed=1002, environment_num=3, num_envs=n, traffic_mode="respawn"))
    ref = Oracle(env.T, n, auto_reset=True, num_slots=env.engine.num_slots)
    _free_run(env, ref, n, 150, hold_lane)
    env.close()
    # lidar noise + dropout, incremental steering
    env = VecPGDriveEnv(dict(start_seed=1000, environment_num=20, num_envs=n, noise_seed=5,
                             vehicle_config=dict(increment_steering=True, lidar=dict(gaussian_noise=0.05, dropout_prob=0.1))))
    ref = Oracle(env.T, n, auto_reset=True, num_slots=env.engine.num_slots, lidar_gaussian_noise=0.05,
                 lidar_dropout_prob=0.1, noise_seed=5, increment_steering=True)
    _free_run(env, ref, n, 150, lambda t, obs: np.c_[rs.uniform(-1, 1, n), rs.uniform(0, 1, n)])
    assert float((env.obs[:, 34:] == 0).float().mean()) > 0.05
    env.close()
    # accident scenes
    for safe in (False, True):
        env = VecPGDriveEnv(dict(start_seed=100, environment_num=12, num_envs=n, traffic_density=0.05, accident_prob=0.8,
                                 object_manager=True, safe_rl_env=safe,
                                 vehicle_config=dict(spawn_lane_index=(">", ">>", 2))))
        assert (env.T["slots"]["type"] >= 5).sum() > 40
        ref = Oracle(env.T, n, auto_reset=True, num_slots=env.engine.num_slots, safe_rl_env=safe)
        seen = _free_run(env, ref, n, 400, hold_lane)
        assert seen & 2048, "no traffic object was touched"
        env.close()


def test_random_traffic_redraws_the_traffic_at_every_reset():
    """Mirrors test_random_engine.py:75-99: random_traffic + respawn mode, 20 resets of the same seed -- the traffic
    vehicles do not stand where they stood the episode before."""
    from pgdrive_b200 import PGDriveEnv
    env = PGDriveEnv({"random_traffic": True, "traffic_mode": "respawn", "traffic_density": 0.3, "start_seed": 5})
    try:
        last, moved, has_traffic = None, 0, False
        for i in range(20):
            env.reset(force_seed=5)
            st = env.get_state()["veh"][0]
            n = int(env._parts[5]["max_slots"])
            pos = [(float(v["x"]), float(v["y"])) for v in st[1:n]]
            has_traffic = has_traffic or len(pos) > 0
            if last is not None and len(pos) == len(last):
                moved += sum(np.hypot(a[0] - b[0], a[1] - b[1]) >= 0.5 for a, b in zip(last, pos)) > 0
            last = pos
        assert has_traffic and moved >= 15
        # without the option the same seed gives the same traffic
        fixed = PGDriveEnv({"traffic_mode": "respawn", "traffic_density": 0.3, "start_seed": 5})
        fixed.reset(force_seed=5)
        a = fixed.get_state()["veh"][0][1:n]["x"].copy()
        fixed.reset(force_seed=5)
        assert np.array_equal(a, fixed.get_state()["veh"][0][1:n]["x"])
        fixed.close()
    finally:
        env.close()


def test_packed_rows_round_trip_bit_for_bit():
    """pgd_pack_rows / pgd_expand_rows (the gather's wire format: head + 240-bit hit mask + beams that are not 1.0):
    real observation rows, rows with every beam hit, rows with none, and special bit patterns come back exactly."""
    import torch
    from pgdrive_b200 import VecPGDriveEnv, cabi
    env = VecPGDriveEnv(dict(num_envs=300, start_seed=1000, environment_num=20))
    try:
        env.reset()
        g = torch.Generator(device="cuda"); g.manual_seed(3)
        for t in range(150):
            a = torch.rand((300, 2), generator=g, device="cuda") * 2 - 1
            a[:, 1] = a[:, 1].abs(); a[:, 0] *= 0.1
            obs = env.step(a)[0]
        rows = obs.clone()
        assert bool((rows[:, -240:] < 1.0).any()), "no lidar hit in 300 environments after 150 steps"
        d = rows.shape[1]
        extra = torch.rand((5, d), device="cuda")
        extra[0, -240:] = 1.0                      # no hit at all
        extra[1, -240:] = 0.25                     # every beam hit
        extra[2, -240:] = torch.where(torch.arange(240, device="cuda") % 2 == 0, 1.0, 0.0)  # 0.0 is a hit
        extra[3, -240:] = 1.0; extra[3, -1] = 0.5; extra[3, -240] = 0.75                    # first and last beam only
        extra[4, -240:] = torch.nextafter(torch.ones(240, device="cuda"), torch.zeros(240, device="cuda"))  # 1 - ulp
        rows = torch.cat([rows, extra]).contiguous()
        n = rows.shape[0]
        lib = env.engine.lib
        stride = lib.pgd_packed_row_words(d)
        assert stride == (d + 8 + 31) // 32 * 32 and lib.pgd_packed_row_words(100) == -1
        packed = torch.full((n, stride), float("nan"), device="cuda")
        back = torch.full((n, d), float("nan"), device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        cabi.check(lib, lib.pgd_pack_rows(rows.data_ptr(), packed.data_ptr(), n, d, st))
        cabi.check(lib, lib.pgd_expand_rows(packed.data_ptr(), back.data_ptr(), n, d, st))
        torch.cuda.synchronize()
        same = rows.view(torch.int32) == back.view(torch.int32)
        assert bool(same.all()), "rows differ after the round trip: %s" % (~same).nonzero()[:8].tolist()
        # what crosses the link: head + mask + hits
        hits = (rows[:, -240:].view(torch.int32) != 0x3f800000).sum(1)
        mask_words = packed[:, d - 240:d - 232].view(torch.int32)
        pop = sum(((mask_words >> b) & 1) for b in range(32)).sum(1)
        assert torch.equal(pop.to(torch.int64), hits.to(torch.int64)), (pop[:8].tolist(), hits[:8].tolist())
        # what is written: head + mask + hits, zero-filled to a 32-byte boundary; nothing behind it
        used = d - 240 + 8 + hits
        col = torch.arange(stride, device="cuda")[None, :]
        pad = (col >= used[:, None]) & (col < ((used + 7) // 8 * 8)[:, None])
        assert bool((packed[pad] == 0).all()) and bool(torch.isnan(packed[col >= ((used + 7) // 8 * 8)[:, None]]).all())
        # the unaligned paths: rows that start 4 bytes off a 16-byte boundary on either side
        flat_r = torch.empty(n * d + 1, device="cuda"); flat_p = torch.empty(n * stride + 1, device="cuda")
        flat_b = torch.full((n * d + 1, ), float("nan"), device="cuda")
        flat_r[1:] = rows.reshape(-1)
        cabi.check(lib, lib.pgd_pack_rows(flat_r[1:].data_ptr(), flat_p[1:].data_ptr(), n, d, st))
        cabi.check(lib, lib.pgd_expand_rows(flat_p[1:].data_ptr(), flat_b[1:].data_ptr(), n, d, st))
        torch.cuda.synchronize()
        assert torch.equal(flat_b[1:].view(torch.int32), rows.reshape(-1).view(torch.int32))
        assert lib.pgd_pack_rows(None, packed.data_ptr(), n, d, st) == -1
        assert lib.pgd_expand_rows(packed.data_ptr(), None, n, d, st) == -1
    finally:
        env.close()


def test_delta_expansion_equals_full_expansion():
    """pgd_expand_rows_delta: expanding a sequence of steps into the SAME buffer, storing only what changed since the rows
    the buffer holds, gives bit for bit what pgd_expand_rows gives -- hits appearing, moving and disappearing, rows with
    every beam hit, and a buffer that starts as NaN (first use: full)."""
    import torch
    from pgdrive_b200 import VecPGDriveEnv, cabi
    n = 1000
    env = VecPGDriveEnv(dict(num_envs=n, start_seed=1000, environment_num=20))
    try:
        lib = env.engine.lib
        st = torch.cuda.current_stream().cuda_stream
        env.reset()
        g = torch.Generator(device="cuda"); g.manual_seed(11)
        d = env.obs.shape[1]
        stride = lib.pgd_packed_row_words(d)
        packed = torch.full((n, stride), float("nan"), device="cuda")
        dense = torch.full((n, d), float("nan"), device="cuda")      # the buffer that is expanded into, again and again
        want = torch.empty((n, d), device="cuda")
        state = torch.zeros((n, 8), dtype=torch.int32, device="cuda")
        seen_hits = 0
        for t in range(120):
            a = torch.rand((n, 2), generator=g, device="cuda") * 2 - 1
            a[:, 1] = a[:, 1].abs(); a[:, 0] *= 0.1
            rows = env.step(a)[0].clone()
            if t % 10 == 3:
                rows[::7, -240:] = torch.rand((len(rows[::7]), 240), generator=g, device="cuda")  # every beam a hit
            if t % 10 == 4:
                rows[::7, -240:] = 1.0                                                              # ... and gone again
            seen_hits += int((rows[:, -240:] != 1.0).sum())
            cabi.check(lib, lib.pgd_pack_rows(rows.data_ptr(), packed.data_ptr(), n, d, st))
            cabi.check(lib, lib.pgd_expand_rows(packed.data_ptr(), want.data_ptr(), n, d, st))
            cabi.check(lib, lib.pgd_expand_rows_delta(packed.data_ptr(), dense.data_ptr(), state.data_ptr(), n, d,
                                                      1 if t == 0 else 0, st))
            torch.cuda.synchronize()
            assert torch.equal(want.view(torch.int32), rows.view(torch.int32))
            same = dense.view(torch.int32) == rows.view(torch.int32)
            assert bool(same.all()), "step %d: %s" % (t, (~same).nonzero()[:8].tolist())
            assert torch.equal(state, packed[:, d - 240:d - 232].view(torch.int32))
        assert seen_hits > 1000
        assert lib.pgd_expand_rows_delta(packed.data_ptr(), dense.data_ptr(), None, n, d, 0, st) == -1
    finally:
        env.close()


def test_rows_to_host_equals_dense_copy():
    """pgd_rows_to_host (packed rows over PCIe, delta expansion on the host) for a batch that is already in HBM --
    bit for bit what a dense copy gives: over many steps into one array, into a second array, with rows of many hits."""
    import numpy as np
    import torch
    from pgdrive_b200 import VecPGDriveEnv
    n = 20000 + 8  # 4 chunks, the last one partial
    env = VecPGDriveEnv(dict(num_envs=n, start_seed=1000, environment_num=20))
    try:
        env.reset()
        g = torch.Generator(device="cuda"); g.manual_seed(2)
        host = np.full((n, env.obs_dim), np.nan, np.float32)
        other = np.full((n, env.obs_dim), np.nan, np.float32)
        hr, hd = np.zeros(n, np.float32), np.zeros(n, np.uint8)
        for t in range(40):
            a = torch.rand((n, 2), generator=g, device="cuda") * 2 - 1
            a[:, 1] = a[:, 1].abs(); a[:, 0] *= 0.1
            obs, rew, done, _ = env.step(a)
            if t % 9 == 4:
                obs = obs.clone()
                obs[::5, -240:] = torch.rand((len(obs[::5]), 240), generator=g, device="cuda")  # more hits than expected
            dst = other if t % 7 == 3 else host
            env.rows_to_host(obs, rew, done, dst, hr, hd)
            assert np.array_equal(dst.view(np.uint32), obs.cpu().numpy().view(np.uint32)), t
            assert np.array_equal(hr, rew.cpu().numpy()) and np.array_equal(hd, done.cpu().numpy())
            h2d, d2h = env.host_transfer_bytes()
            assert h2d == 0 and 0 < d2h
        env.rows_to_host(obs[:5000].contiguous(), None, None, host[:5000])  # another row count: new buffers, full
        assert np.array_equal(host[:5000].view(np.uint32), obs[:5000].cpu().numpy().view(np.uint32))
        # a gathered batch of eight shards: 17 chunks of 16 384 rows
        big = obs.repeat(14, 1)[:270000].contiguous()
        big_host = np.empty((270000, env.obs_dim), np.float32)
        for rep in range(3):
            if rep == 1:
                big[1::3, -240:] = 1.0
            if rep == 2:
                big[:, :34] += 1.0
            env.rows_to_host(big, None, None, big_host)
            assert np.array_equal(big_host.view(np.uint32), big.cpu().numpy().view(np.uint32)), rep
        with pytest.raises(ValueError):
            env.rows_to_host(obs, rew, None, host, hr, hd)
    finally:
        env.close()
