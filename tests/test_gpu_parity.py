"""CUDA step vs the CPU oracle, through the C-ABI (run on the B200 box: pytest -m gpu).

Bar: BIT-EXACT -- observations, rewards, dones, contact flags and info records are compared with array_equal
(north_star asks for 1e-3 on floats and exact flags; both sides share include/pgd_math.h, so equality holds)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FLAG_MASK = 0x7ff


def _pair(n, seeds, density=0.1, auto_reset=True, detectors=None, **cfg):
    import torch  # noqa: F401
    from oracle.oracle import Oracle
    from pgdrive_b200 import VecPGDriveEnv
    okw = {}
    if detectors is not None:  # (side lasers, side distance, lane-line lasers, lane-line distance)
        ns, ds, nl, dl = detectors
        cfg["vehicle_config"] = dict(side_detector=dict(num_lasers=ns, distance=ds),
                                     lane_line_detector=dict(num_lasers=nl, distance=dl))
        okw = dict(n_side=ns, side_distance=ds, n_lane_line=nl, lane_line_distance=dl)
    env = VecPGDriveEnv(
        dict(start_seed=seeds[0], environment_num=len(seeds), num_envs=n, traffic_density=density,
             auto_reset=auto_reset, **cfg)
    )
    ref = Oracle(env.T, n, auto_reset=auto_reset, num_slots=env.engine.num_slots,
                 horizon=cfg.get("horizon", 0) or 0, **okw)
    return env, ref


def _reset_both(env, ref):
    obs = env.reset().cpu().numpy().copy()
    ro = ref.reset(range(env.num_envs), [env.episode_of_seed[int(s)] for s in env.env_seeds]).copy()
    return obs, ro


def _rollout(env, ref, steps, action_fn, check_state_every=0):
    """Free-running comparison, BIT-EXACT: every transcendental of the step comes from include/pgd_math.h (explicit
    fmaf chains, same bits from gcc and nvcc), everything else is IEEE add / mul / div / sqrt without contraction on
    both sides, so observations, rewards, dones, flags and the whole info record must be equal, not close.  (Round 1
    compared within 1e-3 and re-synchronised environments whose IDM sat on a discrete threshold, because CUDA's
    sincosf and glibc's differ in the last bit; that allowance is gone.)"""
    import torch
    n = env.num_envs
    dones = 0
    for t in range(steps):
        a = action_fn(t).astype(np.float32)
        o, r, d, _ = env.step(torch.from_numpy(a).cuda())
        o, r, d = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
        info = env.info_numpy()
        ro, rr, rd, rinfo = ref.step(a)
        bad = np.nonzero(d != rd)[0]
        assert len(bad) == 0, "step %d: done differs in envs %s" % (t, bad[:8])
        fl = (info["flags"] & FLAG_MASK) != (rinfo["flags"] & FLAG_MASK)
        assert not fl.any(), "step %d: flags differ in envs %s: %s vs %s" % (
            t, np.nonzero(fl)[0][:8], info["flags"][fl][:8], rinfo["flags"][fl][:8])
        if not np.array_equal(o, ro):
            err = np.abs(o - ro)
            raise AssertionError("step %d: obs differs by %g at %s" % (
                t, err.max(), np.unravel_index(err.argmax(), err.shape)))
        np.testing.assert_array_equal(r, rr, err_msg="step %d reward" % t)
        for f in info.dtype.names:
            np.testing.assert_array_equal(info[f], rinfo[f], err_msg="step %d info %s" % (t, f))
        dones += int(d.sum())
        if check_state_every and t % check_state_every == 0:
            for e in range(0, n, max(1, n // 8)):
                sg, sr = env.get_state(e)["veh"][0], ref.get_state(e)["veh"][0]
                k = env.T["episodes"][env.episode_of_seed[int(env.env_seeds[e])]]["n_slots"]
                alive = (sr["flags"][:k] & 1) != 0  # a removed vehicle's record is dead storage
                for f in ("lane", "ck0", "ck1", "rt_lane", "timer", "rnd_n", "airborne", "flags", "x", "y", "heading",
                          "speed", "yaw_rate"):
                    np.testing.assert_array_equal(sg[f][:k][alive], sr[f][:k][alive],
                                                  err_msg="state %s env %d step %d" % (f, e, t))
    return dones


def _check_lidar(gpu, ref, t):
    assert np.array_equal(gpu, ref), "step %d: lidar differs by %g" % (t, np.abs(gpu - ref).max())
    return 0


def test_reset_observation_matches_oracle():
    seeds = list(range(1000, 1100))
    env, ref = _pair(100, seeds)
    obs, ro = _reset_both(env, ref)
    assert np.array_equal(obs, ro)
    assert obs.min() >= 0.0 and obs.max() <= 1.0
    # spawn: lane 0 centre of a 3 x 3.5 m road, heading aligned, at rest
    np.testing.assert_allclose(obs[:, 0], 1.75 / 18, atol=1e-6)
    np.testing.assert_allclose(obs[:, 1], 8.75 / 18, atol=1e-6)
    np.testing.assert_allclose(obs[:, 2], 0.5, atol=1e-6)
    env.close()


def test_random_policy_rollout_with_autoreset():
    seeds = list(range(1000, 1100))
    n = 400
    env, ref = _pair(n, seeds)
    _reset_both(env, ref)
    rs = np.random.RandomState(1)
    # BASELINE.json's action distribution: uniform [-1, 1]^2 (half the throttles brake, so cars crawl)
    _rollout(env, ref, 100, lambda t: rs.uniform(-1, 1, (n, 2)), check_state_every=50)

    def forward(t):  # same steering noise but never braking: leaves the road within tens of steps
        a = rs.uniform(-1, 1, (n, 2))
        a[:, 1] = np.abs(a[:, 1])
        return a

    dones = _rollout(env, ref, 250, forward, check_state_every=50)
    assert dones > n  # the auto-reset path is exercised many times per env
    env.close()


def test_lane_following_rollout_meets_traffic():
    """Gentle steering + throttle keeps the ego on the road for hundreds of steps, so traffic is triggered,
    IDM runs, vehicles are removed at their destination and the lidar sees chassis."""
    seeds = list(range(1000, 1100))
    n = 200
    env, ref = _pair(n, seeds)
    _reset_both(env, ref)
    rs = np.random.RandomState(2)

    def act(t):
        a = np.zeros((n, 2))
        a[:, 0] = rs.uniform(-0.05, 0.05, n)
        a[:, 1] = rs.uniform(0.2, 1.0, n)
        return a

    _rollout(env, ref, 400, act, check_state_every=40)
    env.close()


def test_config1_single_env_seed_1000_no_traffic():
    """BASELINE.json configs[0]: 1 env, seed 1000, traffic_density 0, RandomState(0) actions, 1000 steps,
    reset(force_seed=1000) on done -- through the PGDriveEnv drop-in class."""
    from oracle.oracle import Oracle
    from pgdrive_b200 import PGDriveEnv
    from pgdrive_b200.env import build_seed_tables, default_config, parse_map_config
    env = PGDriveEnv(dict(start_seed=1000, environment_num=100, traffic_density=0.0))
    T = build_seed_tables([1000], parse_map_config(default_config()), 0.0, ((">", ">>", 0), 5.0, 0.0))
    ref = Oracle(T, 1, auto_reset=False)
    o = env.reset(force_seed=1000)
    ro = ref.reset([0], [0])[0]
    assert o.dtype == np.float64 and o.shape == (274, )
    assert np.array_equal(o, ro)
    actions = np.random.RandomState(0).uniform(-1, 1, (1000, 2)).astype(np.float32)
    episodes = 0
    for t in range(1000):
        o, r, d, info = env.step(actions[t])
        ro, rr, rd, rinfo = ref.step(actions[t][None])
        assert d == bool(rd[0]), t
        assert np.array_equal(o, ro[0].astype(np.float64)), t
        assert r == float(rr[0]), t
        assert info["out_of_road"] == bool(rinfo["flags"][0] & 2)
        assert env.observation_space.contains(o.astype(np.float32))
        if d:
            episodes += 1
            o = env.reset(force_seed=1000)
            ro = ref.reset([0], [0])[0]
            assert np.array_equal(o, ro)
    assert episodes >= 1
    env.close()


def test_nan_action_is_treated_as_minus_one():
    """cutils_clip(nan, -1, 1) == -1 (tests/test_component/test_utils.py, test_ego_vehicle.py:78-84)."""
    import torch
    env, ref = _pair(4, [1000, 1001])
    _reset_both(env, ref)
    a = np.array([[np.nan, 1.0], [0.0, np.nan], [np.nan, np.nan], [5.0, -7.0]], np.float32)
    for _ in range(12):
        o, r, d, _ = env.step(torch.from_numpy(a).cuda())
        ro, rr, rd, _ = ref.step(a)
        assert np.isfinite(o.cpu().numpy()).all()
        assert np.array_equal(o.cpu().numpy(), ro)
    info = env.info_numpy()
    np.testing.assert_allclose(info["steering"], [-1, 0, -1, 1])
    np.testing.assert_allclose(info["acceleration"], [1, -1, -1, -1])
    env.close()


@pytest.mark.parametrize("n", [64, 8192 + 24])  # the larger batch takes the chunked two-stream pipeline
def test_host_buffer_step_equals_device_step(n):
    import torch
    env_a, _ = _pair(n, [1000, 1001, 1002, 1003])
    env_b, _ = _pair(n, [1000, 1001, 1002, 1003])
    env_a.reset()
    env_b.reset()
    rs = np.random.RandomState(5)
    for _ in range(30):
        a = rs.uniform(-1, 1, (n, 2)).astype(np.float32)
        a[:, 1] = np.abs(a[:, 1])
        o1, r1, d1, i1 = env_a.step(a)
        o2, r2, d2, _ = env_b.step(torch.from_numpy(a).cuda())
        np.testing.assert_array_equal(o1, o2.cpu().numpy())
        np.testing.assert_array_equal(r1, r2.cpu().numpy())
        np.testing.assert_array_equal(d1, d2.cpu().numpy())
        np.testing.assert_array_equal(i1["flags"], env_b.info_numpy()["flags"])
    env_a.close()
    env_b.close()


def test_reset_then_host_buffer_step_are_ordered():
    """pgd_step_host runs on the handle's own streams; it must wait for the reset kernels that the caller's stream
    still has in flight (a large batch makes the reset pass long enough to overlap if it did not)."""
    from oracle.oracle import Oracle
    n = 32768
    env, _ = _pair(n, list(range(1000, 1100)))
    ref = Oracle(env.T, 200, auto_reset=True)
    rs = np.random.RandomState(7)
    for rep in range(3):
        env.reset()  # asynchronous on torch's current stream
        idx = np.arange(n) % 100  # replicas of a seed get the same action
        a = rs.uniform(-1, 1, (100, 2)).astype(np.float32)[idx]
        o, r, d, info = env.step(a)  # host buffers: the handle's own streams
        ref.reset(range(200), [i % 100 for i in range(200)])
        ro, rr, rd, rinfo = ref.step(a[:200])
        assert np.array_equal(o[:200], ro) and np.array_equal(r[:200], rr) and np.array_equal(d[:200], rd)
        assert not (info["flags"] & 1024).any()  # no environment treated the step as a second reset
        assert np.array_equal(o, o[:100][idx])
    env.close()


def test_horizon_and_partial_reset():
    import torch
    env, ref = _pair(8, [1000, 1001], horizon=5, auto_reset=False)
    _reset_both(env, ref)
    a = np.zeros((8, 2), np.float32)
    for t in range(5):
        o, r, d, _ = env.step(torch.from_numpy(a).cuda())
    assert d.cpu().numpy().all()
    assert (env.info_numpy()["flags"] & 8).all()  # max_step
    before = env.obs.cpu().numpy().copy()
    env.reset(env_ids=[1, 3], seeds=[1001, 1000])
    after = env.obs.cpu().numpy()
    assert np.array_equal(before[[0, 2, 4, 5, 6, 7]], after[[0, 2, 4, 5, 6, 7]])
    assert (env.info_numpy()["flags"][[1, 3]] & 1024).all()  # was_reset
    o, r, d, _ = env.step(torch.from_numpy(a).cuda())
    d = d.cpu().numpy()
    assert not d[1] and not d[3] and d[0]  # done is sticky for the others
    env.close()


def test_state_round_trip_and_v32_slots():
    import torch
    env, ref = _pair(6, [1000, 1001, 1002], num_slots=32)
    _reset_both(env, ref)
    rs = np.random.RandomState(3)
    _rollout(env, ref, 40, lambda t: np.c_[rs.uniform(-0.1, 0.1, 6), rs.uniform(0.3, 1, 6)])
    s = env.get_state(2)
    env.set_state(4, s)
    s2 = env.get_state(4)
    assert s.tobytes() == s2.tobytes()
    env.close()


def test_full_size_65536_envs_replicas_and_oracle():
    """BASELINE.json configs[2] size.  Environments i and j with i = j (mod 100) play the same seed; fed the same
    actions they must stay BIT-identical (size-independent property), and the first 100 are checked against the
    oracle every step, so all 65 536 are pinned transitively."""
    import torch
    seeds = list(range(1000, 1100))
    n = 65536
    env, _ = _pair(n, seeds)
    from oracle.oracle import Oracle
    ref = Oracle(env.T, 100, auto_reset=True)
    env.reset()
    ref.reset(range(100), range(100))
    rs = np.random.RandomState(11)
    idx = torch.arange(n, device="cuda") % 100
    total_done = 0
    for t in range(150):
        a100 = rs.uniform(-1, 1, (100, 2)).astype(np.float32)
        a100[:, 1] = np.abs(a100[:, 1])
        a100[:, 0] *= 0.3
        a = torch.from_numpy(a100).cuda()[idx].contiguous()
        o, r, d, _ = env.step(a)
        assert torch.equal(o, o[:100][idx]), "step %d: replicas of a seed diverged (obs)" % t
        assert torch.equal(r, r[:100][idx]) and torch.equal(d, d[:100][idx]), "step %d: replicas diverged" % t
        ro, rr, rd, _ = ref.step(a100)
        o100, r100, d100 = o[:100].cpu().numpy(), r[:100].cpu().numpy(), d[:100].cpu().numpy()
        assert np.array_equal(d100, rd), "step %d" % t
        assert np.array_equal(o100, ro), "step %d" % t
        np.testing.assert_array_equal(r100, rr)
        assert float(o.min()) >= 0.0 and float(o.max()) <= 1.0
        total_done += int(d.sum().item())
    assert total_done > n // 4
    env.close()


def test_1000envs_config_with_24_slots():
    """BASELINE.json configs[3]: PGDrive-1000envs-v0, 1000 distinct maps, up to 17 vehicles -> 24 slots."""
    seeds = list(range(1000, 2000))
    n = 2000
    env, ref = _pair(n, seeds)
    assert env.engine.num_slots == 24
    o, ro = _reset_both(env, ref)
    assert np.array_equal(o, ro)
    rs = np.random.RandomState(4)

    def act(t):
        a = rs.uniform(-1, 1, (n, 2))
        a[:, 0] *= 0.15
        a[:, 1] = np.abs(a[:, 1])
        return a

    dones = _rollout(env, ref, 120, act, check_state_every=30)
    assert dones > 0
    env.close()


def test_config5_size_524288_envs_on_one_gpu():
    """BASELINE.json configs[4] total size on a single GPU (maximum size): replicas of a seed stay bit-identical."""
    import torch
    n = 524288
    env, _ = _pair(n, list(range(1000, 1100)))
    env.reset()
    idx = torch.arange(n, device="cuda") % 100
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    for t in range(12):
        a100 = torch.rand((100, 2), generator=g, device="cuda") * 2 - 1
        a100[:, 1] = a100[:, 1].abs()
        o, r, d, _ = env.step(a100[idx].contiguous())
        assert torch.equal(o, o[:100][idx]) and torch.equal(r, r[:100][idx]) and torch.equal(d, d[:100][idx]), t
    assert float(o.min()) >= 0.0 and float(o.max()) <= 1.0
    env.close()


def test_side_and_lane_line_detectors_parity():
    """SURVEY 8 row S3 (off in PGDrive-v0): 24 side beams / 50 m and 12 lane-line beams / 20 m -> 306-float rows."""
    seeds = list(range(1000, 1040))
    n = 160
    env, ref = _pair(n, seeds, detectors=(24, 50.0, 12, 20.0))
    assert env.obs_dim == ref.obs_dim == 24 + 6 + 12 + 10 + 16 + 240
    o, ro = _reset_both(env, ref)
    assert np.array_equal(o, ro)
    rs = np.random.RandomState(6)

    def act(t):
        a = rs.uniform(-1, 1, (n, 2))
        a[:, 0] *= 0.2
        a[:, 1] = np.abs(a[:, 1])
        return a

    _rollout(env, ref, 150, act)
    env.close()


def test_long_soak_with_a_feedback_policy():
    """1000 steps of a lane-keeping feedback policy (steer towards the checkpoint, hold ~25 km/h): episodes end by
    arriving, by crashing into traffic and by leaving the road, and every step is compared with the oracle."""
    import torch
    seeds = list(range(1000, 1100))
    n = 200
    env, ref = _pair(n, seeds)
    _reset_both(env, ref)
    rs = np.random.RandomState(9)
    last = {"obs": ref.obs.copy()}
    seen = {"arrive": 0, "crash": 0, "out": 0}

    def act(t):
        o = last["obs"]
        a = np.zeros((n, 2))
        a[:, 0] = np.clip(-(o[:, 9] - 0.5) * 6.0 + rs.uniform(-0.05, 0.05, n), -1, 1)
        a[:, 1] = np.where(o[:, 3] < 0.3, 0.6, 0.0)
        return a

    for chunk in range(10):
        def act_and_track(t):
            a = act(t)
            return a
        # _rollout steps both; refresh the policy input from the oracle's observation after every step
        for t in range(100):
            _rollout(env, ref, 1, act_and_track)
            last["obs"] = ref.obs.copy()
            fl = ref.info["flags"]
            seen["arrive"] += int(((fl & 4) != 0).sum())
            seen["crash"] += int(((fl & 1) != 0).sum())
            seen["out"] += int(((fl & 2) != 0).sum())
    assert seen["arrive"] > 10 and seen["crash"] > 5 and seen["out"] > 50, seen
    env.close()


@pytest.mark.parametrize("variant", ["default", "second_copy", "dense", "one_thread"])
def test_host_buffer_step_variants_equal_device_step(variant, monkeypatch):
    """pgd_step_host ships packed rows and expands them on the host as a delta against the previous step's rows.  Every
    way through it gives the device step's results bit for bit: hits travelling with the first copy / fetched by a
    second one, dense rows, one host thread, a changed destination array (full expansion), detector fans in the head,
    and an invalidated state."""
    import ctypes
    import torch
    from pgdrive_b200 import cabi
    if variant == "second_copy":
        monkeypatch.setenv("PGDRIVE_B200_HOST_FIRST_HITS", "0")
        monkeypatch.setenv("PGDRIVE_B200_HOST_CHUNKS", "3")
    if variant == "one_thread":
        monkeypatch.setenv("PGDRIVE_B200_HOST_THREADS", "1")
    n = 4096 + 40
    det = (8, 50.0, 4, 20.0) if variant == "default" else None  # detector fans: a longer head
    env_a, _ = _pair(n, [1000, 1001, 1002, 1003, 1004], detectors=det)
    env_b, _ = _pair(n, [1000, 1001, 1002, 1003, 1004], detectors=det)
    env_a.reset()
    env_b.reset()
    rs = np.random.RandomState(9)
    other = np.full((n, env_a.obs_dim), np.nan, np.float32)  # a second destination, used through the C-ABI directly
    rew, done = np.zeros(n, np.float32), np.zeros(n, np.uint8)
    hits = 0
    for t in range(60):
        a = rs.uniform(-1, 1, (n, 2)).astype(np.float32)
        a[:, 1] = np.abs(a[:, 1])
        a[:, 0] *= 0.1
        if variant == "dense" and t == 20:
            monkeypatch.setenv("PGDRIVE_B200_HOST_DENSE", "1")
        if variant == "dense" and t == 40:
            monkeypatch.delenv("PGDRIVE_B200_HOST_DENSE")
        if t % 13 == 5:  # another destination array: everything is rewritten there, and again when we come back
            e = env_a.engine
            cabi.check(e.lib, e.lib.pgd_step_host(e.h, a.ctypes.data, other.ctypes.data, rew.ctypes.data,
                                                  done.ctypes.data, None))
            o1, r1, d1 = other, rew, done
        else:
            if t % 13 == 9:
                env_a._h_obs[::3] = -5.0  # the caller scribbled over the staging rows ... and says so
                cabi.check(env_a.engine.lib, env_a.engine.lib.pgd_host_invalidate(env_a.engine.h))
            o1, r1, d1, _ = env_a.step(a, copy=False)
            assert not o1.flags.writeable
        o2, r2, d2, _ = env_b.step(torch.from_numpy(a).cuda())
        np.testing.assert_array_equal(o1.view(np.uint32), o2.cpu().numpy().view(np.uint32), err_msg="step %d" % t)
        np.testing.assert_array_equal(r1, r2.cpu().numpy())
        np.testing.assert_array_equal(d1, d2.cpu().numpy())
        hits += int((o1[:, -240:] != 1.0).sum())
    assert hits > 5 * n  # the lidar saw things: the hit path was exercised
    env_a.close()
    env_b.close()
