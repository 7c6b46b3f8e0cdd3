"""The role-per-warp step (pgdrive_b200/csrc/pgd_step.cuh), HOST build, against the independent CPU oracle
(oracle/pgd_oracle.c).  The host build runs the kernel's phases in order over all (role, lane) pairs of a 32-environment
CTA -- shared-memory exchanges included -- and shares include/pgd_math.h with the oracle, so the bar is bit-identical
observations, rewards, done flags and info records over free-running rollouts, runnable without a GPU."""
import numpy as np
import pytest

V0 = dict(type="block_num", config=3, lane_num=3, lane_width=3.5, exit_length=50)
SPAWN = ((">", ">>", 0), 5.0, 0.0)
ROLES = [4]  # warps per CTA emulated by the host build (test_role_counts varies it)
ENVS_PER_CTA = [32]  # lanes of a warp that carry an environment (the launcher picks fewer to fill whole waves)


def _tables(seeds, density=0.1):
    from pgdrive_b200 import env as E
    return E.merge_tables([E._seed_tables((s, V0, density, SPAWN)) for s in seeds])


def _pair(T, n, **cfg):
    from oracle.oracle import Oracle
    from oracle.step_host import HostStep
    slots = 16 if T["max_slots"] <= 16 else 32
    return Oracle(T, n, num_slots=slots, **cfg), HostStep(T, n, roles=ROLES[0], envs_per_cta=ENVS_PER_CTA[0], num_slots=slots, **cfg)


def _actions(rs, n, mode):
    a = rs.uniform(-1, 1, (n, 2)).astype(np.float32)
    if mode == "forward":
        a[:, 1] = np.abs(a[:, 1])
        a[:, 0] *= 0.1
    elif mode == "lane":
        a[:, 0] = 0.0
        a[:, 1] = 0.6
    return a


def _same(x, y):
    o1, r1, d1, i1 = x
    o2, r2, d2, i2 = y
    return (np.array_equal(o1, o2) and np.array_equal(r1, r2) and np.array_equal(d1, d2)
            and i1.tobytes() == i2.tobytes())


@pytest.mark.parametrize("mode,n_seeds,n,steps,density", [
    ("uniform", 30, 120, 150, 0.1),
    ("forward", 30, 120, 300, 0.1),
    ("lane", 50, 150, 300, 0.1),
    ("lane", 12, 48, 300, 0.2),   # 32 vehicle slots
])
def test_free_running_rollouts_are_bit_identical(mode, n_seeds, n, steps, density):
    T = _tables(range(1000, 1000 + n_seeds), density)
    a, b = _pair(T, n, auto_reset=True)
    eps = [i % n_seeds for i in range(n)]
    assert np.array_equal(a.reset(range(n), eps), b.reset(range(n), eps))
    rs = np.random.RandomState(3)
    dones = 0
    for t in range(steps):
        act = _actions(rs, n, mode)
        ra, rb = a.step(act), b.step(act)
        assert _same(ra, rb), (mode, t)
        dones += int(ra[2].sum())
    assert dones > 0 or mode == "uniform"  # episodes ended and restarted inside the rollout
    a.close()
    b.close()


def test_horizon_sticky_done_and_partial_reset():
    T = _tables([1000, 1001, 1002])
    a, b = _pair(T, 6, auto_reset=False, horizon=7)
    eps = [0, 1, 2, 0, 1, 2]
    a.reset(range(6), eps)
    b.reset(range(6), eps)
    rs = np.random.RandomState(0)
    for t in range(12):
        act = _actions(rs, 6, "forward")
        if t == 3:
            act[2] = np.nan  # NaN action -> -1 like the compiled cutils_clip
        ra, rb = a.step(act), b.step(act)
        assert _same(ra, rb), t
        if t >= 6:
            assert ra[2].all() and rb[2].all()  # max_step reached and done stays set without auto-reset
    ia = a.reset([1, 4], [2, 0]).copy()
    ib = b.reset([1, 4], [2, 0]).copy()
    assert np.array_equal(ia[[1, 4]], ib[[1, 4]])
    ra, rb = a.step(np.zeros((6, 2), np.float32)), b.step(np.zeros((6, 2), np.float32))
    assert np.array_equal(ra[2], rb[2]) and np.array_equal(ra[0][[1, 4]], rb[0][[1, 4]])
    assert not ra[2][1] and not ra[2][4] and ra[2][0]
    a.close()
    b.close()


def test_reward_scheme_options():
    T = _tables(range(1000, 1010))
    cfg = dict(auto_reset=True, use_lateral=True, out_of_route_done=True, success_reward=20.0, driving_reward=2.0,
               speed_reward=0.3, out_of_road_penalty=7.0, crash_vehicle_penalty=3.0, decision_repeat=3)
    a, b = _pair(T, 40, **cfg)
    eps = [i % 10 for i in range(40)]
    a.reset(range(40), eps)
    b.reset(range(40), eps)
    rs = np.random.RandomState(5)
    for t in range(200):
        act = _actions(rs, 40, "forward")
        assert _same(a.step(act), b.step(act)), t
    a.close()
    b.close()


@pytest.mark.parametrize("ns,ds,nl,dl", [(12, 50.0, 8, 20.0), (0, 50.0, 16, 30.0), (120, 40.0, 0, 20.0)])
def test_side_and_lane_line_detectors(ns, ds, nl, dl):
    T = _tables(range(1000, 1016))
    cfg = dict(auto_reset=True, n_side=ns, side_distance=ds, n_lane_line=nl, lane_line_distance=dl)
    a, b = _pair(T, 48, **cfg)
    eps = [i % 16 for i in range(48)]
    assert np.array_equal(a.reset(range(48), eps), b.reset(range(48), eps))
    assert a.obs.shape[1] == (ns or 2) + 6 + nl + 266
    rs = np.random.RandomState(8)
    for t in range(160):
        act = _actions(rs, 48, "forward")
        assert _same(a.step(act), b.step(act)), t
    a.close()
    b.close()


def test_soak_with_a_feedback_policy_arrivals_crashes_and_exits():
    """A lane-keeping feedback policy (steer towards the checkpoint, hold ~25 km/h) for 700 steps: episodes end by
    arriving, by crashing into traffic and by leaving the road; every step is bit-identical to the oracle."""
    n_seeds, n = 100, 200
    T = _tables(range(1000, 1000 + n_seeds))
    a, b = _pair(T, n, auto_reset=True)
    eps = [i % n_seeds for i in range(n)]
    assert np.array_equal(a.reset(range(n), eps), b.reset(range(n), eps))
    rs = np.random.RandomState(9)
    obs = a.obs.copy()
    seen = dict(arrive=0, crash=0, out=0)
    for t in range(700):
        act = np.zeros((n, 2), np.float32)
        act[:, 0] = np.clip(-(obs[:, 9] - 0.5) * 6.0 + rs.uniform(-0.05, 0.05, n), -1, 1)
        act[:, 1] = np.where(obs[:, 3] < 0.3, 0.6, 0.0)
        ra, rb = a.step(act, threads=4), b.step(act)
        assert _same(ra, rb), t
        obs = ra[0].copy()
        fl = ra[3]["flags"]
        seen["arrive"] += int(((fl & 4) != 0).sum())
        seen["crash"] += int(((fl & 1) != 0).sum())
        seen["out"] += int(((fl & 2) != 0).sum())
    assert seen["arrive"] > 5 and seen["crash"] > 3 and seen["out"] > 30, seen
    a.close()
    b.close()


def test_property_random_step_configurations():
    """Randomly drawn step configurations (sub-steps per decision, step size, reward scheme, detectors, horizon,
    auto-reset, policy): still bit-identical to the oracle."""
    from hypothesis import given, settings, strategies as st
    T = _tables(range(1000, 1012))

    @settings(max_examples=14, deadline=None, derandomize=True)
    @given(st.integers(1, 8), st.sampled_from([0.01, 0.02, 0.04]), st.booleans(), st.booleans(), st.booleans(),
           st.sampled_from([0, 0, 6, 24]), st.sampled_from([0, 0, 5]), st.sampled_from([0, 40]),
           st.sampled_from(["forward", "lane", "uniform"]), st.integers(0, 10**6))
    def check(repeat, dt, use_lateral, route_done, auto_reset, n_side, n_lane_line, horizon, mode, seed):
        cfg = dict(decision_repeat=repeat, dt=dt, use_lateral=use_lateral, out_of_route_done=route_done,
                   auto_reset=auto_reset, n_side=n_side, n_lane_line=n_lane_line, horizon=horizon,
                   side_distance=45.0, lane_line_distance=25.0)
        a, b = _pair(T, 24, **cfg)
        eps = [i % 12 for i in range(24)]
        assert np.array_equal(a.reset(range(24), eps), b.reset(range(24), eps))
        rs = np.random.RandomState(seed)
        for t in range(70):
            act = _actions(rs, 24, mode)
            assert _same(a.step(act), b.step(act)), (cfg, mode, t)
        a.close()
        b.close()

    check()


def test_stored_state_equals_the_oracles_state():
    """What the layout writes back lazily (parked traffic: only the drop counter; removed vehicles: nothing) is, field
    for field, the state the oracle holds -- checked every 20 steps of a rollout in which traffic wakes up and dies."""
    T = _tables(range(1000, 1030))
    n = 90
    a, b = _pair(T, n, auto_reset=True)
    eps = [i % 30 for i in range(n)]
    a.reset(range(n), eps)
    b.reset(range(n), eps)
    rs = np.random.RandomState(4)
    removed = 0
    obs = a.obs.copy()
    for t in range(600):
        act = np.zeros((n, 2), np.float32)  # lane-keeping feedback: long episodes, traffic wakes up and arrives
        act[:, 0] = np.clip(-(obs[:, 9] - 0.5) * 6.0 + rs.uniform(-0.05, 0.05, n), -1, 1)
        act[:, 1] = np.where(obs[:, 3] < 0.3, 0.6, 0.0)
        ra = a.step(act, threads=4)
        assert _same(ra, b.step(act)), t
        obs = ra[0].copy()
        if t % 25 == 0 or t == 599:
            for e in range(n):
                sa, sb = a.get_state(e), b.get_state(e)
                k = int(T["episodes"][eps[e]]["n_slots"])
                for f in ("episode", "next_group", "done", "ep_len", "prev_steer", "prev_throttle", "ep_reward", "energy"):
                    assert sa[f][0] == sb[f][0], (t, e, f)
                va, vb = sa["veh"][0][:k], sb["veh"][0][:k]
                assert va.tobytes() == vb.tobytes(), (t, e, [f for f in va.dtype.names if not np.array_equal(va[f], vb[f])])
                removed += int(((va["flags"] & 1) == 0).sum())
    assert removed > 0  # some traffic left its road and was removed during the rollout
    a.close()
    b.close()


def test_random_agent_model_observation_and_dynamics():
    """random_agent_model: the ego is one of five vehicle types per seed and the observation gains its length / 10 and
    width / 2.5 after the (optional) lane-line beams (obs/state_obs.py:18-23,103-105)."""
    from pgdrive_b200 import env as E
    from pgdrive_b200.episode import VEHICLE_BODY, TYPE_KEYS
    seeds = list(range(1000, 1020))
    T = E.merge_tables([E._seed_tables((s, V0, 0.1, SPAWN, None, True)) for s in seeds])
    types = [TYPE_KEYS[int(T["slots"][int(e["slot_off"])]["type"])] for e in T["episodes"]]
    assert len(set(types)) >= 3
    for cfg in (dict(), dict(n_side=6, n_lane_line=4, side_distance=50.0, lane_line_distance=20.0)):
        a, b = _pair(T, 40, auto_reset=True, random_agent_model=True, **cfg)
        eps = [i % 20 for i in range(40)]
        oa, ob = a.reset(range(40), eps), b.reset(range(40), eps)
        assert np.array_equal(oa, ob)
        n_first, nl = cfg.get("n_side", 0) or 2, cfg.get("n_lane_line", 0)
        assert oa.shape[1] == n_first + 6 + nl + 2 + 266
        for i in range(40):
            body = VEHICLE_BODY[types[eps[i]]]
            assert abs(oa[i, n_first + 6 + nl] - body[0] / 10) < 1e-6 and abs(oa[i, n_first + 6 + nl + 1] - body[1] / 2.5) < 1e-6
        rs = np.random.RandomState(6)
        for t in range(200):
            act = _actions(rs, 40, "forward")
            assert _same(a.step(act), b.step(act)), t
        a.close()
        b.close()


@pytest.mark.parametrize("roles", [2, 3, 8])
def test_role_counts(roles):
    """The slot -> role assignment and the task split depend on the number of warps per CTA; the result must not."""
    ROLES[0] = roles
    try:
        T = _tables(range(1000, 1020))
        n = 70  # not a multiple of 32: the last CTA is partly empty
        a, b = _pair(T, n, auto_reset=True)
        eps = [i % 20 for i in range(n)]
        assert np.array_equal(a.reset(range(n), eps), b.reset(range(n), eps))
        rs = np.random.RandomState(12)
        for t in range(250):
            act = _actions(rs, n, "lane" if t % 2 else "forward")
            assert _same(a.step(act), b.step(act)), t
        a.close()
        b.close()
    finally:
        ROLES[0] = 4


@pytest.mark.parametrize("epc", [28, 6])
def test_environments_per_cta(epc):
    """pgd_launch_step gives a CTA fewer than 32 environments when that fills whole waves of the GPU; the unused lanes of
    every warp then stay idle through all phases (work list, named barrier, lidar scatter, row store)."""
    ENVS_PER_CTA[0] = epc
    try:
        T = _tables(range(1000, 1020))
        n = 70
        a, b = _pair(T, n, auto_reset=True)
        eps = [i % 20 for i in range(n)]
        assert np.array_equal(a.reset(range(n), eps), b.reset(range(n), eps))
        rs = np.random.RandomState(13)
        for t in range(250):
            act = _actions(rs, n, "lane" if t % 2 else "forward")
            assert _same(a.step(act), b.step(act)), t
        a.close()
        b.close()
    finally:
        ENVS_PER_CTA[0] = 32


def test_respawn_traffic_mode_all_traffic_awake_from_the_first_step():
    """traffic_mode="respawn": every respawn lane carries a vehicle per 10 m and all of them run IDM from step 0
    (PgdSlot.group = PGD_GROUP_AWAKE).  Seeds whose maps need at most 32 vehicle slots."""
    from pgdrive_b200 import env as E
    seeds = [1002, 1003, 1004, 1006, 1010, 1012, 1017, 1018, 1022, 1028]
    T = E.merge_tables([E._seed_tables((s, V0, 0.1, SPAWN, None, False, "respawn")) for s in seeds])
    assert 16 < T["max_slots"] <= 32 and (T["slots"]["group"] == -2).sum() == len(T["slots"]) - len(seeds)
    n = 40
    a, b = _pair(T, n, auto_reset=True)
    eps = [i % len(seeds) for i in range(n)]
    assert np.array_equal(a.reset(range(n), eps), b.reset(range(n), eps))
    rs = np.random.RandomState(21)
    moved = 0
    for t in range(220):
        act = _actions(rs, n, "lane" if t < 150 else "forward")
        ra, rb = a.step(act, threads=4), b.step(act)
        assert _same(ra, rb), t
        if t == 10:  # traffic is already driving (bumper to bumper: one vehicle per 10 m), no trigger needed
            moved = sum(int((b.get_state(e)["veh"][0]["speed"][1:] > 0.5).sum()) for e in range(10))
    assert moved >= 10
    a.close()
    b.close()


def test_lidar_noise_dropout_and_increment_steering():
    """Lidar gaussian_noise / dropout_prob (obs/state_obs.py:165-182) and increment_steering (base_vehicle.py:351-358):
    kernel source and oracle share the counter-based generator, so they stay bit-identical; the noise itself is pinned
    statistically (the reference draws from numpy's global generator) and steering increments are checked against the
    reference's formula."""
    T = _tables(range(1000, 1012))
    n = 48
    eps = [i % 12 for i in range(n)]
    clean_a, clean_b = _pair(T, n, auto_reset=True)
    cfg = dict(auto_reset=True, lidar_gaussian_noise=0.05, lidar_dropout_prob=0.1, noise_seed=3, increment_steering=True)
    a, b = _pair(T, n, **cfg)
    assert np.array_equal(a.reset(range(n), eps), b.reset(range(n), eps))
    clean_a.reset(range(n), eps)
    rs = np.random.RandomState(31)
    steer = np.zeros(n, np.float32)
    dropped = total = 0
    resid = []
    for t in range(120):
        act = _actions(rs, n, "forward")
        act[:, 0] *= 5.0
        ra, rb = a.step(act), b.step(act)
        assert _same(ra, rb), t
        fresh = (ra[3]["flags"] & 1024) != 0
        steer = np.where(fresh, 0.0, np.clip(steer + np.clip(act[:, 0], -1, 1) * np.float32(0.05), -1, 1)).astype(np.float32)
        assert np.array_equal(ra[3]["steering"][~fresh], steer[~fresh]), t
        lid = ra[0][:, 34:]
        assert lid.min() >= 0.0 and lid.max() <= 1.0
        dropped += int((lid == 0.0).sum())
        total += lid.size
        resid.append(lid[(lid > 0.0) & (lid < 1.0)].ravel())
    assert abs(dropped / total - 0.1) < 0.004  # dropout sets ~10 % of the beams to exactly 0
    # most beams see nothing (1.0); after noise + clip they are 1 - |N(0, 0.05)| or 1: the unclipped ones are a half normal
    r = 1.0 - np.concatenate(resid)
    r = r[r < 0.3]
    assert abs(np.sqrt((r ** 2).mean()) - 0.05) < 0.003
    # the same call with another seed gives other noise; without noise the rows are the clean ones
    for x in (a, b, clean_a, clean_b):
        x.close()


def test_accident_scenes_crash_object_and_safe_rl_env():
    """SafePGDriveEnv's accident scenes as static slots (cones, tripods, barriers: PGD_TYPE_OBJECT; broken-down vehicles):
    obstacles for IDM and the lidar, crash_object charged once per object, crash_vehicle for the broken-down vehicle,
    and with safe_rl_env a crash costs but does not end the episode (safe_pgdrive_env.py:44-51)."""
    from pgdrive_b200 import env as E
    seeds = [100, 101, 104, 106, 110, 112, 0, 1000]
    # half of the environments start on the right-most lane, where most coned-off lane ends and barriers are
    T = E.merge_tables([E._seed_tables((s, V0, 0.05, ((">", ">>", 2 * (k % 2)), 5.0, 0.0), None, False, "trigger", 0.8))
                        for k, s in enumerate(seeds + seeds)])
    n_obj = int((T["slots"]["type"] >= 5).sum())
    assert n_obj > 60 and (T["slots"]["group"] == -3).sum() >= n_obj and T["max_slots"] <= 32
    n = 64
    eps = [i % (2 * len(seeds)) for i in range(n)]
    for safe in (False, True):
        a, b = _pair(T, n, auto_reset=True, safe_rl_env=safe, crash_object_penalty=4.0, crash_object_cost=2.0)
        assert np.array_equal(a.reset(range(n), eps), b.reset(range(n), eps))
        rs = np.random.RandomState(17)
        obs = a.obs.copy()
        hits = done_on_hit = lidar_sees = 0
        for t in range(500):
            act = np.zeros((n, 2), np.float32)  # holds its lane (heading error only): drives into whatever blocks it
            act[:, 0] = np.clip((obs[:, 2] - 0.5) * 8.0 + rs.uniform(-0.02, 0.02, n), -1, 1)
            act[:, 1] = np.where(obs[:, 3] < 0.25, 0.5, 0.0)
            ra, rb = a.step(act, threads=4), b.step(act)
            assert _same(ra, rb), (safe, t)
            obs = ra[0].copy()
            fl = ra[3]["flags"]
            hit = (fl & 2048) != 0
            hits += int(hit.sum())
            done_on_hit += int((ra[2][hit] != 0).sum())
            if hit.any():
                only = hit & ((fl & (2 | 1 | 4)) == 0)  # crash_object alone: its reward and cost
                assert np.all(ra[1][only] == -4.0) and np.all(ra[3]["cost"][only] == 2.0)
            lidar_sees += int((ra[0][:, 34:] < 1.0).any(axis=1).sum())
        assert hits > 5 and lidar_sees > 100, (safe, hits, lidar_sees)
        assert (done_on_hit == 0) if safe else (done_on_hit > 0), (safe, hits, done_on_hit)
        a.close()
        b.close()


def test_oracle_fast_queries_equal_brute_force():
    """The oracle's timed configuration (bucket-grid queries, bench.py's cpu_baseline / --impl reference) gives the same
    bits as its brute-force loops over every primitive."""
    from oracle.oracle import Oracle
    T = _tables(range(1000, 1030))
    n = 90
    a, b = Oracle(T, n, auto_reset=True, num_slots=16), Oracle(T, n, auto_reset=True, num_slots=16, fast=True)
    eps = [i % 30 for i in range(n)]
    assert np.array_equal(a.reset(range(n), eps), b.reset(range(n), eps))
    rs = np.random.RandomState(8)
    for t in range(300):
        act = _actions(rs, n, "forward" if t % 3 else "lane")
        assert _same(a.step(act, threads=4), b.step(act, threads=4)), t
    a.close()
    b.close()
