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
