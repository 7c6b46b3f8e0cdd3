"""The CUDA step against the committed reference-generated fixtures (not via the oracle): IDM actions,
navigation info, state observation, neighbour features, reward (tests/golden/step_v0.json.gz) and the
240-beam lidar (tests/golden/lidar_v0.json.gz).  Same tolerances as tests/test_oracle_golden.py."""
import base64

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _env_for(seed, density=0.1):
    from pgdrive_b200 import VecPGDriveEnv, tables
    T = tables.build_tables([seed], density=density).finish()
    T["max_slots"] = int(T["episodes"]["n_slots"].max())
    return VecPGDriveEnv(dict(start_seed=seed, environment_num=1, num_envs=1, auto_reset=False,
                              traffic_density=density), tables_dict=T)


def test_step_fixtures_through_the_cuda_path():
    import torch
    from pgdrive_b200 import cabi
    records = load_golden("step_v0.json.gz") + load_golden("step_dense.json.gz")
    envs = {}
    n_idm = 0
    for rec in records:
        key = (rec["seed"], rec.get("density", 0.1))
        env = envs.get(key) or envs.setdefault(key, _env_for(*key))
        if not getattr(env, "_was_reset", False):
            env.reset()
            env._was_reset = True
        s0 = np.frombuffer(base64.b64decode(rec["s0"]), dtype=cabi.ENV_STATE_DT).copy()
        env.set_state(0, s0)
        obs, rew, done, _ = env.step(torch.tensor([rec["action"]], dtype=torch.float32, device="cuda"))
        obs = obs.cpu().numpy()[0]
        info = env.info_numpy()[0]
        veh = env.get_state(0)["veh"][0]
        tag = (rec["seed"], rec["t"])
        for g in rec["idm"]:
            if g.get("tie"):
                continue
            v = veh[g["slot"]]
            assert int(v["rt_lane"]) == g["rt_lane"] and int(v["timer"]) == g["timer"], tag
            assert float(v["target_speed"]) == g["target_speed"], tag
            np.testing.assert_allclose(float(v["steer"]), g["steering"], rtol=2e-3, atol=2e-4, err_msg=str(tag))
            np.testing.assert_allclose(float(v["throttle"]), g["acc"], rtol=2e-3, atol=2e-4, err_msg=str(tag))
            n_idm += 1
        assert [int(veh[0]["ck0"]), int(veh[0]["ck1"])] == rec["ck"], tag
        np.testing.assert_allclose(obs[8:18], rec["navi"], atol=2e-5, err_msg=str(tag))
        np.testing.assert_allclose(obs[:7], rec["state"][:7], atol=2e-5, err_msg=str(tag))
        np.testing.assert_allclose(obs[7], rec["state"][7], atol=2e-4, err_msg=str(tag))
        if not any(g.get("tie") for g in rec["idm"]):  # a tied vehicle may have taken the other, equally valid action
            np.testing.assert_allclose(obs[18:34], rec["neighbours"], atol=2e-5, err_msg=str(tag))
        np.testing.assert_allclose(float(info["step_reward"]), rec["step_reward"], rtol=1e-3, atol=2e-4, err_msg=str(tag))
        assert bool(int(info["flags"]) & cabi.F_ARRIVE_DEST) == rec["arrive_dest"], tag
    assert n_idm > 6500
    for e in envs.values():
        e.close()


def test_lidar_fixture_through_the_cuda_path():
    import torch
    from pgdrive_b200 import cabi
    scenes = load_golden("lidar_v0.json.gz")
    env = _env_for(1003)
    env.reset()
    zero = torch.zeros((1, 2), dtype=torch.float32, device="cuda")
    for sc in scenes:
        s = np.frombuffer(base64.b64decode(sc["state"]), dtype=cabi.ENV_STATE_DT).copy()
        s["veh"]["airborne"] = 5  # nothing moves during the step: the observation is of exactly these poses
        env.set_state(0, s)
        obs = env.step(zero)[0].cpu().numpy()[0]
        cloud = np.array(sc["cloud"])
        got = obs[34:]
        close = np.abs(got - cloud) < 2e-4
        for i in np.nonzero(~close)[0]:
            nb = [cloud[(i - 1) % 240], cloud[(i + 1) % 240]]
            assert (min(nb) < 1.0) != (max(nb) < 1.0) or abs(got[i] - cloud[i]) < 5e-3, (i, got[i], cloud[i])
        assert close.mean() > 0.995
    env.close()


def test_detector_fixture_through_the_cuda_path():
    """tests/golden/detectors_v0.json.gz (reference beam loop over reference-built line ghosts) vs the kernel."""
    import torch
    from pgdrive_b200 import VecPGDriveEnv, cabi, tables
    scenes = load_golden("detectors_v0.json.gz")
    envs = {}
    zero = torch.zeros((1, 2), dtype=torch.float32, device="cuda")
    for sc in scenes:
        seed = sc["seed"]
        if seed not in envs:
            T = tables.build_tables([seed]).finish()
            envs[seed] = VecPGDriveEnv(
                dict(start_seed=seed, environment_num=1, num_envs=1, auto_reset=False,
                     vehicle_config=dict(side_detector=dict(num_lasers=120, distance=50.0),
                                         lane_line_detector=dict(num_lasers=40, distance=20.0))), tables_dict=T)
            envs[seed].reset()
        env = envs[seed]
        s = np.frombuffer(base64.b64decode(sc["state"]), dtype=cabi.ENV_STATE_DT).copy()
        s["veh"]["airborne"] = 5
        env.set_state(0, s)
        obs = env.step(zero)[0].cpu().numpy()[0]
        assert obs.shape == (120 + 6 + 40 + 266, )
        for name, got in (("side", obs[:120]), ("lane_line", obs[126:166])):
            want = np.array(sc[name])
            close = np.abs(got - want) < 5e-4
            for i in np.nonzero(~close)[0]:
                nb = [want[(i - 1) % len(want)], want[(i + 1) % len(want)]]
                assert abs(got[i] - want[i]) < 0.2 and min(nb) < 1.0, (name, i, got[i], want[i])
            assert close.mean() > 0.97
    for e in envs.values():
        e.close()
