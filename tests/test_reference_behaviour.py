"""The reference's own behavioural tests, replayed on the drop-in PGDriveEnv (B200 box).

Each test names the reference test it mirrors (paths under /root/reference/pgdrive/tests).  Only deviation: the
reference uses traffic_density 1.0 / 20 on map "SSS", which needs more than the 32 vehicle slots of this
simulator; 0.3 is used instead (19 vehicles on seed 0)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REWARDS = dict(success_reward=1111, out_of_road_penalty=2222, crash_vehicle_penalty=3333, crash_object_penalty=4444,
               out_of_road_cost=5555, crash_vehicle_cost=6666, crash_object_cost=7777)


def _env(cfg):
    from pgdrive_b200 import PGDriveEnv
    return PGDriveEnv(cfg)


def test_collision_with_vehicle():
    """test_functionality/test_collision.py:4-19"""
    env = _env({"traffic_density": 0.3, "map": "SSS"})
    env.reset()
    try:
        hit = False
        for _ in range(1, 500):
            o, r, d, info = env.step([0, 1])
            if info["crash_vehicle"]:
                hit = True
                break
        assert hit, "Collision function is broken!"
    finally:
        env.close()


def test_collision_with_sidewalk_and_line_contact():
    """test_functionality/test_collision.py:22-49 (both tests drive [-0.5, 1] for 100 steps)"""
    env = _env({"traffic_density": .0})
    env.reset()
    try:
        sidewalk = broken = white = False
        for _ in range(1, 100):
            o, r, d, info = env.step([-0.5, 1])
            sidewalk |= info["crash_sidewalk"]
            broken |= info["on_broken_line"]
            white |= info["on_white_continuous_line"]
        assert sidewalk and broken and white, "Collision function is broken!"
    finally:
        env.close()


def test_reward_cost_done():
    """test_functionality/test_reward_cost_done.py:6-74, including the success / out-of-road cases that the
    reference keeps commented out."""
    cfg = dict(REWARDS, map="SSS", traffic_density=0.3)
    env = _env(cfg)
    try:
        env.reset()
        for _ in range(1000):
            o, r, d, i = env.step([0, 1])
            if d:
                break
        assert i["crash"] and i["crash_vehicle"]
        assert i["cost"] == REWARDS["crash_vehicle_cost"]
        assert r == -REWARDS["crash_vehicle_penalty"]
    finally:
        env.close()
    env = _env(dict(REWARDS, map="S", traffic_density=0))
    try:
        env.reset()
        for _ in range(1000):
            o, r, d, i = env.step([0, 1])
            if d:
                break
        assert i["arrive_dest"] and i["cost"] == 0 and r == REWARDS["success_reward"]
        env.reset()
        for _ in range(1000):
            o, r, d, i = env.step([1, 1])
            if d:
                break
        assert i["out_of_road"] and i["cost"] == REWARDS["out_of_road_cost"] and r == -REWARDS["out_of_road_penalty"]
    finally:
        env.close()


def test_obs_action_space_and_info_keys():
    """test_functionality/test_obs_action_space.py:10-14, test_obs_noise.py:7-20 (info keys)"""
    env = _env(dict(start_seed=1000, environment_num=100))
    try:
        o = env.reset(force_seed=1007)
        assert env.observation_space.contains(o.astype(np.float32))
        assert env.action_space.contains(env.action_space.sample())
        o, r, d, i = env.step(env.action_space.sample())
        assert env.observation_space.contains(o.astype(np.float32))
        assert isinstance(r, float) and isinstance(d, bool)
        for k in ("cost", "velocity", "steering", "acceleration", "step_reward", "crash_vehicle", "out_of_road",
                  "arrive_dest", "crash", "crash_object", "crash_building", "max_step", "episode_reward",
                  "episode_length", "raw_action", "step_energy", "episode_energy", "overtake_vehicle_num"):
            assert k in i, k
        assert env.current_seed == 1007
    finally:
        env.close()


def test_nan_actions_do_not_break_the_vehicle():
    """test_functionality/test_ego_vehicle.py:78-84"""
    env = _env({"traffic_density": .0})
    try:
        env.reset()
        for a in ([np.nan, np.nan], [np.nan, 1.0], [1.0, np.nan], [np.inf, -np.inf]):
            o, r, d, i = env.step(a)
            assert np.isfinite(o).all() and np.isfinite(r)
    finally:
        env.close()


def test_same_force_seed_same_episode_regardless_of_environment_num():
    """test_functionality/test_random_engine.py:20-72: a forced seed gives the same map and the same traffic no
    matter how many environments the env was configured with."""
    rolls = []
    for num in (1, 10, 100):
        env = _env(dict(start_seed=1000, environment_num=num if num > 1 else 1, traffic_density=0.1))
        if num == 1:
            env = _env(dict(start_seed=1005, environment_num=1, traffic_density=0.1))
        try:
            o0 = env.reset(force_seed=1005)
            traj = [o0]
            for _ in range(60):
                traj.append(env.step([0.0, 0.6])[0])
            state = env.get_state()["veh"][0]
            alive = (state["flags"] & 1) != 0
            rolls.append((np.array(traj), np.c_[state["x"], state["y"], state["heading"]][alive]))
        finally:
            env.close()
    for traj, state in rolls[1:]:
        assert np.array_equal(traj, rolls[0][0])
        assert np.array_equal(state, rolls[0][1])  # every traffic vehicle ends at the same pose as well


def test_lane_following_for_2000_steps():
    """test_functionality/test_navigation.py:24-90: a PID on obs-derived errors keeps the car on the road; the test
    reads the observation layout (o[0] left distance, navigation info at o[8:18])."""
    env = _env(dict(start_seed=1000, environment_num=5, traffic_density=0.0))
    try:
        o = env.reset(force_seed=1002)
        steps_alive = 0
        for _ in range(2000):
            # steer towards the first checkpoint: o[9] is its lateral projection (0.5 = straight ahead)
            steering = float(np.clip(-(o[9] - 0.5) * 6.0, -1, 1))
            throttle = 0.5 if o[3] < 0.35 else 0.0
            o, r, d, i = env.step([steering, throttle])
            steps_alive += 1
            if d:
                assert i["arrive_dest"] or i["out_of_road"]
                o = env.reset(force_seed=1002)
        assert steps_alive == 2000
    finally:
        env.close()


def test_out_of_road_side_detector():
    """test_functionality/test_out_of_road.py:7-36: drifting off an 11-block straight road, the side detector's
    nearest reading at the moment of `done` is below vehicle-diagonal / range."""
    import math
    for steering in (-0.01, 0.01):
        for distance in (10, 50, 100):
            env = _env(dict(map="SSSSSSSSSSS",
                            vehicle_config=dict(side_detector=dict(num_lasers=120, distance=distance))))
            try:
                obs = env.reset()
                assert obs.shape == (120 + 6 + 266, )
                tolerance = math.sqrt(1.852**2 + 4.51**2) / distance
                for _ in range(4000):
                    o, r, d, i = env.step([steering, 1])
                    if d:
                        assert i["out_of_road"] or i["crash_vehicle"]
                        if i["out_of_road"]:
                            assert min(o[:120]) < tolerance, (min(o[:120]), tolerance)
                        break
                assert d
            finally:
                env.close()
