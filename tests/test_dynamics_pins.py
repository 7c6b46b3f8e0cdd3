"""Behavioural pins for what stands in for Bullet (DESIGN.md "What stands in for Bullet").

The reference ships no numeric trajectories; its only dynamics check is a script that PRINTS acceleration time, brake
distance and turning displacement (tests/scripts/benchmark_brake.py:9-103).  This file restates that protocol on the
CPU oracle (the CUDA step is bit-identical to it) and checks the planar model against what the reference's own
constants imply: engine force on four wheels, per-wheel brake impulse, friction circle, speed limit, steering lock --
plus the vehicle-type table itself against the reference's (tests/golden/vehicle_types.json.gz, tools/make_golden.py
vehicle_types)."""
import math

import numpy as np
import pytest

from conftest import load_golden

V0 = dict(type="block_sequence", config="SSSSSSSS", lane_num=3, lane_width=3.5, exit_length=50)
SPAWN = ((">", ">>", 1), 5.0, 0.0)
DT, REPEAT, G = 0.02, 5, 9.81


@pytest.fixture(scope="module")
def straight_world():
    from oracle.oracle import Oracle
    from pgdrive_b200 import env as E
    T = E._seed_tables((4, V0, 0.0, SPAWN))  # benchmark_brake.py: start_seed 4, "SSSSSSSSSS", no traffic
    ref = Oracle(T, 1, auto_reset=False, num_slots=16)
    yield T, ref
    ref.close()


def _run(ref, action, steps, until=None):
    out = []
    for t in range(steps):
        o, r, d, info = ref.step(np.array([action], np.float32))
        s = ref.get_state(0)["veh"][0][0]
        out.append((float(s["x"]), float(s["y"]), float(s["heading"]), float(s["speed"]), bool(d[0])))
        if until is not None and until(out[-1]):
            break
    return out


def test_vehicle_type_table_matches_reference():
    from pgdrive_b200 import episode
    gold = load_golden("vehicle_types.json.gz")
    for key in ("s", "m", "l", "xl", "default"):
        g = gold[key]
        body = episode.VEHICLE_BODY[key]
        assert body == (g["LENGTH"], g["WIDTH"], g["HEIGHT"], float(g["MASS"]), g["FRONT_WHEELBASE"], g["REAR_WHEELBASE"],
                        g["TIRE_RADIUS"], g["LATERAL_TIRE_TO_CENTER"]), key
        for name, sp in g["space"].items():
            mine = episode.VEHICLE_SPACE[key][name]
            if sp["type"] == "ConstantSpace":
                assert mine == ("c", sp["fields"][0]), (key, name)
            else:  # BoxSpace = namedtuple("max min") written positionally as (750, 850), i.e. max=750, min=850:
                # sampled as uniform(low=min, high=max), reproduced literally
                assert mine == ("f", sp["fields"][1], sp["fields"][0]), (key, name)
    assert gold["_base"] == dict(MAX_LENGTH=10, MAX_WIDTH=2.5, MAX_STEERING=60, STEERING_INCREMENT=0.05)
    ob = gold["_objects"]
    assert episode.OBJECT_BODY["TrafficCone"][:2] == (2 * ob["TrafficCone"]["RADIUS"], ) * 2
    assert episode.OBJECT_BODY["TrafficWarning"][:2] == (2 * ob["TrafficWarning"]["RADIUS"], ) * 2
    # the barrier's LENGTH lies across the lane (setH without the vehicle's -90 degrees)
    assert episode.OBJECT_BODY["TrafficBarrier"][:2] == (ob["TrafficBarrier"]["WIDTH"], ob["TrafficBarrier"]["LENGTH"])


def test_acceleration_brake_and_coasting_follow_the_reference_constants(straight_world):
    """benchmark_brake.py's protocol: rest for 20 steps, full throttle to the speed limit, full brake to rest."""
    T, ref = straight_world
    slot = T["slots"][0]
    mass, f_max, b_max, mu = float(slot["mass"]), float(slot["max_engine"]), float(slot["max_brake"]), float(slot["friction"])
    ref.reset([0], [0])
    rest = _run(ref, [0.0, 0.0], 20)
    assert all(abs(p[3]) < 1e-6 for p in rest) and abs(rest[-1][0] - rest[0][0]) < 1e-6  # dropped onto its wheels, at rest
    acc = _run(ref, [0.0, 1.0], 400, until=lambda p: p[3] * 3.6 >= 79.0)
    a_expected = min(4.0 * f_max / mass, mu * G)  # applyEngineForce(max_engine_force * throttle) on all 4 wheels
    t_expected = (79.0 / 3.6) / a_expected
    assert abs(len(acc) * DT * REPEAT - t_expected) < 0.25, (len(acc) * 0.1, t_expected)
    cruise = _run(ref, [0.0, 1.0], 50)
    v_top = max(p[3] for p in cruise) * 3.6
    assert 80.0 <= v_top < 81.5  # the engine cuts out above max_speed = 80 km/h (base_vehicle.py:364-366)
    x0, v0 = cruise[-1][0], cruise[-1][3]
    brk = _run(ref, [0.0, -1.0], 200, until=lambda p: p[3] * 3.6 <= 1.0)
    # setBrake(|throttle| * max_brake_force) per wheel as an impulse per sub-step, capped by tyre friction mu * g
    a_brake = min(4.0 * b_max / mass / DT, mu * G)
    d_expected = v0 * v0 / (2.0 * a_brake)
    assert abs((brk[-1][0] - x0) - d_expected) < 0.08 * d_expected + 1.0, (brk[-1][0] - x0, d_expected)
    assert all(abs(p[1] - brk[0][1]) < 1e-3 for p in brk)  # straight line
    # idle: throttle 0 keeps setBrake(2.0) on every wheel -> 4 * 2 / (m * dt) of rolling deceleration
    _run(ref, [0.0, 1.0], 60)
    c0 = _run(ref, [0.0, 0.0], 1)[-1]
    c1 = _run(ref, [0.0, 0.0], 30)[-1]
    a_idle = (c0[3] - c1[3]) / (30 * DT * REPEAT)
    assert abs(a_idle - 4.0 * 2.0 / mass / DT) < 0.02


def test_steering_lock_friction_circle_and_sign(straight_world):
    T, ref = straight_world
    slot = T["slots"][0]
    lf, lr, mu, max_steer = float(slot["lf"]), float(slot["lr"]), float(slot["friction"]), float(slot["max_steer"])
    assert abs(max_steer - math.radians(40.0)) < 1e-6  # max_steering = 40 degrees at the Panda API
    ref.reset([0], [0])
    _run(ref, [0.0, 0.0], 8)
    _run(ref, [0.0, 0.3], 25)
    # +steering turns LEFT: the heading (clockwise positive in PGDrive's frame, +y = right) decreases
    left = _run(ref, [1.0, 0.0], 12)
    assert left[-1][2] < left[0][2] - 0.05 and left[-1][1] < left[0][1]
    # slow, full lock: the yaw rate settles at the kinematic bicycle's v * sin(beta) / lr, beta = atan(lr / L * tan(delta))
    ref.reset([0], [0])
    _run(ref, [0.0, 0.0], 8)
    _run(ref, [0.0, 0.2], 12)
    turn = _run(ref, [-1.0, 0.05], 15)
    v = turn[-1][3]
    yaw = (turn[-1][2] - turn[-2][2]) / (DT * REPEAT)
    beta = math.atan(lr / (lf + lr) * math.tan(max_steer))
    assert abs(yaw - v * math.sin(beta) / lr) < 0.08 * abs(yaw) + 0.02
    # fast: lateral acceleration never exceeds mu * g (friction circle)
    ref.reset([0], [0])
    _run(ref, [0.0, 0.0], 8)
    _run(ref, [0.0, 1.0], 60)
    fast = _run(ref, [1.0, 1.0], 10)
    for p0, p1 in zip(fast[:-1], fast[1:]):
        lat_acc = p1[3] * abs(p1[2] - p0[2]) / (DT * REPEAT)
        assert lat_acc <= mu * G * 1.02
