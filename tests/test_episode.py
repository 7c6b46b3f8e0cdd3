"""reset(force_seed=s) decisions against the reference's own traffic manager / navigation / parameter
sampling (fixtures from tools/make_golden.py).  Mirrors test_random_engine.py:54-188."""
from pgdrive_b200 import episode, mapgen


def test_episode_templates_match_reference(golden_resets):
    for s, d in golden_resets.items():
        m = mapgen.generate_map(int(s))
        ep = episode.make_episode(m, int(s), d["density"])
        assert ep.ego_seed == d["ego_seed"]
        assert ep.ego_params == d["ego_params"], s
        assert ep.ego_checkpoints == d["ego_checkpoints"], s
        assert len(ep.block_vehicles) == len(d["block_vehicles"])
        for (trigger, slots), g in zip(ep.block_vehicles, d["block_vehicles"]):
            assert list(trigger) == g["trigger_road"]
            assert len(slots) == len(g["vehicles"]), s
            for v, gv in zip(slots, g["vehicles"]):
                assert (v.type, list(v.lane), v.long, v.seed) == (gv["type"], gv["lane"], gv["long"], gv["seed"]), s
                assert (v.idm_seed, v.overtake_timer) == (gv["idm_seed"], gv["overtake_timer"]), s
                assert v.params == gv["params"], s
                assert v.checkpoints == gv["checkpoints"], s


def test_ego_params_known_answer():
    # SURVEY.md appendix A
    m = mapgen.generate_map(1000)
    ep = episode.make_episode(m, 1000, 0.0)
    assert ep.ego_params["max_engine_force"] == 783.99267578125
    assert ep.ego_params["max_brake_force"] == 113.99264526367188
    assert ep.block_vehicles == []


def test_random_agent_model_type_and_parameters_match_reference():
    """random_agent_model (manager/agent_manager.py:63-71): ego vehicle type drawn per seed and its sampled
    parameters, against the reference's own random_vehicle_type / parameter sampling
    (tests/golden/reset_random_agent.json.gz, tools/make_golden.py random_agent)."""
    from conftest import load_golden
    from pgdrive_b200 import episode, mapgen
    gold = load_golden("reset_random_agent.json.gz")
    seen = set()
    for s, rec in gold.items():
        seed = int(s)
        m = mapgen.generate_map(seed) if seed in (1000, 1001, 5) else None
        if m is not None:
            ep = episode.make_episode(m, seed, 0.0, random_agent_model=True)
            assert ep.ego_type == rec["type"], s
            for k, v in rec["params"].items():
                assert float(ep.ego_params[k]) == v, (s, k)
            assert episode.make_episode(m, seed, 0.0).ego_type == "default"
        else:  # the draw alone (no map needed)
            from pgdrive_b200 import rng
            t = episode.TYPE_KEYS[int(rng.seeded(seed).choice(5, p=[1 / 5] * 5))]
            assert t == rec["type"], s
            p = episode.sample_vehicle(t, rng.draw_seed(rng.seeded(seed)))
            for k, v in rec["params"].items():
                assert float(p[k]) == v, (s, k)
        body = episode.VEHICLE_BODY[rec["type"]]
        assert (body[0], body[1]) == (rec["length"], rec["width"]) and rec["max_length"] == 10 and rec["max_width"] == 2.5
        seen.add(rec["type"])
    assert len(seen) >= 4


def test_respawn_traffic_mode_matches_reference():
    """traffic_mode="respawn" (traffic_manager.py:63-66,188-237,292-309): which lanes are filled, the shuffled order of
    the 10 m slots on each, vehicle types / seeds / parameters / IDM seeds / routes -- against the reference's own
    _create_respawn_vehicles (tests/golden/reset_respawn.json.gz, tools/make_golden.py reset_respawn).  "hybrid" is the
    trigger code path in this version of the reference."""
    from conftest import load_golden
    gold = load_golden("reset_respawn.json.gz")
    counts = []
    for s, d in gold.items():
        m = mapgen.generate_map(int(s))
        ep = episode.make_episode(m, int(s), d["density"], traffic_mode="respawn")
        assert ep.ego_seed == d["ego_seed"] and ep.ego_params == d["ego_params"], s
        assert len(ep.block_vehicles) == 1 and ep.block_vehicles[0][0] is None
        slots, want = ep.block_vehicles[0][1], d["block_vehicles"][0]["vehicles"]
        assert len(slots) == len(want), s
        for v, gv in zip(slots, want):
            assert (v.type, list(v.lane), v.long, v.seed) == (gv["type"], gv["lane"], gv["long"], gv["seed"]), s
            assert (v.idm_seed, v.overtake_timer) == (gv["idm_seed"], gv["overtake_timer"]), s
            assert v.params == gv["params"] and v.checkpoints == gv["checkpoints"], s
        counts.append(len(slots))
        hy = episode.make_episode(m, int(s), d["density"], traffic_mode="hybrid")
        tr = episode.make_episode(m, int(s), d["density"], traffic_mode="trigger")
        assert [(r, [(v.lane, v.long, v.seed) for v in vs]) for r, vs in hy.block_vehicles] == \
               [(r, [(v.lane, v.long, v.seed) for v in vs]) for r, vs in tr.block_vehicles]
    assert min(counts) >= 12 and max(counts) > 31  # some maps need more than the 32 vehicle slots the kernel has


def test_accident_scenes_match_reference():
    """SafePGDriveEnv's accident scenes (manager/object_manager.py:40-124): which blocks get one, cones / tripods /
    barriers / broken-down vehicles with their lanes and Frenet coordinates, the engine seeds they consume before the ego
    draws its own, the vehicle type taken from the traffic stream, and the traffic that avoids the coned-off lanes --
    against the reference's own TrafficObjectManager.reset + TrafficManager (tests/golden/reset_accidents.json.gz)."""
    from conftest import load_golden
    gold = load_golden("reset_accidents.json.gz")
    kinds = set()
    for s, d in gold.items():
        m = mapgen.generate_map(int(s))
        ep = episode.make_episode(m, int(s), d["density"], accident_prob=d["accident_prob"])
        assert len(ep.objects) == len(d["objects"]), s
        for o, g in zip(ep.objects, d["objects"]):
            assert (o.kind, list(o.lane), o.seed) == (g["kind"], g["lane"], g["seed"]), s
            assert abs(o.long - g["long"]) < 1e-9 and abs(o.lat - g["lat"]) < 1e-9, s
            if o.kind == "vehicle":
                assert o.type == g["type"] and o.params == g["params"], s
            kinds.add(o.kind)
        assert ep.ego_seed == d["ego_seed"] and ep.ego_params == d["ego_params"], s
        assert len(ep.block_vehicles) == len(d["block_vehicles"])
        for (trigger, slots), g in zip(ep.block_vehicles, d["block_vehicles"]):
            assert list(trigger) == g["trigger_road"] and len(slots) == len(g["vehicles"]), s
            for v, gv in zip(slots, g["vehicles"]):
                assert (v.type, list(v.lane), v.long, v.seed, v.idm_seed) == \
                    (gv["type"], gv["lane"], gv["long"], gv["seed"], gv["idm_seed"]), s
                assert v.params == gv["params"] and v.checkpoints == gv["checkpoints"], s
    assert kinds == {"TrafficCone", "TrafficWarning", "TrafficBarrier", "vehicle"}


def test_random_traffic_draws_from_one_stream_across_resets():
    """random_traffic (traffic_manager.py:348-350): the traffic manager is not re-seeded at reset, so every visit of a
    map draws new traffic from the same generator; without it a seed always gives the same traffic
    (test_random_engine.py:75-99 is the reference's own check of this)."""
    import numpy as np
    m = mapgen.generate_map(5)
    fixed = [episode.make_episode(m, 5, 0.3, traffic_mode="respawn") for _ in range(2)]
    place = lambda ep: [(v.lane, v.long, v.type) for _, vs in ep.block_vehicles for v in vs]
    assert place(fixed[0]) == place(fixed[1])
    rs = np.random.RandomState(123)
    drawn = [place(episode.make_episode(m, 5, 0.3, traffic_mode="respawn", traffic_rs=rs)) for _ in range(4)]
    assert all(len(d) == len(drawn[0]) > 0 for d in drawn)  # respawn mode: the number of slots per lane is fixed
    assert len({tuple(d) for d in drawn}) == 4              # ... what stands on them is not
    assert all(d != place(fixed[0]) for d in drawn)
    # the ego does not depend on the traffic stream
    assert episode.make_episode(m, 5, 0.3, traffic_rs=rs).ego_params == fixed[0].ego_params
