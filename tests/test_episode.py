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
