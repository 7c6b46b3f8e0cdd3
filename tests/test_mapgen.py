"""Seed -> map determinism against fixtures made by the unmodified reference (tools/make_golden.py).

Mirrors the reference's own pins: test_loading_map_from_json.py:8-50 (JSON == live BIG) and
test_random_engine.py:20-52 (same seed => same map)."""
import pytest

from pgdrive_b200 import mapgen, rng


def _lanes(m):
    return [(f, t, i, ln) for (f, t), ls in m.net.roads() for i, ln in enumerate(ls)]


def test_known_seed_hashes():
    # SURVEY.md appendix A known-answer values, computed with the reference's random_utils
    assert rng.hash_seed(1000) == 6064680319747761938
    assert rng.hash_seed(1001) == 12325696033393922275
    assert rng.draw_seed(rng.seeded(1000)) == 61626
    assert rng.draw_seed(rng.seeded(1001)) == 40898


def test_block_sequences_match_reference(golden_maps):
    for s, d in golden_maps.items():
        seq = mapgen.search_sequence(int(s))
        gold = d["block_sequence"]
        assert [b["id"] for b in seq] == [b["id"] for b in gold], s
        for a, b in zip(seq, gold):
            assert a["pre_block_socket_index"] == b["pre_block_socket_index"]
            for k, v in b.items():
                if k not in ("id", "pre_block_socket_index"):
                    assert float(a[k]) == float(v), (s, k)


def test_lanes_match_reference_bit_for_bit(golden_maps):
    for s, d in golden_maps.items():
        m = mapgen.generate_map(int(s))
        mine = _lanes(m)
        assert len(mine) == len(d["lanes"]), s
        for (f, t, i, ln), g in zip(mine, d["lanes"]):
            assert (f, t, i, ln.kind) == (g["frm"], g["to"], g["idx"], g["kind"]), s
            assert [str(x) for x in ln.line_types] == g["line_types"], (s, f, t, i)
            colour = ["Y" if abs(c[0] - 1) > 1e-6 else "G" for c in g["line_color"]]
            assert list(ln.line_color) == colour, (s, f, t, i)
            a = [ln.sx, ln.sy, ln.ex, ln.ey, ln.length, ln.width, ln.speed_limit]
            b = g["start"] + g["end"] + [g["length"], g["width"], g["speed_limit"]]
            if ln.kind == "C":
                a += [ln.cx, ln.cy, ln.radius, ln.ph0, ln.ph1, ln.dir]
                b += g["center"] + [g["radius"], g["start_phase"], g["end_phase"], g["direction"]]
            assert [float(x) for x in a] == [float(x) for x in b], (s, f, t, i)


def test_sockets_spawn_lanes_trigger_roads(golden_maps):
    for s, d in golden_maps.items():
        m = mapgen.generate_map(int(s))
        index = {id(ln): [f, t, i] for f, t, i, ln in _lanes(m)}
        for b, g in zip(m.blocks, d["blocks"]):
            socks = [dict(index=x.index, pos=list(x.pos), neg=list(x.neg)) for x in b.sockets.values()]
            assert socks == g["sockets"], (s, b.name)
            assert [list(r) for r in b.respawn] == g["respawn_roads"], (s, b.name)
            assert list(b.pre_socket.pos) == g["trigger_road"]
            if b.idx:
                assert [[index[id(ln)] for ln in ls] for ls in b.spawn_lanes()] == g["spawn_lanes"], (s, b.name)


def test_same_seed_same_map_and_named_sequence():
    a = mapgen.generate_map(1003)
    b = mapgen.generate_map(1003)
    assert [(x.sx, x.sy, x.ex, x.ey) for *_, x in _lanes(a)] == [(x.sx, x.sy, x.ex, x.ey) for *_, x in _lanes(b)]
    m = mapgen.generate_map(0, sequence="SSS")  # the reference's map="SSS" shorthand
    assert [b.id for b in m.blocks] == list("ISSS")
    with pytest.raises(KeyError):
        mapgen.generate_map(0, sequence="Q")


def test_thousand_seeds_match_reference_digests():
    """PGDrive-1000envs-v0 (seeds 1000..1999): lanes, sockets, spawn lanes and every reset decision hashed bit-exactly
    (tools/golden_hash.py) against digests computed from the unmodified reference (tools/make_golden.py digests)."""
    import os
    import sys
    from multiprocessing import get_context
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from conftest import load_golden
    gold = load_golden("digests_1000_1999.json.gz")
    seeds = sorted(int(s) for s in gold)
    with get_context("fork").Pool(min(8, os.cpu_count() or 1)) as pool:
        got = pool.map(_digest_of_seed, seeds, chunksize=16)
    bad = [s for s, g in zip(seeds, got) if list(g) != gold[str(s)]]
    assert not bad, "seeds whose map / reset digests differ from the reference: %s" % bad[:10]


def _digest_of_seed(seed):
    import golden_hash
    from pgdrive_b200 import episode
    m = mapgen.generate_map(seed)
    index = {id(ln): [f, t, i] for f, t, i, ln in _lanes(m)}
    lanes = []
    for f, t, i, ln in _lanes(m):
        rec = dict(frm=f, to=t, idx=i, kind=ln.kind, line_types=[str(x) for x in ln.line_types],
                   colours=list(ln.line_color), start=[ln.sx, ln.sy], end=[ln.ex, ln.ey], length=ln.length,
                   width=ln.width, speed_limit=ln.speed_limit)
        if ln.kind == "C":
            rec.update(center=[ln.cx, ln.cy], radius=ln.radius, start_phase=ln.ph0, end_phase=ln.ph1, direction=ln.dir)
        lanes.append(rec)
    blocks = []
    for b in m.blocks:
        spawn = [[index[id(ln)] for ln in ls] for ls in b.spawn_lanes()] if b.idx else []
        blocks.append([b.id, [[x.index, list(x.pos), list(x.neg)] for x in b.sockets.values()],
                       [list(r) for r in b.respawn], list(b.pre_socket.pos), spawn])
    ep = episode.make_episode(m, seed, 0.1)
    rec = dict(ego_seed=ep.ego_seed, ego_params=ep.ego_params, ego_checkpoints=ep.ego_checkpoints,
               block_vehicles=[(list(tr), [dict(type=v.type, lane=list(v.lane), long=v.long, seed=v.seed,
                                                params=v.params, idm_seed=v.idm_seed,
                                                overtake_timer=v.overtake_timer, checkpoints=v.checkpoints)
                                           for v in vs]) for tr, vs in ep.block_vehicles])
    return golden_hash.map_digest(lanes, blocks), golden_hash.episode_digest(rec), "".join(b.id for b in m.blocks)


def test_map_file_round_trip_in_the_reference_format(golden_maps, tmp_path):
    """dump_all_maps -> JSON -> load (test_loading_map_from_json.py:8-63): the dumped block sequences are the
    reference's, and tables built from the file equal tables built by the live block search."""
    import json
    import numpy as np
    from pgdrive_b200 import PGDriveEnv
    from pgdrive_b200.env import build_seed_tables, load_map_file, parse_map_config
    env = PGDriveEnv(dict(start_seed=1000, environment_num=4))
    data = env.dump_all_maps()
    assert set(data) == {"map_config", "map_data"} and sorted(data["map_data"]) == [1000, 1001, 1002, 1003]
    for s in data["map_data"]:
        gold = golden_maps[str(s)]["block_sequence"]
        mine = data["map_data"][s]["block_sequence"]
        assert [b["id"] for b in mine] == [b["id"] for b in gold]
        assert json.loads(json.dumps(mine)) == gold  # same keys, same float values after a JSON round trip
    path = tmp_path / "maps.json"
    path.write_text(json.dumps(data))
    mc = parse_map_config(env.config)
    stored = load_map_file(str(path), mc, [1000, 1001, 1002, 1003])
    assert stored is not None
    spawn = ((">", ">>", 0), 5.0, 0.0)
    a = build_seed_tables([1000, 1001, 1002, 1003], mc, 0.1, spawn, stored=stored)
    b = build_seed_tables([1000, 1001, 1002, 1003], mc, 0.1, spawn)
    for k in a:
        if hasattr(a[k], "tobytes"):
            assert a[k].tobytes() == b[k].tobytes(), k
    # a file made for another map_config, or not covering the seeds, is ignored (the reference falls back to BIG)
    assert load_map_file(str(path), dict(mc, lane_num=2), [1000]) is None
    assert load_map_file(str(path), mc, [1000, 1999]) is None
    env2 = PGDriveEnv(dict(start_seed=1000, environment_num=4, _load_map_from_json=str(path)))
    assert env2._stored is not None and 1002 in env2._stored


REF_JSON = "/root/reference/pgdrive/assets/maps/20210814_generated_maps_start_seed_0_environment_num_30000.json"


def _search(seed):
    return seed, mapgen.search_sequence(seed)


@pytest.mark.skipif(not __import__("os").path.exists(REF_JSON), reason="reference assets only exist in the build container")
def test_block_search_against_the_shipped_30000_map_file():
    """The reference ships the block sequences of seeds 0..29999 (its golden asset, proven identical to live BIG by
    test_loading_map_from_json.py).  Every tenth seed (3000 maps) must come out of our block search identically, and the
    file itself loads through load_map_file."""
    import json
    import os
    from multiprocessing import get_context
    from pgdrive_b200.env import default_config, load_map_file, parse_map_config
    with open(REF_JSON) as f:
        data = json.load(f)
    seeds = list(range(0, 30000, 10))
    with get_context("fork").Pool(min(8, os.cpu_count() or 1)) as pool:
        got = dict(pool.map(_search, seeds, chunksize=16))
    bad = [s for s in seeds if json.loads(json.dumps(got[s])) != data["map_data"][str(s)]["block_sequence"]]
    assert not bad, "block sequences differ from the shipped file for seeds %s" % bad[:10]
    stored = load_map_file(data, parse_map_config(default_config()), range(1000, 1100))
    assert stored is not None and len(stored) == 30000


def test_random_lane_width_and_number_match_reference():
    """random_lane_width / random_lane_num (manager/map_manager.py:157-169): the per-seed lane configuration and the
    map generated with it, against the reference's own add_random_to_map + live block search
    (tests/golden/maps_random_lane.json.gz, tools/make_golden.py random_lane)."""
    from conftest import load_golden
    from pgdrive_b200.env import seed_map_config
    base = dict(type="block_num", config=3, lane_width=3.5, lane_num=3, exit_length=50)
    gold = load_golden("maps_random_lane.json.gz")
    assert len(gold) >= 12
    for s, rec in gold.items():
        seed = int(s)
        for name, flags in (("both", (True, True)), ("width", (True, False)), ("num", (False, True))):
            mc = seed_map_config(base, seed, *flags)
            assert mc["lane_width"] == rec[name]["lane_width"] and mc["lane_num"] == rec[name]["lane_num"], (s, name)
        assert seed_map_config(base, seed) is base
        mc = seed_map_config(base, seed, True, True)
        m = mapgen.generate_map(seed, block_num=3, lane_num=mc["lane_num"], lane_width=mc["lane_width"])
        assert [b["id"] for b in m.block_sequence] == [b["id"] for b in rec["block_sequence"]], s
        mine = _lanes(m)
        assert len(mine) == len(rec["lanes"]), s
        for (f, t, i, ln), g in zip(mine, rec["lanes"]):
            assert (f, t, i, ln.kind) == (g["frm"], g["to"], g["idx"], g["kind"]), s
            assert [str(x) for x in ln.line_types] == g["line_types"], (s, f, t, i)
            a = [ln.sx, ln.sy, ln.ex, ln.ey, ln.length, ln.width]
            b = g["start"] + g["end"] + [g["length"], g["width"]]
            if ln.kind == "C":
                a += [ln.cx, ln.cy, ln.radius, ln.ph0, ln.ph1, ln.dir]
                b += g["center"] + [g["radius"], g["start_phase"], g["end_phase"], g["direction"]]
            assert [float(x) for x in a] == [float(x) for x in b], (s, f, t, i)
        # the simulator drives on the search-time map with these options: sockets and spawn roads must agree too
        for b, g in zip(m.blocks, rec["blocks"]):
            socks = [dict(index=x.index, pos=list(x.pos), neg=list(x.neg)) for x in b.sockets.values()]
            assert socks == g["sockets"], (s, b.name)
            assert [list(r) for r in b.respawn] == g["respawn_roads"], (s, b.name)


def test_random_lane_options_need_live_map_generation():
    from pgdrive_b200.config import check_supported, default_config
    cfg = default_config()
    cfg.update(dict(random_lane_width=True))
    with pytest.raises(AssertionError):
        check_supported(cfg)
    cfg.update(dict(load_map_from_json=False))
    check_supported(cfg)
