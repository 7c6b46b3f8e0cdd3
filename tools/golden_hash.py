"""Canonical digests of a generated map / episode template, shared by tools/make_golden.py (reference side) and
tests/test_mapgen.py (this repo's side).  Floats are hashed by repr(), i.e. bit-exactly."""
import hashlib
import json


def map_digest(lanes, blocks):
    """lanes: list of dicts (frm, to, idx, kind, line_types, colours 'G'/'Y', start, end, length, width, speed_limit and
    for arcs center, radius, start_phase, end_phase, direction); blocks: list of (id, sockets, respawn, trigger)."""
    rows = []
    for l in lanes:
        row = [l["frm"], l["to"], l["idx"], l["kind"], list(l["line_types"]), list(l["colours"])]
        nums = list(l["start"]) + list(l["end"]) + [l["length"], l["width"], l["speed_limit"]]
        if l["kind"] == "C":
            nums += list(l["center"]) + [l["radius"], l["start_phase"], l["end_phase"], l["direction"]]
        row.append([repr(float(x)) for x in nums])
        rows.append(row)
    return hashlib.sha1(json.dumps([rows, blocks], sort_keys=True).encode()).hexdigest()


def episode_digest(rec):
    """rec: dict(ego_seed, ego_params, ego_checkpoints, block_vehicles=[(trigger, [vehicle dicts])])."""
    def fl(d):
        return {k: repr(float(v)) for k, v in sorted(d.items())}

    rows = [rec["ego_seed"], fl(rec["ego_params"]), list(rec["ego_checkpoints"])]
    for trigger, vs in rec["block_vehicles"]:
        rows.append([list(trigger), [[v["type"], list(v["lane"]), repr(float(v["long"])), v["seed"], fl(v["params"]),
                                      v["idm_seed"], v["overtake_timer"], list(v["checkpoints"])] for v in vs]])
    return hashlib.sha1(json.dumps(rows, sort_keys=True).encode()).hexdigest()
