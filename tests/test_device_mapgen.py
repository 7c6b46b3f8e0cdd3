"""The device reset path (pgdrive_b200/csrc/pgd_mapgen.cuh): reference RNG streams, correctly rounded trigonometry,
block search, tables and episode templates.

CPU tests run the HOST build of the generator's source (oracle/mapgen_host.cpp) against numpy's RandomState, glibc
and the reference-pinned Python reset path (pgdrive_b200/mapgen.py + episode.py + tables.py, themselves bit-exact
against fixtures of the unmodified reference).  GPU tests run the sm_100a kernel through the C-ABI and require its
tables to equal the host build's bit for bit.

Known, documented gap: the reference decides discrete things on the last bit of libm results (an intersection exit is
30.000000000000004 m or 29.999999999999993 m long and holds int(length / 10) = 3 or 2 traffic spawn slots); the
generator's trigonometry is correctly rounded, glibc's is not always (0.13 % of calls differ in the last bit), so about
one map in two thousand draws different traffic (17 of the 29 000 seeds 1000-29999 differ in any table byte).  Such
seeds are listed below, not silently skipped.
"""
import math

import numpy as np
import pytest

V0 = dict(type="block_num", config=3, lane_num=3, lane_width=3.5, exit_length=50)
SPAWN = ((">", ">>", 0), 5.0, 0.0)
LIBM_TIE_SEEDS = {1055, 245}  # last-bit libm ties (see module docstring); 1055 at 3 lanes, 245 at 2 lanes
TABLE_KEYS = ("maps", "roads", "lanes", "boxes", "cell_start", "cell_entries", "episodes", "slots", "route_nodes",
              "route_roads")


def _host():
    from oracle import mapgen_host
    return mapgen_host


def _same(a, b):
    return all(len(a[k]) == len(b[k]) and a[k].tobytes() == b[k].tobytes() for k in TABLE_KEYS)


def _python_tables(seed, mc=V0, density=0.1, spawn=SPAWN):
    from pgdrive_b200 import env
    return env._seed_tables((seed, mc, density, spawn))


# ---------------------------------------------------------------------------------------------------- CPU
def test_seed_hash_matches_the_reference_hash():
    from pgdrive_b200 import rng
    lib = _host().lib()
    for v in [0, 1, 9, 10, 99, 1000, 4242, 65535, 999999, 2**31 - 1, 2**40 + 17]:
        assert lib.pgd_host_hash_seed(v) == rng.hash_seed(v)


def test_random_streams_match_numpy_legacy_randomstate():
    from pgdrive_b200 import rng
    probs = [0.3, 0.1, 0.1, 0.1, 0.15, 0.15, 0.1, 0, 0, 0, 0, 0, 0]
    ops = [(0, 13), (1, 0), (0, 10000), (0, 1), (0, 1000000), (2, 13), (0, 65536), (3, 57), (0, 50), (0, 25), (1, 0),
           (0, 3), (3, 200), (0, 2), (2, 13), (3, 2), (3, 1)]
    for seed in [0, 5, 1000, 4242, 999999]:
        rs = rng.seeded(seed)
        want = []
        for op, a in ops:
            if op == 0:
                want.append(int(rs.randint(0, a)))
            elif op == 1:
                want.append(rs.random_sample())
            elif op == 2:
                want.append(int(rs.choice(13, p=probs)))
            else:
                lst = list(range(a))
                rs.shuffle(lst)
                want += lst
        got = _host().rng_script(seed, ops, probs)
        assert np.array_equal(got, np.array(want, np.float64)), seed
    # choice over a list of names draws like randint(0, len), and a single-element list consumes nothing
    rs = rng.seeded(3)
    want = [["a", "b", "c"].index(str(rs.choice(["a", "b", "c"]))), 0 * len(str(rs.choice(["a"]))),
            int(rs.randint(0, 7))]
    assert list(_host().rng_script(3, [(0, 3), (0, 1), (0, 7)])) == want


def test_trigonometry_is_correctly_rounded():
    """Against exact rational Taylor sums, and close to glibc (which is not always correctly rounded)."""
    from fractions import Fraction
    lib = _host().lib()

    def taylor(x, first, k0):
        x = Fraction(x)
        term, total, k = first, first, k0
        while abs(term) > Fraction(1, 10**45):
            term = -term * x * x / (k * (k + 1))
            total += term
            k += 2
        return total

    rs = np.random.RandomState(0)
    for x in list(rs.uniform(-3, 3, 40)) + [0.58456527268280567, -0.11020694993154656, -2.9902229133126887]:
        x = float(x)
        s, c = lib.pgd_host_sin(x), lib.pgd_host_cos(x)
        assert abs(Fraction(s) - taylor(x, Fraction(x), 2)) <= Fraction(math.ulp(s)) / 2, x
        assert abs(Fraction(c) - taylor(x, Fraction(1), 1)) <= Fraction(math.ulp(c)) / 2, x
    xs = rs.uniform(-10, 10, 20000)
    mism = sum(lib.pgd_host_sin(float(x)) != math.sin(x) for x in xs) + sum(
        lib.pgd_host_cos(float(x)) != math.cos(x) for x in xs)
    assert mism < 0.004 * 2 * len(xs)  # ~0.13 % expected: glibc's own misroundings
    for x in xs[:2000]:
        assert abs(lib.pgd_host_sin(float(x)) - math.sin(x)) <= math.ulp(math.sin(x))
        assert abs(lib.pgd_host_cos(float(x)) - math.cos(x)) <= math.ulp(math.cos(x))
    ys, zs = rs.uniform(-400, 400, 5000), rs.uniform(-400, 400, 5000)
    for y, z in zip(ys, zs):
        assert abs(lib.pgd_host_atan2(float(y), float(z)) - math.atan2(y, z)) <= math.ulp(math.atan2(y, z))
    assert lib.pgd_host_atan2(0.0, -1.0) == math.pi and lib.pgd_host_atan2(1.0, 0.0) == math.pi / 2
    assert lib.pgd_host_atan2(0.0, 2.0) == 0.0 and lib.pgd_host_atan2(-3.0, 0.0) == -math.pi / 2
    assert lib.pgd_host_cos(math.pi / 2) == math.cos(math.pi / 2)


@pytest.mark.parametrize("lo,n", [(1000, 30), (0, 10), (5000, 10)])
def test_host_build_equals_reference_pinned_tables(lo, n):
    from pgdrive_b200 import devgen
    gc = devgen.make_gen_config(V0, 0.1, SPAWN)
    caps = devgen.caps_for(gc)
    for s in range(lo, lo + n):
        rc, got, seq = _host().generate(s, gc, caps)
        assert rc == 0, (s, devgen.GEN_ERRORS.get(rc))
        assert _same(_python_tables(s), got) or s in LIBM_TIE_SEEDS, s


def test_host_build_block_sequence_matches_the_reference_fixture():
    """The block sequence the device search settles on (ids, socket indices, float32 parameters) equals what the
    unmodified reference produced (tests/golden/maps_v0_1000_1099.json.gz, made by tools/make_golden.py)."""
    from conftest import load_golden
    from pgdrive_b200 import devgen
    gold = load_golden("maps_v0_1000_1099.json.gz")
    gc = devgen.make_gen_config(V0, 0.1, SPAWN)
    caps = devgen.caps_for(gc)
    for seed in range(1000, 1040):
        ref = gold[str(seed)]["block_sequence"]
        rc, _, seq = _host().generate(seed, gc, caps)
        assert rc == 0 and len(seq) == len(ref)
        for i, (r, b) in enumerate(zip(seq, ref)):
            assert devgen.CODE_BLOCK[int(r[0])] == b["id"], (seed, i)
            if i == 0:
                continue
            owner = int(r[1])
            assert b["pre_block_socket_index"] == "%d%s-socket%d" % (owner, devgen.CODE_BLOCK[int(seq[owner][0])], r[2])
            f = r[8:13].view(np.float32)
            for name, val in (("length", f[0]), ("radius", f[1]), ("angle", f[2]), ("exit_radius", f[3]),
                              ("inner_radius", f[4])):
                if name in b:
                    assert np.float32(b[name]) == val, (seed, i, name)
            for name, val in (("dir", r[3]), ("change_lane_num", r[4]), ("decrease_increase", r[5]), ("t_type", r[6])):
                if name in b:
                    assert int(b[name]) == int(val), (seed, i, name)


@pytest.mark.parametrize("mc,density,seeds,spawn", [
    (dict(V0, lane_num=2), 0.1, range(200, 212), SPAWN),
    (dict(V0, lane_num=1), 0.2, range(0, 8), SPAWN),
    (dict(V0, lane_num=4, lane_width=3.0), 0.1, range(0, 8), SPAWN),
    (dict(V0, config=7), 0.1, range(0, 6), SPAWN),
    (V0, 0.0, range(0, 6), SPAWN),
    (dict(V0, type="block_sequence", config="XTO"), 0.1, range(0, 6), SPAWN),
    (dict(V0, type="block_sequence", config="SCrRXTO"), 0.1, range(0, 5), SPAWN),
    (dict(V0, exit_length=70), 0.1, range(0, 6), ((">", ">>", 1), 8.0, 0.5)),
])
def test_host_build_other_map_configs(mc, density, seeds, spawn):
    from pgdrive_b200 import devgen
    gc = devgen.make_gen_config(mc, density, spawn)
    caps = devgen.caps_for(gc)
    for s in seeds:
        rc, got, _ = _host().generate(s, gc, caps)
        assert rc == 0, (s, devgen.GEN_ERRORS.get(rc))
        assert _same(_python_tables(s, mc, density, spawn), got) or s in LIBM_TIE_SEEDS, s


def test_generator_reports_overflow_instead_of_truncating():
    from pgdrive_b200 import devgen
    gc = devgen.make_gen_config(V0, 0.1, SPAWN)
    caps = devgen.caps_for(gc)
    caps.boxes = 100
    rc, _, _ = _host().generate(1000, gc, caps)
    assert rc == 3
    caps = devgen.caps_for(gc)
    caps.lanes = 20
    assert _host().generate(1000, gc, caps)[0] == 1
    gc2 = devgen.make_gen_config(V0, 0.5, SPAWN)  # dense traffic: more than 32 vehicle slots
    assert _host().generate(5006, gc2, devgen.caps_for(gc2))[0] == 9
    with pytest.raises(ValueError):
        devgen.make_gen_config(dict(V0, type="block_sequence", config="SZ"), 0.1, SPAWN)


# ---------------------------------------------------------------------------------------------------- GPU
def _device_tables(seeds, mc=V0, density=0.1, spawn=SPAWN, slots=32):
    from pgdrive_b200 import devgen
    from pgdrive_b200.config import ENGINE_CONFIG, default_config
    from pgdrive_b200.env import _Engine
    cfg = default_config()
    cfg.update(ENGINE_CONFIG)
    eng = _Engine(cfg, 1, slots, 0, True)
    gc = devgen.make_gen_config(mc, density, spawn)
    status, counts = devgen.generate(eng, list(seeds), gc)
    T = devgen.download(eng)
    eng.close()
    return gc, T, status, counts


@pytest.mark.gpu
def test_device_tables_equal_host_build_bit_for_bit():
    from pgdrive_b200 import devgen
    seeds = list(range(1000, 1100))
    gc, T, status, counts = _device_tables(seeds)
    assert (status == 0).all()
    caps = devgen.caps_for(gc)
    pinned = 0
    for m, s in enumerate(seeds):
        rc, want, _ = _host().generate(s, gc, caps)
        assert rc == 0
        got = devgen.compact(T, m)
        assert _same(want, got), s
        assert list(counts[m][:3]) == [len(want["lanes"]), len(want["roads"]), len(want["boxes"])]
        if m % 4 == 0:  # and against the reference-pinned Python path
            assert _same(_python_tables(s), got) or s in LIBM_TIE_SEEDS, s
            pinned += 1
    assert pinned == 25


@pytest.mark.gpu
def test_device_tables_other_configs_and_1000_seeds():
    from pgdrive_b200 import devgen
    for mc, density, seeds in [(dict(V0, lane_num=2), 0.1, range(200, 230)), (dict(V0, config=7), 0.1, range(0, 12)),
                               (dict(V0, type="block_sequence", config="CrXRO"), 0.1, range(0, 12))]:
        gc, T, status, _ = _device_tables(list(seeds), mc, density)
        caps = devgen.caps_for(gc)
        assert (status == 0).all()
        for m, s in enumerate(seeds):
            assert _same(_host().generate(s, gc, caps)[1], devgen.compact(T, m)), (mc, s)
    # PGDrive-1000envs-v0's seeds in one launch: every 10th checked against the host build
    seeds = list(range(1000, 2000))
    gc, T, status, counts = _device_tables(seeds)
    assert (status == 0).all() and (counts[:, 5] <= 32).all()
    caps = devgen.caps_for(gc)
    for m in range(0, 1000, 10):
        assert _same(_host().generate(seeds[m], gc, caps)[1], devgen.compact(T, m)), seeds[m]


@pytest.mark.gpu
def test_env_on_device_generated_maps_matches_env_on_host_tables():
    """Same seeds, same actions: an environment whose maps were generated on the GPU reproduces, bit for bit, the one
    whose tables were built by the reference-pinned host path -- including seed 1055, a libm tie seed, which the device
    path takes from the host builder (devgen.tie_seeds) -- and the CPU oracle agrees on the downloaded tables."""
    import torch
    from oracle.oracle import Oracle
    from pgdrive_b200 import VecPGDriveEnv
    n, steps = 200, 150
    common = dict(start_seed=1000, environment_num=100, num_envs=n, traffic_density=0.1)
    dev = VecPGDriveEnv(dict(common))  # the default: maps and episode templates generated on the GPU
    host = VecPGDriveEnv(dict(common, device_mapgen=False))
    assert dev.reset_path == "device" and host.reset_path == "host" and dev.device_mapgen_patched == [1055]
    assert dev.engine.num_slots == host.engine.num_slots
    ref = Oracle(dev.T, n, auto_reset=True, num_slots=dev.engine.num_slots)
    o1, o2 = dev.reset().cpu().numpy(), host.reset().cpu().numpy()
    ro = ref.reset(range(n), [dev.episode_of_seed[int(s)] for s in dev.env_seeds])
    assert np.array_equal(o1, o2) and np.array_equal(o1, ro)
    rs = np.random.RandomState(1)
    for t in range(steps):
        a = rs.uniform(-1, 1, (n, 2)).astype(np.float32)
        a[:, 1] = np.abs(a[:, 1]) * 0.8 + 0.2
        at = torch.from_numpy(a).cuda()
        r1 = [x.cpu().numpy().copy() for x in dev.step(at)[:3]]
        r2 = [x.cpu().numpy().copy() for x in host.step(at)[:3]]
        for x, y in zip(r1, r2):
            assert np.array_equal(x, y), t
        oo, rr, dd, _ = ref.step(a)
        assert np.array_equal(r1[2], dd) and np.array_equal(r1[0], oo) and np.array_equal(r1[1], rr), t
    dev.close()
    host.close()


@pytest.mark.gpu
def test_device_generation_errors_are_reported():
    from pgdrive_b200 import devgen
    from pgdrive_b200.config import ENGINE_CONFIG, default_config
    from pgdrive_b200.env import _Engine
    cfg = default_config()
    cfg.update(ENGINE_CONFIG)
    eng = _Engine(cfg, 1, 16, 0, True)
    gc = devgen.make_gen_config(V0, 0.5, SPAWN)
    with pytest.raises(RuntimeError, match="vehicle slots"):
        devgen.generate(eng, [5006, 5007], gc)
    caps = devgen.caps_for(gc)
    caps.boxes = 64
    with pytest.raises(RuntimeError, match="box table full"):
        devgen.generate(eng, [1000], devgen.make_gen_config(V0, 0.1, SPAWN), caps)
    eng.close()


def test_host_build_property_random_map_configs():
    """Randomly drawn map configurations (1-4 lanes, 1-6 blocks, lane width, exit length, density, seed): the
    generator's host build equals the reference-pinned Python reset path bit for bit (or reports the same overflow)."""
    from hypothesis import given, settings, strategies as st
    from pgdrive_b200 import devgen

    @settings(max_examples=30, deadline=None, derandomize=True)
    @given(st.integers(1, 4), st.integers(1, 6), st.sampled_from([3.0, 3.25, 3.5, 4.0, 4.5]),
           st.sampled_from([40, 50, 70]), st.sampled_from([0.0, 0.05, 0.1, 0.15]), st.integers(0, 20000))
    def check(lane_num, blocks, width, exit_length, density, seed):
        mc = dict(type="block_num", config=blocks, lane_num=lane_num, lane_width=width, exit_length=exit_length)
        gc = devgen.make_gen_config(mc, density, SPAWN)
        rc, got, _ = _host().generate(seed, gc, devgen.caps_for(gc))
        want = _python_tables(seed, mc, density, SPAWN)
        if int(want["max_slots"]) > 32:
            assert rc == 9
            return
        assert rc == 0, devgen.GEN_ERRORS.get(rc)
        assert _same(want, got), (lane_num, blocks, width, exit_length, density, seed)

    check()


def test_host_build_random_lane_width_and_number():
    """random_lane_width / random_lane_num decided per seed inside the generator (manager/map_manager.py:157-169)."""
    from pgdrive_b200 import devgen
    from pgdrive_b200.env import seed_map_config
    for flags in ((True, True), (True, False), (False, True)):
        gc = devgen.make_gen_config(V0, 0.1, SPAWN, random_lane=flags)
        caps = devgen.caps_for(gc)
        for s in list(range(1000, 1010)) + [5, 77, 2500]:
            rc, got, _ = _host().generate(s, gc, caps)
            assert rc == 0
            want = _python_tables(s, seed_map_config(V0, s, *flags))
            assert _same(want, got), (flags, s)
            if flags[1]:
                assert int(got["maps"]["lane_num"][0]) == 2
