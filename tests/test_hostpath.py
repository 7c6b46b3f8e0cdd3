"""Host half of pgd_step_host (pgdrive_b200/csrc/pgd_hostpath.cu) without a GPU: the row expansion against a numpy
restatement of the packed format, full and delta, and the thread pool's self-test."""
import ctypes

import numpy as np
import pytest

HEAD, BEAMS = 34, 240


def _lib():
    import __graft_entry__
    __graft_entry__.build()
    from pgdrive_b200 import cabi
    return cabi.load_library()


def _pack(rows):
    """numpy restatement of what pgd_pack_compact_kernel writes for one group: [head | 8 mask words] rows + hit values in
    row and beam order."""
    n, d = rows.shape
    head = d - BEAMS
    bits = rows[:, head:].view(np.uint32) != 0x3f800000
    padded = np.zeros((n, 256), bool)
    padded[:, :BEAMS] = bits
    words = (padded.reshape(n, 8, 32) * (1 << np.arange(32, dtype=np.uint64))).sum(2).astype(np.uint32)
    base = np.concatenate([rows[:, :head], words.view(np.float32)], 1).copy()
    hits = rows[:, head:][bits].copy()  # row-major boolean indexing = row order, beams in order
    return base, hits, words


def _rows(rs, n, d, p_hit):
    rows = rs.uniform(0, 1, (n, d)).astype(np.float32)
    beams = np.where(rs.uniform(size=(n, BEAMS)) < p_hit, rs.uniform(0, 1, (n, BEAMS)), 1.0).astype(np.float32)
    rows[:, d - BEAMS:] = beams
    return rows


@pytest.mark.parametrize("d", [274, 240, 274 + 60])
def test_host_expansion_full_and_delta(d):
    lib = _lib()
    rs = np.random.RandomState(3)
    n = 300
    dense = np.full((n, d), np.nan, np.float32)
    state = np.zeros((n, 8), np.uint32)
    for t in range(12):
        rows = _rows(rs, n, d, [0.0, 0.002, 0.05, 1.0, 0.0, 0.3][t % 6])
        if t == 5:
            rows[7, d - 1] = 0.0           # 0.0 and 1 - ulp are hits
            rows[8, d - BEAMS] = np.nextafter(np.float32(1), np.float32(0))
        base, hits, words = _pack(rows)
        pad = np.concatenate([np.full(5, np.nan, np.float32), hits, np.full(3, np.nan, np.float32)])  # offset 5
        used = lib.pgd_host_expand_rows(base.ctypes.data, pad.ctypes.data, 5, n, d, dense.ctypes.data, state.ctypes.data,
                                        1 if t == 0 else 0)
        assert used == len(hits)
        assert np.array_equal(dense.view(np.uint32), rows.view(np.uint32)), t
        assert np.array_equal(state, words)
    # a full expansion does not depend on what the destination or the state held
    dense[:] = 7.0
    state[:] = 0
    rows = _rows(rs, n, d, 0.1)
    base, hits, _ = _pack(rows)
    assert lib.pgd_host_expand_rows(base.ctypes.data, hits.ctypes.data, 0, n, d, dense.ctypes.data, state.ctypes.data,
                                    1) == len(hits)
    assert np.array_equal(dense.view(np.uint32), rows.view(np.uint32))
    assert lib.pgd_host_expand_rows(None, hits.ctypes.data, 0, n, d, dense.ctypes.data, state.ctypes.data, 1) == -1
    assert lib.pgd_host_expand_rows(base.ctypes.data, hits.ctypes.data, 0, n, 100, dense.ctypes.data, state.ctypes.data,
                                    1) == -1


def test_host_pool_runs_every_item_exactly_once():
    lib = _lib()
    for workers, items, rounds in ((1, 50, 20), (4, 1, 200), (8, 128, 300), (16, 7, 300), (3, 0, 10)):
        assert lib.pgd_host_pool_selftest(workers, items, rounds) == 0, (workers, items, rounds)
    assert lib.pgd_host_pool_selftest(0, 1, 1) == -1
