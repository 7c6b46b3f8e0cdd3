"""One GPU: time the gather's row kernels alone on real observation rows of the steady state -- pgd_pack_rows,
pgd_expand_rows, pgd_words_checksum (the consumer's read) -- and check the round trip bit for bit.
Usage: python tools/rows_bench.py [n_rows=458752] [preroll=2048]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from pgdrive_b200 import VecPGDriveEnv, cabi
    n_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 7 * 65536
    preroll = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    n = 65536
    env = VecPGDriveEnv(dict(num_envs=n, start_seed=1000, environment_num=100, traffic_density=0.1))
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    acts = torch.rand((256, n, 2), generator=g, device="cuda") * 2 - 1
    for t in range(preroll):
        obs = env.step(acts[t % 256])[0]
    d = obs.shape[1]
    rows = obs.repeat((n_rows + n - 1) // n, 1)[:n_rows].contiguous()
    env.step(acts[0])
    rows2 = env.step(acts[1])[0].repeat((n_rows + n - 1) // n, 1)[:n_rows].contiguous()  # the same rows two steps on
    lib, h = env.engine.lib, env.engine.h
    stride = lib.pgd_packed_row_words(d)
    packed = torch.zeros((n_rows, stride), device="cuda")
    back = torch.zeros((n_rows, d), device="cuda")
    acc = torch.zeros(1, dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    hits = (rows[:, -240:].view(torch.int32) != 0x3f800000).sum(1).float()
    wire = (((d - 240 + 8) + hits + 7) // 8 * 8 * 4).mean().item()

    def timed(fn, reps=10):
        ms = []
        for _ in range(reps):
            flush.fill_(1)  # the kernels are timed on cold rows, as in the gather (503 MB per step do not stay in the L2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        return ms[len(ms) // 2]

    t_pack = timed(lambda: cabi.check(lib, lib.pgd_pack_rows(rows.data_ptr(), packed.data_ptr(), n_rows, d, st)))
    t_exp = timed(lambda: cabi.check(lib, lib.pgd_expand_rows(packed.data_ptr(), back.data_ptr(), n_rows, d, st)))
    # delta expansion: the buffer holds the rows of two steps ago (two gather buffers alternate)
    packed2 = torch.zeros((n_rows, stride), device="cuda")
    state = torch.zeros((n_rows, 8), dtype=torch.int32, device="cuda")
    cabi.check(lib, lib.pgd_pack_rows(rows2.data_ptr(), packed2.data_ptr(), n_rows, d, st))
    cabi.check(lib, lib.pgd_expand_rows_delta(packed.data_ptr(), back.data_ptr(), state.data_ptr(), n_rows, d, 1, st))
    flip = [0]

    def delta():
        flip[0] ^= 1
        src = packed2 if flip[0] else packed
        cabi.check(lib, lib.pgd_expand_rows_delta(src.data_ptr(), back.data_ptr(), state.data_ptr(), n_rows, d, 0, st))

    t_delta = timed(delta)
    ok_delta = torch.equal((rows2 if flip[0] else rows).view(torch.int32), back.view(torch.int32))
    cabi.check(lib, lib.pgd_expand_rows(packed.data_ptr(), back.data_ptr(), n_rows, d, st))
    t_sum = timed(lambda: cabi.check(lib, lib.pgd_words_checksum(h, back.data_ptr(), n_rows * d * 4, acc.data_ptr(), st)))
    ok = torch.equal(rows.view(torch.int32), back.view(torch.int32))
    dense = n_rows * d * 4
    print(json.dumps(dict(
        rows=n_rows, obs_dim=d, packed_stride_words=stride, mean_hits=hits.mean().item(), wire_bytes_per_row=wire,
        round_trip_bit_exact=ok,
        pack=dict(ms=t_pack, ns_per_row=t_pack * 1e6 / n_rows, gbs=(dense + wire * n_rows) / t_pack / 1e6),
        expand=dict(ms=t_exp, ns_per_row=t_exp * 1e6 / n_rows, gbs=(dense + wire * n_rows) / t_exp / 1e6),
        expand_delta=dict(ms=t_delta, ns_per_row=t_delta * 1e6 / n_rows, bit_exact=ok_delta),
        consume=dict(ms=t_sum, ns_per_row=t_sum * 1e6 / n_rows, gbs=dense / t_sum / 1e6))))
    env.close()


if __name__ == "__main__":
    main()
