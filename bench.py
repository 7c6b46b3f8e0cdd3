#!/usr/bin/env python
"""Benchmark of the hot path: env-steps/s of the fused step kernel (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--envs E] [--workload v0|1000envs]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Own arm: every rank steps `--envs` (default 65 536) PGDrive-v0 environments (seeds 1000..1099, traffic density
0.1, 240-beam lidar, 16 vehicle slots); a "step" is ONE kernel launch advancing all of them by one decision step
(5 physics sub-steps + observation + reward/done, auto-reset of finished episodes).  Every timed region starts
after an untimed pre-roll of the same policy (2 048 steps for the random policy, whose episodes last ~1 000 steps;
256 for the driving policy), so the number does not depend on where in the episode distribution the timer starts
(fresh resets are the cheapest state of the simulator: 0.10 ms per step against 0.29 ms in the steady state).  Weak scaling: with N ranks
the job simulates N x 65 536 environments and every step rank 0 receives the whole observation / reward / done
batch (--gather, DESIGN.md "Multi-GPU") and READS it (a checksum kernel stands for the policy network).
  value        device-resident, random policy of BASELINE.md: actions pre-generated in HBM, CUDA-event time of K
               steps, max over ranks
  driving      the same kernel under a policy that drives (traffic awake, lidar hits, frequent resets), with its
               own roofline block
  e2e          same steps through the public host-buffer API: N = 1 VecPGDriveEnv.step(numpy) -> pgd_step_host
               (pinned H2D of the actions, D2H of obs / reward / done / info inside the timed region); N > 1 the
               actions of the WHOLE batch start on rank 0's host and the gathered results end there
  roofline     algorithmic bytes per env-step (DESIGN.md "Bytes") x envs / average kernel time vs measured HBM peak;
               `issue` = the secondary bound: warp instructions per launch (ncu, profiles/) against the SMs' issue rate
  cpu_baseline the CPU oracle (oracle/pgd_oracle.c, a scalar port of the same step) on all host cores, bounded sample
  sim_only     N > 1: the same K steps without the gather (does the kernel itself scale?)

Reference arm (--impl reference): the reference's own step runs on Panda3D/Bullet, which is neither vendored nor
installable offline, so this arm times the CPU oracle port on all host cores on the same config and metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

OBS_DIM = 274
INFO_BYTES = 40
# (PGDRIVE_B200_BENCH_PREROLL shortens both pre-rolls for the ncu launch-list run only: that run is about WHICH kernels
# launch, never about a number)
PREROLL = int(os.environ.get("PGDRIVE_B200_BENCH_PREROLL", 2048))  # untimed random-policy steps before the timed region: the cost of a step climbs from 0.10 ms right
# after a reset to a 0.33 ms peak near step 400 and settles by step ~2000 (mean episode ~1000 steps,
# profiles/r02i_cost_curve_*.log); the driving policy (mean episode ~50 steps) is steady after 256
PREROLL_DRIVING = min(256, PREROLL)
SMS, ISSUE_PER_SM_CLK = 148, 4  # B200: 148 SMs x 4 warp schedulers, one warp instruction per scheduler per clock

WORKLOADS = {
    # name: (first seed, number of seeds, vehicle slots, description)
    "v0": (1000, 100, 16, "%d envs PGDrive-v0 (seeds 1000-1099), traffic_density=0.1, 240-beam lidar, 16 vehicle slots"),
    "1000envs": (1000, 1000, 24, "%d envs PGDrive-1000envs-v0 (seeds 1000-1999, 1000 distinct maps), "
                                 "traffic_density=0.1, 240-beam lidar, 24 vehicle slots"),
}


def algorithmic_bytes_per_env_step(num_slots):
    """DESIGN.md 'Bytes moved per env-step': action 8 R + obs 1096 W + reward 4 W + done 1 W + info 40 W, the
    structure-of-arrays state read and written once (80 B per vehicle slot + 32 B per env, each way), and the
    map / template records amortised over the environments sharing a map (~80 B, L2-resident)."""
    io = 8 + 4 * OBS_DIM + 4 + 1 + INFO_BYTES
    state = 2 * (80 * num_slots + 32)
    return io + state + 80


class ClockSampler:
    """nvidia-smi clocks and throttle reasons under load (B200_PROFILING.md recipe).  nvidia-smi needs up to a second to
    deliver its first line and the timed region is 40 ms, so the sampler starts early, `mark()` is called when the GPU
    starts running the step kernel back to back (pre-roll -> warm-up -> timed steps, no idle gap), and only samples
    that arrive between the mark and `stop()` -- called right after the timed region -- count."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.t0 = [], None, index, None

    def mark(self):
        self.t0 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms",
                 "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True
            )
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_end = time.time()
        rows = [r for t, r in self.rows if self.t0 is None or self.t0 + 0.05 <= t <= t_end]
        window = "pre-roll + warm-up + timed steps (the step kernel back to back)"
        if not rows:  # nvidia-smi too slow to start: better the samples around the region than none
            rows, window = [r for _, r in self.rows], "all samples of the run (none arrived inside the loaded window)"
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm), window=window)


def host_threads():
    return len(os.sched_getaffinity(0))


def cpu_oracle_rate(T, n_envs, steps, warmup, threads, episode_ids, num_slots, seed=0):
    """env-steps/s of the CPU oracle on `threads` host threads (uniform [-1,1]^2 actions, auto-reset)."""
    from oracle.oracle import Oracle
    ref = Oracle(T, n_envs, auto_reset=True, num_slots=num_slots, fast=True)  # bucket-grid queries: not a strawman
    ref.reset(range(n_envs), episode_ids)
    rs = np.random.RandomState(seed)
    acts = rs.uniform(-1, 1, (warmup + steps, n_envs, 2)).astype(np.float32)
    for t in range(warmup):
        ref.step(acts[t], threads=threads)
    t0 = time.perf_counter()
    for t in range(warmup, warmup + steps):
        ref.step(acts[t], threads=threads)
    dt = time.perf_counter() - t0
    ref.close()
    return n_envs * steps / dt, dt


def build_tables(workload="v0"):
    from pgdrive_b200.env import build_seed_tables, default_config, parse_map_config
    first, count, _, _ = WORKLOADS[workload]
    return build_seed_tables(range(first, first + count), parse_map_config(default_config()), 0.1,
                             ((">", ">>", 0), 5.0, 0.0))


def line_config(args, n_gpus, n_slots, desc):
    """The `config` object of the JSON line: names the workload; identical for both arms given the same arguments
    (everything that describes HOW an arm ran it is in `details`)."""
    return dict(
        workload=desc % args.envs, envs_per_gpu=args.envs, total_envs=n_gpus * args.envs,
        parallelism="env-sharded x%d" % n_gpus,
        l2=l2_note((2 * (80 * n_slots + 32) + 4 * OBS_DIM) * args.envs / 1e6, args.envs))


def l2_note(mb, envs):
    if mb > 126:
        return "no flush: state + observations touched per step = %.0f MB per %d environments > 126 MB L2" % (mb, envs)
    return ("no flush: state + observations touched per step = %.0f MB per %d environments fit the 126 MB L2 -- an "
            "L2-warm figure; the headline configuration (65 536 environments, 244 MB) does not fit") % (mb, envs)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as orc
    orc.build()
    T = build_tables(args.workload)
    threads = host_threads()
    _, n_seeds, n_slots, desc = WORKLOADS[args.workload]
    # calibrate, then size the per-step sample so that warmup + steps finish in about 100 s
    rate0, _ = cpu_oracle_rate(T, 1024, 4, 1, threads, [i % n_seeds for i in range(1024)], n_slots)
    n = int(min(args.envs, max(256, rate0 * 100.0 / (args.steps + args.warmup))))
    n = max(100, n // 100 * 100)
    rate, dt = cpu_oracle_rate(T, n, args.steps, args.warmup, threads, [i % n_seeds for i in range(n)], n_slots)
    sample = "%d of %d envs per step x %d steps, %d host threads" % (n, args.envs, args.steps, threads)
    line = dict(
        impl="reference", metric="env-steps/s", value=rate, unit="env-steps/s", n_gpus=args.gpus, steps=args.steps,
        warmup=args.warmup, ms_per_step=dt / args.steps * 1e3, higher_is_better=True, scaling="weak",
        vs_baseline=None, dtype="f32", data="synthetic",
        config=line_config(args, args.gpus, n_slots, desc),
        details=dict(actions="uniform[-1,1]^2, RandomState(0)",
                     note="reference step needs Panda3D/Bullet (not installable offline): CPU oracle port timed instead"),
        cpu_baseline=dict(value=rate, unit="env-steps/s", cores=threads, kind="port", sample=sample),
        e2e=dict(value=rate, unit="env-steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
    )
    print(json.dumps(line))
    return 0


def issue_bound(kernel_ms, sm_mhz, policy, workload, envs=65536):
    """Secondary (honest) bound, SURVEY 8(d): the kernel is latency / issue bound, not HBM bound.  Warp instructions
    per launch come from the committed ncu capture of this workload (profiles/issue_slots.json); the rate they are
    issued at is live: kernel time from this run, SM clock from nvidia-smi during it."""
    path = os.path.join(ROOT, "profiles", "issue_slots.json")
    if not os.path.exists(path) or not sm_mhz or not kernel_ms:
        return None
    rec = json.load(open(path)).get(workload, {}).get(policy)
    if not rec:
        return None
    peak = SMS * ISSUE_PER_SM_CLK * sm_mhz * 1e6  # warp instructions / s
    insts = rec["warp_instructions_per_launch"] * envs / 65536.0  # the capture is of a 65 536-environment launch
    achieved = insts / (kernel_ms * 1e-3)
    return dict(bound="issue", achieved=achieved / 1e9, peak=peak / 1e9, unit="G warp-inst/s", frac=achieved / peak,
                warp_instructions_per_launch=insts,
                threads_per_instruction=rec.get("threads_per_instruction"), source=rec.get("source"))


def run_own(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs `python -m torch.distributed.run --nproc-per-node %d`" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's INFO log (version banner, rings, "comm ... nranks N") goes to stderr so that stdout stays the one
        # JSON line; it is not switched off
        if os.environ.get("NCCL_DEBUG"):
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        # the collective runs on a high-priority stream: its few CTAs are scheduled as soon as a step-kernel CTA
        # retires instead of queueing behind the whole step kernel
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)

    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    from pgdrive_b200 import VecPGDriveEnv, cabi
    from pgdrive_b200.sharding import GatherBuffers, PeerGather, balanced_sizes, sizes_with_rank0
    K, W = args.steps, args.warmup
    first_seed, n_seeds, n_slots, desc = WORKLOADS[args.workload]
    T = build_tables(args.workload)
    # ---- how the batch is cut.  Weak scaling: N ranks share N * --envs environments.  "Rank 0 holds the batch" gives
    # rank 0 extra work every step (expanding the other ranks' packed rows, reading the whole gathered batch), so by
    # default it simulates fewer environments than the others (sharding.balanced_sizes); --balance equal = N equal shards.
    gather_mode = "none"
    sizes = [args.envs] * world
    if world > 1:
        gather_mode = args.gather
        if gather_mode == "auto":
            # With every result funnelled into rank 0 the dense gathers are bound by that GPU's NVLink ingress from 4
            # GPUs on (copy engine 721 GB/s, the kernel's own remote row stores 653 GB/s: profiles/r03i_*, r03n_*), so
            # the rows cross the link packed (profiles/r03o_*, r03p_*: 4 GPUs 792 -> 924 M, 8 GPUs 750 -> 1 153 M
            # env-steps/s).  At 2 GPUs the link is not the bound and the packing kernels only cost SM time (679 vs 614 M).
            gather_mode = "sparse" if world >= 3 else "copy"
        if gather_mode in ("peer", "copy", "sparse"):
            ok = torch.tensor([1 if local_rank == 0 or torch.cuda.can_device_access_peer(local_rank, 0) else 0],
                              dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                gather_mode = "nccl (peer access unavailable)"
        if gather_mode in ("peer", "copy", "sparse") and args.balance != "equal":
            # ns per simulated env-step in the steady state of the random policy: 65 536 / 433 M (v0, 16 slots),
            # 65 536 / 343 M (1000 maps, 24 slots) -- profiles/r03m_bench_*.json
            cost = dict(step=args.step_ns or {"v0": 2.3, "1000envs": 2.9}[args.workload])
            if args.rank0_envs:
                sizes = sizes_with_rank0(world * args.envs, world, args.rank0_envs)
            else:
                sizes = balanced_sizes(world * args.envs, world, gather_mode, cost=cost)
    n, n_max, total_envs = sizes[rank], max(sizes), sum(sizes)
    balance_note = "equal shards" if len(set(sizes)) == 1 else (
        "rank 0 holds, expands and reads the whole batch every step, so it simulates fewer environments than the other "
        "ranks (sharding.balanced_sizes: both sides of the gather on the same clock); total = N x envs_per_gpu")
    peer_wanted = gather_mode in ("peer", "copy", "sparse")

    def make_bufs():  # whole-batch buffers of the NCCL all-gather (equal shards); one local buffer at N = 1
        return [GatherBuffers(torch, n, world, rank, dev, obs_dim=OBS_DIM) for _ in range(2 if world > 1 else 1)]

    bufs = None if peer_wanted else make_bufs()
    side = torch.cuda.Stream(device=dev, priority=-1) if world > 1 else None
    env = VecPGDriveEnv(
        dict(start_seed=first_seed, environment_num=n_seeds, num_envs=n, traffic_density=0.1, device=local_rank,
             num_slots=n_slots),
        tables_dict=T, obs_out=None if bufs is None else bufs[0].local(bufs[0].obs)
    )
    peer = None
    if world > 1 and gather_mode in ("peer", "copy", "sparse"):
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            peer = PeerGather(env, torch, dist, sizes, world, rank, mode=gather_mode)
        except Exception as exc:  # noqa: BLE001
            sys.stderr.write("rank %d: peer gather unavailable (%s)\n" % (rank, exc))
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            if len(set(sizes)) > 1:
                raise SystemExit("peer mapping failed with unequal shards: run with --balance equal (NCCL all-gather)")
            peer, gather_mode, bufs = None, "nccl (peer mapping unavailable)", make_bufs()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()  # early: nvidia-smi takes its time to deliver the first line (see ClockSampler)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)  # Philox counter-based stream, one per rank
    n_act = max(W + K, 256)
    actions = torch.rand((n_act, n, 2), generator=gen, device=dev, dtype=torch.float32) * 2 - 1
    fwd = actions.clone()  # a policy that drives: never brakes, small steering noise
    fwd[..., 1] = fwd[..., 1].abs()
    fwd[..., 0] *= 0.1

    # ---- gather plumbing.  The host enqueues every step's work in ~60 us (one kernel launch, three event operations,
    # one 4-byte all-reduce, one consumer reduction): pre-created events, pre-resolved pointers, ONE reduction kernel
    # for the consumer -- with 15+ torch calls per step the loop was bound by Python, not by the GPUs.
    use_ev = [torch.cuda.Event(), torch.cuda.Event()]
    ar_ev = [torch.cuda.Event(), torch.cuda.Event()]
    ready_ev = [torch.cuda.Event(), torch.cuda.Event()]
    armed = dict(use=[False, False], ar=[False, False])
    counter = [0]
    consumed = torch.zeros(1, dtype=torch.int64, device=dev)   # rank 0: checksum of what the consumer read (timed steps)
    verified = torch.zeros(1, dtype=torch.int64, device=dev)   # rank 0: the same over the verification steps
    local_sum = torch.zeros(1, dtype=torch.int64, device=dev)  # every rank: checksum of its OWN rows, same steps
    direct = peer is not None and (peer.mode == "peer" or rank == 0)  # rank 0's own rows are local memory either way
    step_ptrs, whole = [None, None], [None, None]
    if peer is not None:
        for i in range(2):
            step_ptrs[i] = peer.pointers(i) if direct else peer.local_pointers(i)
            if rank == 0:
                whole[i] = peer.words(i)  # the whole gathered buffer (obs | reward | done) as int32 words
    elif world > 1:
        for i in range(2):
            whole[i] = [bufs[i].obs.view(torch.int32), bufs[i].reward.view(torch.int32), bufs[i].done.view(torch.int32)]

    def words_sum(acc, tensors):
        """Integer checksum (sum of the buffer's 32-bit words, 64-bit accumulator; pgd_words_checksum reads at HBM speed):
        exact and order-independent, so the consumer's sum over the whole gathered batch must EQUAL the total of the
        ranks' sums over their own rows."""
        e = env.engine
        st = torch.cuda.current_stream(dev).cuda_stream
        for t in tensors:
            nbytes = t.numel() * t.element_size()
            if nbytes % 16 == 0 and t.data_ptr() % 16 == 0:
                cabi.check(e.lib, e.lib.pgd_words_checksum(e.h, t.data_ptr(), nbytes, acc.data_ptr(), st))
            else:  # odd-sized views (never the whole-batch buffers)
                acc += t.sum(dtype=torch.int64)

    def own_rows(i):
        if peer is not None:
            return [t.view(torch.int32) for t in peer.local_views(i, remote=direct)]
        b = bufs[i]
        return [b.local(b.obs).view(torch.int32), b.local(b.reward).view(torch.int32), b.local(b.done).view(torch.int32)]

    def step_and_gather(a, check=False):
        """One step of this rank + the gather of everybody's results to rank 0 (+ rank 0 reading the batch).

        Two gather buffers alternate.  Before the kernel of step t writes buffer i = t % 2 it waits for
          use_ev[i]     the last LOCAL reader of buffer i from step t - 2 (copy engine push / all-gather of this rank's
                        rows; on rank 0 the consumer that read the whole batch), and -- only when the kernel stores
                        straight into rank 0's memory (peer mode, ranks > 0) --
          ar_ev[1 - i]  this rank's completion barrier of step t - 1: rank 0 enqueues its read of step t - 2 before it
                        joins that barrier, so the barrier having completed means buffer i has been read.
        Nothing else is ordered: the gather of step t overlaps the kernel of step t + 1."""
        if world == 1:
            env.step(a)
            return
        i = counter[0] & 1
        counter[0] += 1
        cur = torch.cuda.current_stream(dev)
        if armed["use"][i]:
            cur.wait_event(use_ev[i])
        if peer is not None:
            if peer.mode == "peer" and rank != 0 and armed["ar"][1 - i]:
                cur.wait_event(ar_ev[1 - i])
            env.step_into(a, *step_ptrs[i])
            if check == "verify":  # untimed: this rank's own rows (read back over NVLink in peer mode)
                words_sum(local_sum, own_rows(i))
            ready_ev[i].record(cur)
            with torch.cuda.stream(side):
                side.wait_event(ready_ev[i])
                if not direct:
                    peer.push(i)  # copy engine (or the packing kernel): local rows -> rank 0's buffer over NVLink
                    use_ev[i].record(side)
                    armed["use"][i] = True
                peer.completion_barrier()
                ar_ev[i].record(side)
                armed["ar"][i] = True
                if rank == 0:
                    peer.expand(i)  # "sparse": packed rows of the other ranks -> the whole-batch buffer
                    if check:  # the consumer reads the whole gathered batch, every step
                        words_sum(verified if check == "verify" else consumed, whole[i])
                    use_ev[i].record(side)
                    armed["use"][i] = True
            return
        b = bufs[i]
        env.step(a, out=(b.local(b.obs), b.local(b.reward), b.local(b.done)))
        if check == "verify":
            words_sum(local_sum, own_rows(i))
        ready_ev[i].record(cur)
        with torch.cuda.stream(side):
            side.wait_event(ready_ev[i])
            b.all_gather(dist)
            if rank == 0 and check:
                words_sum(verified if check == "verify" else consumed, whole[i])
            use_ev[i].record(side)
            armed["use"][i] = True

    def drain():
        if side is not None:
            torch.cuda.current_stream(dev).wait_stream(side)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(acts, k, fn):
        """K steps bracketed by barrier + synchronize; returns (total ms, mean per-step ms between events)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        barrier()
        ev0.record()
        for t in range(k):
            k_ev[t][0].record()
            fn(acts[t])
            k_ev[t][1].record()
        drain()
        ev1.record()
        barrier()
        return ev0.elapsed_time(ev1), float(np.mean([a.elapsed_time(b) for a, b in k_ev]))

    # ---- headline: BASELINE.md's random policy, gather included -------------------------------------------------------
    env.reset()
    torch.cuda.synchronize()
    if rank == 0:
        clocks.mark()  # from here to the end of the timed region the GPU runs the step kernel back to back
    for t in range(PREROLL):  # untimed pre-roll: not tied to --steps / --warmup
        env.step(actions[t % n_act])
    for t in range(W):
        step_and_gather(actions[t])
    drain()
    barrier()
    launches0 = env.launch_count
    ms, kernel_ms = timed(actions[W:W + K], K, lambda a: step_and_gather(a, check=(world > 1)))
    launches = env.launch_count - launches0
    clk = clocks.stop() if rank == 0 else None
    if world > 1:  # untimed: 4 more steps in which every rank also sums its own rows; rank 0's sums must equal their total
        for t in range(4):
            step_and_gather(actions[t], check="verify")
        drain()
        barrier()
    done_rate = float(env.done.float().mean().item()) if world == 1 else None

    # ---- N > 1: the same steps without the gather (does the kernel itself scale?) -----------------------------------
    sim_ms = None
    if world > 1:
        sim_ms, _ = timed(actions[W:W + K], K, lambda a: env.step(a))

    if args.headline_only:  # tuning aid (gather modes / shard sizes at N > 1): not a bench line
        times = torch.tensor([ms, kernel_ms, sim_ms or 0.0], dtype=torch.float64, device=dev)
        sums = [torch.zeros_like(local_sum) for _ in range(world)]
        if world > 1:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
            dist.all_gather(sums, local_sum)
        if rank == 0:
            ms, kernel_ms, sim_ms = [float(x) for x in times.tolist()]
            print(json.dumps(dict(headline_only=True, n_gpus=world, gather=gather_mode, shard_sizes=sizes,
                                  value=total_envs * K / (ms * 1e-3), ms_per_step=ms / K, kernel_ms=kernel_ms,
                                  sim_only_ms_per_step=sim_ms / K if sim_ms else None,
                                  gather_ok=bool(int(verified.item()) == int(torch.stack(sums).sum().item())))))
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- the same kernel under a policy that drives (own pre-roll, own roofline) -----------------------------------------
    env.reset()
    for t in range(PREROLL_DRIVING):
        env.step(fwd[t % n_act])
    fwd_k = min(K, fwd.shape[0])
    fwd_ms, fwd_kernel_ms = timed(fwd[:fwd_k], fwd_k, lambda a: env.step(a))
    fwd_done_rate = float(env.done.float().mean().item())

    # ---- end to end through the public host-buffer API ------------------------------------------------------------------
    env.reset()
    for t in range(PREROLL):
        env.step(actions[t % n_act])
    e2e_steps = min(K, 64)
    h_actions = actions[W:W + e2e_steps].cpu().numpy()
    for t in range(min(W, 4)):
        env.step(h_actions[t % e2e_steps], copy=False)
    barrier()
    t0 = time.perf_counter()
    for t in range(e2e_steps):
        o, r, d, i = env.step(h_actions[t], copy=False)  # pinned staging arrays: valid until the next step
    torch.cuda.synchronize()
    per_rank_e2e_s = time.perf_counter() - t0
    checksum = float(o[:, :8].sum())
    e2e_s = per_rank_e2e_s
    e2e_api = ("VecPGDriveEnv.step(numpy) -> pgd_step_host: observation rows cross PCIe packed (head + hit mask + beams "
               "that are not 1.0) and host threads expand them into the caller's array, bit-identical to the dense copy")
    e2e_h2d, e2e_d2h = env.host_transfer_bytes()  # counted by the library: the copies it enqueued in the last step
    e2e_dense_d2h = n * (4 * OBS_DIM + 4 + 1 + INFO_BYTES)
    if world > 1 and peer is not None:
        # north_star's path: the actions of the WHOLE batch start in rank 0's host memory, the gathered observation /
        # reward / done batch ends there.  H2D on rank 0, broadcast over NVLink, step + gather, one D2H on rank 0.
        rows, row0 = total_envs, sum(sizes[:rank])
        all_act_dev = torch.empty((rows, 2), dtype=torch.float32, device=dev)
        if rank == 0:
            h_all = torch.empty((rows, 2), dtype=torch.float32, pin_memory=True)
            h_all.uniform_(-1, 1)
            h_obs = torch.empty((rows, OBS_DIM), dtype=torch.float32, pin_memory=True)
            h_rew = torch.empty(rows, dtype=torch.float32, pin_memory=True)
            h_done = torch.empty(rows, dtype=torch.uint8, pin_memory=True)

        packed_host = [True]

        def host_step():
            if rank == 0:
                all_act_dev.copy_(h_all, non_blocking=True)
            dist.broadcast(all_act_dev, src=0)
            step_and_gather(all_act_dev[row0:row0 + n])
            drain()
            if rank == 0:
                go, gr, gd = peer.tensors(counter[0] - 1)
                if packed_host[0]:
                    try:  # the gathered batch crosses PCIe packed and is expanded by host threads (pgd_rows_to_host)
                        env.rows_to_host(go, gr, gd, h_obs.numpy(), h_rew.numpy(), h_done.numpy())
                    except Exception as exc:  # noqa: BLE001 -- fall back to dense copies, say so
                        sys.stderr.write("pgd_rows_to_host failed (%s): dense copies\n" % exc)
                        packed_host[0] = False
                if not packed_host[0]:
                    h_obs.copy_(go, non_blocking=True)
                    h_rew.copy_(gr, non_blocking=True)
                    h_done.copy_(gd, non_blocking=True)
            torch.cuda.synchronize()

        e2e_steps_n = min(e2e_steps, 16)
        host_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps_n):
            host_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) * e2e_steps / e2e_steps_n  # normalised to e2e_steps below
        e2e_api = ("rank 0 host actions [N*n, 2] -> H2D -> NCCL broadcast -> step + gather to rank 0 -> D2H of the "
                   "gathered obs / reward / done to rank 0's host" + (
                       " (rows packed over PCIe, expanded by host threads: pgd_rows_to_host)" if packed_host[0] else ""))
        e2e_h2d, e2e_d2h = rows * 8, rows * (4 * OBS_DIM + 4 + 1)
        if packed_host[0] and rank == 0:
            e2e_d2h = env.host_transfer_bytes()[1]

    times = torch.tensor([ms, e2e_s * 1e3, kernel_ms, fwd_ms, fwd_kernel_ms, per_rank_e2e_s * 1e3, sim_ms or 0.0],
                         dtype=torch.float64, device=dev)
    sums = None
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        sums = [torch.zeros_like(local_sum) for _ in range(world)]
        dist.all_gather(sums, local_sum)
    ms, e2e_ms, kernel_ms, fwd_ms, fwd_kernel_ms, per_rank_e2e_ms, sim_ms = [float(x) for x in times.tolist()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = total_envs * K / (ms * 1e-3)
    b_step = algorithmic_bytes_per_env_step(n_slots)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("bytes_per_launch")

    def roofline(k_ms, policy):
        achieved = b_step * n_max / (k_ms * 1e-3) / 1e9  # k_ms is the max over ranks: the largest shard's kernel
        return dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                    traffic=traffic if (policy == "random" and args.workload == "v0" and n_max == 65536) else None,
                    kernel="pgd_step_kernel<%d, 4>" % n_slots, kernel_ms=k_ms, bytes_per_env_step=b_step,
                    peak_source=peak_src,
                    issue=issue_bound(k_ms, (clk or {}).get("sm_mhz"), policy, args.workload, envs=n_max))

    gather_check = None
    if world > 1:
        want = int(torch.stack(sums).sum().item())
        got = int(verified.item())
        gather_check = dict(
            consumer="rank 0 reads the whole gathered batch after every step's gather: integer checksum = sum of its "
                     "32-bit words (consumed_timed).  Over 4 more untimed steps every rank also sums its own rows; the "
                     "consumer's checksum must EQUAL the total of the ranks' (verified == sum_of_rank_local)",
            consumed_timed=int(consumed.item()), verified=got, sum_of_rank_local=want, ok=bool(got == want))

    # ---- the reset path: maps + episode templates of the workload's seeds generated ON the device -------------------
    reset_path = None
    if world == 1:
        try:
            from pgdrive_b200 import devgen
            from pgdrive_b200.env import _seed_tables, default_config, parse_map_config
            mc = parse_map_config(default_config())
            gc = devgen.make_gen_config(mc, 0.1)
            seeds = list(range(first_seed, first_seed + n_seeds))
            devgen.generate(env.engine, seeds[:4], gc)  # warm-up: module load, local-memory allocation
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            devgen.generate(env.engine, seeds, gc)
            torch.cuda.synchronize()
            dev_s = time.perf_counter() - t0
            k = min(len(seeds), 20)
            t0 = time.perf_counter()
            for sd in seeds[:k]:
                _seed_tables((sd, mc, 0.1, ((">", ">>", 0), 5.0, 0.0)))
            host_rate = k / (time.perf_counter() - t0)
            reset_path = dict(seeds=len(seeds), device_maps_per_s=len(seeds) / dev_s, device_ms=dev_s * 1e3,
                              host_python_maps_per_s=host_rate, host_sample="%d seeds, 1 process" % k)
            env.engine.load(T)
            env.reset()
        except Exception as e:  # the headline must not depend on this leg
            reset_path = dict(error=str(e)[:200])

    cpu = None
    if world == 1 and not args.no_cpu:
        threads = host_threads()
        cn, cs = 32768, 128  # ~4.2M env-steps: 10-30 s of CPU work on a 16-thread host
        rate, dt = cpu_oracle_rate(T, cn, cs, 2, threads, [i % n_seeds for i in range(cn)], n_slots)
        cpu = dict(value=rate, unit="env-steps/s", cores=threads, kind="port",
                   sample="%d of %d envs x %d steps (%.1f s), CPU oracle (bucket-grid queries) on %d host threads" % (
                       cn, args.envs, cs, dt, threads))

    collective = {
        "none": "none",
        "peer": "gather to rank 0 fused into the step kernel: the 32 rows of a CTA leave with one bulk (TMA) store "
                "straight into rank 0's HBM through CUDA-IPC peer mappings over NVLink; a 4-byte all-reduce per step on "
                "a side stream is the completion barrier; a buffer is rewritten only after rank 0 has read it",
        "copy": "kernel writes this rank's rows locally, the copy engine pushes them into rank 0's buffer (CUDA-IPC peer "
                "mapping) on a side stream while the next step's kernel runs; 4-byte all-reduce as completion barrier",
        "sparse": "kernel writes this rank's rows locally; on a side stream, while the next step's kernel runs, "
                  "pgd_pack_rows stores them into rank 0's HBM (CUDA-IPC peer mapping) in the wire format head + 240-bit "
                  "hit mask + beams that are not 1.0, and after the 4-byte all-reduce completion barrier rank 0's "
                  "pgd_expand_rows restores them bit for bit into the whole-batch buffer, which rank 0 then reads",
    }.get(gather_mode, "in-place NCCL all-gather of obs/reward/done every step, double-buffered on a high-priority "
                       "side stream so that it overlaps the next step's kernel [%s]" % gather_mode)
    line = dict(
        metric="env-steps/s", value=value, unit="env-steps/s", n_gpus=world, steps=K, warmup=W,
        ms_per_step=ms / K, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype="f32", data="synthetic",
        config=line_config(args, world, n_slots, desc),
        details=dict(
            shard_sizes=sizes if world > 1 else None, balance=None if world == 1 else balance_note,
            actions="uniform[-1,1]^2, Philox, pre-generated in HBM", preroll_steps=PREROLL,
            preroll_steps_driving=PREROLL_DRIVING,
            arithmetic="float32 throughout; the reference's Python side computes in float64, Bullet in float32",
            collective=collective, done_rate_last_step=done_rate,
        ),
        # N > 1: the step kernel's own time is that of the simulation-only leg (largest shard); the per-step time of the
        # gathered loop includes waiting for buffers
        roofline=roofline(kernel_ms if world == 1 else sim_ms / K, "random"),
        driving=dict(value=n_max * fwd_k / (fwd_ms * 1e-3), unit="env-steps/s per GPU", steps=fwd_k,
                     policy="throttle |u|, steering 0.1 u: traffic awake, lidar hits, frequent resets",
                     done_rate_last_step=fwd_done_rate, roofline=roofline(fwd_kernel_ms, "driving")),
        cpu_baseline=cpu,
        reset_path=reset_path,
        e2e=dict(value=total_envs * e2e_steps / (e2e_ms * 1e-3), unit="env-steps/s",
                 h2d_bytes_per_step=e2e_h2d, d2h_bytes_per_step=e2e_d2h, steps=e2e_steps, api=e2e_api,
                 checksum=checksum, dense_d2h_bytes_per_step=e2e_dense_d2h if world == 1 else None,
                 per_rank=dict(value=total_envs * e2e_steps / (per_rank_e2e_ms * 1e-3),
                               api="every rank: VecPGDriveEnv.step(numpy) -> pgd_step_host on its own shard")),
        gpu_launches=int(launches), clocks=clk,
    )
    if world > 1:
        line["sim_only"] = dict(value=total_envs * K / (sim_ms * 1e-3), unit="env-steps/s", ms_per_step=sim_ms / K,
                                note="the same K steps without the gather (with balanced shards rank 0 idles most of "
                                     "this leg: the figure is that of the N - 1 larger shards)")
        line["gather_check"] = gather_check
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--envs", type=int, default=65536, help="environments per GPU (4096 = BASELINE.json configs[1])")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--gather", default="auto", choices=["auto", "peer", "copy", "sparse", "nccl"],
                    help="N > 1: how rank 0 gets the whole batch (copy: local rows pushed into rank 0's HBM by the copy "
                         "engine on a side stream while the next step's kernel runs; sparse: the same, but the rows cross "
                         "NVLink packed -- head + hit mask + beams that are not 1.0 -- and rank 0 expands them; peer: rows "
                         "stored by the step kernel straight into rank 0's HBM; nccl: in-place all-gather, also the "
                         "fall-back when peer mapping is not permitted; auto = copy at 2 GPUs, sparse from 3 on)")
    ap.add_argument("--headline-only", action="store_true", help="tuning aid: time the headline (and the simulation-only "
                    "leg) and print a short record instead of the bench line")
    ap.add_argument("--balance", default="auto", choices=["auto", "equal"],
                    help="N > 1: auto = rank 0, which expands and reads the whole gathered batch every step, simulates fewer "
                         "environments than the others (sharding.balanced_sizes); equal = N shards of --envs")
    ap.add_argument("--rank0-envs", type=int, default=0, help="N > 1: fix rank 0's shard (multiple of 32); the rest is "
                    "shared evenly")
    ap.add_argument("--step-ns", type=float, default=0.0, help="override the cost model's ns per simulated env-step")
    ap.add_argument("--workload", default="v0", choices=sorted(WORKLOADS), help="v0 = BASELINE.json configs[2] (the metric's "
                    "configuration); 1000envs = configs[3]")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_own(args)


if __name__ == "__main__":
    sys.exit(main())
