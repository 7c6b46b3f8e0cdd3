"""2+ GPU check of the gather to rank 0 (run under torchrun on a multi-GPU box): environment i of EVERY rank plays the
same seed with the same actions, and rank 0 additionally steps a reference environment batch as large as the largest
shard.  Every row a rank delivered into rank 0's whole-batch buffer -- by the step kernel's remote row stores, by the copy
engine, or packed / expanded -- must be bit-identical to the reference row of the same local index.  Shards are equal
or unequal (rank 0 with a smaller one: sharding.balanced_sizes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from pgdrive_b200 import VecPGDriveEnv
from pgdrive_b200.sharding import PeerGather, sizes_with_rank0
all_ok = True
for mode, sizes in (("peer", [4096] * world), ("copy", [4096] * world), ("sparse", [4096] * world),
                    ("sparse", sizes_with_rank0(4096 * world, world, 1024)),
                    ("copy", sizes_with_rank0(4096 * world, world, 2048))):
    ok = True
    n, first = sizes[rank], [sum(sizes[:r]) for r in range(world)]
    env = VecPGDriveEnv(dict(start_seed=1000, environment_num=20, num_envs=n, device=lr))
    ref = VecPGDriveEnv(dict(start_seed=1000, environment_num=20, num_envs=max(sizes), device=lr)) if rank == 0 else None
    pg = PeerGather(env, torch, dist, sizes, world, rank, mode=mode)
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    acts = torch.rand((60, max(sizes), 2), generator=g, device="cuda") * 2 - 1
    acts[..., 1] = acts[..., 1].abs()
    if ref is not None:
        ref.reset()
    for t in range(60):
        direct = mode == "peer" or rank == 0
        env.step_into(acts[t, :n].contiguous(), *(pg.pointers(t) if direct else pg.local_pointers(t)))
        if not direct:
            pg.push(t)
        pg.completion_barrier()
        if rank == 0:
            pg.expand(t)
            want_obs, want_rew, want_done = ref.step(acts[t])[:3]
        torch.cuda.synchronize()
        if rank == 0:
            obs, rew, done = pg.tensors(t)
            for r in range(world):
                k, f = sizes[r], first[r]
                ok &= torch.equal(obs[f:f + k], want_obs[:k]) and torch.equal(rew[f:f + k], want_rew[:k]) \
                    and torch.equal(done[f:f + k], want_done[:k])
            ok &= bool(obs.abs().sum(dim=1).min() > 0)
        dist.barrier()
    if rank == 0:
        print("%s gather ok" % mode if ok else "%s GATHER MISMATCH" % mode, "world", world, "shards", sizes, flush=True)
    all_ok &= ok
    pg.close(); env.close()
    if ref is not None:
        ref.close()
if rank == 0 and all_ok:
    print("peer gather ok", flush=True)
dist.destroy_process_group()
sys.exit(0 if all_ok else 1)
