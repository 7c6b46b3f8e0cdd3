"""2+ GPU check of the fused peer gather (run under torchrun on a multi-GPU box):
every rank plays the SAME seeds with the SAME actions, so the rows each rank wrote into rank 0's buffer over NVLink
must be bit-identical to rank 0's own rows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from pgdrive_b200 import VecPGDriveEnv
from pgdrive_b200.sharding import PeerGather
n = 4096
env = VecPGDriveEnv(dict(start_seed=1000, environment_num=20, num_envs=n, device=lr))
pg = PeerGather(env, torch, dist, n, world, rank)
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(5)
acts = torch.rand((60, n, 2), generator=g, device="cuda") * 2 - 1
acts[..., 1] = acts[..., 1].abs()
ok = True
for t in range(60):
    env.step_into(acts[t], *pg.pointers(t))
    pg.completion_barrier()
    torch.cuda.synchronize()
    if rank == 0:
        obs, rew, done = pg.tensors(t)
        for r in range(1, world):
            ok &= torch.equal(obs[:n], obs[r * n:(r + 1) * n]) and torch.equal(rew[:n], rew[r * n:(r + 1) * n]) \
                and torch.equal(done[:n], done[r * n:(r + 1) * n])
        ok &= bool(obs[:n].abs().sum() > 0)
    dist.barrier()
if rank == 0:
    print("peer gather ok" if ok else "PEER GATHER MISMATCH", "world", world)
pg.close(); env.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
