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
ok = True
for mode in ("peer", "copy", "sparse"):  # kernel's row stores / copy engine / packed rows (sharding.PeerGather)
    env = VecPGDriveEnv(dict(start_seed=1000, environment_num=20, num_envs=n, device=lr))
    pg = PeerGather(env, torch, dist, n, world, rank, mode=mode)
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    acts = torch.rand((60, n, 2), generator=g, device="cuda") * 2 - 1
    acts[..., 1] = acts[..., 1].abs()
    for t in range(60):
        direct = mode == "peer" or rank == 0
        env.step_into(acts[t], *(pg.pointers(t) if direct else pg.local_pointers(t)))
        if not direct:
            pg.push(t)
        pg.completion_barrier()
        if rank == 0:
            pg.expand(t)
        torch.cuda.synchronize()
        if rank == 0:
            obs, rew, done = pg.tensors(t)
            for r in range(1, world):
                ok &= torch.equal(obs[:n], obs[r * n:(r + 1) * n]) and torch.equal(rew[:n], rew[r * n:(r + 1) * n]) \
                    and torch.equal(done[:n], done[r * n:(r + 1) * n])
            ok &= bool(obs[:n].abs().sum() > 0)
        dist.barrier()
    if rank == 0:
        print("%s gather ok" % mode if ok else "%s GATHER MISMATCH" % mode, "world", world, flush=True)
    pg.close(); env.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
