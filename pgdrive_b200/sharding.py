"""Environment sharding across ranks (one process per GPU) and the per-step gather of results.

Environments never interact (the reference hosts exactly one engine per process,
/root/reference/pgdrive/engine/engine_utils.py:8-15), so the batch is cut into contiguous index ranges and
the only collective is the all-gather that hands the whole observation / reward / done batch to rank 0
(BASELINE.json north_star).  The functions here are pure rank arithmetic plus thin wrappers over
``torch.distributed`` so that they run under gloo on CPU (tests) and NCCL on GPUs (bench.py).
"""


def shard_range(total_envs, world_size, rank):
    """Contiguous range [lo, hi) of global environment indices owned by ``rank`` (sizes differ by at most one)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(total_envs, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def seed_of_env(global_env, start_seed, environment_num):
    """Seed played by a global environment index (env i -> start_seed + i mod environment_num)."""
    return start_seed + global_env % environment_num


class GatherBuffers:
    """Whole-batch result buffers laid out [world * n_local, ...]; a rank's own rows are a view the step kernel
    writes into directly, so ``all_gather`` below is in place (no packing copy)."""
    def __init__(self, torch, n_local, world_size, rank, device, obs_dim=274):
        self.torch, self.n, self.world, self.rank = torch, n_local, world_size, rank
        self.obs = torch.empty((world_size * n_local, obs_dim), dtype=torch.float32, device=device)
        self.reward = torch.empty(world_size * n_local, dtype=torch.float32, device=device)
        self.done = torch.empty(world_size * n_local, dtype=torch.uint8, device=device)

    def local(self, t):
        return t[self.rank * self.n:(self.rank + 1) * self.n]

    def all_gather(self, dist, group=None):
        if self.world == 1:
            return
        for t in (self.obs, self.reward, self.done):
            dist.all_gather_into_tensor(t, self.local(t), group=group)
