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


# Cost of one row (= one environment's results of one step) on a B200, nanoseconds, steady state of the random policy:
# simulate it (profiles/r03m_*); pack it into the wire format and store it into rank 0's HBM (r04d: 2 GPUs); expand it on
# rank 0 (delta expansion); have rank 0's consumer read it (profiles/r04g_rows_bench_delta.json).  "fixed": what rank 0's step costs beyond
# its kernels (completion barrier, stream hand-overs), less the same on a sending rank (r04b / r04d shard sweeps).
ROW_COST_NS = dict(step=2.3, pack=0.39, expand=0.23, consume=0.19, fixed=12000.0)


def balanced_sizes(total_envs, world_size, mode="sparse", granule=32, min_rank0=4096, cost=None):
    """Shard sizes for "rank 0 holds the batch": rank 0 simulates FEWER environments than the others, by as much as
    its extra work per step -- expanding the other ranks' packed rows and reading the whole gathered batch -- takes.

    With equal shards every rank finishes its step at the same time and then waits for rank 0 to expand and read
    ``total_envs`` rows (8 GPUs: 0.15 ms of simulation, 0.27 ms of gather work on rank 0).  Solving
    ``step*n0 + expand*(T - n0) + consume*T + fixed  =  (step + pack) * (T - n0) / (world - 1)`` for n0 puts both sides
    on the same clock.  ``mode``: "sparse" (pack / expand kernels) or "copy" / "peer" (no SM work for the transfer itself).
    Sizes are multiples of ``granule`` (32 = the environments of one CTA: every bulk row store stays whole and 16-byte
    aligned); ranks 1.. differ by at most one granule; the sizes sum to ``total_envs``."""
    if world_size < 1 or total_envs % granule or total_envs < world_size * granule:
        raise ValueError("total_envs must be a multiple of %d with at least one granule per rank" % granule)
    if world_size == 1:
        return [total_envs]
    c = dict(ROW_COST_NS, **(cost or {}))
    pack, expand = (c["pack"], c["expand"]) if mode == "sparse" else (0.0, 0.0)
    T, others = float(total_envs), world_size - 1
    per_other = (c["step"] + pack) / others
    n0 = ((per_other - expand - c["consume"]) * T - c["fixed"]) / (c["step"] - expand + per_other)
    equal = total_envs // world_size
    n0 = int(min(max(n0, min(min_rank0, equal)), equal) + 0.5 * granule) // granule * granule  # nearest granule
    if n0 >= equal and total_envs % (world_size * granule) == 0:
        return [equal] * world_size
    return sizes_with_rank0(total_envs, world_size, max(n0, granule), granule)


def sizes_with_rank0(total_envs, world_size, rank0_envs, granule=32):
    """Rank 0 gets ``rank0_envs``; the other ranks share the rest in granules, sizes differing by at most one granule."""
    others = world_size - 1
    if world_size < 2 or rank0_envs % granule or total_envs % granule or rank0_envs <= 0 or \
            total_envs - rank0_envs < others * granule:
        raise ValueError("rank 0's shard and the total must be multiples of %d and leave a granule per rank" % granule)
    base, extra = divmod((total_envs - rank0_envs) // granule, others)
    return [rank0_envs] + [(base + (1 if r < extra else 0)) * granule for r in range(others)]


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


class _DevArray:
    """Lets torch wrap a raw device pointer (``torch.as_tensor``) through ``__cuda_array_interface__``."""
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=2)


class PeerGather:
    """Gather to rank 0 over CUDA-IPC peer mappings: rank 0 owns ``depth`` whole-batch buffers, every other rank maps
    them.  Two ways to fill them:

    ``mode="peer"``  fused into the step kernel: a rank hands the kernel pointers to ITS rows of rank 0's buffer, so
                     observations / rewards / dones are written over NVLink as they are produced (one bulk store per
                     32 environments); no collective carries payload.
    ``mode="copy"``  the kernel writes local rows; ``push(i)`` copies them into rank 0's buffer with the copy engine
                     (enqueue it on a side stream: it overlaps the next step's kernel and the SMs are free meanwhile).
    ``mode="sparse"`` like "copy", but the observation rows cross NVLink in the wire format of ``pgd_pack_rows``
                     (head + 240-bit hit mask + the beams that are not 1.0: about a quarter of the bytes): ``push(i)``
                     packs the local rows straight into a staging area in rank 0's HBM, ``expand(i)`` on rank 0
                     restores them, bit for bit, into the whole-batch buffer.  With every result funnelled into one GPU
                     the gather is bound by that GPU's NVLink ingress; this is what moves the bound.

    In both, ``completion_barrier`` (a 4-byte all-reduce) tells rank 0 that every rank's rows of that buffer have
    landed, and -- because rank 0 enqueues its reads of buffer i before it joins the barrier of step i + 1 -- a rank
    that waits for barrier i + 1 before writing buffer i again (step i + 2) never overwrites unread rows."""
    def __init__(self, env, torch, dist, n_local, world_size, rank, obs_dim=274, depth=2, mode="peer", delta=True):
        """``n_local``: rows per rank -- one int (equal shards) or the list of every rank's size (``balanced_sizes``;
        multiples of 32 so that every rank's rows start 16-byte aligned).  ``delta`` (mode "sparse"): rank 0 expands
        into a buffer that still holds the rows of ``depth`` steps ago and stores only what changed
        (pgd_expand_rows_delta); needs every step's ``expand`` and nobody else writing the other ranks' rows."""
        import ctypes as C
        from . import cabi
        if mode not in ("peer", "copy", "sparse"):
            raise ValueError("mode must be 'peer', 'copy' or 'sparse'")
        self.env, self.torch, self.dist, self.mode = env, torch, dist, mode
        self.sizes = [int(n_local)] * world_size if isinstance(n_local, int) else [int(v) for v in n_local]
        if len(self.sizes) != world_size or min(self.sizes) <= 0:
            raise ValueError("one positive shard size per rank")
        if len(set(self.sizes)) > 1 and any(v % 32 for v in self.sizes):
            raise ValueError("unequal shard sizes must be multiples of 32")
        self.first = [sum(self.sizes[:r]) for r in range(world_size)]  # first global row of every rank
        self.n, self.row0 = self.sizes[rank], self.first[rank]
        self.world, self.rank, self.depth, self.obs_dim = world_size, rank, depth, obs_dim
        rows = sum(self.sizes)
        self.rows = rows
        self.obs_bytes, self.rew_bytes = rows * obs_dim * 4, rows * 4
        self.done_bytes = (rows + 255) // 256 * 256
        self.stride = self.obs_bytes + self.rew_bytes + self.done_bytes
        e = env.engine
        self._lib, self._h = e.lib, e.h
        # "sparse": per buffer a staging area of packed rows (fixed stride pgd_packed_row_words) behind the dense batch
        self.packed_words = int(e.lib.pgd_packed_row_words(obs_dim)) if mode == "sparse" else 0
        self.packed_bytes = (rows * self.packed_words * 4 + 255) // 256 * 256
        self.stride += self.packed_bytes
        base = C.c_void_p()
        handle = C.create_string_buffer(64)
        if rank == 0:
            cabi.check(e.lib, e.lib.pgd_peer_alloc(e.h, self.stride * depth, C.byref(base), handle))
        blob = [handle.raw if rank == 0 else None]
        if world_size > 1:
            dist.broadcast_object_list(blob, src=0)
        if rank != 0:
            cabi.check(e.lib, e.lib.pgd_peer_open(e.h, blob[0], C.byref(base)))
        self.base = base.value
        self._flag = torch.zeros(1, dtype=torch.int32, device=e.device)
        # delta expansion: per buffer, the hit masks (8 words per row) of the rows of ranks 1.. that the buffer holds
        self._mask_state = None
        if mode == "sparse" and delta and rank == 0 and world_size > 1:
            self._mask_state = [torch.zeros((rows - self.sizes[0], 8), dtype=torch.int32, device=e.device)
                                for _ in range(depth)]
            self._state_valid = [False] * depth
        self._local = None
        if mode in ("copy", "sparse") and rank != 0:
            dev = e.device
            self._local = [(torch.empty((self.n, obs_dim), dtype=torch.float32, device=dev),
                            torch.empty(self.n, dtype=torch.float32, device=dev),
                            torch.empty(self.n, dtype=torch.uint8, device=dev)) for _ in range(depth)]

    def _offsets(self, i):
        b = self.base + (i % self.depth) * self.stride
        return (b + self.row0 * self.obs_dim * 4, b + self.obs_bytes + self.row0 * 4,
                b + self.obs_bytes + self.rew_bytes + self.row0)

    def pointers(self, i):
        """(obs, reward, done) device pointers of THIS rank's rows in rank 0's buffer ``i``."""
        return self._offsets(i)

    def local_pointers(self, i):
        """mode "copy", rank > 0: pointers of the local staging rows the kernel writes before ``push``."""
        return tuple(t.data_ptr() for t in self._local[i % self.depth])

    def _wrap(self, ptr, shape, typestr):
        return self.torch.as_tensor(_DevArray(ptr, shape, typestr), device=self.env.engine.device)

    def remote_views(self, i):
        """THIS rank's rows of rank 0's buffer ``i`` as tensors (peer-mapped memory on ranks > 0)."""
        o, r, d = self._offsets(i)
        return (self._wrap(o, (self.n, self.obs_dim), "<f4"), self._wrap(r, (self.n, ), "<f4"),
                self._wrap(d, (self.n, ), "|u1"))

    def local_views(self, i, remote=False):
        """The rows this rank's kernel wrote for buffer ``i`` (local staging, or its rows of rank 0's buffer)."""
        return self.remote_views(i) if remote or self._local is None else self._local[i % self.depth]

    def _packed_ptr(self, i, rank):
        b = self.base + (i % self.depth) * self.stride + (self.stride - self.packed_bytes)
        return b + self.first[rank] * self.packed_words * 4

    def push(self, i):
        """modes "copy" / "sparse", rank > 0: enqueue (current stream) the transfer of the local rows into rank 0's
        buffer ``i``.  "copy": device-to-device copies -- contiguous, so the driver hands them to the copy engine.
        "sparse": the observation rows are packed by a kernel that stores into rank 0's staging area."""
        views, local = self.remote_views(i), self._local[i % self.depth]
        if self.mode == "sparse":
            from . import cabi
            st = self.torch.cuda.current_stream(self.env.engine.device).cuda_stream
            cabi.check(self._lib, self._lib.pgd_pack_rows(local[0].data_ptr(), self._packed_ptr(i, self.rank), self.n,
                                                          self.obs_dim, st))
            views, local = views[1:], local[1:]
        for dst, src in zip(views, local):
            dst.copy_(src, non_blocking=True)

    def expand(self, i):
        """mode "sparse", rank 0, after the completion barrier of buffer ``i``: restore the rows of ranks 1.. from the
        staging area into the whole-batch buffer (current stream).  With ``delta`` only the head, the beams that were
        or are hits and the new hit mask of a row are stored: the buffer still holds the rows of step i - depth, and a
        beam that was 1.0 then and is 1.0 now needs no store."""
        if self.mode != "sparse" or self.world == 1:
            return
        assert self.rank == 0
        from . import cabi
        st = self.torch.cuda.current_stream(self.env.engine.device).cuda_stream
        b = self.base + (i % self.depth) * self.stride
        src, dst, n = self._packed_ptr(i, 1), b + self.sizes[0] * self.obs_dim * 4, self.rows - self.sizes[0]
        if self._mask_state is None:
            cabi.check(self._lib, self._lib.pgd_expand_rows(src, dst, n, self.obs_dim, st))
            return
        k = i % self.depth
        cabi.check(self._lib, self._lib.pgd_expand_rows_delta(src, dst, self._mask_state[k].data_ptr(), n, self.obs_dim,
                                                              0 if self._state_valid[k] else 1, st))
        self._state_valid[k] = True

    def completion_barrier(self):
        """Enqueue (on the current stream) a barrier after which rank 0 may read the buffer written last."""
        if self.world > 1:
            self.dist.all_reduce(self._flag)

    def tensors(self, i):
        """Rank 0 only: the whole-batch (obs, reward, done) tensors of buffer ``i``."""
        assert self.rank == 0
        rows = self.rows
        b = self.base + (i % self.depth) * self.stride
        return (self._wrap(b, (rows, self.obs_dim), "<f4"), self._wrap(b + self.obs_bytes, (rows, ), "<f4"),
                self._wrap(b + self.obs_bytes + self.rew_bytes, (rows, ), "|u1"))

    def words(self, i):
        """Rank 0 only: buffer ``i`` (observations | rewards | dones + padding) as ONE int32 tensor -- a consumer that
        wants to touch every byte of the gathered batch can do it with a single reduction."""
        assert self.rank == 0
        b = self.base + (i % self.depth) * self.stride
        return [self._wrap(b, ((self.stride - self.packed_bytes) // 4, ), "<i4")]

    def close(self):
        import ctypes as C
        if self.base:
            self._lib.pgd_peer_release(self._h, C.c_void_p(self.base), 1 if self.rank == 0 else 0)
            self.base = 0
