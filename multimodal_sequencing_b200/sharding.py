"""Data-parallel evaluation across GPUs: manuals are independent units (models/berson/eval.py:85-129 is a
plain `for batch` loop), so rank r of G takes manuals r, r+G, ... and NO collective sits on the data path.
The only exchange is the optional merge of the [B_r, N] int32 predictions for rank-0 metrics."""
import torch
import torch.distributed as dist


def shard_indices(n_manuals, rank, world):
    """Indices of the manuals rank `rank` orders (round-robin keeps equal-N shards balanced)."""
    return list(range(rank, n_manuals, world))


def merge_predictions(local_perm, local_idx, n_manuals, group=None):
    """all_gather the per-rank permutations into dataset order.  local_perm [B_r, N] int tensor (CPU for gloo,
    CUDA for nccl), local_idx the indices from shard_indices.  Returns [n_manuals, N] on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        out = torch.empty(n_manuals, local_perm.shape[1], dtype=local_perm.dtype, device=local_perm.device)
        out[torch.as_tensor(local_idx, device=local_perm.device)] = local_perm
        return out
    N = local_perm.shape[1]
    cap = (n_manuals + world - 1) // world
    dev = local_perm.device
    buf = torch.full((cap, N + 1), -1, dtype=torch.int32, device=dev)  # column 0 = dataset index (-1 = padding)
    buf[:len(local_idx), 0] = torch.as_tensor(local_idx, dtype=torch.int32, device=dev)
    buf[:len(local_idx), 1:] = local_perm.to(torch.int32)
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf, group=group)
    out = torch.full((n_manuals, N), -1, dtype=torch.int32, device=dev)
    for g in gathered:
        keep = g[:, 0] >= 0
        out[g[keep, 0].long()] = g[keep, 1:]
    assert (out >= 0).all(), "some manuals were not ordered by any rank"
    return out


def allreduce_gradients(flat_grads, group=None):
    """Data-parallel fine-tuning (BASELINE configs[3]): every rank holds the gradients of its own manuals in ONE flat fp32
    buffer (OrderingEngine.new_grad_buffer), so the whole exchange is a single all-reduce (NCCL over NVLink on the GPU
    box, gloo in the CPU tests).  Sums in place and returns the factor that turns the sum into the mean over ranks --
    pass it to adamw_step(grad_scale=...), which folds it into the clip coefficient instead of spending a pass on it."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


class GradientReducer:
    """Bucketed gradient all-reduce overlapped with the backward pass (what DistributedDataParallel does for the reference,
    trainers/train.py:217-221).  msq_train_step records a CUDA event per region of the flat gradient buffer when that region
    is final (heads, BERT layers top -> bottom, embeddings, visn_fc, ViT blocks top -> bottom, ViT stem); regions are merged
    into buckets of >= bucket_mb, and each bucket is all-reduced (NCCL, SUM) on a side stream that waits for the bucket's
    last event only -- the host has enqueued the whole step long before the GPU has executed it, so the collectives run
    under the rest of the backward pass.  `allreduce()` returns 1 / world for adamw_step(grad_scale=...).

    Falls back to one all-reduce of the whole buffer if the regions do not cover every parameter slot."""

    def __init__(self, engine, flat_grads, bucket_mb=64, group=None):
        self.eng, self.flat, self.group = engine, flat_grads, group
        self.bucket_elems = int(bucket_mb * (1 << 20) // 4)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.side = torch.cuda.Stream(device=flat_grads.device) if flat_grads.is_cuda else None
        self.n_buckets = 1
        self._plan = None

    def _regions(self):
        import ctypes as C
        lib, h = self.eng.lib, self.eng._h
        out = []
        for i in range(int(lib.msq_train_ready_count(h))):
            b, e = C.c_int64(), C.c_int64()
            if lib.msq_train_ready_info(h, i, C.byref(b), C.byref(e)) != 0:
                return None
            out.append((int(b.value), int(e.value)))
        return out

    def _make_plan(self):
        regions = self._regions()
        if not regions:
            return None
        # every parameter slot must lie inside a region
        for _, off, numel, _ in self.eng.train_layout():
            if not any(b <= off and off + numel <= e for b, e in regions):
                return None
        plan, cur = [], None   # bucket = [begin, end, index of its last region]
        for i, (b, e) in enumerate(regions):
            if cur is not None and (b == cur[1] or e == cur[0] or abs(b - cur[1]) < 1024 or abs(cur[0] - e) < 1024):
                cur = [min(cur[0], b), max(cur[1], e), i]
            else:
                if cur is not None:
                    plan.append(tuple(cur))
                cur = [b, e, i]
            if cur[1] - cur[0] >= self.bucket_elems:
                plan.append(tuple(cur))
                cur = None
        if cur is not None:
            plan.append(tuple(cur))
        return plan

    def allreduce(self):
        if self.world == 1:
            return 1.0
        if self._plan is None:
            self._plan = self._make_plan() or False
            self.n_buckets = len(self._plan) if self._plan else 1
        if not self._plan or self.side is None:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            return 1.0 / self.world
        import ctypes as C
        lib, h = self.eng.lib, self.eng._h
        works = []
        for b, e, last in self._plan:
            lib.msq_train_ready_wait(h, last, C.c_void_p(self.side.cuda_stream))
            with torch.cuda.stream(self.side):
                works.append(dist.all_reduce(self.flat[b:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for w in works:
            w.wait()     # the compute stream waits for the collectives; the host does not block
        return 1.0 / self.world
