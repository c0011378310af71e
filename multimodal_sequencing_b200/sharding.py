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
