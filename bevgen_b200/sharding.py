"""Multi-GPU plumbing: one process per GPU (torch.distributed), scenes sharded contiguously across ranks, weights broadcast
once from rank 0 (NCCL over NVLink on GPUs, gloo in CPU tests).  The data path has NO collective: scenes are independent
(no cross-sample op anywhere on the hot path; GroupNorm / LayerNorm are per sample), and the attention block layout is
built deterministically from the config on every rank, so the reference's only collective — dist.broadcast(master_layout),
modules/transformer/sparse_self_attention.py:50-52 — disappears."""
import torch
import torch.distributed as dist


def scene_shard(n_scenes: int, rank: int, world: int):
    """Contiguous, balanced partition: ranks [0, n % world) get one extra scene."""
    base, extra = divmod(n_scenes, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def broadcast_module_weights(module: torch.nn.Module, src: int = 0, bucket_bytes: int = 256 << 20):
    """Broadcast parameters AND persistent buffers of `module` from `src`, packed into large flat buckets (NVSwitch gives
    every peer full bandwidth, so buckets are sized for launch latency, not for link count).  Returns bytes sent."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    tensors = [p.data for p in module.parameters()] + [b for _, b in module.named_buffers() if b is not None and b.is_floating_point()]
    total, i = 0, 0
    while i < len(tensors):
        dtype, dev = tensors[i].dtype, tensors[i].device
        group, nbytes = [], 0
        while i < len(tensors) and tensors[i].dtype == dtype and tensors[i].device == dev and (not group or nbytes < bucket_bytes):
            group.append(tensors[i])
            nbytes += tensors[i].numel() * tensors[i].element_size()
            i += 1
        flat = torch.cat([t.reshape(-1) for t in group])
        dist.broadcast(flat, src=src)
        o = 0
        for t in group:
            t.copy_(flat[o:o + t.numel()].view_as(t))
            o += t.numel()
        total += nbytes
    return total


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
