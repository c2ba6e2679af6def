"""Multi-GPU plumbing: one process per GPU (torch.distributed), scenes sharded contiguously across ranks, weights broadcast
once from rank 0 (NCCL over NVLink on GPUs, gloo in CPU tests).  The data path has NO collective: scenes are independent
(no cross-sample op anywhere on the hot path; GroupNorm / LayerNorm are per sample).  The attention block layouts are RNG-free
only at density = 1.0; for density < 1 they are drawn with torch.multinomial from the global RNG (like the reference), so the
integer `master_layout` buffers travel with the weights in the same start-up broadcast — the reference's only collective,
dist.broadcast(master_layout) (modules/transformer/sparse_self_attention.py:50-52), folded into it."""
import torch
import torch.distributed as dist


def scene_shard(n_scenes: int, rank: int, world: int):
    """Contiguous, balanced partition: ranks [0, n % world) get one extra scene."""
    base, extra = divmod(n_scenes, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def broadcast_module_weights(module: torch.nn.Module, src: int = 0, bucket_bytes: int = 256 << 20):
    """Broadcast parameters AND buffers (floating point and integer, e.g. `master_layout`) of `module` from `src`, packed into
    large flat buckets per dtype (NVSwitch gives every peer full bandwidth, so buckets are sized for launch latency, not for
    link count).  The copies go through `Tensor.copy_` under no_grad (bumping `_version`) and every pre-packed engine below
    `module` is dropped afterwards (engine_cache.invalidate_all).  Returns bytes sent."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    from .engine_cache import invalidate_all
    tensors = list(module.parameters()) + [b for _, b in module.named_buffers() if b is not None]
    tensors.sort(key=lambda t: str(t.dtype))          # stable: one run of buckets per dtype
    total, i = 0, 0
    while i < len(tensors):
        dtype, dev = tensors[i].dtype, tensors[i].device
        group, nbytes = [], 0
        while i < len(tensors) and tensors[i].dtype == dtype and tensors[i].device == dev and (not group or nbytes < bucket_bytes):
            group.append(tensors[i])
            nbytes += tensors[i].numel() * tensors[i].element_size()
            i += 1
        with torch.no_grad():
            flat = torch.cat([t.detach().reshape(-1) for t in group])
            dist.broadcast(flat, src=src)
            if dist.get_rank() != src:
                for t, piece in zip(group, flat.split([t.numel() for t in group])):
                    t.copy_(piece.view_as(t))
        total += nbytes
    invalidate_all(module)
    return total


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
