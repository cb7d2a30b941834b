"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on
B200, gloo in CPU tests).

The render path shards by ray index and needs NO data-path collective; training adds
exactly one collective per step, the all-reduce of the 595,844 fp32 gradients (2.4 MB,
latency bound) -- SURVEY.md section 8e.  The reference has no multi-GPU code at all.
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(num_items: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of ``num_items`` owned by ``rank`` (sizes differ by <= 1)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(num_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: List[int], rank: Optional[int] = None, world_size: Optional[int] = None) -> List[int]:
    """This rank's slice of a (shuffled) global ray batch -- every rank must pass the same list."""
    lo, hi = shard_range(len(batch), rank, world_size)
    return batch[lo:hi]


def allreduce_gradients(model: torch.nn.Module, average: bool = True):
    """Sum (or mean) the gradients of all ranks with ONE flat all-reduce.

    Use as ``Raycaster.fit(..., grad_sync=allreduce_gradients)``; clipping and Adam then run
    identically on every rank, so the replicas stay bit-identical."""
    rank, ws = world()
    if ws == 1:
        return
    grads = [p.grad for p in model.parameters() if p.requires_grad and p.grad is not None]
    if not grads:
        return
    # the training kernels write every gradient into one flat buffer and autograd hands the views on as .grad:
    # all-reduce that buffer in place (no gather / scatter copies)
    flat = model.__dict__.get("_ffn_flat_grad")
    if flat is not None:
        base = flat.untyped_storage().data_ptr()
        if all(g.untyped_storage().data_ptr() == base for g in grads):
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            trainer = model.__dict__.get("_ffn_trainer")
            if trainer is not None and trainer.flat_grad is flat:
                # FusedTrainer: the mean is taken inside ffn_clip_adam (g * 1/ws before clipping), no extra launch
                trainer.set_grad_scale(1.0 / ws if average else 1.0)
            elif average:
                flat /= ws
            return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat /= ws
    views, off = [], 0
    for g in grads:
        n = g.numel()
        views.append(flat[off:off + n].view_as(g))
        off += n
    torch._foreach_copy_(grads, views)      # one multi-tensor launch instead of one copy per parameter


def broadcast_parameters(model: torch.nn.Module, src: int = 0):
    """Make every rank start from rank ``src``'s weights."""
    _, ws = world()
    if ws == 1:
        return
    for p in model.parameters():
        dist.broadcast(p.data, src)


def gather_render(color: torch.Tensor, alpha: torch.Tensor, counts: List[int]):
    """Collect per-rank pixel slices of a frame on every rank (optional; 16 B/ray)."""
    rank, ws = world()
    if ws == 1:
        return color, alpha
    n = max(counts)
    buf = torch.zeros((n, 4), dtype=torch.float32, device=color.device)
    buf[:len(color), :3] = color
    buf[:len(alpha), 3] = alpha
    out = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(out, buf)
    full = torch.cat([o[:c] for o, c in zip(out, counts)])
    return full[:, :3].contiguous(), full[:, 3].contiguous()
