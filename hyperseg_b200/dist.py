"""Data-parallel inference over the GPUs of one box: one process per GPU, no communication in the forward.

Images are independent (the reference's only multi-device mode is nn.DataParallel over the batch:
hyperseg/test.py:136-137, hyperseg/test_fps.py:155-156), so a batch is split into contiguous shards, every rank
runs its shard through its own model replica, and NCCL (over NVLink / NVSwitch) is needed only when a whole-box
result is requested:

  * gather_logits     all_gather of the per-rank logits (what DataParallel's gather returns on GPU 0);
  * confusion_matrix  per-rank C x C int64 matrix of (label, prediction) pairs, all-reduced with SUM -- the
                      operation the reference sketches in ConfusionMatrix.reduce_from_all_processes
                      (hyperseg/utils/seg_utils.py:38-44) -- from which mIoU follows (seg_utils.py:46-56).
                      It moves C*C*8 bytes instead of B*C*H*W logits and gives the same mIoU.

Everything works with any torch.distributed backend; tests run it with gloo on CPU (world_size 2).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_items: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of n_items for `rank`; the first n_items % world_size ranks get one extra item."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch: torch.Tensor, rank: int | None = None, world_size: int | None = None) -> torch.Tensor:
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    lo, hi = shard_bounds(batch.shape[0], rank, world_size)
    return batch[lo:hi]


def gather_logits(local: torch.Tensor, total_items: int | None = None) -> torch.Tensor:
    """Concatenate the per-rank outputs along dim 0 on every rank.  Shards may be ragged (B % world != 0): the ranks
    first exchange their shard sizes (one tiny all_gather), then gather shards padded to the largest one and trim.
    ``total_items`` is only checked against the exchanged sizes."""
    rank, w = world()
    if w == 1:
        return local
    mine = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
    sizes = [torch.empty_like(mine) for _ in range(w)]
    dist.all_gather(sizes, mine)
    counts = [int(c.item()) for c in sizes]
    if total_items is not None and sum(counts) != total_items:
        raise ValueError(f"shards hold {sum(counts)} items, expected {total_items}")
    m = max(counts)
    if all(c == m for c in counts):
        out = [torch.empty_like(local) for _ in range(w)]
        dist.all_gather(out, local.contiguous())
        return torch.cat(out, 0)
    pad = local.new_zeros((m,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(out, pad)
    return torch.cat([o[:c] for o, c in zip(out, counts)], 0)


def confusion_matrix(pred: torch.Tensor, target: torch.Tensor, num_classes: int, ignore_index: int = 255) -> torch.Tensor:
    """num_classes x num_classes int64 counts, rows = ground truth, columns = prediction (seg_utils.py:10-17)."""
    pred = pred.reshape(-1).long()
    target = target.reshape(-1).long()
    keep = (target != ignore_index) & (target >= 0) & (target < num_classes) & (pred >= 0) & (pred < num_classes)
    idx = target[keep] * num_classes + pred[keep]
    return torch.bincount(idx, minlength=num_classes * num_classes).reshape(num_classes, num_classes)


def all_reduce_confusion(mat: torch.Tensor) -> torch.Tensor:
    """Sum the per-rank matrices (NCCL all-reduce of C*C int64 over NVLink on a GPU box)."""
    _, w = world()
    if w > 1:
        dist.all_reduce(mat, op=dist.ReduceOp.SUM)
    return mat


def miou(mat: torch.Tensor) -> tuple[float, torch.Tensor]:
    """Global accuracy-free summary: mean IoU and per-class IoU from a confusion matrix (seg_utils.py:46-56)."""
    h = mat.double()
    iou = torch.diag(h) / (h.sum(1) + h.sum(0) - torch.diag(h))
    return torch.nanmean(iou).item(), iou
