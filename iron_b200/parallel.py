"""Data-parallel glue of the stage-2 step: rays / patches shard across the ranks of one box,
weights are replicated, and the ONLY exchange is one all-reduce of the flat fp32 gradient buffer per step
(NCCL over NVLink on GPUs; the same code runs over gloo in the CPU tests).  The reference is single-GPU, so this
is new host logic, not a restatement."""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n work items: the first n % world ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def crop_for_rank(rank: int, patch: int, width: int = 512, height: int = 512, stride: int = 0) -> Tuple[int, int]:
    """Upper-left corner of rank's patch: rank 0 is the canonical centre crop, the others are distinct windows `stride` pixels
    apart around it (stride 0 / default = patch: the windows tile; a smaller stride gives overlapping windows that stay near
    the centre, i.e. comparable work per rank)."""
    offs = [(0, 0), (1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (-1, -1), (1, -1)]
    dx, dy = offs[rank % 8]
    ring = rank // 8 + 1
    step = stride if stride > 0 else patch
    cx, cy = width // 2 - patch // 2, height // 2 - patch // 2
    x = min(max(cx + dx * step * ring, 0), width - patch)
    y = min(max(cy + dy * step * ring, 0), height - patch)
    return x, y


def collect_params(modules: Iterable[torch.nn.Module]) -> List[torch.nn.Parameter]:
    out: List[torch.nn.Parameter] = []
    for m in modules:
        out += [p for p in m.parameters() if p.requires_grad]
    return out


def allreduce_flat(flat: torch.Tensor, world: int) -> int:
    """In-place SUM all-reduce of a flat gradient bucket whose producer already scaled it by 1 / world
    (GraphedStage2Step(flat_grads=True, grad_scale=1 / world): `.grad` of every parameter is a view of it).  One collective,
    no cat / div / copy-back passes; returns the number of elements exchanged."""
    if world <= 1:
        return 0
    dist.all_reduce(flat)
    return int(flat.numel())


def allreduce_gradients(params: Sequence[torch.nn.Parameter], world: int, average: bool = True) -> int:
    """One all-reduce over the flat gradient buffer; returns the number of elements exchanged.
    Parameters without a gradient contribute zeros (so every rank's buffer has the same layout)."""
    if world <= 1:
        return 0
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat)
    if average:
        flat.div_(world)
    off = 0
    dsts, srcs = [], []
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            dsts.append(p.grad)
            srcs.append(g)
        off += n
    if dsts:
        torch._foreach_copy_(dsts, srcs)     # one multi-tensor kernel instead of one copy launch per parameter
    return int(flat.numel())
