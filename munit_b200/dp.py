"""Batch data parallelism for the MUNIT step (one process per GPU, torch.distributed/NCCL over NVLink).

The reference has no multi-GPU path.  Sharding is exact here: InstanceNorm / AdaIN / the
MUNIT LayerNorm normalise per sample, the discriminator has no norm, and every loss is a mean over the
batch -- so summing the per-rank gradients (each a mean over B/W samples) and scaling by 1/W reproduces the
single-process gradient of the global batch up to fp summation order.  The 1/W scale is fused into the Adam
kernel (FlatAdam.grad_scale); gradients live in one flat fp32 arena per optimiser, all-reduced in buckets.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist

BUCKET_BYTES = 32 << 20  # ~32 MB: large enough for NVLink bandwidth, small enough to pipeline


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rank r owns the contiguous slice [r*B/W, (r+1)*B/W) of the global batch."""
    b = x.shape[0]
    assert b % world == 0, f"global batch {b} not divisible by world size {world}"
    per = b // world
    return x[rank * per:(rank + 1) * per]


def global_style_noise(global_batch: int, style_dim: int, rank: int, world: int) -> torch.Tensor:
    """Bit-exact style codes under DP: every rank draws the GLOBAL randn(B, style_dim, 1, 1) from the host
    generator (same seed on every rank, the reference's call trainer.py:366) and keeps its slice."""
    return shard_batch(torch.randn(global_batch, style_dim, 1, 1), rank, world)


def bucket_slices(numel: int, bucket_bytes: int = BUCKET_BYTES, elem_bytes: int = 4) -> List[Tuple[int, int]]:
    """[start, end) element ranges covering a flat arena, last-to-first: backward fills the arena from its
    tail (parameters are laid out in forward order), so tail buckets are ready first."""
    per = max(1, bucket_bytes // elem_bytes)
    out = []
    end = numel
    while end > 0:
        start = max(0, end - per)
        out.append((start, end))
        end = start
    return out


def allreduce_arena(arena: torch.Tensor, group=None, bucket_bytes: int = BUCKET_BYTES, async_op: bool = False):
    """Sum-all-reduce a flat gradient arena bucket by bucket.  Returns the work handles when async."""
    works = []
    for s, e in bucket_slices(arena.numel(), bucket_bytes, arena.element_size()):
        w = dist.all_reduce(arena[s:e], op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works
