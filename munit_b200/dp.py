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


class GradSync:
    """Readiness-driven bucketed all-reduce of one flat gradient arena, overlapped with the backward pass
    (the role DDP's reducer plays for autograd-accumulated gradients).

    The kernels accumulate parameter gradients straight into the arena (ops.py), so there is no AccumulateGrad
    node to hook.  Instead every autograd Function that owns parameters reports `note_use(grad_buf)` in its
    forward and `note_done(grad_buf)` after its backward has launched the gradient kernels.  A bucket -- a
    contiguous run of whole parameters, cut from the arena's tail because backward fills it tail first -- is
    complete when all uses of all its parameters are done; it is then all-reduced asynchronously (NCCL runs on
    its own stream, ordered after the compute stream(s) that produced the bucket) while the rest of the backward
    keeps the SMs busy.  Buckets are always launched in the same fixed order on every rank.  `finish()` launches
    what is left (parameters never used this pass) and makes the current stream wait for every reduction.
    Works under CUDA-graph capture: the collectives become nodes of the captured step.
    """

    def __init__(self, arena: torch.Tensor, slices: List[Tuple[int, int]], bucket_bytes: int = BUCKET_BYTES, group=None):
        self.arena, self.group = arena, group
        self.base = arena.data_ptr()
        self.esize = arena.element_size()
        per = max(1, bucket_bytes // self.esize)
        # buckets of whole parameters, last parameter first
        self.buckets: List[Tuple[int, int]] = []
        end = arena.numel()
        start_of = sorted(o for o, _ in slices)
        cur_end = end
        cur_start = end
        for o in reversed(start_of):
            cur_start = o
            if cur_end - cur_start >= per:
                self.buckets.append((cur_start, cur_end))
                cur_end = cur_start
        if cur_end > 0:
            self.buckets.append((0, cur_end))
        self._starts = [s for s, _ in self.buckets]  # descending
        self.enabled = False
        self.comm = None
        self.launched_early = 0  # cumulative count of buckets that went out before finish()
        self.early_pass = 0      # ... in the most recent begin()..finish() pass
        self._reset()

    def _reset(self):
        nb = len(self.buckets)
        self.pending = [0] * nb
        self.used = [False] * nb
        self.ready = [False] * nb
        self.next = 0  # next bucket to launch (fixed order)
        self.works = []
        self.events = [dict() for _ in range(nb)]

    def bucket_of(self, ptr: int) -> int:
        off = (ptr - self.base) // self.esize
        if ptr < self.base or off >= self.arena.numel():
            return -1
        for i, s in enumerate(self._starts):  # few buckets: linear scan
            if off >= s:
                return i
        return -1

    def begin(self):
        """Start of a backward phase (after zero_grad)."""
        self._reset()
        self.early_pass = 0
        self.enabled = True

    def note_use(self, buf: torch.Tensor):
        if not self.enabled:
            return
        b = self.bucket_of(buf.data_ptr())
        if b >= 0:
            self.pending[b] += 1
            self.used[b] = True

    def note_done(self, buf: torch.Tensor):
        if not self.enabled:
            return
        b = self.bucket_of(buf.data_ptr())
        if b < 0:
            return
        self.pending[b] -= 1
        if self.arena.is_cuda:
            cur = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(cur)
            self.events[b][cur.cuda_stream] = ev
        if self.pending[b] == 0 and self.used[b]:
            self.ready[b] = True
            self._launch_ready(early=True)

    def _launch(self, b: int):
        s, e = self.buckets[b]
        if not self.arena.is_cuda:
            self.works.append(dist.all_reduce(self.arena[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            return
        # The collective is issued from a dedicated launch stream that waits for the last gradient kernel of the
        # bucket on every compute stream that contributed: the compute streams themselves never wait for each
        # other or for the network here.
        if self.comm is None:
            self.comm = torch.cuda.Stream()
        for ev in self.events[b].values():
            self.comm.wait_event(ev)
        if _wgrad_streams is not None:
            for ws in _wgrad_streams():
                self.comm.wait_stream(ws)
        with torch.cuda.stream(self.comm):
            self.works.append(dist.all_reduce(self.arena[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _launch_ready(self, early: bool):
        while self.next < len(self.buckets) and self.ready[self.next]:
            self._launch(self.next)
            self.next += 1
            if early:
                self.launched_early += 1
                self.early_pass += 1

    def finish(self):
        """End of the backward phase (all compute streams joined into the current one): reduce what is left,
        then order the current stream after every reduction."""
        if not self.enabled:
            return
        for b in range(len(self.buckets)):
            self.ready[b] = True
        if self.arena.is_cuda and self.next < len(self.buckets):
            # buckets without a recorded use this pass: order them after everything the current stream has seen
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            for b in range(self.next, len(self.buckets)):
                self.events[b][-1] = ev
        self._launch_ready(early=False)
        for w in self.works:
            w.wait()  # the current stream waits for the reduction (NCCL's stream), not the host
        if self.arena.is_cuda and self.comm is not None:
            torch.cuda.current_stream().wait_stream(self.comm)
        self.works = []
        self.enabled = False


_wgrad_streams = None  # set by ops: companion streams with outstanding weight-gradient launches
