"""Step runner: one MUNIT training step (dis_update + gen_update, train.py:182-187) executed as CUDA
graph replays -- a step issues ~2000 kernels, so Python launch overhead would otherwise dominate at
B = 8 -- and batch data parallelism: one process per GPU, replicated weights, gradients of the flat
arenas summed with NCCL all-reduce (torch.distributed) bucket by bucket as the backward pass completes them
(dp.GradSync, captured into the graph; MUNIT_DP_OVERLAP=0: whole arenas between three captured segments), the
1/world scale fused into the Adam kernel.  All norms are per-sample and every loss is a batch mean, so sharding the
batch and averaging gradients equals the single-GPU large-batch step."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

import os

from . import _lib, dp, ops


class StepRunner:
    def __init__(self, trainer, cfg: dict, batch: int, hw: int, use_graph: bool = True, world: int = 1,
                 two_streams: int = 0, reuse_forward: bool = False):
        """two_streams: 0 one stream; 1 the two domains' branches on two streams; 2 additionally every weight
        gradient on a companion stream, joined before the optimizer step.  reuse_forward: gen_update picks up the
        generator pass dis_update already ran on the same batch and weights (trainer.reuse_forward)."""
        self.t, self.cfg, self.batch, self.hw = trainer, cfg, batch, hw
        trainer.reuse_forward = bool(reuse_forward)
        trainer.parallel_streams = bool(two_streams and use_graph)
        from .networks import MsImageDis
        MsImageDis.scale_streams = bool(two_streams and use_graph) and os.environ.get("MUNIT_DIS_SCALE_STREAMS", "1") != "0"
        trainer.wgrad_overlap = bool(int(two_streams) >= 2 and use_graph)
        self.use_graph, self.world = use_graph, world
        # gen_update's generator pass issued under the discriminator update (trainer._early_generator_forward); needs the
        # whole step in one graph (no early pass across the segment boundaries of the split gradient exchange)
        split_exchange = world > 1 and os.environ.get("MUNIT_DP_OVERLAP", "1") == "0"
        trainer.overlap_updates = (bool(two_streams and use_graph) and not reuse_forward and not split_exchange
                                   and os.environ.get("MUNIT_OVERLAP_UPDATES", "1") != "0")
        # Data parallel: the round-2 concurrency features (CTA-pair clusters, third stream, early generator pass) have
        # run at 2 GPUs only -- the one 8-GPU run with all of them on completed the headline and then stopped making
        # progress inside a later graph replay (device-side, cause not isolated before the GPU budget ended;
        # profiles/r2_dp8_hang.md).  Until that is understood, ranks of a multi-GPU job execute the step the way
        # round 1 validated at 2 / 4 / 8 GPUs: two streams, single-CTA tap-GEMMs, sequential updates.
        # MUNIT_DP_FEATURES=1 turns the features back on.
        if world > 1 and os.environ.get("MUNIT_DP_FEATURES", "0") == "0":
            from . import kernels as _K
            trainer.overlap_updates = False
            trainer.style_stream = False
            _K.PAIR = False
        dev = next(trainer.parameters()).device
        self.dev = dev
        sd = trainer.style_dim
        self.x_a = torch.zeros(batch, 3, hw, hw, device=dev)
        self.x_b = torch.zeros(batch, 3, hw, hw, device=dev)
        self.s_a = torch.zeros(batch, sd, 1, 1, device=dev)
        self.s_b = torch.zeros(batch, sd, 1, 1, device=dev)
        self.s_a2 = torch.zeros(batch, sd, 1, 1, device=dev)
        self.s_b2 = torch.zeros(batch, sd, 1, 1, device=dev)
        self.extra = "extra" in cfg["optimizer"]
        self.graphs: Dict[int, list] = {}
        self.loss_refs: Dict[int, dict] = {}
        self.launches_per_step: Optional[int] = None
        for opt in (trainer.dis_opt, trainer.gen_opt):
            opt.build_arena()
            opt.enable_graph_hyper()
            opt.grad_scale = 1.0 / world
        # Gradient exchange.  Default for world > 1: dp.GradSync -- each bucket of a gradient arena is all-reduced
        # as soon as the backward pass has completed it, overlapped with the remaining dgrad / wgrad kernels, and
        # the collectives are captured into the step's CUDA graph.  MUNIT_DP_OVERLAP=0: the whole arena is reduced
        # between three captured segments (no NCCL inside a capture).
        self.overlap = world > 1 and os.environ.get("MUNIT_DP_OVERLAP", "1") != "0"
        ops.SYNCS.clear()
        trainer.grad_sync = {}
        if self.overlap:
            dp._wgrad_streams = ops._active_wgrad_streams
            bb = int(float(os.environ.get("MUNIT_DP_BUCKET_MB", "25")) * (1 << 20))
            for name, opt in (("dis", trainer.dis_opt), ("gen", trainer.gen_opt)):
                gs = dp.GradSync(opt.g_arena, opt.slices, bucket_bytes=bb)
                trainer.grad_sync[name] = gs
                ops.SYNCS.append(gs)
        self.iter = 0
        # Warm-up and capture share ONE side stream: autograd grad-accumulator nodes remember the stream
        # they were created on, and a node surviving from warm-up on another stream would make the
        # engine's end-of-backward stream sync reach outside the capture (cudaErrorStreamCaptureIsolation).
        self.stream = torch.cuda.Stream()

    # ------------------------------------------------------------------ pieces of a step
    def _seg_dis(self):
        self.t._dis_backward(self.x_a, self.x_b, self.cfg, self.s_a, self.s_b,
                             early_gen=(self.s_a2, self.s_b2) if self.t.overlap_updates else None)

    def _seg_mid(self):
        self.t.dis_opt_step()
        self.t._gen_backward(self.x_a, self.x_b, self.cfg, None, None, False, self.s_a2, self.s_b2)

    def _seg_end(self):
        self.t.gen_opt_step()

    def _allreduce(self, opt):
        if self.world > 1 and not self.overlap:
            dp.allreduce_arena(opt.g_arena)

    def _eager_step(self):
        self._seg_dis()
        self._allreduce(self.t.dis_opt)
        self._seg_mid()
        self._allreduce(self.t.gen_opt)
        self._seg_end()

    def _py_state(self):
        t = self.t
        return [(o.step_count, getattr(o, "_have_copy", None)) for o in (t.dis_opt, t.gen_opt)]

    def _restore_py_state(self, st):
        for o, (c, h) in zip((self.t.dis_opt, self.t.gen_opt), st):
            o.step_count = c
            if h is not None:
                o._have_copy = h

    def _capture(self, parity: int):
        """Capture the step for iterations of the given parity (ExtraAdam alternates extrapolation/step)."""
        st = self._py_state()
        if self.extra:  # python-side flag ExtraAdam.step() checks; device state is untouched by capture
            for o in (self.t.dis_opt, self.t.gen_opt):
                o._have_copy = parity == 1
        split = self.world > 1 and not self.overlap
        segs = [self._seg_dis, self._seg_mid, self._seg_end] if split else [self._eager_step]
        graphs = []
        pool = None
        before = _lib.launches
        # captured NCCL collectives: the process group's watchdog thread may touch CUDA while this thread captures
        mode = "thread_local" if self.overlap else "global"
        for fn in segs:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool, stream=self.stream, capture_error_mode=mode):
                fn()
            pool = g.pool()
            graphs.append(g)
        self.launches_per_step = _lib.launches - before
        self._restore_py_state(st)
        self.graphs[parity] = graphs
        # the loss_* tensors this graph writes (static outputs); re-published on the trainer after each replay
        self.loss_refs[parity] = {k: v for k, v in self.t.__dict__.items() if k.startswith("loss_") and torch.is_tensor(v)}

    # ------------------------------------------------------------------ public
    def warmup_and_capture(self, eager_steps: int = 2):
        eager_steps = max(eager_steps, 1)  # the first eager step sets kernel attributes / builds plans and shadows
        if self.extra and (self.iter + eager_steps) % 2:
            eager_steps += 1  # ExtraAdam: capture starts on an extrapolation (even) iteration
        """Eager steps (sets kernel attributes, builds plans/shadows, warms the allocator), then capture."""
        s = self.stream
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(eager_steps):
                self._prepare_host_state()
                before = _lib.launches
                self._eager_step()
                self.launches_per_step = _lib.launches - before
                self._advance()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if self.use_graph:
            n_par = 2 if self.extra else 1
            assert not self.extra or self.iter % 2 == 0
            for par in range(n_par):
                self.t.iterations = self.iter + par
                self._capture(par)
            self.t.iterations = self.iter
            torch.cuda.synchronize()

    def _prepare_host_state(self):
        t = self.t
        t.iterations = self.iter
        if not torch.cuda.is_current_stream_capturing():
            for opt in (t.dis_opt, t.gen_opt):  # H2D of {lr, bias corrections} ahead of the replay, race-free ring
                opt.upload_hyper(opt.step_count + 1)

    def _advance(self):
        self.iter += 1

    def step(self):
        """One dis_update + gen_update on the resident inputs."""
        self._prepare_host_state()
        if not self.use_graph:
            self._eager_step()
        else:
            graphs = self.graphs[self.iter % 2 if self.extra else 0]
            t = self.t
            if len(graphs) == 3:
                graphs[0].replay()
                self._allreduce(t.dis_opt)
                graphs[1].replay()
                self._allreduce(t.gen_opt)
                graphs[2].replay()
            else:
                graphs[0].replay()
            for k, v in self.loss_refs[self.iter % 2 if self.extra else 0].items():
                setattr(t, k, v)
            for opt in (t.dis_opt, t.gen_opt):  # python-side bookkeeping the replay does not run
                opt.step_count += 1
                if self.extra:
                    opt._have_copy = (self.iter % 2 == 0)
        self._advance()

    def release(self):
        """Drop the captured graphs (they hold references to the NCCL communicator's kernels: destroy them before
        torch.distributed.destroy_process_group(), which otherwise waits on them at exit)."""
        import gc

        torch.cuda.synchronize()
        self.graphs.clear()
        self.loss_refs.clear()
        gc.collect()
        torch.cuda.synchronize()

    def load_inputs(self, x_a, x_b, s_a, s_b, s_a2, s_b2):
        """Host (pinned) -> device copies of one step's inputs: images and the four style-code draws
        (dis_update then gen_update, trainer.py:1146-1147,366-367)."""
        self.x_a.copy_(x_a, non_blocking=True)
        self.x_b.copy_(x_b, non_blocking=True)
        self.s_a.copy_(s_a, non_blocking=True)
        self.s_b.copy_(s_b, non_blocking=True)
        self.s_a2.copy_(s_a2, non_blocking=True)
        self.s_b2.copy_(s_b2, non_blocking=True)

    def losses(self):
        t = self.t
        return dict(loss_dis_total=t.loss_dis_total, loss_gen_total=t.loss_gen_total)
