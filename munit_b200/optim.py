"""Flat-arena optimisers: torch.optim.Adam (as resolved by trainer.py:41-45 on torch 2.x) and ExtraAdam
(scripts/extraadam.py:14-168) as ONE multi-tensor kernel launch over a contiguous fp32 arena.

At construction every parameter is re-homed into the arena (`p.data` becomes a view, conv weights keep
their channels_last strides) and receives a persistent `.grad` view of a second arena, so that
  * the wgrad / bias / norm kernels accumulate gradients in place (ops.py),
  * zero_grad() is one memset, the update one launch, and a data-parallel all-reduce sees flat buffers.
"""
from __future__ import annotations

from typing import List

import torch
from torch.optim import Optimizer

from . import kernels as K


def _bump_versions(params):
    """Our kernels write parameters through raw pointers; tell autograd (and the weight-shadow caches
    keyed on `_version`, ops.ConvLayer.refresh) that they changed."""
    ps = tuple(params)
    torch._C._autograd._unsafe_set_version_counter(ps, tuple(p._version + 1 for p in ps))


class FlatAdam(Optimizer):
    """Adam with coupled L2 weight decay; `mode` selects torch's formula (denominator
    sqrt(v)/sqrt(1-b2^t)+eps) -- the optimiser the reference instantiates when `optimizer: adam`."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if amsgrad:
            raise NotImplementedError("amsgrad is not used by any MUNIT config")
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad)
        super().__init__(params, defaults)
        self._arena_built = False
        self.step_count = 0
        self.grad_scale = 1.0  # 1/world_size under data parallelism (fused into the update)
        # CUDA-graph mode: {lr, 1-b1^t, 1-b2^t} live in a device array.  It is refreshed by an H2D copy that is
        # enqueued *outside* the captured graph, from a ring of pinned host slots: a slot is rewritten only after the
        # copy that last read it has completed (event per slot), so un-synchronised replay loops can never pick up the
        # scalars of a later step (the round-1 design re-used one pinned buffer inside the graph and could).
        self.hyper_ring = None
        self.hyper_dev = None
        self._ring_events = []
        self._ring_pos = 0

    # ------------------------------------------------------------------ arena
    def _all_params(self) -> List[torch.nn.Parameter]:
        return [p for g in self.param_groups for p in g["params"]]

    def build_arena(self):
        if self._arena_built:
            return
        ps = self._all_params()
        dev = ps[0].device
        assert dev.type == "cuda", "munit_b200 optimisers run on the GPU only (no CPU fallback)"
        sizes = [((p.numel() + 63) // 64) * 64 for p in ps]  # 256-byte aligned slices
        total = sum(sizes)
        self.p_arena = torch.zeros(total, dtype=torch.float32, device=dev)
        self.g_arena = torch.zeros(total, dtype=torch.float32, device=dev)
        self.m_arena = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v_arena = torch.zeros(total, dtype=torch.float32, device=dev)
        self.saved_arena = None
        off = 0
        self.slices = []
        for p, sz in zip(ps, sizes):
            n = p.numel()

            def view(arena):
                flat = arena[off:off + n]
                if p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous():
                    co, ci, kh, kw = p.shape
                    return flat.view(co, kh, kw, ci).permute(0, 3, 1, 2)
                return flat.view(p.shape)

            pv = view(self.p_arena)
            pv.copy_(p.data)
            p.data = pv
            p.grad = view(self.g_arena)
            self.slices.append((off, n))
            off += sz
        self._arena_built = True

    def zero_grad(self, set_to_none: bool = False):
        if not self._arena_built:
            self.build_arena()
        self.g_arena.zero_()

    # ------------------------------------------------------------------ update
    def _launch(self, mode: int, save: bool):
        if not self._arena_built:
            self.build_arena()
        g = self.param_groups[0]
        self.step_count += 1
        if mode != 0 and self.saved_arena is None:
            self.saved_arena = torch.empty_like(self.p_arena)
        if self.hyper_dev is not None and not torch.cuda.is_current_stream_capturing():
            # eager call (trainer.*_opt_step outside a StepRunner replay, after resume, ...): the device scalars must
            # describe THIS step; inside a capture the runner uploads them before every replay
            self.upload_hyper(self.step_count)
        K.adam(self.p_arena, self.g_arena, self.m_arena, self.v_arena, self.saved_arena, None, mode, save, g["lr"],
               g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], self.step_count, self.grad_scale,
               self.hyper_dev)
        ps = self._all_params()
        _bump_versions(ps)
        from . import ops  # (ops does not import optim)

        ops.refresh_shadows(self, ps)  # every bf16 weight shadow of this arena, one launch

    # ------------------------------------------------------------------ CUDA-graph support
    HYPER_RING = 32

    def enable_graph_hyper(self):
        """Route lr / bias corrections through device memory so a captured step can be replayed."""
        if not self._arena_built:
            self.build_arena()
        if self.hyper_dev is None:
            self.hyper_ring = torch.zeros(self.HYPER_RING, 4, dtype=torch.float32).pin_memory()
            self.hyper_dev = torch.zeros(4, dtype=torch.float32, device=self.p_arena.device)
            self._ring_events = [None] * self.HYPER_RING
            self.upload_hyper(self.step_count + 1)  # never leave the device scalars at 0 (lr / 0 -> NaN)

    def upload_hyper(self, step: int):
        """Enqueue, on the current stream, the H2D copy of the scalars for optimiser step `step` (1-based).  Must be
        called outside a graph capture."""
        g = self.param_groups[0]
        slot = self._ring_pos % self.HYPER_RING
        self._ring_pos += 1
        ev = self._ring_events[slot]
        if ev is not None:
            ev.synchronize()  # the copy that last read this slot (HYPER_RING uploads ago) has finished
        row = self.hyper_ring[slot]
        row[0] = g["lr"]
        row[1] = 1.0 - g["betas"][0] ** step
        row[2] = 1.0 - g["betas"][1] ** step
        self.hyper_dev.copy_(row, non_blocking=True)
        ev = ev or torch.cuda.Event()
        ev.record()
        self._ring_events[slot] = ev

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self._launch(0, False)
        return loss

    # ------------------------------------------------------------------ checkpoint contract (trainer.save/resume)
    def _state_view(self, arena, p, off):
        """The slice of `arena` that belongs to parameter p, shaped and strided like p itself (the arenas keep
        conv weights in channels_last order, so a flat reshape would scramble them)."""
        return arena.as_strided(p.shape, p.stride(), off)

    def state_dict(self):
        """Interchangeable with torch.optim.Adam.state_dict() (what the reference writes to optimizer.pt,
        trainer.py:1401-1429): per-parameter step / exp_avg / exp_avg_sq with the parameter's logical shape."""
        if not self._arena_built:
            self.build_arena()
        state = {}
        for i, (p, (off, n)) in enumerate(zip(self._all_params(), self.slices)):
            state[i] = dict(step=torch.tensor(float(self.step_count)),
                            exp_avg=self._state_view(self.m_arena, p, off).clone(),
                            exp_avg_sq=self._state_view(self.v_arena, p, off).clone())
        groups = [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]
        groups[0]["params"] = list(range(len(self.slices)))
        return dict(state=state, param_groups=groups)

    def load_state_dict(self, sd):
        """Accepts this class's state_dict() and torch.optim.Adam's (a reference optimizer.pt)."""
        if not self._arena_built:
            self.build_arena()
        for i, (p, (off, n)) in enumerate(zip(self._all_params(), self.slices)):
            st = sd["state"].get(i)
            if st is None:
                continue
            for arena, key in ((self.m_arena, "exp_avg"), (self.v_arena, "exp_avg_sq")):
                src = st[key]
                if src.numel() != n:
                    raise ValueError(f"optimizer state {key}[{i}] has {src.numel()} elements, parameter has {n}")
                self._state_view(arena, p, off).copy_(src.reshape(p.shape) if src.dim() != p.dim() else src)
            self.step_count = int(float(st["step"]))
        for g, sg in zip(self.param_groups, sd["param_groups"]):
            for k in ("lr", "betas", "eps", "weight_decay"):
                if k in sg:
                    g[k] = sg[k]


class ExtraAdam(FlatAdam):
    """scripts/extraadam.py: extragradient Adam.  extrapolation(): save p once, p += u;
    step(): p = saved + u.  `u` uses the legacy Adam formula (denominator sqrt(v)+eps,
    extraadam.py:155-168); moments and the step counter advance on both calls."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        super().__init__(params, lr, betas, eps, weight_decay, amsgrad)
        self._have_copy = False

    @torch.no_grad()
    def extrapolation(self):
        self._launch(1, not self._have_copy)
        self._have_copy = True

    @torch.no_grad()
    def step(self, closure=None):
        if not self._have_copy:
            raise RuntimeError("Need to call extrapolation before calling step.")
        loss = closure() if closure is not None else None
        self._launch(2, False)
        self._have_copy = False
        return loss
