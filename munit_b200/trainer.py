"""B200-native drop-in for cc-ai/MUNIT `scripts/trainer.py::MUNIT_Trainer` (hot path).

Same constructor argument (the config dict), methods and attributes as the reference:
forward (trainer.py:307-334), gen_update (:336-561), dis_update (:1133-1186), sample (:773-928),
sample_fid (:1087-1131), update_learning_rate (:1326-1335), save / resume (:1337-1429),
gen_opt_step / dis_opt_step (:252-268), recon_criterion(_mask) (:279-305); attributes gen | gen_a, gen_b,
dis_a, dis_b, gen_opt, dis_opt, s_a, s_b, style_dim, iterations, loss_*.

Differences that do not change results: the generator pass of dis_update runs without autograd (the
reference records it and then .detach()es, trainer.py:1178-1179); gen_update does not compute the
discriminator weight gradients the reference computes and discards (trainer.py:490-491 vs :1145).
Domain-adaptation head: `adaptation.dfeat_lambda > 0` builds domain_classifier_sr_a / _b
(trainer.py:162-179), `adaptation.adv_lambda > 0` adds compute_classifier_sr_loss(fool=True) to gen_update
(:521-525,555) and domain_classifier_sr_update (:1237-1265) trains the classifiers; in gen_update the classifier
weight gradients, which the reference computes and zeroes unused (:1241), are not computed.
Out of scope (raise NotImplementedError): semantic_w, domain_adv_w (the reference's compute_domain_adv_loss returns
None, :669-714), output classifiers, sem_seg_lambda, vgg_w.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import ops
from .heads import domainClassifier
from .networks import AdaINGen, AdaINGen_double, MsImageDis
from .optim import ExtraAdam, FlatAdam
from .ops import Act
from .utils import get_model_list, get_scheduler, weights_init


def _scalar(t):
    return t.reshape(())


class _nvtx:
    """NVTX range around the two update phases when MUNIT_NVTX=1 (the library marks every launch, csrc/api.cu)."""
    on = os.environ.get("MUNIT_NVTX", "0") == "1"

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _nvtx.on:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _nvtx.on:
            torch.cuda.nvtx.range_pop()
        return False


class MUNIT_Trainer(nn.Module):
    def __init__(self, hyperparameters):
        super().__init__()
        lr = hyperparameters["lr"]
        self.gen_state = hyperparameters["gen_state"]
        self.guided = hyperparameters["guided"]
        self.newsize = hyperparameters["crop_image_height"]
        self.semantic_w = hyperparameters["semantic_w"] > 0
        self.recon_mask = hyperparameters["recon_mask"] == 1
        self.dann_scheduler = None
        self.full_adaptation = hyperparameters["adaptation"]["full_adaptation"] == 1
        self.hyperparameters = hyperparameters
        for k, why in (("semantic_w", "needs the Resnet34_8s checkpoint"), ("domain_adv_w", "domainClassifier head"),
                       ("vgg_w", "load_vgg16 always raises in the reference")):
            if hyperparameters.get(k, 0) > 0:
                raise NotImplementedError(f"{k} > 0 is outside the B200 hot path ({why}); see SURVEY.md s2")
        for k in ("sem_seg_lambda", "output_classifier_lambda", "output_adv_lambda"):
            if hyperparameters["adaptation"].get(k, 0) > 0:
                raise NotImplementedError(f"adaptation.{k} > 0 is outside the B200 hot path; see SURVEY.md s8(f)")
        if hyperparameters["adaptation"].get("adv_lambda", 0) > 0 and not hyperparameters["adaptation"].get("dfeat_lambda", 0) > 0:
            # the reference would fail with AttributeError in gen_update (trainer.py:162,521-525,652)
            raise ValueError("adaptation.adv_lambda > 0 needs adaptation.dfeat_lambda > 0 (the classifiers are only built then)")
        optimizer = FlatAdam if "extra" not in hyperparameters["optimizer"] else ExtraAdam
        self.domain_classif_ab = False
        self.use_classifier_sr = hyperparameters["adaptation"].get("dfeat_lambda", 0) > 0
        self.train_seg = False
        self.use_output_classifier_sr = False

        if self.gen_state == 0:
            self.gen_a = AdaINGen(hyperparameters["input_dim_a"], hyperparameters["gen"])  # auto-encoder for domain a
            self.gen_b = AdaINGen(hyperparameters["input_dim_b"], hyperparameters["gen"])  # auto-encoder for domain b
        elif self.gen_state == 1:
            self.gen = AdaINGen_double(hyperparameters["input_dim_a"], hyperparameters["gen"])
        else:
            print("self.gen_state unknown value:", self.gen_state)
        self.dis_a = MsImageDis(hyperparameters["input_dim_a"], hyperparameters["dis"])  # discriminator for domain a
        self.dis_b = MsImageDis(hyperparameters["input_dim_b"], hyperparameters["dis"])  # discriminator for domain b
        self.instancenorm = nn.InstanceNorm2d(512, affine=False)
        self.style_dim = hyperparameters["gen"]["style_dim"]

        # fix the noise used in sampling (host RNG, drawn mid-stream exactly like trainer.py:93-95)
        display_size = int(hyperparameters["display_size"])
        self.s_a = torch.randn(display_size, self.style_dim, 1, 1)
        self.s_b = torch.randn(display_size, self.style_dim, 1, 1)

        beta1 = hyperparameters["beta1"]
        beta2 = hyperparameters["beta2"]
        dis_params = list(self.dis_a.parameters()) + list(self.dis_b.parameters())
        if self.gen_state == 0:
            gen_params = list(self.gen_a.parameters()) + list(self.gen_b.parameters())
        else:
            gen_params = list(self.gen.parameters())
        self.dis_opt = optimizer([p for p in dis_params if p.requires_grad], lr=lr, betas=(beta1, beta2),
                                 weight_decay=hyperparameters["weight_decay"])
        self.gen_opt = optimizer([p for p in gen_params if p.requires_grad], lr=lr, betas=(beta1, beta2),
                                 weight_decay=hyperparameters["weight_decay"])
        self.dis_scheduler = get_scheduler(self.dis_opt, hyperparameters)
        self.gen_scheduler = get_scheduler(self.gen_opt, hyperparameters)

        # Network weight initialization (trainer.py:124-127)
        self.apply(weights_init(hyperparameters["init"]))
        self.dis_a.apply(weights_init("gaussian"))
        self.dis_b.apply(weights_init("gaussian"))
        self.iterations = 0
        # Classifier on the content features for the synthetic / real adaptation (trainer.py:162-179): built after
        # the generator / discriminator initialisation, b before a, initialised a before b -- the reference's order,
        # so a seeded construction draws the same values.
        if self.use_classifier_sr:
            self.domain_classifier_sr_b = domainClassifier(256)
            self.domain_classifier_sr_a = domainClassifier(256)
            dann_params = list(self.domain_classifier_sr_a.parameters()) + list(self.domain_classifier_sr_b.parameters())
            self.classif_opt_sr = optimizer([p for p in dann_params if p.requires_grad], lr=lr, betas=(beta1, beta2),
                                            weight_decay=hyperparameters["weight_decay"])
            self.domain_classifier_sr_a.apply(weights_init("gaussian"))
            self.domain_classifier_sr_b.apply(weights_init("gaussian"))
            self.classif_sr_scheduler = get_scheduler(self.classif_opt_sr, hyperparameters)
        # Two-stream mode (set by engine.StepRunner under CUDA-graph capture): the domain-a and domain-b branches
        # of a step are independent for long stretches; forking them onto two streams gives the captured graph
        # parallel branches, which fills partial waves and hides launch latency of the many small kernels.
        # Data parallelism: dp.GradSync per optimiser arena ("gen" / "dis"), installed by engine.StepRunner; the
        # gradient all-reduce of a bucket starts as soon as the backward pass has completed it.
        self.grad_sync = {}
        self.parallel_streams = False
        self.wgrad_overlap = False  # with parallel_streams: weight gradients on companion streams (ops.wgrad_async)
        # Side streams per stream *context*: 0 = the step's main line, 1 = the early generator pass of gen_update that
        # engine.StepRunner overlaps with the discriminator update (_early_generator_forward).
        self._side = {}
        self._side_used = {}
        self._ctx = 0
        self._gstream = None
        self._early = None
        self.overlap_updates = False
        self.style_stream = True  # shared style encoder on a third stream (engine.StepRunner turns it off under DP)
        self.early_used = 0  # gen_update calls that picked up an early generator pass

    def _fork_join(self, fa, *fbs):
        """fa() on the current stream, every other function on its own side stream, then join.  Fork/join discipline
        keeps the caching allocator safe: every side-stream region starts after all earlier work of the main stream."""
        if not self.parallel_streams:
            return (fa(),) + tuple(f() for f in fbs)
        cur = torch.cuda.current_stream()
        pool = self._side.setdefault(self._ctx, [])
        while len(pool) < len(fbs):
            pool.append(torch.cuda.Stream())
        sides = pool[:len(fbs)]
        self._side_used[self._ctx] = max(self._side_used.get(self._ctx, 0), len(sides))
        for side in sides:
            side.wait_stream(cur)
        res = [fa()]
        for side, f in zip(sides, fbs):
            with torch.cuda.stream(side):
                res.append(f())
        for side in sides:
            cur.wait_stream(side)
        return tuple(res)

    def _join_side(self, ctxs=(0,)):
        if self.parallel_streams:
            # only the streams forked since the last join (backward nodes run on their forward streams); a stream
            # that took no part in the current graph capture must not be waited on
            cur = torch.cuda.current_stream()
            for ctx in ctxs:
                if ctx == 1 and self._gstream is not None:
                    cur.wait_stream(self._gstream)
                for side in self._side.get(ctx, [])[:self._side_used.get(ctx, 0)]:
                    cur.wait_stream(side)
                self._side_used[ctx] = 0
        ops.join_side_streams()
        ops.wgrad_join()

    # ------------------------------------------------------------------ gen_update's forward under the dis update
    # The generator pass of gen_update (encode, decode, encode again: trainer.py:400-419) reads the generator
    # weights, the images and its own style draws -- nothing dis_update writes.  Under the step runner it is issued
    # on its own stream (with its own side streams) right after dis_update's no-grad generator pass, so that it fills
    # the SMs the discriminator's small deep-scale launches leave idle; gen_update joins it before it needs the
    # *updated* discriminator for the adversarial loss, so every result is what the sequential order produces.
    def _early_key(self, x_a, x_b, s_a, s_b):
        return (x_a.data_ptr(), x_a._version, x_b.data_ptr(), x_b._version, s_a.data_ptr(), s_a._version,
                s_b.data_ptr(), s_b._version, self.gen_opt.step_count)

    def _early_generator_forward(self, x_a, x_b, s_a, s_b):
        cur = torch.cuda.current_stream()
        if self._gstream is None:
            self._gstream = torch.cuda.Stream()
        g = self._gstream
        g.wait_stream(cur)
        self._ctx = 1
        try:
            with torch.cuda.stream(g):
                self.gen_opt.zero_grad()
                gsync = self.grad_sync.get("gen")
                if gsync is not None:
                    gsync.begin()
                out = self._gen_forward_batched(x_a, x_b, s_a, s_b, None)
        finally:
            self._ctx = 0
        self._early = (self._early_key(x_a, x_b, s_a, s_b), out)

    # ------------------------------------------------------------------ device
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self.s_a = fn(self.s_a)
        self.s_b = fn(self.s_b)
        return out

    # ------------------------------------------------------------------ optimiser steps (trainer.py:252-268)
    def dis_opt_step(self):
        if "extra" in self.hyperparameters["optimizer"] and (self.iterations % 2 == 0):
            self.dis_opt.extrapolation()
        else:
            self.dis_opt.step()

    def gen_opt_step(self):
        if "extra" in self.hyperparameters["optimizer"] and (self.iterations % 2 == 0):
            self.gen_opt.extrapolation()
        else:
            self.gen_opt.step()

    def classif_opt_sr_step(self):
        """trainer.py:243-250."""
        if "extra" in self.hyperparameters["optimizer"] and (self.iterations % 2 == 0):
            self.classif_opt_sr.extrapolation()
        else:
            self.classif_opt_sr.step()

    # ------------------------------------------------------------------ losses (trainer.py:279-305)
    def recon_criterion(self, input, target):
        """mean |input - target|; accepts NCHW fp32 tensors or Act pairs (content codes)."""
        if isinstance(input, Act):
            return _scalar(ops.L1Fn.apply(input.t, target.t, input.pad, target.pad))
        return _scalar(ops.L1Fn.apply(input.float(), target.float(), 0, 0))

    def recon_criterion_mask(self, input, target, mask):
        """mean |(input - target) * (1 - mask)| over all elements (trainer.py:292-305)."""
        n, _, h, w = input.shape
        keep = (1 - mask.to(input.device, torch.float32)).reshape(n, 1, h, w)
        return _scalar(ops.L1MaskedFn.apply(input.float(), target.float(), keep))

    # ------------------------------------------------------------------ generator plumbing
    def _enc(self, which, x):
        if self.gen_state == 0:
            return (self.gen_a if which == "a" else self.gen_b).encode_act(x)
        return self.gen.encode_act(x, 1 if which == "a" else 2)

    def _dec(self, which, c, s):
        if self.gen_state == 0:
            return (self.gen_a if which == "a" else self.gen_b).decode(c, s)
        return self.gen.decode(c, s, 1 if which == "a" else 2)

    def _gen_forward_stage12(self, x_a, x_b, s_a, s_b):
        """The first two stages of gen_update (trainer.py:400-412) for the shared-style generator -- encode both
        images, decode within and across domains -- with calls that use the SAME weights on independent inputs
        merged into one batch: the style encoder sees [x_a; x_b], each decoder sees its within-domain and
        cross-domain pair together.  Every norm is per sample, so the results are those of the reference's
        one-call-per-tensor sequence."""
        g = self.gen
        b = x_a.shape[0]

        def cat_act(u, v):
            return Act(torch.cat([u.t, v.t], 0), u.pad)

        xab_cat = torch.cat([x_a, x_b], 0)
        if self.style_stream:  # the shared style encoder is independent of both content encoders: third stream
            c_a, c_b, s_both = self._fork_join(lambda: g.enc1_content.forward_act(x_a, 1),
                                               lambda: g.enc2_content.forward_act(x_b, 1),
                                               lambda: g.enc_style(xab_cat))
        else:
            (c_a, s_both), c_b = self._fork_join(lambda: (g.enc1_content.forward_act(x_a, 1), g.enc_style(xab_cat)),
                                                 lambda: g.enc2_content.forward_act(x_b, 1))
        s_a_prime, s_b_prime = s_both[:b], s_both[b:]
        sty_a = s_a if self.guided == 0 else s_a_prime   # style used for the cross-domain decode into a
        sty_b = s_b if self.guided == 0 else s_b_prime
        in1, st1 = cat_act(c_a, c_b), torch.cat([s_a_prime, sty_a], 0)
        in2, st2 = cat_act(c_b, c_a), torch.cat([s_b_prime, sty_b], 0)
        out1, out2 = self._fork_join(lambda: g.decode(in1, st1, 1),    # [x_a_recon; x_ba]
                                     lambda: g.decode(in2, st2, 2))    # [x_b_recon; x_ab]
        x_a_recon, x_ba = out1[:b], out1[b:]
        x_b_recon, x_ab = out2[:b], out2[b:]
        return c_a, s_a_prime, c_b, s_b_prime, x_a_recon, x_b_recon, x_ba, x_ab

    def _gen_forward_batched(self, x_a, x_b, s_a, s_b, stage12=None):
        """Stages 1-3 of gen_update (trainer.py:400-419); `stage12`: the first two stages if they were already
        computed on the same inputs and weights (see _dis_backward)."""
        g = self.gen
        b = x_a.shape[0]
        if stage12 is None:
            stage12 = self._gen_forward_stage12(x_a, x_b, s_a, s_b)
        c_a, s_a_prime, c_b, s_b_prime, x_a_recon, x_b_recon, x_ba, x_ab = stage12
        xre_cat = torch.cat([x_ba, x_ab], 0)
        if self.style_stream:
            c_b_recon, c_a_recon, s_re = self._fork_join(lambda: g.enc1_content.forward_act(x_ba, 1),
                                                         lambda: g.enc2_content.forward_act(x_ab, 1),
                                                         lambda: g.enc_style(xre_cat))
        else:
            (c_b_recon, s_re), c_a_recon = self._fork_join(
                lambda: (g.enc1_content.forward_act(x_ba, 1), g.enc_style(xre_cat)),
                lambda: g.enc2_content.forward_act(x_ab, 1))
        s_a_recon, s_b_recon = s_re[:b], s_re[b:]
        return (c_a, s_a_prime, c_b, s_b_prime, x_a_recon, x_b_recon, x_ba, x_ab, c_b_recon, s_a_recon, c_a_recon,
                s_b_recon)

    # ------------------------------------------------------------------ forward reuse between the two updates
    # With guided == 1 the generator pass of dis_update (encode x_a, x_b; decode across domains with the encoded
    # styles, trainer.py:1149-1165) computes exactly what the first two stages of the following gen_update compute
    # again on the same batch with the same weights (trainer.py:400-412) -- the discriminator step in between does
    # not touch the generator.  With `reuse_forward` set, dis_update runs those stages once *with* autograd (the
    # reference also records them and then detaches, trainer.py:1178-1179) and keeps them; gen_update picks them
    # up if, and only if, it is called on the same tensors (object and version) before any generator weight
    # changed.  Results are those of the separate passes (tests/test_trainer_gpu.py); 2S+2C+2D+2M = 155 of the
    # step's 1395 GMAC are not computed twice.
    reuse_forward = False

    def _reuse_key(self, x_a, x_b):
        return (x_a.data_ptr(), x_a._version, tuple(x_a.shape), x_b.data_ptr(), x_b._version, tuple(x_b.shape),
                self.gen_opt.step_count, self.gen.enc_style.model[0].conv.weight._version)

    def _can_reuse(self, x_a, x_b):
        return (self.reuse_forward and self.gen_state == 1 and self.guided == 1 and x_a.shape == x_b.shape
                and torch.is_grad_enabled())

    def _style_noise(self, x_a, x_b, s_a, s_b):
        """Host-generator draws in the reference's order (trainer.py:366-367,1146-1147) unless supplied."""
        if s_a is None:
            s_a = torch.randn(x_a.size(0), self.style_dim, 1, 1).to(x_a.device)
            s_b = torch.randn(x_b.size(0), self.style_dim, 1, 1).to(x_b.device)
        return s_a, s_b

    def forward(self, x_a, x_b):
        """Translate x_a -> x_ab and x_b -> x_ba with the fixed display codes (trainer.py:307-334)."""
        self.eval()
        with torch.no_grad():
            c_a, _ = self._enc("a", x_a)
            c_b, _ = self._enc("b", x_b)
            x_ba = self._dec("a", c_b, self.s_a)
            x_ab = self._dec("b", c_a, self.s_b)
        self.train()
        return x_ab, x_ba

    # ------------------------------------------------------------------ updates
    def gen_update(self, x_a, x_b, hyperparameters, mask_a=None, mask_b=None, comet_exp=None, synth=False,
                   semantic_gt_a=None, semantic_gt_b=None, s_a=None, s_b=None):
        """One generator update (trainer.py:336-561)."""
        self._gen_backward(x_a, x_b, hyperparameters, mask_a, mask_b, synth, s_a, s_b)
        self.gen_opt_step()
        if comet_exp is not None and self.iterations % 100 == 0:
            for k in ("loss_gen_adv_a", "loss_gen_adv_b", "loss_gen_recon_x_a", "loss_gen_recon_s_a",
                      "loss_gen_recon_c_a", "loss_gen_recon_x_b", "loss_gen_recon_s_b", "loss_gen_recon_c_b",
                      "loss_gen_cycrecon_x_a", "loss_gen_cycrecon_x_b", "loss_gen_total"):
                comet_exp.log_metric(k, getattr(self, k).cpu().detach())

    def _gen_backward(self, x_a, x_b, hyperparameters, mask_a=None, mask_b=None, synth=False, s_a=None, s_b=None):
        with _nvtx("munit gen_update"):
            return self._gen_backward_impl(x_a, x_b, hyperparameters, mask_a, mask_b, synth, s_a, s_b)

    def _gen_backward_impl(self, x_a, x_b, hyperparameters, mask_a=None, mask_b=None, synth=False, s_a=None, s_b=None):
        """Losses + gradients of gen_update (everything up to, not including, the optimiser step)."""
        cache, self._fwd_cache = getattr(self, "_fwd_cache", None), None
        stage12 = None
        if cache is not None and self._can_reuse(x_a, x_b) and cache[0] == self._reuse_key(x_a, x_b):
            stage12 = cache[1]  # gradients and the gradient-sync pass were opened by _dis_backward
        del cache
        gsync = self.grad_sync.get("gen")
        early, self._early = self._early, None
        if early is not None and (s_a is None or early[0] != self._early_key(x_a, x_b, s_a, s_b)):
            # not the step the early pass was issued for: join and drop it, then run the normal order
            torch.cuda.current_stream().wait_stream(self._gstream)
            early = None
        join = (0,)
        if early is not None:
            torch.cuda.current_stream().wait_stream(self._gstream)  # gradients / sync pass were opened by the early pass
            join = (0, 1)  # its backward nodes run on the early streams
            self.early_used += 1
        elif stage12 is None:
            self.gen_opt.zero_grad()
            if gsync is not None:
                gsync.begin()
        ops.WG.enabled = bool(self.parallel_streams and self.wgrad_overlap)
        s_a, s_b = self._style_noise(x_a, x_b, s_a, s_b)
        cyc = hyperparameters["recon_x_cyc_w"] > 0
        if early is not None:
            (c_a, s_a_prime, c_b, s_b_prime, x_a_recon, x_b_recon, x_ba, x_ab, c_b_recon, s_a_recon, c_a_recon,
             s_b_recon) = early[1]
            del early
        elif self.gen_state == 1 and x_a.shape == x_b.shape:
            (c_a, s_a_prime, c_b, s_b_prime, x_a_recon, x_b_recon, x_ba, x_ab, c_b_recon, s_a_recon, c_a_recon,
             s_b_recon) = self._gen_forward_batched(x_a, x_b, s_a, s_b, stage12)
            del stage12
        else:
            # encode
            c_a, s_a_prime = self._enc("a", x_a)
            c_b, s_b_prime = self._enc("b", x_b)
            # decode (within domain)
            x_a_recon = self._dec("a", c_a, s_a_prime)
            x_b_recon = self._dec("b", c_b, s_b_prime)
            # decode (cross domain)
            if self.guided == 0:
                x_ba = self._dec("a", c_b, s_a)
                x_ab = self._dec("b", c_a, s_b)
            elif self.guided == 1:
                x_ba = self._dec("a", c_b, s_a_prime)
                x_ab = self._dec("b", c_a, s_b_prime)
            else:
                print("self.guided unknown value:", self.guided)
            # encode again
            c_b_recon, s_a_recon = self._enc("a", x_ba)
            c_a_recon, s_b_recon = self._enc("b", x_ab)
        # decode again (if needed)
        if cyc:
            x_aba, x_bab = self._fork_join(lambda: self._dec("a", c_a_recon, s_a_prime),
                                           lambda: self._dec("b", c_b_recon, s_b_prime))
        else:
            x_aba = x_bab = None

        # reconstruction loss
        self.loss_gen_recon_x_a = self.recon_criterion(x_a_recon, x_a)
        self.loss_gen_recon_x_b = self.recon_criterion(x_b_recon, x_b)
        if self.guided == 0:
            self.loss_gen_recon_s_a = self.recon_criterion(s_a_recon, s_a)
            self.loss_gen_recon_s_b = self.recon_criterion(s_b_recon, s_b)
        else:
            self.loss_gen_recon_s_a = self.recon_criterion(s_a_recon, s_a_prime)
            self.loss_gen_recon_s_b = self.recon_criterion(s_b_recon, s_b_prime)
        self.loss_gen_recon_c_a = self.recon_criterion(c_a_recon, c_a)
        self.loss_gen_recon_c_b = self.recon_criterion(c_b_recon, c_b)
        # synthetic-pair reconstruction (trainer.py:452-464): pixels identical in the pair must stay aligned
        self.loss_gen_recon_synth = 0
        if synth:
            mask_alignment = (torch.sum(torch.abs(x_a - x_b), 1) == 0).unsqueeze(1).to(torch.float32)
            self.loss_gen_recon_synth = (self.recon_criterion_mask(x_ab, x_b, 1 - mask_alignment)
                                         + self.recon_criterion_mask(x_ba, x_a, 1 - mask_alignment))
        if self.recon_mask:
            self.loss_gen_cycrecon_x_a = self.recon_criterion_mask(x_aba, x_a, mask_a) if cyc else 0
            self.loss_gen_cycrecon_x_b = self.recon_criterion_mask(x_bab, x_b, mask_b) if cyc else 0
        else:
            self.loss_gen_cycrecon_x_a = self.recon_criterion(x_aba, x_a) if cyc else 0
            self.loss_gen_cycrecon_x_b = self.recon_criterion(x_bab, x_b) if cyc else 0
        # GAN loss (discriminator weights frozen: dgrad only)
        self.loss_gen_adv_a, self.loss_gen_adv_b = self._fork_join(
            lambda: self.dis_a.calc_gen_loss(x_ba, frozen=True), lambda: self.dis_b.calc_gen_loss(x_ab, frozen=True))
        self.loss_gen_vgg_a = 0
        self.loss_gen_vgg_b = 0
        self.loss_sem_seg = 0
        self.domain_adv_loss = 0
        # adaptation loss on the content features: fool the synthetic / real classifiers (trainer.py:521-525)
        adv_lambda = hyperparameters["adaptation"].get("adv_lambda", 0)
        self.loss_classifier_sr = (self.compute_classifier_sr_loss(c_a, c_b, domain_synth=synth, fool=True, frozen=True)
                                   if adv_lambda > 0 else 0)
        self.loss_output_classifier_sr = 0
        # total loss (trainer.py:539-558)
        self.loss_gen_total = (
            hyperparameters["gan_w"] * self.loss_gen_adv_a
            + hyperparameters["gan_w"] * self.loss_gen_adv_b
            + hyperparameters["recon_x_w"] * self.loss_gen_recon_x_a
            + hyperparameters["recon_s_w"] * self.loss_gen_recon_s_a
            + hyperparameters["recon_c_w"] * self.loss_gen_recon_c_a
            + hyperparameters["recon_x_w"] * self.loss_gen_recon_x_b
            + hyperparameters["recon_s_w"] * self.loss_gen_recon_s_b
            + hyperparameters["recon_c_w"] * self.loss_gen_recon_c_b
            + hyperparameters["recon_x_cyc_w"] * self.loss_gen_cycrecon_x_a
            + hyperparameters["recon_x_cyc_w"] * self.loss_gen_cycrecon_x_b
            + hyperparameters["recon_synth_w"] * self.loss_gen_recon_synth
        )
        if adv_lambda > 0:
            self.loss_gen_total = self.loss_gen_total + adv_lambda * self.loss_classifier_sr
        self.loss_gen_total.backward()
        self._join_side(join)  # backward nodes ran on their forward streams; the optimiser step waits for all of them
        if gsync is not None:
            gsync.finish()  # reduce the buckets that are still local, wait for the ones already in flight
        self._last = dict(x_ab=x_ab.detach(), x_ba=x_ba.detach())
        self._release_graph()

    # ------------------------------------------------------------------ adaptation head (trainer.py:638-667,1237-1265)
    def compute_classifier_sr_loss(self, c_a, c_b, domain_synth=False, fool=False, frozen=False):
        """mean((D_a(c_a) - t)^2) + mean((D_b(c_b) - t)^2) with t = 0.5 (fool), 0 (synthetic) or 1 (real);
        c_a / c_b: content codes (ops.Act or NCHW fp32).  frozen: no classifier weight gradients."""
        target = 0.5 if fool else (0.0 if domain_synth else 1.0)
        out_a, out_b = self._fork_join(lambda: self.domain_classifier_sr_a(c_a, frozen=frozen),
                                       lambda: self.domain_classifier_sr_b(c_b, frozen=frozen))
        return _scalar(ops.MseConstFn.apply(out_a.float(), target)) + _scalar(ops.MseConstFn.apply(out_b.float(), target))

    def domain_classifier_sr_update(self, x_a, x_b, domain_synth, lambda_classifier, step=None, comet_exp=None):
        """One update of the two content-feature classifiers on detached content codes (trainer.py:1237-1265)."""
        self.classif_opt_sr.zero_grad()
        ops.WG.enabled = False
        with torch.no_grad():
            c_a, _ = self._enc("a", x_a)
            c_b, _ = self._enc("b", x_b)
        loss = self.compute_classifier_sr_loss(c_a, c_b, domain_synth, fool=False)
        loss = lambda_classifier * loss
        loss.backward()
        self._join_side()
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size()
            if world > 1:  # data parallel: sum the per-rank gradients, 1/world fused into the update
                from . import dp
                dp.allreduce_arena(self.classif_opt_sr.g_arena)
                self.classif_opt_sr.grad_scale = 1.0 / world
        self.classif_opt_sr_step()
        self.loss_classifier_sr_update = loss.detach()
        if comet_exp is not None and self.iterations % 100 == 0:
            comet_exp.log_metric("loss_classifier_sr", loss.cpu().detach(), step=step)

    def dis_update(self, x_a, x_b, hyperparameters, comet_exp=None, s_a=None, s_b=None):
        """One discriminator update (trainer.py:1133-1186)."""
        self._dis_backward(x_a, x_b, hyperparameters, s_a, s_b)
        self.dis_opt_step()
        if comet_exp is not None and self.iterations % 100 == 0:
            comet_exp.log_metric("loss_dis_b", self.loss_dis_b.cpu().detach())
            comet_exp.log_metric("loss_dis_a", self.loss_dis_a.cpu().detach())

    def _dis_backward(self, x_a, x_b, hyperparameters, s_a=None, s_b=None, early_gen=None):
        with _nvtx("munit dis_update"):
            return self._dis_backward_impl(x_a, x_b, hyperparameters, s_a, s_b, early_gen)

    def _dis_backward_impl(self, x_a, x_b, hyperparameters, s_a=None, s_b=None, early_gen=None):
        """Losses + gradients of dis_update (everything up to, not including, the optimiser step).  early_gen: the
        (s_a, s_b) style draws of the gen_update that follows on the same batch -- its generator pass is then issued
        here, overlapped with the discriminator work (_early_generator_forward)."""
        self.dis_opt.zero_grad()
        dsync = self.grad_sync.get("dis")
        if dsync is not None:
            dsync.begin()
        ops.WG.enabled = bool(self.parallel_streams and self.wgrad_overlap)
        s_a, s_b = self._style_noise(x_a, x_b, s_a, s_b)
        self._fwd_cache = None
        if self._can_reuse(x_a, x_b):
            # generator pass shared with the gen_update that follows (see reuse_forward above)
            self.gen_opt.zero_grad()
            gsync = self.grad_sync.get("gen")
            if gsync is not None:
                gsync.begin()
            stage12 = self._gen_forward_stage12(x_a, x_b, s_a, s_b)
            x_ba, x_ab = stage12[6], stage12[7]
            self._fwd_cache = (self._reuse_key(x_a, x_b), stage12)
            del stage12
        else:
            with torch.no_grad():
                (c_a, s_a_prime), (c_b, s_b_prime) = self._fork_join(lambda: self._enc("a", x_a),
                                                                     lambda: self._enc("b", x_b))
                if self.guided == 0:
                    sty_a, sty_b = s_a, s_b
                elif self.guided == 1:
                    sty_a, sty_b = s_a_prime, s_b_prime
                else:
                    print("self.guided unknown value:", self.guided)
                x_ba, x_ab = self._fork_join(lambda: self._dec("a", c_b, sty_a), lambda: self._dec("b", c_a, sty_b))
            self._early = None
            if (early_gen is not None and self.overlap_updates and self.parallel_streams and self.gen_state == 1
                    and x_a.shape == x_b.shape):
                self._early_generator_forward(x_a, x_b, *early_gen)
        # D loss
        self.loss_dis_a, self.loss_dis_b = self._fork_join(lambda: self.dis_a.calc_dis_loss(x_ba.detach(), x_a),
                                                           lambda: self.dis_b.calc_dis_loss(x_ab.detach(), x_b))
        self.loss_dis_total = hyperparameters["gan_w"] * self.loss_dis_a + hyperparameters["gan_w"] * self.loss_dis_b
        self.loss_dis_total.backward()
        self._join_side()
        if dsync is not None:
            dsync.finish()
        self._release_graph()

    def _release_graph(self):
        """Drop references into the autograd graph of the finished update: the stored loss_* become
        detached 0-dim tensors and the AdaIN layers forget their (graph-attached) parameter views."""
        for k, v in list(self.__dict__.items()):
            if k.startswith("loss_") and torch.is_tensor(v) and v.grad_fn is not None:
                setattr(self, k, v.detach())
        for m in self.modules():
            if m.__class__.__name__ == "AdaptiveInstanceNorm2d" and m._w is not None:
                m._w, m._b = m._w.detach(), m._b.detach()

    # ------------------------------------------------------------------ sampling (trainer.py:773-928,1087-1131)
    def sample(self, x_a, x_b):
        """(x_a, x_a_recon, x_ab1, x_ab2, x_b, x_b_recon, x_ba1, x_ba2).  Batched: every norm is
        per-sample, so this equals the reference's batch-1 loop."""
        self.eval()
        n = x_a.size(0)
        s_a1, s_b1 = self.s_a[:n], self.s_b[:n]
        s_a2 = torch.randn(x_a.size(0), self.style_dim, 1, 1).to(x_a.device)
        s_b2 = torch.randn(x_b.size(0), self.style_dim, 1, 1).to(x_b.device)
        with torch.no_grad():
            c_a, s_a_fake = self._enc("a", x_a)
            c_b, s_b_fake = self._enc("b", x_b)
            x_a_recon = self._dec("a", c_a, s_a_fake)
            x_b_recon = self._dec("b", c_b, s_b_fake)
            if self.guided == 0:
                x_ba1, x_ba2 = self._dec("a", c_b, s_a1), self._dec("a", c_b, s_a2)
                x_ab1, x_ab2 = self._dec("b", c_a, s_b1), self._dec("b", c_a, s_b2)
            else:
                x_ba1 = x_ba2 = self._dec("a", c_b, s_a_fake)
                x_ab1 = x_ab2 = self._dec("b", c_a, s_b_fake)
        self.train()
        return x_a, x_a_recon, x_ab1, x_ab2, x_b, x_b_recon, x_ba1, x_ba2

    def sample_fid(self, x_a, x_b):
        self.eval()
        with torch.no_grad():
            c_a, _ = self._enc("a", x_a)
            _, s_b_fake = self._enc("b", x_b)
            if self.guided == 1:
                x_ab1 = self._dec("b", c_a, s_b_fake)
            else:
                print("self.guided unknown value:", self.guided)
                x_ab1 = None
        self.train()
        return x_ab1

    def update_learning_rate(self):
        """trainer.py:1326-1335."""
        if self.dis_scheduler is not None:
            self.dis_scheduler.step()
        if self.gen_scheduler is not None:
            self.gen_scheduler.step()

    # ------------------------------------------------------------------ checkpoints (trainer.py:1337-1429)
    def resume(self, checkpoint_dir, hyperparameters):
        last_model_name = get_model_list(checkpoint_dir, "gen")
        state_dict = torch.load(last_model_name, map_location="cpu")
        if self.gen_state == 0:
            self.gen_a.load_state_dict(state_dict["a"])
            self.gen_b.load_state_dict(state_dict["b"])
        else:
            self.gen.load_state_dict(state_dict["2"])
        iterations = int(last_model_name[-11:-3])
        last_model_name = get_model_list(checkpoint_dir, "dis")
        state_dict = torch.load(last_model_name, map_location="cpu")
        self.dis_a.load_state_dict(state_dict["a"])
        self.dis_b.load_state_dict(state_dict["b"])
        state_dict = torch.load(os.path.join(checkpoint_dir, "optimizer.pt"), map_location="cpu")
        self.dis_opt.load_state_dict(state_dict["dis"])
        self.gen_opt.load_state_dict(state_dict["gen"])
        self.dis_scheduler = get_scheduler(self.dis_opt, hyperparameters, iterations)
        self.gen_scheduler = get_scheduler(self.gen_opt, hyperparameters, iterations)
        print("Resume from iteration %d" % iterations)
        return iterations

    def save(self, snapshot_dir, iterations):
        gen_name = os.path.join(snapshot_dir, "gen_%08d.pt" % (iterations + 1))
        dis_name = os.path.join(snapshot_dir, "dis_%08d.pt" % (iterations + 1))
        opt_name = os.path.join(snapshot_dir, "optimizer.pt")
        if self.gen_state == 0:
            torch.save({"a": self.gen_a.state_dict(), "b": self.gen_b.state_dict()}, gen_name)
        else:
            torch.save({"2": self.gen.state_dict()}, gen_name)
        torch.save({"a": self.dis_a.state_dict(), "b": self.dis_b.state_dict()}, dis_name)
        torch.save({"gen": self.gen_opt.state_dict(), "dis": self.dis_opt.state_dict()}, opt_name)
