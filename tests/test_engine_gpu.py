"""engine.StepRunner: the CUDA-graph replayed step (one stream, two-stream fork/join, and fork/join with the weight
gradients on companion streams) must reproduce
the eager dis_update + gen_update sequence -- same losses and same weights after several steps."""
import pytest
import torch

from oracle import munit_oracle as O

pytestmark = pytest.mark.gpu


def _run(mode, steps=3, optimizer="adam"):
    from munit_b200.engine import StepRunner
    from munit_b200.trainer import MUNIT_Trainer

    cfg = O.config_256_core(optimizer=optimizer)
    torch.manual_seed(0)
    t = MUNIT_Trainer(cfg).cuda()
    g = torch.Generator().manual_seed(3)
    xa = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    xb = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    sd = t.style_dim
    styles = [torch.randn(2, sd, 1, 1, generator=g) for _ in range(4)]
    r = StepRunner(t, cfg, 2, 64, use_graph=mode != "eager", two_streams={"graph2": 1, "graph3": 2}.get(mode, 0))
    r.load_inputs(xa, xb, *styles)
    pre = 2 if optimizer == "extraadam" else 1  # real eager steps before capture (ExtraAdam captures an even/odd pair)
    if mode != "eager":
        r.warmup_and_capture(pre)
    else:
        for _ in range(pre):
            r.step()
    losses = []
    for _ in range(steps):
        r.step()
        ls = r.losses()
        losses.append((float(ls["loss_dis_total"]), float(ls["loss_gen_total"])))
    torch.cuda.synchronize()
    return losses, t.gen_opt.p_arena.clone(), t.dis_opt.p_arena.clone()


@pytest.mark.parametrize("optimizer", ["adam", "extraadam"])
def test_graph_replay_matches_eager(optimizer):
    le, ge, de = _run("eager", optimizer=optimizer)
    for mode in ("graph", "graph2", "graph3"):  # one stream; two branches; two branches + wgrad companion streams
        lg, gg, dg = _run(mode, optimizer=optimizer)
        pre = 2 if optimizer == "extraadam" else 1
        for i, ((a, b), (c, d)) in enumerate(zip(le, lg)):
            # runs drift apart step by step (fp32-atomic order noise amplified by bf16 storage); the
            # compared losses start after `pre` un-compared steps (measured: 0.53 % on the third ExtraAdam step)
            tol = 5e-3 * (1 + 2 * (i + pre - 1))
            assert abs(a - c) <= tol * abs(a) and abs(b - d) <= tol * abs(b), (mode, le, lg)
        # weights: split-K wgrad uses fp32 atomics, so two runs differ by ~1e-7 relative in the gradients; bf16
        # storage amplifies that to percent-level gradient differences within a step or two, so
        # only the Adam step budget is a hard bound here -- the losses above are the functional check.
        lr, n_steps = 1e-4, len(le) + 2
        assert float((gg - ge).abs().max()) <= 2.5 * lr * n_steps, mode
        assert float((dg - de).abs().max()) <= 2.5 * lr * n_steps, mode
