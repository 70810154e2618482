"""Drop-in modules (munit_b200/networks.py) on the B200 against the CPU oracle / golden fixtures.
bf16 tensor-core path: rel-L2 <= 1e-2 per layer (BASELINE.json north_star); whole networks stack ~20
bf16 layers, so their end-to-end outputs are held to 3e-2 and reported."""
import pytest
import torch

from oracle import munit_oracle as O
from tests.gpu_util import bf16_round, rel_l2

pytestmark = pytest.mark.gpu

LAYER_TOL = 1e-2
NET_TOL = 3e-2


def _images(seed, b, hw):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(b, 3, hw, hw, generator=g) * 2 - 1, torch.rand(b, 3, hw, hw, generator=g) * 2 - 1


BLOCKS = [
    # cin, cout, k, s, p, norm, act, n, hw
    (3, 64, 7, 1, 3, "none", "relu", 2, 32),
    (3, 64, 7, 1, 3, "in", "relu", 2, 32),
    (3, 64, 4, 2, 1, "none", "lrelu", 2, 32),
    (64, 128, 4, 2, 1, "in", "relu", 2, 32),
    (64, 128, 4, 2, 1, "none", "lrelu", 3, 16),
    (256, 256, 3, 1, 1, "in", "relu", 2, 16),
    (256, 256, 3, 1, 1, "in", "none", 1, 16),
    (128, 64, 5, 1, 2, "ln", "relu", 2, 24),
    (64, 3, 7, 1, 3, "none", "tanh", 2, 32),
]


def _run_block_oracle(sd0, x, gy, s, p, norm, act, image):
    sd = {k_: v.detach().clone().requires_grad_(True) for k_, v in sd0.items()}
    xr = x.clone().requires_grad_(True)
    y = O.conv_block(sd, "", xr, s, p, norm, act, image=image)
    if gy is None:
        gy = bf16_round(torch.randn_like(y))
    y.backward(gy)
    return dict(y=y.detach(), dx=xr.grad, dw=sd["conv.weight"].grad, db=sd["conv.bias"].grad,
                dgamma=sd["norm.gamma"].grad if "norm.gamma" in sd else None,
                dbeta=sd["norm.beta"].grad if "norm.beta" in sd else None), gy


def _run_block_gpu(blk, x, gy, norm):
    blk = blk.cuda()
    xg = x.cuda().requires_grad_(True)
    y = blk(xg)
    y.backward(gy.cuda())
    out = dict(y=y.detach().cpu(), dx=xg.grad.cpu(), dw=blk.conv.weight.grad.cpu())
    if norm in ("none", "ln"):
        out["db"] = blk.conv.bias.grad.cpu()
    if norm == "ln":
        out["dgamma"], out["dbeta"] = blk.norm.gamma.grad.cpu(), blk.norm.beta.grad.cpu()
    return out


@pytest.mark.parametrize("cin,cout,k,s,p,norm,act,n,hw", BLOCKS)
def test_conv2dblock_forward_backward(cin, cout, k, s, p, norm, act, n, hw):
    """Conv2dBlock (networks.py:627-701) forward + backward against the PINNED fp32 oracle (QUANT off) on identical
    weights and inputs, rel-L2 <= 1e-2 for outputs AND gradients (BASELINE.json north_star).  "Identical" for a bf16
    tensor-core path means bf16-representable operands: the weights and the input are rounded to bf16 once, on both
    sides, so that the product terms are exact and only accumulation order and storage precision differ (the effect
    of rounding fp32 master weights to the bf16 operand is measured by test_conv2dblock_fp32_master_weights).  The
    storage-aware (QUANT) figure is printed next to it for information."""
    from munit_b200.networks import Conv2dBlock

    torch.manual_seed(0)
    blk = Conv2dBlock(cin, cout, k, s, p, norm=norm, activation=act, pad_type="reflect")
    torch.nn.init.normal_(blk.conv.bias, 0, 0.1)
    with torch.no_grad():
        blk.conv.weight.copy_(bf16_round(blk.conv.weight))
    sd0 = {k_: v.detach().clone().contiguous() for k_, v in blk.state_dict().items()}
    x = bf16_round(torch.randn(n, cin, hw, hw))
    ref, gy = _run_block_oracle(sd0, x, None, s, p, norm, act, cin < 64)
    O.QUANT = True
    try:
        refq, _ = _run_block_oracle(sd0, x, gy, s, p, norm, act, cin < 64)
    finally:
        O.QUANT = False
    got = _run_block_gpu(blk, x, gy, norm)
    assert got["y"].shape == ref["y"].shape
    errs = {k_: rel_l2(v, ref[k_]) for k_, v in got.items()}
    errs_q = {k_: rel_l2(v, refq[k_]) for k_, v in got.items()}
    print("block", (cin, cout, k, s, norm, act), "vs fp32 oracle", {k_: round(v, 5) for k_, v in errs.items()},
          "| vs storage-aware oracle", {k_: round(v, 5) for k_, v in errs_q.items()})
    assert max(errs.values()) < LAYER_TOL, errs


@pytest.mark.parametrize("cin,cout,k,s,p,norm,act,n,hw", [b for b in BLOCKS if b[5] != "none"])
def test_conv2dblock_fp32_master_weights(cin, cout, k, s, p, norm, act, n, hw):
    """Same blocks with arbitrary fp32 master weights (what training holds).  The tensor cores multiply bf16
    operands, so the conv output differs from the fp32 reference by ~2e-3 per element *before* anything is stored;
    in front of a norm + ReLU that flips the mask of every element within 2e-3 of its channel mean (~1e-3 of them)
    and each flip is a full-size gradient error: rel-L2 ~ sqrt(1e-3) = 3 %.  That floor belongs to bf16 operands, not
    to this implementation: it is measured here by running the fp32 CPU oracle itself with bf16-rounded weights
    (everything else fp32) and the CUDA path must stay within 1.25x of it (outputs stay <= 1e-2)."""
    from munit_b200.networks import Conv2dBlock

    torch.manual_seed(1)
    blk = Conv2dBlock(cin, cout, k, s, p, norm=norm, activation=act, pad_type="reflect")
    sd0 = {k_: v.detach().clone().contiguous() for k_, v in blk.state_dict().items()}
    x = bf16_round(torch.randn(n, cin, hw, hw))
    ref, gy = _run_block_oracle(sd0, x, None, s, p, norm, act, cin < 64)
    sd_r = dict(sd0)
    sd_r["conv.weight"] = bf16_round(sd0["conv.weight"])
    floor, _ = _run_block_oracle(sd_r, x, gy, s, p, norm, act, cin < 64)
    got = _run_block_gpu(blk, x, gy, norm)
    errs = {k_: rel_l2(v, ref[k_]) for k_, v in got.items()}
    fl = {k_: rel_l2(floor[k_], ref[k_]) for k_ in got}
    print("block", (cin, cout, k, s, norm, act), "ours vs fp32", {k_: round(v, 5) for k_, v in errs.items()},
          "| fp32 oracle with bf16-rounded weights vs fp32", {k_: round(v, 5) for k_, v in fl.items()})
    assert errs["y"] < LAYER_TOL, errs
    for k_ in errs:
        assert errs[k_] < max(LAYER_TOL, 1.25 * fl[k_]), (k_, errs[k_], fl[k_])


def test_generator_and_discriminator_vs_golden(golden):
    """AdaINGen_double / AdaINGen / MsImageDis forward on the seeded weights of tests/golden/nets.pt
    (reference outputs) -- also exercises load_state_dict with the reference's keys."""
    from munit_b200.networks import AdaINGen, AdaINGen_double, MsImageDis

    fx = golden("nets.pt")
    cfg = O.config_256_core()
    gsd = O.init_state_dict(O.gen_spec(cfg["gen"], 3, True), fx["seeds"]["gen"], "kaiming")
    dsd = O.init_state_dict(O.dis_spec(cfg["dis"], 3), fx["seeds"]["dis"], "gaussian")
    gen = AdaINGen_double(3, cfg["gen"])
    gen.load_state_dict(gsd)
    dis = MsImageDis(3, cfg["dis"])
    dis.load_state_dict(dsd)
    gen, dis = gen.cuda(), dis.cuda()
    x_a, x_b = _images(fx["seeds"]["img"], fx["b"], fx["hw"])
    x_a, x_b = x_a.cuda(), x_b.cuda()
    with torch.no_grad():
        c_a, s_a = gen.encode(x_a, 1)
        c_b, s_b = gen.encode(x_b, 2)
        x_ab = gen.decode(c_a, s_b, 2)
        x_ba = gen.decode(c_b, s_a, 1)
        ap = gen.get_adain_param(s_b, 2)
        x_ab_ref = fx["x_ab"].cuda()
        d_out = dis(x_ab_ref)
        dl = dis.calc_dis_loss(x_ab_ref, x_b)
        gl = dis.calc_gen_loss(x_ab_ref)
    errs = dict(s_a=rel_l2(s_a.cpu(), fx["s_a"]), s_b=rel_l2(s_b.cpu(), fx["s_b"]),
                x_ab=rel_l2(x_ab.cpu(), fx["x_ab"]), x_ba=rel_l2(x_ba.cpu(), fx["x_ba"]))
    for i, (o, r) in enumerate(zip(d_out, fx["d_out"])):
        assert o.shape == r.shape
        errs[f"d{i}"] = rel_l2(o.cpu(), r)
    print("whole-network rel-L2:", {k: round(v, 5) for k, v in errs.items()})
    assert max(errs.values()) < NET_TOL, errs
    assert c_a.shape == tuple(fx["c_a"]["shape"]) and ap.shape == tuple(fx["adain_params"]["shape"])
    assert abs(float(dl) - fx["dis_loss"]) < 2e-2 * abs(fx["dis_loss"])
    assert abs(float(gl) - fx["gen_loss"]) < 2e-2 * abs(fx["gen_loss"])
    # single generator with sampled style (gen_state 0 / test_batch.py semantics)
    g0 = AdaINGen(3, cfg["gen"])
    g0.load_state_dict(O.init_state_dict(O.gen_spec(cfg["gen"], 3, False), fx["seeds"]["gen0"], "kaiming"))
    g0 = g0.cuda()
    with torch.no_grad():
        c0, s0 = g0.encode(x_a)
        y0 = g0.decode(c0, fx["s_rand"].cuda())
    assert rel_l2(y0.cpu(), fx["y0"]) < NET_TOL and rel_l2(s0.cpu(), fx["s0"]) < NET_TOL


def test_adain_assignment_is_exact():
    """assign_adain_params slicing (networks.py:230-239) must be bit-exact."""
    from munit_b200.networks import AdaINGen

    cfg = O.config_256_core()
    g = AdaINGen(3, cfg["gen"])
    n_ad = g.get_num_adain_params(g.dec)
    assert n_ad == 4096
    p = torch.arange(2 * n_ad, dtype=torch.float32).view(2, n_ad)
    g.assign_adain_params(p, g.dec)
    ref = O.split_adain_params(p, 8, 256)
    mods = [m for m in g.dec.modules() if m.__class__.__name__ == "AdaptiveInstanceNorm2d"]
    assert len(mods) == 8
    for m, (b, w) in zip(mods, ref):
        assert torch.equal(m.bias, b) and torch.equal(m.weight, w)


def test_standalone_norm_modules():
    from munit_b200.networks import AdaptiveInstanceNorm2d, LayerNorm

    x = torch.randn(2, 64, 8, 8)
    ln = LayerNorm(64)
    ref = O.layer_norm_munit(x, ln.gamma.detach(), ln.beta.detach())
    assert rel_l2(ln.cuda()(x.cuda()).cpu(), ref) < LAYER_TOL
    ad = AdaptiveInstanceNorm2d(64).cuda()
    w, b = torch.randn(128), torch.randn(128)
    ad.weight, ad.bias = w.cuda(), b.cuda()
    assert rel_l2(ad(x.cuda()).cpu(), O.instance_norm(x, w, b)) < LAYER_TOL


def test_style_sampled_inference_matches_oracle():
    """test_batch.py semantics (encode once, decode per random style) vs the oracle, batched over images."""
    from munit_b200.networks import AdaINGen

    cfg = O.config_256_core()
    sd_a = O.init_state_dict(O.gen_spec(cfg["gen"], 3, False), 31, "kaiming")
    sd_b = O.init_state_dict(O.gen_spec(cfg["gen"], 3, False), 32, "kaiming")
    ga, gb = AdaINGen(3, cfg["gen"]), AdaINGen(3, cfg["gen"])
    ga.load_state_dict(sd_a)
    gb.load_state_dict(sd_b)
    ga, gb = ga.cuda().eval(), gb.cuda().eval()
    x, _ = _images(9, 3, 64)
    torch.manual_seed(4)
    styles = torch.randn(2, cfg["gen"]["style_dim"], 1, 1)
    oa, ob = O.Gen(sd_a, cfg["gen"], False), O.Gen(sd_b, cfg["gen"], False)
    with torch.no_grad():
        c_ref, _ = oa.encode(x)
        c, _ = ga.encode_act(x.cuda())
        for j in range(2):
            s = styles[j:j + 1].expand(3, -1, -1, -1).contiguous()
            y_ref = ob.decode(c_ref, s)
            y = gb.decode(c, s.cuda())
            # batched decode == per-image decode (all norms are per-sample)
            y1 = gb.decode(ops_act_slice(c, 1), s[:1].cuda())
            assert rel_l2(y.cpu(), y_ref) < NET_TOL, rel_l2(y.cpu(), y_ref)
            assert torch.equal(y[1:2], y1)


def ops_act_slice(a, i):
    from munit_b200.ops import Act

    return Act(a.t[i:i + 1].contiguous(), a.pad)
