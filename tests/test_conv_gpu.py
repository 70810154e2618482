"""tcgen05 tap-GEMM kernels (forward, input gradient, weight gradient) against torch fp32 conv2d on
identical bf16-rounded operands.  Tolerance: rel-L2 <= 1e-2 (bf16 tensor-core path, BASELINE.json)."""
import pytest
import torch
import torch.nn.functional as F

from tests.gpu_util import bf16_round, error_flag, nhwc, rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-2

CASES = [
    # n, h, w, cin, cout, k, s, pad
    (2, 16, 16, 64, 64, 3, 1, 1),
    (1, 16, 32, 128, 256, 3, 1, 1),
    (2, 8, 8, 256, 256, 3, 1, 1),
    (1, 12, 20, 64, 128, 5, 1, 2),
    (2, 16, 16, 64, 128, 4, 2, 1),
    (3, 8, 8, 256, 512, 4, 2, 1),
    (1, 20, 12, 64, 16, 7, 1, 3),
    (5, 4, 4, 512, 64, 1, 1, 0),
    (8, 64, 64, 256, 256, 3, 1, 1),
    (2, 24, 24, 128, 64, 5, 1, 2),  # swapped wgrad orientation (cout < 128 <= cin)
    (2, 16, 16, 256, 64, 4, 2, 1),
]


def _setup(n, h, w, cin, cout, k, s, pad, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = bf16_round(torch.randn(n, cin, h, w, device="cuda", generator=g))
    wt = bf16_round(torch.randn(cout, cin, k, k, device="cuda", generator=g) * (2.0 / (cin * k * k)) ** 0.5)
    bias = torch.randn(cout, device="cuda", generator=g)
    xp = F.pad(x, (pad,) * 4, mode="reflect") if pad else x
    return xp.contiguous(), wt, bias


@pytest.mark.parametrize("n,h,w,cin,cout,k,s,pad", CASES)
@pytest.mark.parametrize("act", ["none", "lrelu"])
def test_tapgemm_forward(n, h, w, cin, cout, k, s, pad, act):
    from munit_b200 import geometry as G, kernels as K

    xp, wt, bias = _setup(n, h, w, cin, cout, k, s, pad)
    y_ref = F.conv2d(xp, wt, bias, stride=s)
    if act == "lrelu":
        y_ref = F.leaky_relu(y_ref, 0.2)
    hp, wp = xp.shape[2:]
    ho, wo = y_ref.shape[2:]
    plan = G.plan_fwd(n, hp, wp, cin, k, k, s, s, cout, (ho * wo * cout, wo * cout, cout, 0, 0))
    a = nhwc(xp).to(torch.bfloat16)
    b = wt.permute(0, 2, 3, 1).reshape(cout, -1).contiguous().to(torch.bfloat16)
    out = torch.zeros(n, ho, wo, cout, dtype=torch.bfloat16, device="cuda")
    K.tapgemm(plan, a, b, out, bias, act)
    torch.cuda.synchronize()
    assert error_flag() == 0, "kernel pipeline timed out"
    err = rel_l2(out, nhwc(y_ref))
    assert err < TOL, err


@pytest.mark.parametrize("n,h,w,cin,cout,k,s,pad", CASES)
def test_tapgemm_dgrad(n, h, w, cin, cout, k, s, pad):
    from munit_b200 import geometry as G, kernels as K

    xp, wt, _ = _setup(n, h, w, cin, cout, k, s, pad, 1)
    xp = xp.requires_grad_(True)
    y = F.conv2d(xp, wt, None, stride=s)
    gy = bf16_round(torch.randn_like(y))
    y.backward(gy)
    hp, wp = xp.shape[2:]
    ck = max(64, cout)
    plan = G.plan_dgrad(n, hp, wp, cin, k, k, s, s, cout)
    idx = G.dgrad_index_map(cout, cin, k, k, s, s, cin, ck).cuda()
    wsrc = wt.permute(0, 2, 3, 1).contiguous().reshape(-1)
    wd = torch.empty(cin, idx.numel() // cin, dtype=torch.bfloat16, device="cuda")
    K.gather_cast(wsrc, idx, wd)
    dxp = torch.full((n, hp, wp, cin), float("nan"), dtype=torch.bfloat16, device="cuda")
    K.tapgemm(plan, nhwc(gy).to(torch.bfloat16), wd, dxp)
    torch.cuda.synchronize()
    assert error_flag() == 0
    err = rel_l2(dxp, nhwc(xp.grad))
    assert err < TOL, err


@pytest.mark.parametrize("n,h,w,cin,cout,k,s,pad", CASES)
def test_wgrad(n, h, w, cin, cout, k, s, pad):
    from munit_b200 import geometry as G, kernels as K

    xp, wt, _ = _setup(n, h, w, cin, cout, k, s, pad, 2)
    wt = wt.requires_grad_(True)
    y = F.conv2d(xp, wt, None, stride=s)
    gy = bf16_round(torch.randn_like(y))
    y.backward(gy)
    hp, wp = xp.shape[2:]
    plan = G.plan_wgrad(n, hp, wp, cin, k, k, s, s, cout, cout, k * k * cin, cin, 1)
    dw = torch.zeros(cout, k, k, cin, dtype=torch.float32, device="cuda")
    K.wgrad(plan, nhwc(gy).to(torch.bfloat16), nhwc(xp).to(torch.bfloat16), dw)
    torch.cuda.synchronize()
    assert error_flag() == 0
    err = rel_l2(dw, wt.grad.permute(0, 2, 3, 1))
    assert err < TOL, err


def test_forward_into_padded_buffer_and_halo():
    from munit_b200 import geometry as G, kernels as K

    n, h, w, c, k, pad, po = 2, 16, 16, 64, 3, 1, 2
    xp, wt, bias = _setup(n, h, w, c, c, k, 1, pad, 3)
    y = F.relu(F.conv2d(xp, wt, bias))
    hop, wop = h + 2 * po, w + 2 * po
    plan = G.plan_fwd(n, h + 2, w + 2, c, k, k, 1, 1, c, (hop * wop * c, wop * c, c, po, po))
    out = torch.zeros(n, hop, wop, c, dtype=torch.bfloat16, device="cuda")
    K.tapgemm(plan, nhwc(xp).to(torch.bfloat16), wt.permute(0, 2, 3, 1).reshape(c, -1).contiguous().to(torch.bfloat16),
              out, bias, "relu")
    K.halo_fill(out, po)
    torch.cuda.synchronize()
    ref = nhwc(F.pad(y, (po,) * 4, mode="reflect"))
    assert rel_l2(out, ref) < TOL


HALO_CASES = [
    (2, 16, 16, 64, 64, 3, 1, 1),
    (8, 64, 64, 256, 256, 3, 1, 1),
    (1, 24, 40, 128, 128, 5, 1, 2),
    (2, 32, 32, 128, 64, 5, 1, 2),
    (1, 20, 28, 64, 64, 7, 1, 3),
]


@pytest.mark.parametrize("n,h,w,cin,cout,k,s,pad", HALO_CASES)
def test_tapgemm_halo_variant(n, h, w, cin, cout, k, s, pad):
    """Halo-resident kernel variant (one activation box per channel chunk, shifted UMMA descriptors per tap):
    forward and input-gradient against torch conv2d."""
    from munit_b200 import geometry as G, kernels as K

    xp, wt, bias = _setup(n, h, w, cin, cout, k, s, pad, 5)
    xp = xp.requires_grad_(True)
    y_ref = F.conv2d(xp, wt, bias, stride=s)
    gy = bf16_round(torch.randn_like(y_ref))
    y_ref.backward(gy)
    hp, wp = xp.shape[2:]
    ho, wo = y_ref.shape[2:]
    plan = G.plan_fwd(n, hp, wp, cin, k, k, s, s, cout, (ho * wo * cout, wo * cout, cout, 0, 0), halo=1)
    assert plan.halo == 1
    a = nhwc(xp.detach()).to(torch.bfloat16)
    b = wt.permute(0, 2, 3, 1).reshape(cout, -1).contiguous().to(torch.bfloat16)
    out = torch.zeros(n, ho, wo, cout, dtype=torch.bfloat16, device="cuda")
    K.tapgemm(plan, a, b, out, bias, "none")
    torch.cuda.synchronize()
    assert error_flag() == 0
    assert rel_l2(out, nhwc(y_ref.detach())) < TOL, rel_l2(out, nhwc(y_ref.detach()))
    dplan = G.plan_dgrad(n, hp, wp, cin, k, k, s, s, cout, halo=1)
    assert dplan.halo == 1
    idx = G.dgrad_index_map(cout, cin, k, k, s, s, cin, max(64, cout)).cuda()
    wd = torch.empty(cin, idx.numel() // cin, dtype=torch.bfloat16, device="cuda")
    K.gather_cast(wt.permute(0, 2, 3, 1).contiguous().reshape(-1), idx, wd)
    dxp = torch.full((n, hp, wp, cin), float("nan"), dtype=torch.bfloat16, device="cuda")
    K.tapgemm(dplan, nhwc(gy).to(torch.bfloat16), wd, dxp)
    torch.cuda.synchronize()
    assert error_flag() == 0
    assert rel_l2(dxp, nhwc(xp.grad)) < TOL, rel_l2(dxp, nhwc(xp.grad))


STATS_CASES = [
    # n, h, w, cin, cout, k, s, pad, halo
    (2, 64, 64, 256, 256, 3, 1, 1, 0),   # residual-block conv: tile (64, 2, 1)
    (3, 32, 32, 64, 128, 4, 2, 1, 0),    # stride-2 down-sampling conv
    (2, 32, 64, 128, 64, 5, 1, 2, 1),    # halo-resident 5x5 (8 x 16 tiles), BN = 64
    (2, 16, 32, 256, 128, 5, 1, 2, 1),   # halo-resident, BN = 128
]


@pytest.mark.parametrize("n,h,w,cin,cout,k,s,pad,halo", STATS_CASES)
@pytest.mark.parametrize("mode", ["in", "adain", "ln"])
def test_epilogue_statistics(n, h, w, cin, cout, k, s, pad, halo, mode):
    """Normalisation partials written by the conv epilogue -> munit_norm_finalize_parts must equal the
    statistics of the stored bf16 output (what the stand-alone statistics pass would have computed)."""
    from munit_b200 import geometry as G, kernels as K

    xp, wt, bias = _setup(n, h, w, cin, cout, k, s, pad, 11)
    bias = bias * 3 + 2  # a mean well away from zero: the unshifted sums must still give the variance
    hp, wp = xp.shape[2:]
    ho, wo = (hp - k) // s + 1, (wp - k) // s + 1
    plan = G.plan_fwd(n, hp, wp, cin, k, k, s, s, cout, (ho * wo * cout, wo * cout, cout, 0, 0), halo=halo)
    assert plan.halo == halo
    kind = 2 if mode == "ln" else 1
    splits = K.stats_splits(plan, kind)
    assert splits > 0
    a = nhwc(xp).to(torch.bfloat16)
    b = wt.permute(0, 2, 3, 1).reshape(cout, -1).contiguous().to(torch.bfloat16)
    out = torch.zeros(n, ho, wo, cout, dtype=torch.bfloat16, device="cuda")
    part = torch.full((n * splits * (cout if kind == 1 else 1) * 2,), float("nan"), device="cuda")
    K.tapgemm(plan, a, b, out, bias, "none", stats=part, stats_kind=kind)
    torch.cuda.synchronize()
    assert error_flag() == 0
    assert rel_l2(out, nhwc(F.conv2d(xp, wt, bias, stride=s))) < TOL
    g = torch.Generator(device="cuda").manual_seed(3)
    p_w = p_b = None
    ldw = 0
    if mode == "adain":
        p_w, p_b, ldw = torch.randn(n, cout, device="cuda", generator=g), torch.randn(n, cout, device="cuda", generator=g), cout
    elif mode == "ln":
        p_w, p_b = torch.rand(cout, device="cuda", generator=g), torch.randn(cout, device="cuda", generator=g)
    coef = K.norm_finalize_parts(part, kind, mode, p_w, p_b, ldw, n, ho * wo, cout)
    stats, shift = K.norm_stats(out)
    ref = K.norm_finalize(stats, shift, mode, p_w, p_b, ldw, ho * wo)
    for i, name in enumerate(("mean", "rinv", "a", "b")):
        assert rel_l2(coef[i], ref[i]) < 2e-5, (name, rel_l2(coef[i], ref[i]))


SPLITK_CASES = [
    # n, h, w, cin, cout, k, s, pad
    (8, 8, 8, 256, 512, 4, 2, 1),    # deep discriminator layer: 1 M tile x 2 N tiles, 64 K blocks
    (8, 16, 16, 256, 512, 4, 2, 1),
    (2, 16, 16, 64, 64, 3, 1, 1),
    (3, 8, 8, 256, 512, 4, 2, 1),
]


@pytest.mark.parametrize("n,h,w,cin,cout,k,s,pad", SPLITK_CASES)
@pytest.mark.parametrize("ksplit", [0, 3, 8])
def test_tapgemm_split_k(n, h, w, cin, cout, k, s, pad, ksplit):
    """Split-K tap-GEMM (fp32 red.add into a scratch + munit_splitk_finish) for launches with too few output
    tiles: forward (bias + LeakyReLU in the finish kernel) and input gradient (stride-2: four phases) must match
    torch conv2d like the unsplit kernel does.  ksplit=0 lets kernels.auto_ksplit decide."""
    from munit_b200 import geometry as G, kernels as K

    xp, wt, bias = _setup(n, h, w, cin, cout, k, s, pad, 7)
    xp = xp.requires_grad_(True)
    y_lin = F.conv2d(xp, wt, bias, stride=s)
    y_ref = F.leaky_relu(y_lin, 0.2)
    gy = bf16_round(torch.randn_like(y_lin))
    y_lin.backward(gy)
    hp, wp = xp.shape[2:]
    ho, wo = y_ref.shape[2:]
    plan = G.plan_fwd(n, hp, wp, cin, k, k, s, s, cout, (ho * wo * cout, wo * cout, cout, 0, 0))
    a = nhwc(xp.detach()).to(torch.bfloat16)
    b = wt.permute(0, 2, 3, 1).reshape(cout, -1).contiguous().to(torch.bfloat16)
    out = torch.full((n, ho, wo, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    launches0 = K._lib.launches
    K.tapgemm(plan, a, b, out, bias, "lrelu", ksplit=ksplit)
    split = K._lib.launches - launches0 == 2
    assert split == (ksplit > 1 or K.auto_ksplit(plan) > 1)
    torch.cuda.synchronize()
    assert error_flag() == 0
    assert rel_l2(out, nhwc(y_ref.detach())) < TOL, rel_l2(out, nhwc(y_ref.detach()))
    dplan = G.plan_dgrad(n, hp, wp, cin, k, k, s, s, cout)
    idx = G.dgrad_index_map(cout, cin, k, k, s, s, cin, max(64, cout)).cuda()
    wd = torch.empty(cin, idx.numel() // cin, dtype=torch.bfloat16, device="cuda")
    K.gather_cast(wt.permute(0, 2, 3, 1).contiguous().reshape(-1), idx, wd)
    dxp = torch.full((n, hp, wp, cin), float("nan"), dtype=torch.bfloat16, device="cuda")
    K.tapgemm(dplan, nhwc(gy).to(torch.bfloat16), wd, dxp, ksplit=min(ksplit, dplan.num_taps * dplan.chunks))
    torch.cuda.synchronize()
    assert error_flag() == 0
    assert rel_l2(dxp, nhwc(xp.grad)) < TOL, rel_l2(dxp, nhwc(xp.grad))
