"""Bandwidth kernels (layout, halo, norms fwd/bwd, activations, MLP, pooling, losses, Adam) against the
CPU oracle formulas / torch autograd.  fp32 statistic paths: rel-L2 <= 1e-4 on the fp32 quantities;
bf16 activations add their own 2^-9 rounding, so tensors stored in bf16 are held to 1e-2."""
import pytest
import torch
import torch.nn.functional as F

from oracle import munit_oracle as O
from tests.gpu_util import bf16_round, nchw, nhwc, rel_l2

pytestmark = pytest.mark.gpu


def test_image_to_act_and_back():
    from munit_b200 import kernels as K

    x = torch.randn(2, 3, 12, 10, device="cuda")
    act = K.image_to_act(x, 3, 8)
    ref = nhwc(F.pad(x, (3,) * 4, mode="reflect"))
    assert torch.equal(act[..., :3].float(), bf16_round(ref))
    assert float(act[..., 3:].float().abs().max()) == 0
    back = K.act_to_nchw(act, 3, 3)
    assert torch.equal(back, bf16_round(x))


@pytest.mark.parametrize("k,s,pad,kwp,cp", [(7, 1, 3, 8, 8), (4, 2, 1, 4, 16)])
def test_kwexp_roundtrip(k, s, pad, kwp, cp):
    from munit_b200 import kernels as K

    n, c, h, w = 2, 3, 12, 16
    x = torch.randn(n, c, h, w, device="cuda")
    e = K.image_to_kwexp(x, pad, k, s, kwp, cp)
    xp = F.pad(x, (pad,) * 4, mode="reflect")
    wo = (w + 2 * pad - k) // s + 1
    ref = torch.zeros(n, h + 2 * pad, wo, kwp, cp, device="cuda")
    for kw in range(k):
        ref[:, :, :, kw, :c] = xp[:, :, :, kw: kw + s * (wo - 1) + 1: s].permute(0, 2, 3, 1)
    assert torch.equal(e.view(ref.shape).float(), bf16_round(ref))
    # adjoint: <E(x), dE> == <x, E^T(dE)>
    de = bf16_round(torch.randn_like(ref))
    de[:, :, :, k:, :] = 0
    de[..., c:] = 0
    dx = K.kwexp_to_image_grad(de.to(torch.bfloat16).view(e.shape), n, c, h, w, pad, k, s, kwp, cp)
    xr = x.clone().requires_grad_(True)
    xpr = F.pad(xr, (pad,) * 4, mode="reflect")
    tot = 0
    for kw in range(k):
        tot = tot + (xpr[:, :, :, kw: kw + s * (wo - 1) + 1: s].permute(0, 2, 3, 1) * de[:, :, :, kw, :c]).sum()
    tot.backward()
    assert rel_l2(dx, xr.grad) < 1e-5


def test_nchw_act_roundtrip_and_halo():
    from munit_b200 import kernels as K

    x = torch.randn(2, 64, 9, 7, device="cuda")
    act = K.nchw_to_act(x, 2)
    ref = nhwc(F.pad(x, (2,) * 4, mode="reflect"))
    assert torch.equal(act.float(), bf16_round(ref))
    assert torch.equal(K.act_to_nchw(act, 64, 2), bf16_round(x))


@pytest.mark.parametrize("mode", ["in", "adain", "ln"])
@pytest.mark.parametrize("shape", [(2, 64, 16, 16), (3, 256, 8, 8), (1, 128, 32, 32), (2, 32, 64, 64), (1, 16, 128, 96)])
@pytest.mark.parametrize("relu,res,up", [(True, False, 1), (False, True, 1), (True, False, 2)])
@pytest.mark.parametrize("f16", [0, 1])
def test_norm_forward_backward(mode, shape, relu, res, up, f16):
    """Statistics / finalize / apply (and reduce / finalize / apply backward) against the oracle formulas.  f16=1: the
    raw conv output is stored as fp16 (what ops.ConvFn leaves in front of a norm), f16=0: bf16 (BatchNorm head)."""
    from munit_b200 import kernels as K

    n, c, h, w = shape
    out_pad = 2 if up == 2 else 1
    g = torch.Generator(device="cuda").manual_seed(1)
    y = bf16_round(torch.randn(n, c, h, w, device="cuda", generator=g) * 1.5 + 0.7).requires_grad_(True)
    resid = bf16_round(torch.randn(n, c, h, w, device="cuda", generator=g)).requires_grad_(True) if res else None
    if mode == "adain":
        params = torch.randn(n, 2 * c + 5, device="cuda", generator=g).requires_grad_(True)
        p_b, p_w = params[:, 3:3 + c], params[:, 3 + c:3 + 2 * c]
        ref = O.instance_norm(y, p_w.contiguous().view(-1), p_b.contiguous().view(-1))
        ldw = params.stride(0)
    elif mode == "ln":
        gamma = torch.rand(c, device="cuda", generator=g).requires_grad_(True)
        beta = torch.randn(c, device="cuda", generator=g).requires_grad_(True)
        ref = O.layer_norm_munit(y, gamma, beta)
        p_w, p_b, ldw = gamma, beta, 0
    else:
        ref = O.instance_norm(y)
        p_w = p_b = None
        ldw = 0
    if relu:
        ref = torch.relu(ref)
    if res:
        ref = ref + resid
    if up == 2:
        ref = F.interpolate(ref, scale_factor=2, mode="nearest")
    ref = F.pad(ref, (out_pad,) * 4, mode="reflect")
    # ---- forward through the kernels
    yb = nhwc(y.detach()).to(torch.float16 if f16 else torch.bfloat16)
    rb = K.nchw_to_act(resid.detach(), 1) if res else None
    stats, shift = K.norm_stats(yb)
    coef = K.norm_finalize(stats, shift, mode, p_w.detach() if p_w is not None else None,
                           p_b.detach() if p_b is not None else None, ldw, h * w)
    out = K.norm_apply(yb, coef[2], coef[3], relu, rb, 1, out_pad, up)
    # fp32 statistics path: mean / rinv against the formula
    yf = y.detach()
    if mode == "ln":
        mu = yf.reshape(n, -1).mean(1)
        sd = yf.reshape(n, -1).std(1)
        assert rel_l2(coef[0][:, 0], mu) < 1e-4 and rel_l2(coef[1][:, 0], 1 / (sd + 1e-5)) < 1e-4
    else:
        mu = yf.mean(dim=(2, 3))
        var = yf.var(dim=(2, 3), unbiased=False)
        assert rel_l2(coef[0], mu) < 1e-4 and rel_l2(coef[1], torch.rsqrt(var + 1e-5)) < 1e-4
    assert rel_l2(out, nhwc(ref.detach())) < 1e-2
    # ---- backward
    g_out = bf16_round(torch.randn_like(ref))
    ref.backward(g_out)
    gb = nhwc(g_out).to(torch.bfloat16)
    if mode == "adain":
        gparams = torch.zeros_like(params)
        g_b, g_w, ldg = gparams[:, 3:3 + c], gparams[:, 3 + c:3 + 2 * c], gparams.stride(0)
    elif mode == "ln":
        g_w, g_b, ldg = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda"), 0
    else:
        g_w = g_b = None
        ldg = 0
    if res:  # poison the block the allocator will hand out for g_res: its halo must be zeroed by the kernels
        poison = torch.full((n, h + 2, w + 2, c), float("nan"), dtype=torch.bfloat16, device="cuda")
        del poison
    dy, g_res = K.norm_bwd(gb, out_pad, up, yb, coef, relu, mode, p_w.detach() if p_w is not None else None, ldw,
                           g_w, g_b, ldg, res, 1)
    assert rel_l2(dy, nhwc(y.grad)) < 1e-2
    if res:
        assert rel_l2(g_res[:, 1:-1, 1:-1], nhwc(resid.grad)) < 1e-2
        for edge in (g_res[:, 0], g_res[:, -1], g_res[:, :, 0], g_res[:, :, -1]):
            assert float(edge.float().abs().max()) == 0
    if mode == "adain":
        assert rel_l2(gparams, params.grad) < 1e-3
    if mode == "ln":
        assert rel_l2(g_w, gamma.grad) < 1e-3 and rel_l2(g_b, beta.grad) < 1e-3


@pytest.mark.parametrize("act", ["none", "relu", "lrelu", "tanh"])
def test_act_bwd_and_colsum(act):
    from munit_b200 import kernels as K

    n, c, h, w, pad = 2, 64, 10, 12, 1
    pre = torch.randn(n, c, h, w, device="cuda", requires_grad=True)
    post = O.activation(pre, act)
    outp = F.pad(post, (pad,) * 4, mode="reflect")
    g = bf16_round(torch.randn_like(outp))
    outp.backward(g)
    out_act = nhwc(outp.detach()).to(torch.bfloat16)
    dy = K.act_bwd(nhwc(g).to(torch.bfloat16), out_act, pad, act)
    assert rel_l2(dy, nhwc(pre.grad)) < 1e-2
    db = torch.zeros(c, device="cuda")
    K.colsum(dy, db)
    assert rel_l2(db, dy.float().sum(dim=(0, 1, 2))) < 1e-4
    # ... and the same bias gradient from the act_bwd pass itself (first 48 channels only: c_out < C)
    for c_out, cc in ((c, 64), (48, 64), (256, 256)):
        g2 = torch.randn(n, h + 2 * pad, w + 2 * pad, cc, device="cuda").to(torch.bfloat16)
        o2 = torch.randn_like(g2)
        db2 = torch.full((cc,), 1.0, device="cuda")
        dy2 = K.act_bwd(g2, o2, pad, act, db2, c_out)
        assert torch.equal(dy2, K.act_bwd(g2, o2, pad, act))
        ref = 1.0 + dy2.float().sum(dim=(0, 1, 2))
        assert rel_l2(db2[:c_out], ref[:c_out]) < 1e-4 and float((db2[c_out:] - 1.0).abs().max() if c_out < cc else 0) == 0


def test_gather_cast_multi_matches_per_segment_gathers():
    """munit_gather_cast_multi (all weight shadows of an arena in one launch): segments of ragged sizes, with and
    without an index map, must equal munit_gather_cast / munit_cast_bf16 bit for bit."""
    from munit_b200 import kernels as K

    g = torch.Generator(device="cuda").manual_seed(3)
    arena = torch.randn(50000, device="cuda", generator=g)
    segs, refs = [], []
    off = 0
    for i, (n_src, n_dst) in enumerate([(1000, 4096), (7, 1), (2048, 2048), (5000, 2049), (12345, 30000), (64, 64)]):
        src = arena[off:off + n_src]
        off += n_src
        idx = None if (i % 3 == 2) else torch.randint(-1, n_src, (n_dst,), device="cuda", dtype=torch.int32, generator=g)
        n_dst = n_dst if idx is not None else n_src
        dst = torch.full((n_dst,), 7.0, dtype=torch.bfloat16, device="cuda")
        ref = torch.empty_like(dst)
        if idx is None:
            K.cast_bf16(src, ref)
        else:
            K.gather_cast(src, idx, ref)
        segs.append((src, idx, dst))
        refs.append(ref)
    table, nseg, nblocks = K.gather_seg_table(segs, "cuda")
    assert nseg == 6 and nblocks == sum((d.numel() + K.GATHER_BLOCK - 1) // K.GATHER_BLOCK for _, _, d in segs)
    K.gather_cast_multi(table, nseg, nblocks)
    torch.cuda.synchronize()
    for (_, _, dst), ref in zip(segs, refs):
        assert torch.equal(dst, ref)


def test_gather_cast_add():
    from munit_b200 import kernels as K

    src = torch.randn(1000, device="cuda")
    idx = torch.randint(-1, 1000, (4096,), device="cuda", dtype=torch.int32)
    dst = torch.empty(4096, dtype=torch.bfloat16, device="cuda")
    K.gather_cast(src, idx, dst)
    ref = torch.where(idx >= 0, src[idx.clamp(min=0).long()], torch.zeros((), device="cuda"))
    assert torch.equal(dst.float(), bf16_round(ref))
    acc = torch.ones(4096, device="cuda")
    K.gather_add(src, idx, acc)
    assert torch.allclose(acc, 1 + ref)


def test_linear_fwd_bwd():
    from munit_b200 import kernels as K

    b, i, o = 5, 16, 256
    x = torch.randn(b, i, device="cuda", requires_grad=True)
    w = torch.randn(o, i, device="cuda", requires_grad=True)
    bias = torch.randn(o, device="cuda", requires_grad=True)
    y = torch.relu(F.linear(x, w, bias))
    gy = torch.randn_like(y)
    y.backward(gy)
    yk = K.linear_fwd(x.detach(), w.detach(), bias.detach(), True)
    assert rel_l2(yk, y) < 1e-5
    dw, db = torch.zeros_like(w), torch.zeros_like(bias)
    dx = K.linear_bwd(x.detach(), w.detach(), yk, gy, True, True, dw, db)
    assert rel_l2(dx, x.grad) < 1e-5 and rel_l2(dw, w.grad) < 1e-5 and rel_l2(db, bias.grad) < 1e-5


@pytest.mark.parametrize("b,i,d,o", [(16, 8, 256, 4096), (5, 8, 256, 4096), (40, 16, 64, 100), (1, 8, 256, 2048)])
def test_style_mlp_one_kernel(b, i, d, o):
    """networks.MLP (networks.py:583-597) with the shipped shape runs its forward pass as ONE kernel
    (munit_mlp3_fwd, ops.Mlp3Fn); outputs and all gradients against the same MLP in plain torch fp32."""
    from munit_b200 import _lib
    from munit_b200.networks import MLP

    torch.manual_seed(0)
    mlp = MLP(i, o, d, 3, norm="none", activ="relu").cuda()
    x = torch.randn(b, i, 1, 1, device="cuda", requires_grad=True)
    before = _lib.launches
    y = mlp(x)
    assert _lib.launches - before == 1
    gy = torch.randn_like(y)
    y.backward(gy)
    got = [x.grad.clone()] + [p.grad.clone() for p in mlp.parameters()]
    x.grad = None
    mlp.zero_grad()
    fcs = [blk.fc for blk in mlp.model]
    h = x.reshape(b, -1)
    ref = F.linear(torch.relu(F.linear(torch.relu(F.linear(h, fcs[0].weight, fcs[0].bias)), fcs[1].weight, fcs[1].bias)),
                   fcs[2].weight, fcs[2].bias)
    ref.backward(gy)
    assert rel_l2(y, ref) < 1e-5
    want = [x.grad] + [p.grad for p in mlp.parameters()]
    for g, w in zip(got, want):
        assert rel_l2(g, w) < 1e-4, (g.shape, rel_l2(g, w))


def test_gap_and_dis_head():
    from munit_b200 import kernels as K

    y = bf16_round(torch.randn(3, 6, 5, 256, device="cuda"))
    out = K.gap_fwd(y.to(torch.bfloat16))
    assert rel_l2(out, y.mean(dim=(1, 2))) < 1e-5
    g = torch.randn(3, 256, device="cuda")
    dy = K.gap_bwd(g, (3, 6, 5, 256))
    assert rel_l2(dy, (g / 30)[:, None, None, :].expand(3, 6, 5, 256)) < 1e-2
    # 1x1 conv C->1 + LSGAN
    yv = y.clone().requires_grad_(True)
    w = torch.randn(256, device="cuda", requires_grad=True)
    b = torch.randn(1, device="cuda", requires_grad=True)
    o = (yv * w).sum(-1) + b
    loss_ref = 3.0 * torch.mean((o - 1.0) ** 2)
    loss_ref.backward()
    loss = torch.zeros(1, device="cuda")
    ok = K.dis_head_fwd(y.to(torch.bfloat16), w.detach(), b.detach(), 1.0, loss, 3.0)
    assert rel_l2(ok, o.reshape(-1)) < 1e-5 and abs(float(loss) - float(loss_ref)) < 1e-4 * abs(float(loss_ref))
    dw, db = torch.zeros(256, device="cuda"), torch.zeros(1, device="cuda")
    dyk = K.dis_head_bwd(y.to(torch.bfloat16), w.detach(), ok, 1.0, None, 3.0, dw, db)
    assert rel_l2(dyk, yv.grad) < 1e-2 and rel_l2(dw, w.grad) < 1e-4 and rel_l2(db, b.grad) < 1e-4


@pytest.mark.parametrize("h,w", [(16, 16), (9, 12), (4, 4)])
def test_avgpool(h, w):
    from munit_b200 import kernels as K

    x = torch.randn(2, 3, h, w, device="cuda", requires_grad=True)
    ref = F.avg_pool2d(x, 3, 2, 1, count_include_pad=False)
    gy = torch.randn_like(ref)
    ref.backward(gy)
    y = K.avgpool_fwd(x.detach())
    assert rel_l2(y, ref) < 1e-6
    gx = torch.zeros_like(x)
    K.avgpool_bwd(gy, gx)
    assert rel_l2(gx, x.grad) < 1e-6


def test_l1():
    from munit_b200 import kernels as K

    a = torch.randn(3, 7, 11, device="cuda")
    b = torch.randn(3, 7, 11, device="cuda")
    loss = torch.zeros(1, device="cuda")
    K.l1_fwd(a, b, loss, 2.0 / a.numel())
    assert abs(float(loss) - 2.0 * float((a - b).abs().mean())) < 1e-5
    ga, gb = torch.empty_like(a), torch.empty_like(a)
    gs = torch.tensor([0.5], device="cuda")
    K.l1_bwd(a, b, gs, 2.0 / a.numel(), ga, gb)
    ref = 0.5 * 2.0 / a.numel() * torch.sign(a - b)
    assert torch.allclose(ga, ref) and torch.allclose(gb, -ref)


def test_adam_modes(golden):
    from munit_b200 import kernels as K

    fx = golden("adam.pt")
    for name in ("adam", "extraadam"):
        p = fx["p0"].cuda().clone()
        m, v, saved = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
        pb = torch.empty(p.numel(), dtype=torch.bfloat16, device="cuda")
        for it, g in enumerate(fx["grads"]):
            if name == "adam":
                K.adam(p, g.cuda(), m, v, None, pb, 0, False, 1e-3, 0.5, 0.999, 1e-8, 1e-4, it + 1)
            elif it % 2 == 0:
                K.adam(p, g.cuda(), m, v, saved, pb, 1, True, 1e-3, 0.5, 0.999, 1e-8, 1e-4, it + 1)
            else:
                K.adam(p, g.cuda(), m, v, saved, pb, 2, False, 1e-3, 0.5, 0.999, 1e-8, 1e-4, it + 1)
            assert torch.allclose(p.cpu(), fx[name][it], rtol=1e-5, atol=1e-6), (name, it)
            assert torch.equal(pb.float(), bf16_round(p))


@pytest.mark.parametrize("shape", [(2, 3, 16, 24), (3, 3, 64, 64)])
def test_l1_masked(shape):
    """recon_criterion_mask (trainer.py:292-305): value and both gradients against the oracle expression."""
    from munit_b200 import ops

    n, c, h, w = shape
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(n, c, h, w, device="cuda", generator=g, requires_grad=True)
    b = torch.randn(n, c, h, w, device="cuda", generator=g, requires_grad=True)
    mask = (torch.rand(n, 1, h, w, device="cuda", generator=g) > 0.5).float()
    ref = O.l1_masked(a, b, mask)
    ref.backward()
    ga_ref, gb_ref = a.grad.clone(), b.grad.clone()
    a.grad = b.grad = None
    out = ops.L1MaskedFn.apply(a, b, 1 - mask)
    (out * 1.0).sum().backward()
    assert abs(float(out) - float(ref)) <= 1e-5 * abs(float(ref))
    assert torch.allclose(a.grad, ga_ref, atol=1e-9) and torch.allclose(b.grad, gb_ref, atol=1e-9)


@pytest.mark.parametrize("flip", [0, 1])
def test_u8_crop_normalize_is_bit_exact(flip):
    """GPU tail of the input pipeline == transforms.RandomCrop / hflip / ToTensor / Normalize."""
    from munit_b200 import kernels as K

    g = torch.Generator().manual_seed(4)
    img = torch.randint(0, 256, (37, 53, 3), dtype=torch.uint8, generator=g)
    top, left, ch, cw = 5, 11, 24, 32
    ref = img[top:top + ch, left:left + cw].permute(2, 0, 1).float().div(255)
    if flip:
        ref = ref.flip(2)
    ref = (ref - 0.5) / 0.5
    out = torch.empty(3, ch, cw, device="cuda")
    K.u8_crop_normalize(img.cuda(), top, left, flip, out)
    assert torch.equal(out.cpu(), ref)


def test_gpu_preproc_loader_matches_host_pipeline(tmp_path):
    """data.GpuPreprocLoader (workers: decode + flip + resize, device: crop + normalise) yields the batches the host
    pipeline (data.get_data_loader_folder, the reference's transforms) yields for the same seed."""
    from PIL import Image

    from munit_b200 import data as D

    g = torch.Generator().manual_seed(0)
    for i in range(6):
        arr = torch.randint(0, 256, (40 + 3 * i, 70 - 2 * i, 3), dtype=torch.uint8, generator=g).numpy()
        Image.fromarray(arr).save(tmp_path / f"im{i}.png")
    kw = dict(batch_size=2, train=True, new_size=36, height=32, width=32, num_workers=0, crop=True)
    torch.manual_seed(11)
    host = [b.clone() for b in D.get_data_loader_folder(str(tmp_path), **kw)]
    torch.manual_seed(11)
    dev = [b.cpu() for b in D.get_gpu_data_loader_folder(str(tmp_path), **kw)]
    assert len(host) == len(dev) == 3
    for a, b in zip(host, dev):
        assert torch.equal(a, b)
