"""CPU fp32 oracle for the MUNIT hot path -- TEST INFRASTRUCTURE, not product code.

An independent, *functional* restatement (state_dict in, tensors out) of the
reference's algorithm.  Every function cites the reference lines it follows
(paths relative to /root/reference/).  The arithmetic below is plain PyTorch
fp32 on the CPU; the reference's own third-party arithmetic is PyTorch ATen
(pinned torch==0.4.1 in requirements.txt:84, executed here as torch 2.11).

Parity pinning: the reference ships no tests / golden vectors, so
this oracle is pinned against outputs of the reference itself executed in the
build container (oracle/ref_loader.py + oracle/make_golden.py ->
tests/golden/*.pt, checked by tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

EPS_NORM = 1e-5

# ---------------------------------------------------------------------------------------------
# Optional "storage-aware" mode (tests only): QUANT = True makes the oracle round, with straight-through
# gradients, at exactly the points where the B200 path stores bf16 (conv weights, conv outputs before a
# norm, block outputs, images entering a first layer, gradients flowing through those points).  It
# separates kernel errors from the ReLU-mask flips that bf16 storage itself causes; the golden fixtures
# pin the default fp32 mode (QUANT = False), which is the reference's arithmetic.
# ---------------------------------------------------------------------------------------------
QUANT = False


class _RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, round_grad):
        ctx.round_grad = round_grad
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return (g.to(torch.bfloat16).to(g.dtype) if ctx.round_grad else g), None


def _q(t):
    """bf16 storage point of an activation (value and incoming gradient rounded)."""
    return _RoundSTE.apply(t, True) if QUANT else t


def _qw(t):
    """bf16 GEMM operand made from fp32 master data (gradient passes through unrounded)."""
    return _RoundSTE.apply(t, False) if QUANT else t


# --------------------------------------------------------------------------
# default hyper-parameters (configs/config_256.yaml "core": semantic_w=0,
# recon_mask=0, adaptation.* = 0)
# --------------------------------------------------------------------------
def config_256_core(**over) -> dict:
    cfg = dict(
        lr=1e-4, beta1=0.5, beta2=0.999, weight_decay=1e-4, init="kaiming",
        lr_policy="step", step_size=100000, gamma=0.5,
        gan_w=3, recon_x_w=12, recon_s_w=1, recon_c_w=2, recon_x_cyc_w=12, vgg_w=0,
        semantic_w=0, recon_mask=0, domain_adv_w=0, recon_synth_w=0,
        adaptation=dict(full_adaptation=0, output_classifier_lambda=0, output_adv_lambda=0,
                        output_classif_freq=1, adv_lambda=0, dfeat_lambda=0,
                        classif_frequency=15, sem_seg_lambda=0),
        gen_state=1, guided=1, optimizer="adam", display_size=8,
        gen=dict(dim=64, mlp_dim=256, style_dim=16, activ="relu", n_downsample=2, n_res=4,
                 pad_type="reflect"),
        dis=dict(dim=64, norm="none", activ="lrelu", n_layer=4, gan_type="lsgan",
                 num_scales=3, pad_type="reflect"),
        input_dim_a=3, input_dim_b=3, new_size=256, crop_image_height=256,
        crop_image_width=256, batch_size=1, ratio_disc_gen=5,
    )
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(cfg.get(k), dict):
            cfg[k] = {**cfg[k], **v}
        else:
            cfg[k] = v
    return cfg


# --------------------------------------------------------------------------
# parameter inventories (names/shapes = the reference's state_dict contract)
# --------------------------------------------------------------------------
def _conv(p: str, cin: int, cout: int, k: int) -> List[Tuple[str, Tuple[int, ...]]]:
    return [(p + "weight", (cout, cin, k, k)), (p + "bias", (cout,))]


def style_encoder_spec(prefix, input_dim, dim, style_dim, n_downsample=4):
    """networks.py:442-477 (StyleEncoder)."""
    out = _conv(f"{prefix}model.0.conv.", input_dim, dim, 7)
    idx = 1
    for _ in range(2):
        out += _conv(f"{prefix}model.{idx}.conv.", dim, 2 * dim, 4)
        dim *= 2
        idx += 1
    for _ in range(n_downsample - 2):
        out += _conv(f"{prefix}model.{idx}.conv.", dim, dim, 4)
        idx += 1
    idx += 1  # AdaptiveAvgPool2d holds no parameters (model.<idx>)
    out += _conv(f"{prefix}model.{idx}.", dim, style_dim, 1)
    return out


def content_encoder_spec(prefix, n_down, n_res, input_dim, dim):
    """networks.py:480-512 (ContentEncoder)."""
    out = _conv(f"{prefix}model.0.conv.", input_dim, dim, 7)
    idx = 1
    for _ in range(n_down):
        out += _conv(f"{prefix}model.{idx}.conv.", dim, 2 * dim, 4)
        dim *= 2
        idx += 1
    for r in range(n_res):
        for j in range(2):
            out += _conv(f"{prefix}model.{idx}.model.{r}.model.{j}.conv.", dim, dim, 3)
    return out, dim


def decoder_spec(prefix, n_up, n_res, dim, output_dim):
    """networks.py:515-563 (Decoder); AdaIN dummy buffers networks.py:820-821."""
    out = []
    for r in range(n_res):
        for j in range(2):
            p = f"{prefix}model.0.model.{r}.model.{j}."
            out += _conv(p + "conv.", dim, dim, 3)
            out += [(p + "norm.running_mean", (dim,)), (p + "norm.running_var", (dim,))]
    idx = 1
    for _ in range(n_up):
        idx += 1  # nn.Upsample at model.<idx-1>
        p = f"{prefix}model.{idx}."
        out += _conv(p + "conv.", dim, dim // 2, 5)
        out += [(p + "norm.gamma", (dim // 2,)), (p + "norm.beta", (dim // 2,))]
        dim //= 2
        idx += 1
    out += _conv(f"{prefix}model.{idx}.conv.", dim, output_dim, 7)
    return out


def mlp_spec(prefix, in_dim, out_dim, dim, n_blk=3):
    """networks.py:583-597 (MLP of LinearBlocks)."""
    out = [(f"{prefix}model.0.fc.weight", (dim, in_dim)), (f"{prefix}model.0.fc.bias", (dim,))]
    for i in range(n_blk - 2):
        out += [(f"{prefix}model.{i+1}.fc.weight", (dim, dim)), (f"{prefix}model.{i+1}.fc.bias", (dim,))]
    out += [(f"{prefix}model.{n_blk-1}.fc.weight", (out_dim, dim)),
            (f"{prefix}model.{n_blk-1}.fc.bias", (out_dim,))]
    return out


def gen_spec(gp: dict, input_dim: int = 3, double: bool = True):
    """networks.py:170-209 (AdaINGen) / :262-323 (AdaINGen_double) registration order."""
    dim, sd_, nd, nr, md = gp["dim"], gp["style_dim"], gp["n_downsample"], gp["n_res"], gp["mlp_dim"]
    out = style_encoder_spec("enc_style.", input_dim, dim, sd_)
    cdim = dim * (2 ** nd)
    n_adain = nr * 2 * 2 * cdim
    if double:
        for n in ("enc1_content.", "enc2_content."):
            out += content_encoder_spec(n, nd, nr, input_dim, dim)[0]
        for n in ("dec1.", "dec2."):
            out += decoder_spec(n, nd, nr, cdim, input_dim)
        for n in ("mlp1.", "mlp2."):
            out += mlp_spec(n, sd_, n_adain, md)
    else:
        out += content_encoder_spec("enc_content.", nd, nr, input_dim, dim)[0]
        out += decoder_spec("dec.", nd, nr, cdim, input_dim)
        out += mlp_spec("mlp.", sd_, n_adain, md)
    return out


def dis_spec(dp: dict, input_dim: int = 3):
    """networks.py:20-70 (MsImageDis._make_net per scale)."""
    out = []
    for s in range(dp["num_scales"]):
        dim = dp["dim"]
        out += _conv(f"cnns.{s}.0.conv.", input_dim, dim, 4)
        for i in range(dp["n_layer"] - 1):
            out += _conv(f"cnns.{s}.{i+1}.conv.", dim, 2 * dim, 4)
            dim *= 2
        out += _conv(f"cnns.{s}.{dp['n_layer']}.", dim, 1, 1)
    return out


def init_state_dict(spec, seed: int, kind: str) -> SD:
    """Seeded init with the *distributions* of utils.py:1093-1115 (weights_init):
    generator nets kaiming_normal_(a=0, fan_in), discriminators N(0, 0.02), biases 0,
    LayerNorm gamma ~ U(0,1) (networks.py:859), beta 0, AdaIN buffers 0/1.
    (The RNG stream is ours -- fixtures load these tensors into the reference.)"""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for name, shape in spec:
        if name.endswith("running_mean") or name.endswith("beta") or name.endswith("bias"):
            sd[name] = torch.zeros(shape)
        elif name.endswith("running_var"):
            sd[name] = torch.ones(shape)
        elif name.endswith("gamma"):
            sd[name] = torch.rand(shape, generator=g)
        elif name.endswith("weight"):
            if kind == "gaussian":
                sd[name] = torch.randn(shape, generator=g) * 0.02
            else:
                fan_in = 1
                for d in shape[1:]:
                    fan_in *= d
                sd[name] = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        else:
            raise KeyError(name)
    return sd


# --------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------
def reflect_pad(x, p):
    """nn.ReflectionPad2d (networks.py:643): index -i -> i, H-1+i -> H-1-i."""
    return F.pad(x, (p, p, p, p), mode="reflect") if p > 0 else x


def instance_norm(x, weight=None, bias=None, eps=EPS_NORM):
    """nn.InstanceNorm2d(affine=False) networks.py:657 and AdaIN's F.batch_norm on the
    (1, B*C, H, W) view networks.py:832-845: per-(b,c) mean, *biased* variance,
    (x-mu)*rsqrt(var+eps) [*w + b with w,b of shape (B*C,), b-major]."""
    b, c = x.shape[:2]
    mu = x.mean(dim=(2, 3), keepdim=True)
    var = ((x - mu) ** 2).mean(dim=(2, 3), keepdim=True)
    y = (x - mu) / torch.sqrt(var + eps)
    if weight is not None:
        y = y * weight.view(b, c, 1, 1) + bias.view(b, c, 1, 1)
    return y


def layer_norm_munit(x, gamma, beta, eps=EPS_NORM):
    """networks.py:851-878: per-sample mean and *unbiased* std over C*H*W, eps added to
    std (outside the sqrt), then per-channel gamma/beta."""
    b = x.shape[0]
    flat = x.reshape(b, -1)
    mu = flat.mean(1).view(b, 1, 1, 1)
    n = flat.shape[1]
    std = torch.sqrt(((flat - flat.mean(1, keepdim=True)) ** 2).sum(1) / (n - 1)).view(b, 1, 1, 1)
    y = (x - mu) / (std + eps)
    return y * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)


def activation(x, kind):
    """networks.py:667-681."""
    if kind == "relu":
        return torch.relu(x)
    if kind == "lrelu":
        return F.leaky_relu(x, 0.2)
    if kind == "tanh":
        return torch.tanh(x)
    if kind == "none":
        return x
    raise ValueError(kind)


def conv_block(sd: SD, p: str, x, stride, pad, norm="none", act="relu", adain=None, residual=None, image=False):
    """Conv2dBlock.forward networks.py:695-701: pad -> conv(bias) -> norm -> activation
    (+ the ResBlock residual add of networks.py:623 when `residual` is given)."""
    if image:
        x = _qw(x)
    bias = sd[p + "conv.bias"]
    if QUANT and norm in ("in", "adain"):
        bias = None  # cancelled exactly by the mean subtraction; the B200 path skips it
    y = F.conv2d(reflect_pad(x, pad), _qw(sd[p + "conv.weight"]), bias, stride=stride)
    if norm != "none":
        y = _q(y)
    if norm == "in":
        y = instance_norm(y)
    elif norm == "adain":
        y = instance_norm(y, adain[1], adain[0])  # (bias, weight) pair -> weight, bias
    elif norm == "ln":
        y = layer_norm_munit(y, sd[p + "norm.gamma"], sd[p + "norm.beta"])
    elif norm != "none":
        raise ValueError(norm)
    y = activation(y, act)
    if residual is not None:
        y = y + residual
    return _q(y)


def style_encoder(sd: SD, p: str, x, activ="relu", n_downsample=4):
    """StyleEncoder networks.py:442-477: 7x7 -> 4x4s2 x2 (doubling) -> 4x4s2 x(n-2) -> GAP -> 1x1."""
    y = conv_block(sd, f"{p}model.0.", x, 1, 3, "none", activ, image=True)
    for i in range(1, n_downsample + 1):
        y = conv_block(sd, f"{p}model.{i}.", y, 2, 1, "none", activ)
    y = y.mean(dim=(2, 3), keepdim=True)
    k = n_downsample + 2
    return F.conv2d(y, sd[f"{p}model.{k}.weight"], sd[f"{p}model.{k}.bias"])


def res_blocks(sd: SD, p: str, x, n_res, norm, activ, adain_list=None):
    """ResBlocks/ResBlock networks.py:569-624: (conv3x3+norm+act, conv3x3+norm) + residual."""
    for r in range(n_res):
        a0 = adain_list[2 * r] if adain_list is not None else None
        a1 = adain_list[2 * r + 1] if adain_list is not None else None
        y = conv_block(sd, f"{p}model.{r}.model.0.", x, 1, 1, norm, activ, a0)
        x = conv_block(sd, f"{p}model.{r}.model.1.", y, 1, 1, norm, "none", a1, residual=x)
    return x


def content_encoder(sd: SD, p: str, x, n_down=2, n_res=4, activ="relu"):
    """ContentEncoder networks.py:480-512."""
    y = conv_block(sd, f"{p}model.0.", x, 1, 3, "in", activ, image=True)
    for i in range(1, n_down + 1):
        y = conv_block(sd, f"{p}model.{i}.", y, 2, 1, "in", activ)
    return res_blocks(sd, f"{p}model.{n_down+1}.", y, n_res, "in", activ)


def mlp(sd: SD, p: str, style, n_blk=3):
    """MLP networks.py:583-597: view(B,-1) -> (Linear+ReLU) x (n_blk-1) -> Linear."""
    y = style.reshape(style.shape[0], -1)
    for i in range(n_blk):
        y = F.linear(y, sd[f"{p}model.{i}.fc.weight"], sd[f"{p}model.{i}.fc.bias"])
        if i < n_blk - 1:
            y = torch.relu(y)
    return y


def split_adain_params(adain_params, n_layers, c):
    """assign_adain_params networks.py:230-239: per AdaIN layer in module order take
    columns [:C] -> bias ("mean"), [C:2C] -> weight ("std"), flatten b-major, drop 2C."""
    out = []
    for l in range(n_layers):
        bias = adain_params[:, 2 * c * l: 2 * c * l + c].contiguous().view(-1)
        weight = adain_params[:, 2 * c * l + c: 2 * c * (l + 1)].contiguous().view(-1)
        out.append((bias, weight))
    return out


def decoder(sd: SD, p: str, content, adain_params, n_up=2, n_res=4, activ="relu"):
    """Decoder networks.py:515-563: AdaIN ResBlocks -> [nearest x2 -> 5x5 conv + LN + act] x n_up
    -> 7x7 conv + tanh."""
    c = content.shape[1]
    ad = split_adain_params(adain_params, 2 * n_res, c)
    y = res_blocks(sd, f"{p}model.0.", content, n_res, "adain", activ, ad)
    idx = 1
    for _ in range(n_up):
        y = F.interpolate(y, scale_factor=2, mode="nearest")
        y = conv_block(sd, f"{p}model.{idx+1}.", y, 1, 2, "ln", activ)
        idx += 2
    return conv_block(sd, f"{p}model.{idx}.", y, 1, 3, "none", "tanh")


class Gen:
    """Functional AdaINGen / AdaINGen_double (networks.py:170-388). `name` in {1,2} selects
    enc{name}_content/dec{name}/mlp{name} for the double generator; ignored otherwise."""

    def __init__(self, sd: SD, gp: dict, double: bool):
        self.sd, self.gp, self.double = sd, gp, double

    def _n(self, base, name):
        if not self.double:
            return {"enc": "enc_content.", "dec": "dec.", "mlp": "mlp."}[base]
        return {"enc": f"enc{name}_content.", "dec": f"dec{name}.", "mlp": f"mlp{name}."}[base]

    def encode(self, x, name=None):
        gp = self.gp
        s = style_encoder(self.sd, "enc_style.", x, gp["activ"])
        c = content_encoder(self.sd, self._n("enc", name), x, gp["n_downsample"], gp["n_res"], gp["activ"])
        return c, s

    def decode(self, c, s, name=None):
        gp = self.gp
        ap = mlp(self.sd, self._n("mlp", name), s)
        return decoder(self.sd, self._n("dec", name), c, ap, gp["n_downsample"], gp["n_res"], gp["activ"])


def avg_pool_3s2(x):
    """nn.AvgPool2d(3, stride=2, padding=1, count_include_pad=False) networks.py:32-34:
    divisor = number of in-bounds taps (4 / 6 / 9)."""
    ones = torch.ones_like(x[:1, :1])
    num = F.avg_pool2d(x, 3, 2, 1, count_include_pad=True) * 9.0
    cnt = F.avg_pool2d(ones, 3, 2, 1, count_include_pad=True) * 9.0
    return num / cnt


def dis_forward(sd: SD, dp: dict, x) -> List[torch.Tensor]:
    """MsImageDis.forward networks.py:72-77."""
    outs = []
    for s in range(dp["num_scales"]):
        y = conv_block(sd, f"cnns.{s}.0.", x, 2, 1, "none", dp["activ"], image=True)
        for i in range(1, dp["n_layer"]):
            y = conv_block(sd, f"cnns.{s}.{i}.", y, 2, 1, "none", dp["activ"])
        k = dp["n_layer"]
        outs.append(F.conv2d(y, sd[f"cnns.{s}.{k}.weight"], sd[f"cnns.{s}.{k}.bias"]))
        x = avg_pool_3s2(x)
    return outs


def calc_dis_loss(sd, dp, fake, real):
    """networks.py:79-101 (lsgan)."""
    loss = 0
    for o0, o1 in zip(dis_forward(sd, dp, fake), dis_forward(sd, dp, real)):
        loss = loss + torch.mean((o0 - 0) ** 2) + torch.mean((o1 - 1) ** 2)
    return loss


def calc_gen_loss(sd, dp, fake):
    """networks.py:103-115 (lsgan)."""
    loss = 0
    for o0 in dis_forward(sd, dp, fake):
        loss = loss + torch.mean((o0 - 1) ** 2)
    return loss


def l1(a, b):
    """recon_criterion trainer.py:279-290."""
    return torch.mean(torch.abs(a - b))


def synthetic_pair(seed, b, hw):
    """Seeded synthetic pair for the masked / synthetic-pair losses: x_b equals x_a except inside a box (so the
    alignment mask of trainer.py:455 is non-trivial), plus two binary masks [b,1,hw,hw]."""
    g = torch.Generator().manual_seed(seed)
    x_a = torch.rand(b, 3, hw, hw, generator=g) * 2 - 1
    x_b = x_a.clone()
    x_b[:, :, hw // 4: hw // 2 + 5, hw // 8: hw // 2] = torch.rand(b, 3, hw // 4 + 5, hw // 2 - hw // 8, generator=g) * 2 - 1
    mask_a = (torch.rand(b, 1, hw, hw, generator=g) > 0.6).float()
    mask_b = torch.zeros(b, 1, hw, hw)
    mask_b[:, :, hw // 3:, : hw // 2] = 1.0
    return x_a, x_b, mask_a, mask_b


def l1_masked(a, b, mask):
    """recon_criterion_mask trainer.py:292-305 (mean over *all* elements)."""
    return torch.mean(torch.abs((a - b) * (1 - mask)))


# --------------------------------------------------------------------------
# domain-adaptation head (utils.py:1238-1392, trainer.py:638-667)
# --------------------------------------------------------------------------
def domain_classifier_spec():
    """Parameter / buffer inventory of utils.py:1370-1392 domainClassifier(256) (state_dict order)."""
    spec = []
    for blk, cin, cout in (("BasicBlock1.", 256, 128), ("BasicBlock2.", 128, 64)):
        spec.append((blk + "conv1.weight", (cout, cin, 3, 3)))
        spec += _bn_spec(blk + "bn1.", cout)
        spec.append((blk + "conv2.weight", (cout, cout, 3, 3)))
        spec += _bn_spec(blk + "bn2.", cout)
        spec.append((blk + "downsample.0.weight", (cout, cin, 1, 1)))
        spec += _bn_spec(blk + "downsample.1.", cout)
    spec += [("fc.weight", (1, 64)), ("fc.bias", (1,))]
    return spec


def _bn_spec(p, c):
    return [(p + "weight", (c,)), (p + "bias", (c,)), (p + "running_mean", (c,)), (p + "running_var", (c,)),
            (p + "num_batches_tracked", ())]


def init_classifier_state_dict(seed: int) -> SD:
    """weights_init("gaussian") on domainClassifier (trainer.py:175-176): Conv / Linear weights N(0, 0.02), fc bias
    0; BatchNorm keeps nn.BatchNorm2d's defaults (weight 1, bias 0, running 0 / 1) -- here gamma / beta are drawn
    away from 1 / 0 so that the fixtures exercise them."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for name, shape in domain_classifier_spec():
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.zeros((), dtype=torch.long)
        elif name.endswith("running_mean"):
            sd[name] = torch.zeros(shape)
        elif name.endswith("running_var"):
            sd[name] = torch.ones(shape)
        elif ".bn" in name or "downsample.1." in name:
            sd[name] = (1.0 + 0.2 * torch.randn(shape, generator=g)) if name.endswith("weight") else 0.1 * torch.randn(shape, generator=g)
        elif name == "fc.bias":
            sd[name] = torch.zeros(shape)
        else:
            sd[name] = torch.randn(shape, generator=g) * 0.02
    return sd


def batch_norm(sd: SD, p: str, x, training: bool, momentum=0.1, eps=1e-5):
    """nn.BatchNorm2d (utils.py:1300,1303,1308): training -> per-channel mean / *biased* variance over (N,H,W),
    running statistics updated in place with the *unbiased* variance; eval -> running statistics."""
    if training:
        mu = x.mean(dim=(0, 2, 3))
        var = ((x - mu.view(1, -1, 1, 1)) ** 2).mean(dim=(0, 2, 3))
        with torch.no_grad():
            m = x.numel() / x.shape[1]
            sd[p + "running_mean"].mul_(1 - momentum).add_(momentum * mu.detach())
            sd[p + "running_var"].mul_(1 - momentum).add_(momentum * var.detach() * m / max(m - 1, 1))
            if p + "num_batches_tracked" in sd:
                sd[p + "num_batches_tracked"] += 1
    else:
        mu, var = sd[p + "running_mean"], sd[p + "running_var"]
    y = (x - mu.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + eps)
    return y * sd[p + "weight"].view(1, -1, 1, 1) + sd[p + "bias"].view(1, -1, 1, 1)


def basic_block(sd: SD, p: str, x, training: bool):
    """BasicBlock.forward utils.py:1311-1331 with the conv1x1 + bn shortcut (inplanes != planes)."""
    out = _q(F.conv2d(x, _qw(sd[p + "conv1.weight"]), None, padding=1))
    out = _q(torch.relu(batch_norm(sd, p + "bn1.", out, training)))
    out = _q(F.conv2d(out, _qw(sd[p + "conv2.weight"]), None, padding=1))
    out = _q(batch_norm(sd, p + "bn2.", out, training))
    idn = _q(F.conv2d(x, _qw(sd[p + "downsample.0.weight"]), None))
    idn = _q(batch_norm(sd, p + "downsample.1.", idn, training))
    return _q(torch.relu(out + idn))


def domain_classifier(sd: SD, x, training: bool = True):
    """domainClassifier.forward utils.py:1380-1392: content code [N,256,64,64] -> [N,1] ([1] for N == 1)."""
    y = F.max_pool2d(x, 2)
    y = basic_block(sd, "BasicBlock1.", y, training)
    y = F.max_pool2d(y, 2)
    y = basic_block(sd, "BasicBlock2.", y, training)
    y = F.avg_pool2d(y, (16, 16))
    return F.linear(y.squeeze(), sd["fc.weight"], sd["fc.bias"])


def classifier_sr_loss(sd_a: SD, sd_b: SD, c_a, c_b, domain_synth=False, fool=False, training=True):
    """compute_classifier_sr_loss trainer.py:638-667."""
    o_a = domain_classifier(sd_a, c_a, training)
    o_b = domain_classifier(sd_b, c_b, training)
    t = 0.5 if fool else (0.0 if domain_synth else 1.0)
    return torch.mean((o_a - t) ** 2) + torch.mean((o_b - t) ** 2)


# --------------------------------------------------------------------------
# optimisers
# --------------------------------------------------------------------------
def adam_torch_step(p, g, m, v, step, lr, b1, b2, eps, wd):
    """torch.optim.Adam (torch 2.11, the optimiser trainer.py:41-45 resolves to):
    g += wd*p; m,v EMA; denom = sqrt(v)/sqrt(1-b2^t) + eps; p -= lr/(1-b1^t) * m/denom."""
    g = g + wd * p
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p.addcdiv_(m, denom, value=-lr / bc1)


def adam_legacy_update(p, g, m, v, step, lr, b1, b2, eps, wd):
    """ExtraAdam.update extraadam.py:119-168: denom = sqrt(v)+eps;
    step_size = lr*sqrt(1-b2^t)/(1-b1^t); returns u = -step_size*m/denom."""
    g = g + wd * p
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    return -(lr * math.sqrt(bc2) / bc1) * m / (v.sqrt() + eps)


class OracleOpt:
    """Adam or ExtraAdam (extrapolation/step pair, extraadam.py:30-74) over a dict of tensors."""

    def __init__(self, params: SD, lr, b1, b2, wd, extra: bool, eps=1e-8):
        self.params, self.lr, self.b1, self.b2, self.wd, self.extra, self.eps = params, lr, b1, b2, wd, extra, eps
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.step_n = 0
        self.copy: Optional[SD] = None

    def apply(self, grads: SD, iterations: int):
        """trainer.py:252-268: ExtraAdam extrapolates on even `iterations`, steps on odd."""
        self.step_n += 1
        for k, p in self.params.items():
            g = grads.get(k)
            if g is None:
                continue
            if not self.extra:
                adam_torch_step(p, g, self.m[k], self.v[k], self.step_n, self.lr, self.b1, self.b2, self.eps, self.wd)
        if self.extra:
            if iterations % 2 == 0:
                first = self.copy is None
                if first:
                    self.copy = {}
                for k, p in self.params.items():
                    u = adam_legacy_update(p, grads[k], self.m[k], self.v[k], self.step_n, self.lr, self.b1, self.b2, self.eps, self.wd)
                    if first:
                        self.copy[k] = p.clone()
                    p.add_(u)
            else:
                if self.copy is None:
                    raise RuntimeError("Need to call extrapolation before calling step.")
                for k, p in self.params.items():
                    u = adam_legacy_update(p, grads[k], self.m[k], self.v[k], self.step_n, self.lr, self.b1, self.b2, self.eps, self.wd)
                    p.copy_(self.copy[k] + u)
                self.copy = None


def step_lr(base_lr, step_size, gamma, n_sched_steps):
    """StepLR(step_size, gamma) utils.py:1066-1090 after n scheduler.step() calls
    (update_learning_rate is called at the top of every iteration, train.py:172)."""
    return base_lr * (gamma ** (n_sched_steps // step_size))


# --------------------------------------------------------------------------
# trainer
# --------------------------------------------------------------------------
class OracleTrainer:
    """Functional MUNIT_Trainer hot path: trainer.py:28-127 (setup), :336-561 (gen_update),
    :1133-1186 (dis_update).  Weights come in as state_dicts with the reference's keys."""

    def __init__(self, cfg: dict, gen_sd, dis_a_sd: SD, dis_b_sd: SD, cls_a_sd: Optional[SD] = None,
                 cls_b_sd: Optional[SD] = None):
        self.cfg = cfg
        self.gen_state, self.guided = cfg["gen_state"], cfg["guided"]
        self.style_dim = cfg["gen"]["style_dim"]
        req = lambda sd: {k: (v.clone().requires_grad_(True)
                              if not k.endswith(("running_mean", "running_var", "num_batches_tracked")) else v.clone())
                          for k, v in sd.items()}
        if self.gen_state == 1:
            self.gen_sd = {"": req(gen_sd)}
            g = Gen(self.gen_sd[""], cfg["gen"], True)
            self.enc_a = lambda x: g.encode(x, 1)
            self.enc_b = lambda x: g.encode(x, 2)
            self.dec_a = lambda c, s: g.decode(c, s, 1)
            self.dec_b = lambda c, s: g.decode(c, s, 2)
        else:
            self.gen_sd = {"a": req(gen_sd["a"]), "b": req(gen_sd["b"])}
            ga = Gen(self.gen_sd["a"], cfg["gen"], False)
            gb = Gen(self.gen_sd["b"], cfg["gen"], False)
            self.enc_a, self.enc_b, self.dec_a, self.dec_b = ga.encode, gb.encode, ga.decode, gb.decode
        self.dis_a, self.dis_b = req(dis_a_sd), req(dis_b_sd)
        extra = "extra" in cfg["optimizer"]
        flat = lambda d: {f"{n}/{k}": v for n, sd in d.items() for k, v in sd.items() if v.requires_grad}
        self.gen_params = flat(self.gen_sd)
        self.dis_params = flat({"a": self.dis_a, "b": self.dis_b})
        mk = lambda ps: OracleOpt({k: v.data for k, v in ps.items()}, cfg["lr"], cfg["beta1"], cfg["beta2"],
                                  cfg["weight_decay"], extra)
        self.gen_opt, self.dis_opt = mk(self.gen_params), mk(self.dis_params)
        self.cls_a = self.cls_b = None
        if cls_a_sd is not None:  # trainer.py:162-179
            self.cls_a, self.cls_b = req(cls_a_sd), req(cls_b_sd)
            self.cls_params = flat({"a": self.cls_a, "b": self.cls_b})
            self.cls_opt = mk(self.cls_params)
        self.iterations = 0
        self.losses: Dict[str, float] = {}

    def _grads(self, loss, params):
        names = list(params.keys())
        gs = torch.autograd.grad(loss, [params[n] for n in names], allow_unused=True)
        return {n: g for n, g in zip(names, gs) if g is not None}

    def dis_update(self, x_a, x_b, s_a=None, s_b=None):
        cfg, B = self.cfg, x_a.shape[0]
        if s_a is None:  # trainer.py:1146-1147 draw order: s_a then s_b, host generator
            s_a = torch.randn(B, self.style_dim, 1, 1)
            s_b = torch.randn(x_b.shape[0], self.style_dim, 1, 1)
        with torch.no_grad():
            c_a, s_a_p = self.enc_a(x_a)
            c_b, s_b_p = self.enc_b(x_b)
            if self.guided == 0:
                x_ba, x_ab = self.dec_a(c_b, s_a), self.dec_b(c_a, s_b)
            else:
                x_ba, x_ab = self.dec_a(c_b, s_a_p), self.dec_b(c_a, s_b_p)
        dp = cfg["dis"]
        la = calc_dis_loss(self.dis_a, dp, x_ba, x_a)
        lb = calc_dis_loss(self.dis_b, dp, x_ab, x_b)
        total = cfg["gan_w"] * la + cfg["gan_w"] * lb
        grads = self._grads(total, self.dis_params)
        self.dis_opt.apply(grads, self.iterations)
        self.losses.update(loss_dis_a=float(la), loss_dis_b=float(lb), loss_dis_total=float(total))
        self.dis_grads = grads
        return total.detach()

    def gen_update(self, x_a, x_b, s_a=None, s_b=None, mask_a=None, mask_b=None, synth=False):
        """trainer.py:336-561 incl. the masked cycle loss (recon_mask == 1, trainer.py:466-488) and the
        synthetic-pair reconstruction loss (synth=True, trainer.py:452-464)."""
        cfg, B = self.cfg, x_a.shape[0]
        if s_a is None:  # trainer.py:366-367
            s_a = torch.randn(B, self.style_dim, 1, 1)
            s_b = torch.randn(x_b.shape[0], self.style_dim, 1, 1)
        c_a, s_a_p = self.enc_a(x_a)
        c_b, s_b_p = self.enc_b(x_b)
        x_a_recon, x_b_recon = self.dec_a(c_a, s_a_p), self.dec_b(c_b, s_b_p)
        if self.guided == 0:
            x_ba, x_ab = self.dec_a(c_b, s_a), self.dec_b(c_a, s_b)
        else:
            x_ba, x_ab = self.dec_a(c_b, s_a_p), self.dec_b(c_a, s_b_p)
        c_b_recon, s_a_recon = self.enc_a(x_ba)
        c_a_recon, s_b_recon = self.enc_b(x_ab)
        cyc = cfg["recon_x_cyc_w"] > 0
        x_aba = self.dec_a(c_a_recon, s_a_p) if cyc else None
        x_bab = self.dec_b(c_b_recon, s_b_p) if cyc else None
        L = {}
        L["loss_gen_recon_x_a"], L["loss_gen_recon_x_b"] = l1(x_a_recon, x_a), l1(x_b_recon, x_b)
        if self.guided == 0:
            L["loss_gen_recon_s_a"], L["loss_gen_recon_s_b"] = l1(s_a_recon, s_a), l1(s_b_recon, s_b)
        else:
            L["loss_gen_recon_s_a"], L["loss_gen_recon_s_b"] = l1(s_a_recon, s_a_p), l1(s_b_recon, s_b_p)
        L["loss_gen_recon_c_a"], L["loss_gen_recon_c_b"] = l1(c_a_recon, c_a), l1(c_b_recon, c_b)
        if cfg.get("recon_mask", 0) == 1:
            L["loss_gen_cycrecon_x_a"] = l1_masked(x_aba, x_a, mask_a) if cyc else 0
            L["loss_gen_cycrecon_x_b"] = l1_masked(x_bab, x_b, mask_b) if cyc else 0
        else:
            L["loss_gen_cycrecon_x_a"] = l1(x_aba, x_a) if cyc else 0
            L["loss_gen_cycrecon_x_b"] = l1(x_bab, x_b) if cyc else 0
        L["loss_gen_recon_synth"] = 0
        if synth:
            mask_alignment = (torch.sum(torch.abs(x_a - x_b), 1) == 0).unsqueeze(1).float()
            L["loss_gen_recon_synth"] = (l1_masked(x_ab, x_b, 1 - mask_alignment)
                                         + l1_masked(x_ba, x_a, 1 - mask_alignment))
        dp = cfg["dis"]
        L["loss_gen_adv_a"] = calc_gen_loss(self.dis_a, dp, x_ba)
        L["loss_gen_adv_b"] = calc_gen_loss(self.dis_b, dp, x_ab)
        total = (cfg["gan_w"] * (L["loss_gen_adv_a"] + L["loss_gen_adv_b"])
                 + cfg["recon_x_w"] * (L["loss_gen_recon_x_a"] + L["loss_gen_recon_x_b"])
                 + cfg["recon_s_w"] * (L["loss_gen_recon_s_a"] + L["loss_gen_recon_s_b"])
                 + cfg["recon_c_w"] * (L["loss_gen_recon_c_a"] + L["loss_gen_recon_c_b"])
                 + cfg["recon_x_cyc_w"] * (L["loss_gen_cycrecon_x_a"] + L["loss_gen_cycrecon_x_b"])
                 + cfg.get("recon_synth_w", 0) * L["loss_gen_recon_synth"])
        adv_lambda = cfg["adaptation"].get("adv_lambda", 0)
        if adv_lambda > 0:  # trainer.py:521-525,555
            L["loss_classifier_sr"] = classifier_sr_loss(self.cls_a, self.cls_b, c_a, c_b, synth, fool=True)
            total = total + adv_lambda * L["loss_classifier_sr"]
        grads = self._grads(total, self.gen_params)
        self.gen_opt.apply(grads, self.iterations)
        self.losses.update({k: float(v) for k, v in L.items()})
        self.losses["loss_gen_total"] = float(total)
        self.gen_grads = grads
        self.last = dict(x_ab=x_ab.detach(), x_ba=x_ba.detach())
        return total.detach()

    def domain_classifier_sr_update(self, x_a, x_b, domain_synth, lambda_classifier):
        """trainer.py:1237-1265: the classifiers learn synthetic (0) / real (1) on detached content codes."""
        with torch.no_grad():
            c_a, _ = self.enc_a(x_a)
            c_b, _ = self.enc_b(x_b)
        loss = lambda_classifier * classifier_sr_loss(self.cls_a, self.cls_b, c_a, c_b, domain_synth, fool=False)
        grads = self._grads(loss, self.cls_params)
        self.cls_opt.apply(grads, self.iterations)
        self.losses["loss_classifier_sr_update"] = float(loss)
        self.cls_grads = grads
        return loss.detach()

    @torch.no_grad()
    def translate(self, x_a, styles):
        """test_batch.py:146-164 semantics: encode once, decode once per style code."""
        c_a, _ = self.enc_a(x_a)
        return [self.dec_b(c_a, styles[j:j + 1].expand(x_a.shape[0], -1, -1, -1)) for j in range(styles.shape[0])]
