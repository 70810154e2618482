"""B200-native drop-in for cc-ai/MUNIT `scripts/networks.py` (hot-path classes only).

Same class names, constructor signatures, sub-module attribute names and state_dict keys/shapes as the
reference (MsImageDis networks.py:20-115, AdaINGen :170-254, AdaINGen_double :262-388, StyleEncoder
:442-477, ContentEncoder :480-512, Decoder :515-563, ResBlocks :569-580, MLP :583-597, ResBlock :603-624,
Conv2dBlock :627-701, LinearBlock :704-749, AdaptiveInstanceNorm2d :810-848, LayerNorm :851-878).
Public tensors are NCHW fp32; between our own layers activations travel as `ops.Act` (NHWC bf16 with
the consumer's reflect halo) and all arithmetic runs in the sm_100a kernels of libmunit_b200.so --
nn.Conv2d / nn.Linear instances only *hold* the fp32 master parameters, their forward is never called.
Out of scope (raise): VAEGen, Vgg16, SpectralNorm, bn/sn norms, prelu/selu, zero/replicate padding > 0.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn

from . import ops
from .ops import Act

_ACTS = {"relu": "relu", "lrelu": "lrelu", "tanh": "tanh", "none": "none"}


def _as_cl(conv: nn.Conv2d):
    """Store the conv weight in channels_last memory ([Cout][KH][KW][Cin]) -- the GEMM K order."""
    conv.weight.data = conv.weight.data.contiguous(memory_format=torch.channels_last)
    return conv


##################################################################################
# Normalization layers
##################################################################################
class AdaptiveInstanceNorm2d(nn.Module):
    """networks.py:810-848.  `weight` / `bias` are assigned externally (b-major, B*C values)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.num_features = num_features
        self.eps = eps
        self.momentum = momentum
        self._w = None
        self._b = None
        # just dummy buffers, not used (kept: they are part of the reference state_dict)
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))

    @property
    def weight(self):
        return None if self._w is None else self._w.contiguous().view(-1)

    @weight.setter
    def weight(self, v):
        self._w = v

    @property
    def bias(self):
        return None if self._b is None else self._b.contiguous().view(-1)

    @bias.setter
    def bias(self, v):
        self._b = v

    def params2d(self, b):
        assert self._w is not None and self._b is not None, "Please assign weight and bias before calling AdaIN!"
        c = self.num_features
        w = self._w if self._w.dim() == 2 else self._w.view(b, c)
        bb = self._b if self._b.dim() == 2 else self._b.view(b, c)
        if w.stride(1) != 1 or bb.stride(1) != 1 or w.stride(0) != bb.stride(0):
            w, bb = w.contiguous(), bb.contiguous()
        return w.float(), bb.float()

    def forward(self, x):
        b = x.size(0)
        w, bb = self.params2d(b)
        y = ops.ToActFn.apply(x, 0)
        out = ops.NormFn.apply(y, w, bb, None, "adain", False, 0, 0, 1, self.eps)
        return ops.FromActFn.apply(out, self.num_features, 0)

    def __repr__(self):
        return self.__class__.__name__ + "(" + str(self.num_features) + ")"


class LayerNorm(nn.Module):
    """networks.py:851-878: per-sample mean / unbiased std over C*H*W, eps added to std."""

    def __init__(self, num_features, eps=1e-5, affine=True):
        super().__init__()
        self.num_features = num_features
        self.affine = affine
        self.eps = eps
        if self.affine:
            self.gamma = nn.Parameter(torch.Tensor(num_features).uniform_())
            self.beta = nn.Parameter(torch.zeros(num_features))

    def affine_params(self, dev):
        if self.affine:
            return self.gamma, self.beta
        return (torch.ones(self.num_features, device=dev), torch.zeros(self.num_features, device=dev))

    def forward(self, x):
        g, b = self.affine_params(x.device)
        y = ops.ToActFn.apply(x, 0)
        out = ops.NormFn.apply(y, g, b, None, "ln", False, 0, 0, 1, self.eps)
        return ops.FromActFn.apply(out, self.num_features, 0)


##################################################################################
# Basic Blocks
##################################################################################
class Conv2dBlock(nn.Module):
    """networks.py:627-701: pad -> conv(bias) -> norm -> activation."""

    def __init__(self, input_dim, output_dim, kernel_size, stride, padding=0, norm="none", activation="relu",
                 pad_type="zero"):
        super().__init__()
        self.use_bias = True
        if pad_type == "reflect":
            self.pad = nn.ReflectionPad2d(padding)
        elif pad_type == "zero" and padding == 0:
            self.pad = nn.ZeroPad2d(padding)
        else:
            assert 0, "Unsupported padding type: {}".format(pad_type)
        self.padding = padding
        norm_dim = output_dim
        if norm == "in":
            self.norm = nn.InstanceNorm2d(norm_dim)  # parameter-free marker; never called
        elif norm == "ln":
            self.norm = LayerNorm(norm_dim)
        elif norm == "adain":
            self.norm = AdaptiveInstanceNorm2d(norm_dim)
        elif norm == "none":
            self.norm = None
        else:
            assert 0, "Unsupported normalization: {}".format(norm)
        self.norm_type = norm
        if activation == "relu":
            self.activation = nn.ReLU(inplace=True)
        elif activation == "lrelu":
            self.activation = nn.LeakyReLU(0.2, inplace=True)
        elif activation == "tanh":
            self.activation = nn.Tanh()
        elif activation == "none":
            self.activation = None
        else:
            assert 0, "Unsupported activation: {}".format(activation)
        self.act_type = activation
        if norm != "none":
            assert activation in ("relu", "none"), "norm layers are fused with relu/none only"
            assert output_dim % 64 == 0, "normalised layers need a multiple of 64 output channels"
        self.conv = _as_cl(nn.Conv2d(input_dim, output_dim, kernel_size, stride, bias=self.use_bias))
        self.layer = ops.ConvLayer(input_dim, output_dim, kernel_size, stride, padding)

    # -- internal fast path ---------------------------------------------------------------
    def forward_act(self, x, out_pad: int = 0, upsample: int = 1, residual: Optional[Act] = None,
                    frozen: bool = False) -> Act:
        """x: Act (any halo) or an NCHW fp32 image for the first (Cin<64) layers."""
        layer = self.layer
        if isinstance(x, Act):
            if layer.first:
                raise ValueError("image-space layer expects an NCHW tensor")
            if x.pad != self.padding:
                x = Act(ops.RepadFn.apply(x.t, x.pad, self.padding), self.padding)
            xin = x.t
        elif layer.first:
            xin = x
        else:
            xin = ops.ToActFn.apply(x, self.padding)
        w = self.conv.weight.detach() if frozen else self.conv.weight
        b = self.conv.bias.detach() if frozen else self.conv.bias
        if layer.last and self.norm_type == "none":
            # narrow-output layer: produces the public NCHW fp32 image directly (must end a chain)
            assert upsample == 1 and residual is None and out_pad == 0
            return ops.ConvOutFn.apply(xin, w, b, layer, self.act_type)
        if self.norm_type == "none":
            assert upsample == 1 and residual is None
            out = ops.ConvFn.apply(xin, w, b, layer, self.act_type, out_pad, self.padding)
            return Act(out, out_pad)
        # a per-channel bias is cancelled exactly by the mean subtraction of IN / AdaIN: skip it
        bias = b if self.norm_type == "ln" else None
        y_f16 = True  # raw conv output in front of a norm: fp16 bits (ops.ConvFn)
        y, part = ops.ConvFn.apply(xin, w, bias, layer, "none", 0, self.padding, 2 if self.norm_type == "ln" else 1)
        n = y.shape[0]
        if self.norm_type == "in":
            p_w = p_b = None
        elif self.norm_type == "adain":
            p_w, p_b = self.norm.params2d(n)
        else:
            p_w, p_b = self.norm.affine_params(y.device)
        res_t = residual.t if residual is not None else None
        res_pad = residual.pad if residual is not None else 0
        out = ops.NormFn.apply(y, p_w, p_b, res_t, self.norm_type, self.act_type == "relu", res_pad, out_pad,
                               upsample, self.norm.eps if self.norm_type != "in" else 1e-5, part, y_f16)
        return Act(out, out_pad)

    # -- public path (tensor in, tensor out) ----------------------------------------------
    def forward(self, x):
        a = self.forward_act(x, 0)
        if torch.is_tensor(a):
            return a
        return ops.FromActFn.apply(a.t, self.conv.out_channels, 0)


class ResBlock(nn.Module):
    """networks.py:603-624."""

    def __init__(self, dim, norm="in", activation="relu", pad_type="zero"):
        super().__init__()
        model = []
        model += [Conv2dBlock(dim, dim, 3, 1, 1, norm=norm, activation=activation, pad_type=pad_type)]
        model += [Conv2dBlock(dim, dim, 3, 1, 1, norm=norm, activation="none", pad_type=pad_type)]
        self.model = nn.Sequential(*model)
        assert norm in ("in", "adain", "ln"), "ResBlock needs a normalised second conv to fuse the residual add"

    def forward_act(self, x: Act, out_pad: int = 0, upsample: int = 1) -> Act:
        c1, c2 = self.model[0], self.model[1]
        if x.pad != c1.padding:
            x = Act(ops.RepadFn.apply(x.t, x.pad, c1.padding), c1.padding)
        y = c1.forward_act(x, c2.padding)
        return c2.forward_act(y, out_pad, upsample, residual=x)  # out = norm(conv(y)) + x   (networks.py:623)

    def forward(self, x):
        a = Act(ops.ToActFn.apply(x, 1), 1)
        out = self.forward_act(a, 0)
        return ops.FromActFn.apply(out.t, x.shape[1], 0)


class LinearBlock(nn.Module):
    """networks.py:704-749 (norm none, activation relu/none)."""

    def __init__(self, input_dim, output_dim, norm="none", activation="relu"):
        super().__init__()
        self.fc = nn.Linear(input_dim, output_dim, bias=True)
        if norm != "none":
            assert 0, "Unsupported normalization: {}".format(norm)
        self.norm = None
        if activation == "relu":
            self.activation = nn.ReLU(inplace=True)
        elif activation == "none":
            self.activation = None
        else:
            assert 0, "Unsupported activation: {}".format(activation)

    def forward(self, x):
        return ops.LinearFn.apply(x.float(), self.fc.weight, self.fc.bias, self.activation is not None)


##################################################################################
# Sequential Models
##################################################################################
def _required_pad(m) -> int:
    if isinstance(m, Conv2dBlock):
        return m.padding
    if isinstance(m, ResBlock):
        return m.model[0].padding
    if isinstance(m, ResBlocks):
        return m.model[0].model[0].padding
    return 0


def _flatten(mods) -> List[nn.Module]:
    out = []
    for m in mods:
        if isinstance(m, ResBlocks):
            out += list(m.model)
        else:
            out.append(m)
    return out


def run_chain(mods, x, final_pad: int = 0, frozen: bool = False) -> Act:
    """Run Conv2dBlock / ResBlock(s) / nn.Upsample modules; every producer writes the halo (and the
    nearest-2x upsample, networks.py:534) that its consumer needs, so neither is ever a separate op."""
    flat = _flatten(mods)
    i = 0
    while i < len(flat):
        m = flat[i]
        j, up = i + 1, 1
        if j < len(flat) and isinstance(flat[j], nn.Upsample):
            up, j = 2, j + 1
        next_pad = _required_pad(flat[j]) if j < len(flat) else final_pad
        if isinstance(m, Conv2dBlock):
            x = m.forward_act(x, next_pad, up, frozen=frozen)
        elif isinstance(m, ResBlock):
            x = m.forward_act(x, next_pad, up)
        elif isinstance(m, nn.Upsample):
            raise ValueError("nn.Upsample must follow a normalised block (it is fused into its producer)")
        else:
            raise ValueError(f"unsupported module in chain: {type(m).__name__}")
        i = j
    return x


class ResBlocks(nn.Module):
    """networks.py:569-580."""

    def __init__(self, num_blocks, dim, norm="in", activation="relu", pad_type="zero"):
        super().__init__()
        self.model = []
        for i in range(num_blocks):
            self.model += [ResBlock(dim, norm=norm, activation=activation, pad_type=pad_type)]
        self.model = nn.Sequential(*self.model)

    def forward(self, x):
        a = Act(ops.ToActFn.apply(x, 1), 1)
        out = run_chain([self], a, 0)
        return ops.FromActFn.apply(out.t, x.shape[1], 0)


class MLP(nn.Module):
    """networks.py:583-597."""

    def __init__(self, input_dim, output_dim, dim, n_blk, norm="none", activ="relu"):
        super().__init__()
        self.model = []
        self.model += [LinearBlock(input_dim, dim, norm=norm, activation=activ)]
        for i in range(n_blk - 2):
            self.model += [LinearBlock(dim, dim, norm=norm, activation=activ)]
        self.model += [LinearBlock(dim, output_dim, norm="none", activation="none")]  # no output activations
        self.model = nn.Sequential(*self.model)

    def forward(self, x):
        x = x.reshape(x.size(0), -1)
        m = self.model
        if len(m) == 3 and all(b.norm is None for b in m) and m[0].activation is not None \
                and m[1].activation is not None and m[2].activation is None \
                and isinstance(m[0].activation, nn.ReLU) and isinstance(m[1].activation, nn.ReLU):
            # the shipped shape (3 blocks, ReLU, no norm): one fused forward kernel
            return ops.Mlp3Fn.apply(x.float(), m[0].fc.weight, m[0].fc.bias, m[1].fc.weight, m[1].fc.bias,
                                    m[2].fc.weight, m[2].fc.bias)
        return m(x)


##################################################################################
# Encoder and Decoders
##################################################################################
class StyleEncoder(nn.Module):
    """networks.py:442-477."""

    def __init__(self, n_downsample, input_dim, dim, style_dim, norm, activ, pad_type):
        super().__init__()
        self.model = []
        self.model += [Conv2dBlock(input_dim, dim, 7, 1, 3, norm=norm, activation=activ, pad_type=pad_type)]
        for i in range(2):
            self.model += [Conv2dBlock(dim, 2 * dim, 4, 2, 1, norm=norm, activation=activ, pad_type=pad_type)]
            dim *= 2
        for i in range(n_downsample - 2):
            self.model += [Conv2dBlock(dim, dim, 4, 2, 1, norm=norm, activation=activ, pad_type=pad_type)]
        self.model += [nn.AdaptiveAvgPool2d(1)]  # global average pooling
        self.model += [nn.Conv2d(dim, style_dim, 1, 1, 0)]
        self.model = nn.Sequential(*self.model)
        self.output_dim = dim
        self.style_dim = style_dim

    def forward(self, x):
        mods = list(self.model)
        a = run_chain(mods[:-2], x, 0)
        pooled = ops.GapFn.apply(a.t)  # [B, dim] fp32
        head = mods[-1]
        s = ops.LinearFn.apply(pooled, head.weight, head.bias, False)
        return s.view(s.shape[0], self.style_dim, 1, 1)


class ContentEncoder(nn.Module):
    """networks.py:480-512."""

    def __init__(self, n_downsample, n_res, input_dim, dim, norm, activ, pad_type):
        super().__init__()
        self.model = []
        self.model += [Conv2dBlock(input_dim, dim, 7, 1, 3, norm=norm, activation=activ, pad_type=pad_type)]
        for i in range(n_downsample):
            self.model += [Conv2dBlock(dim, 2 * dim, 4, 2, 1, norm=norm, activation=activ, pad_type=pad_type)]
            dim *= 2
        self.model += [ResBlocks(n_res, dim, norm=norm, activation=activ, pad_type=pad_type)]
        self.model = nn.Sequential(*self.model)
        self.output_dim = dim

    def forward_act(self, x, out_pad=1) -> Act:
        return run_chain(list(self.model), x, out_pad)

    def forward(self, x):
        a = self.forward_act(x, 0)
        return ops.FromActFn.apply(a.t, self.output_dim, 0)


class Decoder(nn.Module):
    """networks.py:515-563."""

    def __init__(self, n_upsample, n_res, dim, output_dim, res_norm="adain", activ="relu", pad_type="zero"):
        super().__init__()
        self.model = []
        self.model += [ResBlocks(n_res, dim, res_norm, activ, pad_type=pad_type)]
        for i in range(n_upsample):
            self.model += [nn.Upsample(scale_factor=2),
                           Conv2dBlock(dim, dim // 2, 5, 1, 2, norm="ln", activation=activ, pad_type=pad_type)]
            dim //= 2
        self.model += [Conv2dBlock(dim, output_dim, 7, 1, 3, norm="none", activation="tanh", pad_type=pad_type)]
        self.model = nn.Sequential(*self.model)
        self.output_dim = output_dim

    def forward(self, x):
        if not isinstance(x, Act):
            x = Act(ops.ToActFn.apply(x, 1), 1)
        a = run_chain(list(self.model), x, 0)
        if torch.is_tensor(a):  # the 7x7 tanh layer wrote the NCHW fp32 image itself
            return a
        return ops.FromActFn.apply(a.t, self.output_dim, 0)


##################################################################################
# Generator
##################################################################################
class _AdaINMixin:
    def assign_adain_params(self, adain_params, model):
        """networks.py:230-239: per AdaIN layer (module order) columns [:C] -> bias, [C:2C] -> weight.
        The slices stay strided views of the MLP output: the norm kernels index them directly."""
        mods = [m for m in model.modules() if m.__class__.__name__ == "AdaptiveInstanceNorm2d"]
        outs = ops.AdainSplitFn.apply(adain_params, tuple(m.num_features for m in mods))
        for i, m in enumerate(mods):
            m.bias = outs[2 * i]
            m.weight = outs[2 * i + 1]

    def get_num_adain_params(self, model):
        """networks.py:241-247."""
        num_adain_params = 0
        for m in model.modules():
            if m.__class__.__name__ == "AdaptiveInstanceNorm2d":
                num_adain_params += 2 * m.num_features
        return num_adain_params


class AdaINGen(nn.Module, _AdaINMixin):
    """networks.py:170-254."""

    def __init__(self, input_dim, params):
        super().__init__()
        dim = params["dim"]
        style_dim = params["style_dim"]
        n_downsample = params["n_downsample"]
        n_res = params["n_res"]
        activ = params["activ"]
        pad_type = params["pad_type"]
        mlp_dim = params["mlp_dim"]
        self.enc_style = StyleEncoder(4, input_dim, dim, style_dim, norm="none", activ=activ, pad_type=pad_type)
        self.enc_content = ContentEncoder(n_downsample, n_res, input_dim, dim, "in", activ, pad_type=pad_type)
        self.dec = Decoder(n_downsample, n_res, self.enc_content.output_dim, input_dim, res_norm="adain",
                           activ=activ, pad_type=pad_type)
        self.mlp = MLP(style_dim, self.get_num_adain_params(self.dec), mlp_dim, 3, norm="none", activ=activ)

    def forward(self, images):
        content, style_fake = self.encode(images)
        return self.decode(content, style_fake)

    def encode(self, images):
        style_fake = self.enc_style(images)
        content = self.enc_content(images)
        return content, style_fake

    def encode_act(self, images):
        """Fast path: content code stays an Act (halo 1) for the decoders / content L1."""
        return self.enc_content.forward_act(images, 1), self.enc_style(images)

    def decode(self, content, style):
        adain_params = self.mlp(style)
        self.assign_adain_params(adain_params, self.dec)
        return self.dec(content)

    def get_adain_param(self, style):
        return self.mlp(style)


class AdaINGen_double(nn.Module, _AdaINMixin):
    """networks.py:262-388: shared style encoder, two content encoders / decoders / MLPs."""

    def __init__(self, input_dim, params):
        super().__init__()
        dim = params["dim"]
        style_dim = params["style_dim"]
        n_downsample = params["n_downsample"]
        n_res = params["n_res"]
        activ = params["activ"]
        pad_type = params["pad_type"]
        mlp_dim = params["mlp_dim"]
        self.enc_style = StyleEncoder(4, input_dim, dim, style_dim, norm="none", activ=activ, pad_type=pad_type)
        self.enc1_content = ContentEncoder(n_downsample, n_res, input_dim, dim, "in", activ, pad_type=pad_type)
        self.enc2_content = ContentEncoder(n_downsample, n_res, input_dim, dim, "in", activ, pad_type=pad_type)
        self.dec1 = Decoder(n_downsample, n_res, self.enc1_content.output_dim, input_dim, res_norm="adain",
                            activ=activ, pad_type=pad_type)
        self.dec2 = Decoder(n_downsample, n_res, self.enc2_content.output_dim, input_dim, res_norm="adain",
                            activ=activ, pad_type=pad_type)
        self.mlp1 = MLP(style_dim, self.get_num_adain_params(self.dec1), mlp_dim, 3, norm="none", activ=activ)
        self.mlp2 = MLP(style_dim, self.get_num_adain_params(self.dec2), mlp_dim, 3, norm="none", activ=activ)

    def forward(self, images, encoder_name):
        content, style_fake = self.encode(images, encoder_name)
        return self.decode(content, style_fake, encoder_name)

    def _enc(self, encoder_name):
        if encoder_name == 1:
            return self.enc1_content
        if encoder_name == 2:
            return self.enc2_content
        print("wrong value for encoder_name, must be 0 or 1")
        return None

    def encode(self, images, encoder_name):
        enc = self._enc(encoder_name)
        if enc is None:
            return None
        style_fake = self.enc_style(images)
        return enc(images), style_fake

    def encode_act(self, images, encoder_name):
        enc = self._enc(encoder_name)
        return enc.forward_act(images, 1), self.enc_style(images)

    def decode(self, content, style, encoder_name):
        if encoder_name == 1:
            mlp, dec = self.mlp1, self.dec1
        elif encoder_name == 2:
            mlp, dec = self.mlp2, self.dec2
        else:
            print("wrong value for encoder_name, must be 0 or 1")
            return None
        adain_params = mlp(style)
        self.assign_adain_params(adain_params, dec)
        return dec(content)

    def get_adain_param(self, style, encoder_name):
        if encoder_name == 1:
            return self.mlp1(style)
        if encoder_name == 2:
            return self.mlp2(style)
        print("wrong value for encoder_name, must be 0 or 1")
        return None


##################################################################################
# Discriminator
##################################################################################
class MsImageDis(nn.Module):
    """networks.py:20-115: multi-scale LSGAN discriminator."""

    def __init__(self, input_dim, params):
        super().__init__()
        self.n_layer = params["n_layer"]
        self.gan_type = params["gan_type"]
        self.dim = params["dim"]
        self.norm = params["norm"]
        self.activ = params["activ"]
        self.num_scales = params["num_scales"]
        self.pad_type = params["pad_type"]
        self.input_dim = input_dim
        assert self.gan_type == "lsgan", "Unsupported GAN type: {}".format(self.gan_type)
        self.downsample = nn.AvgPool2d(3, stride=2, padding=[1, 1], count_include_pad=False)  # marker only
        self.cnns = nn.ModuleList()
        for _ in range(self.num_scales):
            self.cnns.append(self._make_net())

    def _make_net(self):
        dim = self.dim
        cnn_x = []
        cnn_x += [Conv2dBlock(self.input_dim, dim, 4, 2, 1, norm="none", activation=self.activ,
                              pad_type=self.pad_type)]
        for i in range(self.n_layer - 1):
            cnn_x += [Conv2dBlock(dim, dim * 2, 4, 2, 1, norm=self.norm, activation=self.activ,
                                  pad_type=self.pad_type)]
            dim *= 2
        cnn_x += [nn.Conv2d(dim, 1, 1, 1, 0)]
        return nn.Sequential(*cnn_x)

    # Scale-parallel execution (set by the trainer together with its two-stream mode): the per-scale networks share
    # nothing but the image pyramid, and the deeper layers of the smaller scales launch a handful of CTAs each --
    # so the pyramid is built first and the scales run on their own streams (scale 0 on the caller's), forward and,
    # because autograd replays nodes on their forward stream, backward.
    scale_streams = False

    def _pyramid(self, x):
        xs = [x]
        for _ in range(len(self.cnns) - 1):
            xs.append(ops.AvgPoolFn.apply(xs[-1]))
        return xs

    def _fork_scales(self, fns):
        if not MsImageDis.scale_streams or len(fns) == 1 or not torch.cuda.is_available():
            return [f() for f in fns]
        cur = torch.cuda.current_stream()
        if getattr(self, "_streams", None) is None:
            self._streams = [torch.cuda.Stream() for _ in fns[1:]]
        for st in self._streams:
            st.wait_stream(cur)
        out = [fns[0]()]
        for st, f in zip(self._streams, fns[1:]):
            with torch.cuda.stream(st):
                out.append(f())
            ops.note_side_stream(st)
        for st in self._streams:
            cur.wait_stream(st)
        return out

    def _run(self, x, target: float, frozen: bool = False):
        """Returns (per-scale maps, sum over scales of mean((out - target)^2))."""
        def scale(model, xs):
            mods = list(model)
            a = run_chain(mods[:-1], xs, 0, frozen=frozen)
            head = mods[-1]
            hw = head.weight.detach() if frozen else head.weight
            hb = head.bias.detach() if frozen else head.bias
            return ops.DisHeadFn.apply(a.t, hw, hb, target, 1.0)

        res = self._fork_scales([(lambda m=m, xs=xs: scale(m, xs)) for m, xs in zip(self.cnns, self._pyramid(x))])
        outputs, loss = [], None
        for omap, l in res:
            outputs.append(omap)
            loss = l if loss is None else loss + l
        return outputs, loss

    def forward(self, x):
        return self._run(x, 0.0)[0]

    def calc_dis_loss(self, input_fake, input_real):
        """networks.py:79-101 (lsgan): sum_scales mean(out_fake^2) + mean((out_real-1)^2).
        Fake and real share the weights, so they run as ONE batch of 2B through every scale (half the launches,
        twice the rows per GEMM); only the LSGAN heads see the two halves with their own targets."""
        if input_fake.shape != input_real.shape:
            _, l0 = self._run(input_fake, 0.0)
            _, l1 = self._run(input_real, 1.0)
            return (l0 + l1).squeeze(0)
        b = input_fake.shape[0]
        x = torch.cat([input_fake, input_real], 0)

        def scale(model, xs):
            mods = list(model)
            a = run_chain(mods[:-1], xs, 0)
            head = mods[-1]
            ls = None
            for half, target in ((a.t[:b], 0.0), (a.t[b:], 1.0)):
                _, l = ops.DisHeadFn.apply(half, head.weight, head.bias, target, 1.0)
                ls = l if ls is None else ls + l
            return ls

        res = self._fork_scales([(lambda m=m, xs=xs: scale(m, xs)) for m, xs in zip(self.cnns, self._pyramid(x))])
        loss = res[0]
        for l in res[1:]:
            loss = loss + l
        return loss.squeeze(0)

    def calc_gen_loss(self, input_fake, frozen: bool = False):
        """networks.py:103-115 (lsgan): sum_scales mean((out_fake-1)^2).  `frozen` skips the (wasted)
        discriminator weight gradients that the reference computes and discards in gen_update."""
        _, l = self._run(input_fake, 1.0, frozen=frozen)
        return l.squeeze(0)


class VAEGen(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("VAEGen (networks.py:391-434) is outside the MUNIT hot path (SURVEY.md s2)")
