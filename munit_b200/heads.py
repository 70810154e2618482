"""B200-native drop-in for the domain-adaptation head of cc-ai/MUNIT `scripts/utils.py`: conv3x3 / conv1x1
(:1238-1274), BasicBlock (:1276-1331) and domainClassifier (:1370-1392) -- SURVEY.md 8(f).2.

Same class / attribute names and state_dict keys as the reference (`BasicBlock1.conv1.weight`,
`BasicBlock1.bn1.{weight,bias,running_mean,running_var,num_batches_tracked}`, `BasicBlock1.downsample.0.weight`,
..., `fc.weight`, `fc.bias`); the nn.Conv2d / nn.BatchNorm2d / nn.Linear instances only hold parameters and
buffers.  The content code arrives as an `ops.Act` (NHWC bf16 with the decoder's reflect halo, which MaxPool
skips); the zero-padded 3x3 and the 1x1 convolutions run on the tcgen05 tap-GEMM with implicit zero padding,
BatchNorm2d on the statistics / apply kernels of the generator's norms plus csrc/heads.cu.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from .ops import Act


def _as_cl(conv: nn.Conv2d):
    conv.weight.data = conv.weight.data.contiguous(memory_format=torch.channels_last)
    return conv


def conv3x3(in_planes, out_planes, stride=1, groups=1, dilation=1):
    """3x3 convolution with padding (utils.py:1238-1261)."""
    assert groups == 1 and dilation == 1
    return _as_cl(nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=dilation, groups=groups,
                            bias=False, dilation=dilation))


def conv1x1(in_planes, out_planes, stride=1):
    """1x1 convolution (utils.py:1264-1274)."""
    return _as_cl(nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=stride, bias=False))


def _bn(bn: nn.BatchNorm2d, y, relu: bool, training: bool, frozen: bool):
    """bn(y) (+ReLU) on a raw conv output [N,H,W,C] bf16."""
    g, b = (bn.weight.detach(), bn.bias.detach()) if frozen else (bn.weight, bn.bias)
    if training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return ops.BatchNormFn.apply(y, g, b, bn.running_mean, bn.running_var, training, relu,
                                 0.1 if bn.momentum is None else bn.momentum, bn.eps)


def _conv(conv: nn.Conv2d, layer: ops.ConvLayer, x, frozen: bool):
    w = conv.weight.detach() if frozen else conv.weight
    return ops.ConvFn.apply(x, w, None, layer, "none", 0, 0)


class BasicBlock(nn.Module):
    """utils.py:1276-1331 (torchvision's ResNet BasicBlock): conv3x3-bn-relu-conv3x3-bn (+ conv1x1-bn shortcut
    when the channel count changes), add, relu."""

    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=64, dilation=1,
                 norm_layer=None):
        super().__init__()
        if norm_layer is None:
            norm_layer = nn.BatchNorm2d
        if groups != 1 or base_width != 64:
            raise ValueError("BasicBlock only supports groups=1 and base_width=64")
        if dilation > 1:
            raise NotImplementedError("Dilation > 1 not supported in BasicBlock")
        stride = int(stride)  # the reference passes stride=True (utils.py:1375,1377)
        assert stride == 1 and norm_layer is nn.BatchNorm2d, "the B200 path builds the stride-1 BatchNorm block"
        assert inplanes % 64 == 0 and planes % 64 == 0
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = norm_layer(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = norm_layer(planes)
        self.stride = stride
        self.downsample = downsample
        if stride != 1 or inplanes != planes:
            self.downsample = nn.Sequential(conv1x1(inplanes, planes, stride), norm_layer(planes))
        self._l1 = ops.ConvLayer(inplanes, planes, 3, 1, 0, zpad=1)
        self._l2 = ops.ConvLayer(planes, planes, 3, 1, 0, zpad=1)
        self._ld = ops.ConvLayer(inplanes, planes, 1, 1, 0) if self.downsample is not None else None

    def forward_act(self, x: torch.Tensor, frozen: bool = False) -> torch.Tensor:
        """x: [N,H,W,C] bf16 act without halo -> [N,H,W,planes]."""
        tr = self.training
        out = _bn(self.bn1, _conv(self.conv1, self._l1, x, frozen), True, tr, frozen)
        out = _bn(self.bn2, _conv(self.conv2, self._l2, out, frozen), False, tr, frozen)
        identity = x
        if self.downsample is not None:
            identity = _bn(self.downsample[1], _conv(self.downsample[0], self._ld, x, frozen), False, tr, frozen)
        return ops.AddReluFn.apply(out, identity)

    def forward(self, x):
        out = self.forward_act(ops.ToActFn.apply(x, 0))
        return ops.FromActFn.apply(out, out.shape[3], 0)


class domainClassifier(nn.Module):
    """utils.py:1370-1392: MaxPool(2) -> BasicBlock(256,128) -> MaxPool(2) -> BasicBlock(128,64) -> AvgPool(16)
    -> Linear(64,1) on the content code [N,256,64,64]."""

    def __init__(self, dim):
        super().__init__()
        self.max_pool1 = nn.MaxPool2d(2)
        self.BasicBlock1 = BasicBlock(256, 128, True)
        self.max_pool2 = nn.MaxPool2d(2)
        self.BasicBlock2 = BasicBlock(128, 64, True)
        self.avg_pool = nn.AvgPool2d((16, 16))
        self.fc = nn.Linear(64, 1)
        self.output_dim = dim

    def forward(self, x, frozen: bool = False):
        """x: content code as ops.Act or NCHW fp32.  Returns the reference's `fc(avg_pool(.).squeeze())`:
        [N,1] fp32 ([1] when N == 1)."""
        if not isinstance(x, Act):
            x = Act(ops.ToActFn.apply(x, 0), 0)
        t = ops.MaxPool2Fn.apply(x.t, x.pad)
        t = self.BasicBlock1.forward_act(t, frozen)
        t = ops.MaxPool2Fn.apply(t, 0)
        t = self.BasicBlock2.forward_act(t, frozen)
        n, h, w, c = t.shape
        if (h, w) != (16, 16):
            # AvgPool2d((16,16)) followed by Linear(64,1) on the squeezed map only type-checks for a 16x16 map,
            # i.e. 256x256 images (utils.py:1378-1386)
            raise ValueError(f"domainClassifier expects a 64x64 content code (16x16 before the pooling), got {h * 4}x{w * 4}")
        pooled = ops.GapFn.apply(t)  # AvgPool2d(16) of a 16x16 map
        fw, fb = (self.fc.weight.detach(), self.fc.bias.detach()) if frozen else (self.fc.weight, self.fc.bias)
        out = ops.LinearFn.apply(pooled, fw, fb, False)
        return out.view(-1) if n == 1 else out
