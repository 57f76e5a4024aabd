"""3-D conv building blocks with the reference's container layout and a fused CUDA forward.

The reference factories (dmb/modeling/stereo/layers/basic_layers.py:68-216) return
`nn.Sequential(Conv3d|ConvTranspose3d, [BatchNorm3d], [ReLU])`; their child indices (`.0` conv,
`.1` bn) define the checkpoint key names.  `FusedConvUnit` IS such a Sequential (same children,
same keys) but its forward folds the eval-mode BatchNorm into the weights/bias once and runs
    y = relu?( conv(x) + bias + residual? )
as ONE kernel through the C ABI (dmb_b200_conv3d_direct, or the tcgen05 trunk when the
aggregator routes it there).
"""
import torch
import torch.nn as nn

from ....ops import functional as F_
from ....ops.autograd import ConvUnitFn, wants_grad


def consistent_padding_with_dilation(padding, dilation):
    """basic_layers.py:14-28 -- dilation > 1 overrides the padding."""
    if isinstance(dilation, int):
        dilation = (dilation,) * 3
    if isinstance(padding, int):
        padding = (padding,) * 3
    padding = tuple(d if d > 1 else p for p, d in zip(padding, dilation))
    return padding, tuple(dilation)


class FusedConvUnit(nn.Sequential):

    def __init__(self, conv, bn=None, relu=False):
        mods = [conv]
        if bn is not None:
            mods.append(bn)
        if relu:
            mods.append(nn.ReLU(inplace=True))
        super(FusedConvUnit, self).__init__(*mods)
        self._has_bn = bn is not None
        self._has_relu = relu
        self._cache_key = None
        self._cache_val = None
        # training: None = per-process batch statistics; True / a process group = synchronised BatchNorm
        self.sync_group = None

    # -- parameters ---------------------------------------------------------------------
    @property
    def conv(self):
        return self[0]

    @property
    def bn(self):
        return self[1] if self._has_bn else None

    @property
    def transposed(self):
        return isinstance(self[0], nn.ConvTranspose3d)

    def folded(self):
        """(w_packed [K3,Cin,Cout], bias [Cout] or None) with eval-mode BN folded in; cached until
        a parameter/buffer is modified (in place or by load_state_dict) or moved."""
        conv, bn = self.conv, self.bn
        tensors = [conv.weight, conv.bias]
        if bn is not None:
            tensors += [bn.weight, bn.bias, bn.running_mean, bn.running_var]
        key = tuple((t.data_ptr(), t._version, str(t.device)) if t is not None else None for t in tensors)
        if key == self._cache_key:
            return self._cache_val
        with torch.no_grad():
            w = F_.pack_conv_weight(conv.weight.detach(), self.transposed)          # [K3,Cin,Cout]
            b = conv.bias.detach().float() if conv.bias is not None else None
            if bn is not None:
                scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
                w = w * scale.view(1, 1, -1)
                shift = bn.bias.detach().float() - bn.running_mean.float() * scale
                b = shift if b is None else b * scale + shift
            w = w.contiguous()
            b = b.contiguous() if b is not None else None
        self._cache_key, self._cache_val = key, (w, b)
        return w, b

    def geometry(self):
        conv = self.conv
        for name in ("kernel_size", "stride", "padding", "dilation"):
            v = getattr(conv, name)
            if len(set(v)) != 1 and name != "kernel_size":
                raise NotImplementedError("anisotropic %s %s is not supported by the CUDA path" % (name, v))
        if conv.dilation[0] != 1:
            raise NotImplementedError("dilated 3-D convolution is not on the hot path (no shipped config uses it)")
        if conv.groups != 1:
            raise NotImplementedError("grouped 3-D convolution is not supported")
        opad = conv.output_padding[0] if self.transposed else 0
        return tuple(conv.kernel_size), conv.stride[0], conv.padding[0], opad

    # -- forward ------------------------------------------------------------------------
    def forward(self, x, residual=None, relu_after=False):
        """`residual` is added before the (optional) final ReLU requested with `relu_after`;
        the unit's own ReLU (conv3d_bn_relu) cannot be combined with a residual."""
        if self._has_relu and residual is not None:
            raise ValueError("conv+bn+relu unit cannot take a fused residual")
        bn = self.bn
        if self.training and (bn is None or bn.training):
            return self._forward_train(x, residual, relu_after)
        if self.training or wants_grad(x, residual):
            # frozen BatchNorm (`model.train()` followed by `bn.eval()`), or an eval-mode unit whose INPUT carries
            # a gradient: running statistics are folded with differentiable torch ops and the convolution runs
            # through the autograd Function, so input / weight / affine gradients flow like through
            # nn.Sequential(Conv3d, BatchNorm3d.eval()) and no running statistic is touched
            return self._forward_frozen(x, residual, relu_after)
        w, b = self.folded()
        ksize, stride, pad, opad = self.geometry()
        return F_.conv3d_fused(x, w, b, ksize, stride, pad, self.transposed, opad, residual,
                               relu=self._has_relu or relu_after)


def _unit_forward_train(conv, bn, x, residual, relu, sync_group=None, weight=None, bias=None):
    """Training-mode forward of one conv(+bn)(+relu) unit through the autograd Function whose forward and
    backward are the library's kernels (batch statistics, running-stat update as nn.BatchNorm3d).
    `weight` / `bias`: differentiable replacements of the conv's own parameters (frozen-BatchNorm folding)."""
    transposed = isinstance(conv, nn.ConvTranspose3d)
    for name in ("stride", "padding", "dilation"):
        if len(set(getattr(conv, name))) != 1:
            raise NotImplementedError("anisotropic %s is not supported by the CUDA path" % name)
    if conv.dilation[0] != 1 or conv.groups != 1:
        raise NotImplementedError("dilated / grouped 3-D convolution is not on the hot path")
    cfg = dict(transposed=transposed, ksize=tuple(conv.kernel_size), stride=conv.stride[0], pad=conv.padding[0],
               opad=conv.output_padding[0] if transposed else 0, relu=relu, bn=bn, sync_group=sync_group)
    gamma = bn.weight if bn is not None else None
    beta = bn.bias if bn is not None else None
    if weight is None:
        weight, bias = conv.weight, conv.bias
    return ConvUnitFn.apply(x, weight, bias, gamma, beta, residual, cfg)


def _ff_train(self, x, residual=None, relu_after=False):
    return _unit_forward_train(self.conv, self.bn, x, residual, self._has_relu or relu_after, self.sync_group)


def _ff_frozen(self, x, residual=None, relu_after=False):
    conv, bn = self.conv, self.bn
    w, b = conv.weight, conv.bias
    if bn is not None:
        if bn.running_var is None:
            raise NotImplementedError("BatchNorm3d without running statistics cannot run in eval mode on the CUDA path")
        scale = torch.rsqrt(bn.running_var.float() + bn.eps)
        if bn.weight is not None:
            scale = bn.weight * scale
        # Conv3d weight [Cout,Cin,k,k,k]; ConvTranspose3d weight [Cin,Cout,k,k,k]
        w = w * (scale.view(1, -1, 1, 1, 1) if self.transposed else scale.view(-1, 1, 1, 1, 1))
        shift = -bn.running_mean.float() * scale
        if bn.bias is not None:
            shift = shift + bn.bias
        b = shift if b is None else b * scale + shift
    return _unit_forward_train(conv, None, x, residual, self._has_relu or relu_after, None, weight=w, bias=b)


FusedConvUnit._forward_train = _ff_train
FusedConvUnit._forward_frozen = _ff_frozen


def conv3d_bn(batchNorm, in_planes, out_planes, kernel_size=3, stride=1, padding=1, dilation=1, bias=True):
    padding, dilation = consistent_padding_with_dilation(padding, dilation)
    conv = nn.Conv3d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding,
                     dilation=dilation, bias=bias)
    return FusedConvUnit(conv, nn.BatchNorm3d(out_planes) if batchNorm else None, relu=False)


def conv3d_bn_relu(batchNorm, in_planes, out_planes, kernel_size=3, stride=1, padding=1, dilation=1, bias=True):
    padding, dilation = consistent_padding_with_dilation(padding, dilation)
    conv = nn.Conv3d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding,
                     dilation=dilation, bias=bias)
    return FusedConvUnit(conv, nn.BatchNorm3d(out_planes) if batchNorm else None, relu=True)


def deconv3d_bn(batchNorm, in_planes, out_planes, kernel_size=4, stride=2, padding=1, output_padding=0, bias=True):
    conv = nn.ConvTranspose3d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding,
                              output_padding=output_padding, bias=bias)
    return FusedConvUnit(conv, nn.BatchNorm3d(out_planes) if batchNorm else None, relu=False)


def deconv3d_bn_relu(batchNorm, in_planes, out_planes, kernel_size=4, stride=2, padding=1, output_padding=0,
                     bias=True):
    conv = nn.ConvTranspose3d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding,
                              output_padding=output_padding, bias=bias)
    return FusedConvUnit(conv, nn.BatchNorm3d(out_planes) if batchNorm else None, relu=True)


class _PlainCache(object):
    pass


def fused_plain_conv3d(conv, x, residual=None, relu=False):
    """Run a bare nn.Conv3d / nn.ConvTranspose3d (e.g. the 32->1 classifier heads,
    aggregators/PSMNet.py:41-52) through the fused kernel, with an optional residual."""
    if conv.training or wants_grad(x, residual):
        return _unit_forward_train(conv, None, x, residual, relu)
    cache = conv.__dict__.setdefault("_dmb_b200_cache", _PlainCache())
    tensors = [conv.weight, conv.bias]
    key = tuple((t.data_ptr(), t._version, str(t.device)) if t is not None else None for t in tensors)
    if getattr(cache, "key", None) != key:
        transposed = isinstance(conv, nn.ConvTranspose3d)
        with torch.no_grad():
            cache.w = F_.pack_conv_weight(conv.weight.detach(), transposed)
            cache.b = conv.bias.detach().float().contiguous() if conv.bias is not None else None
        cache.key = key
    transposed = isinstance(conv, nn.ConvTranspose3d)
    opad = conv.output_padding[0] if transposed else 0
    return F_.conv3d_fused(x, cache.w, cache.b, tuple(conv.kernel_size), conv.stride[0], conv.padding[0],
                           transposed, opad, residual, relu)
