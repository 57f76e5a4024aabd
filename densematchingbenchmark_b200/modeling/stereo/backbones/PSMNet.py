"""PSMNet feature extractor -- NOT part of the accelerated hot path (2-D
convolutions, 'next' row f.2).  It exists because the reference tree is absent on the GPU box and
BASELINE config 2 times a FULL forward: plain torch/cuDNN modules, laid out so that the state-dict
keys equal the reference's (dmb/modeling/stereo/backbones/PSMNet.py:8-129), producing the
[B,32,H/4,W/4] feature maps the cost-volume path consumes."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _unit(cin, cout, k, stride, pad, dil, relu, bias=False, bn=True):
    pad = dil if dil > 1 else pad          # basic_layers.py:14-28
    mods = [nn.Conv2d(cin, cout, k, stride, pad, dil, bias=bias)]
    if bn:
        mods.append(nn.BatchNorm2d(cout))
    if relu:
        mods.append(nn.ReLU(inplace=True))
    return nn.Sequential(*mods)


class _Residual(nn.Module):
    """Two 3x3 conv units with an identity / projected skip, no ReLU after the add
    (layers/basic_layers.py:219-243)."""

    def __init__(self, bn, cin, cout, stride, downsample, pad, dil):
        super(_Residual, self).__init__()
        self.conv1 = _unit(cin, cout, 3, stride, pad, dil, relu=True, bn=bn)
        self.conv2 = _unit(cout, cout, 3, 1, pad, dil, relu=False, bn=bn)
        self.downsample = downsample

    def forward(self, x):
        y = self.conv2(self.conv1(x))
        return y + (x if self.downsample is None else self.downsample(x))


class PSMNetBackbone(nn.Module):

    def __init__(self, in_planes=3, batch_norm=True):
        super(PSMNetBackbone, self).__init__()
        bn = batch_norm
        self.in_planes = in_planes
        self.batch_norm = bn
        self.firstconv = nn.Sequential(_unit(in_planes, 32, 3, 2, 1, 1, True, bn=bn),
                                       _unit(32, 32, 3, 1, 1, 1, True, bn=bn),
                                       _unit(32, 32, 3, 1, 1, 1, True, bn=bn))
        self._width = 32
        self.layer1 = self._stage(bn, 32, 3, 1, 1, 1)
        self.layer2 = self._stage(bn, 64, 16, 2, 1, 1)
        self.layer3 = self._stage(bn, 128, 3, 1, 1, 1)
        self.layer4 = self._stage(bn, 128, 3, 1, 2, 2)
        for i, win in ((1, 64), (2, 32), (3, 16), (4, 8)):
            setattr(self, "branch%d" % i, nn.Sequential(nn.AvgPool2d((win, win), stride=(win, win)),
                                                        _unit(128, 32, 1, 1, 0, 1, True, bn=bn)))
        self.lastconv = nn.Sequential(_unit(320, 128, 3, 1, 1, 1, True, bn=bn),
                                      nn.Conv2d(128, 32, kernel_size=1, padding=0, stride=1, dilation=1, bias=False))

    def _stage(self, bn, width, blocks, stride, pad, dil):
        down = None
        if stride != 1 or self._width != width:
            down = _unit(self._width, width, 1, stride, 0, 1, relu=False, bias=True, bn=bn)
        layers = [_Residual(bn, self._width, width, stride, down, pad, dil)]
        self._width = width
        layers += [_Residual(bn, width, width, 1, None, pad, dil) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def _forward(self, x):
        half = self.layer1(self.firstconv(x))
        quarter = self.layer2(half)
        deep = self.layer4(self.layer3(quarter))
        size = deep.shape[2:]
        # SPP: the 8/16/32/64 average pools nest exactly (stride == window, floor mode), so each level is a
        # 2x2 pool of the previous one instead of a fresh pass with a 64x64 window over the full map
        # (reference: four independent nn.AvgPool2d, backbones/PSMNet.py:42-57; same windows, same values
        # up to fp32 rounding)
        pooled = []
        level = None
        for i, win in ((4, 8), (3, 16), (2, 32), (1, 64)):
            branch = getattr(self, "branch%d" % i)
            if tuple(branch[0].kernel_size) != (win, win):          # non-standard checkpoint: pool directly
                level_i = branch[0](deep)
            else:
                level = F.avg_pool2d(deep, 8, 8) if level is None else F.avg_pool2d(level, 2, 2)
                level_i = level
            pooled.append(F.interpolate(branch[1](level_i), size, mode='bilinear', align_corners=True))
        return self.lastconv(torch.cat([quarter, deep] + pooled, 1))

    def forward(self, *input):
        if len(input) != 2:
            raise ValueError('expected input length 2 (got {} length input)'.format(len(input)))
        l_img, r_img = input
        if self.training and any(isinstance(m, torch.nn.modules.batchnorm._BatchNorm) and m.training
                                 for m in self.modules()):
            # batch statistics: two passes like the reference (backbones/PSMNet.py:126-127) -- one joint pass would
            # normalise with the statistics of the 2N-image batch and update the running statistics once, not twice
            return self._forward(l_img), self._forward(r_img)
        # eval / frozen BatchNorm: one batched pass over both views instead of two (same result, half the launches)
        both = self._forward(torch.cat([l_img, r_img], 0))
        n = l_img.shape[0]
        return both[:n], both[n:]
