"""Hourglass (reference: cost_processors/utils/hourglass.py:8-86) on fused conv units: every
add / ReLU of the reference forward rides in the epilogue of the conv that produces its operand."""
import torch.nn as nn

from ...layers.basic_layers import conv3d_bn, conv3d_bn_relu, deconv3d_bn


class Hourglass(nn.Module):

    def __init__(self, in_planes, batch_norm=True):
        super(Hourglass, self).__init__()
        self.batch_norm = batch_norm
        c = in_planes
        self.conv1 = conv3d_bn_relu(batch_norm, c, c * 2, kernel_size=3, stride=2, padding=1, bias=False)
        self.conv2 = conv3d_bn(batch_norm, c * 2, c * 2, kernel_size=3, stride=1, padding=1, bias=False)
        self.conv3 = conv3d_bn_relu(batch_norm, c * 2, c * 2, kernel_size=3, stride=2, padding=1, bias=False)
        self.conv4 = conv3d_bn_relu(batch_norm, c * 2, c * 2, kernel_size=3, stride=1, padding=1, bias=False)
        self.conv5 = deconv3d_bn(batch_norm, c * 2, c * 2, kernel_size=3, padding=1, output_padding=1, stride=2,
                                 bias=False)
        self.conv6 = deconv3d_bn(batch_norm, c * 2, c, kernel_size=3, padding=1, output_padding=1, stride=2,
                                 bias=False)

    def forward(self, x, presqu=None, postsqu=None, out_residual=None):
        """Returns (out, pre, post) like the reference.  `out_residual` (extension) is added to
        `out` inside conv6's epilogue -- the aggregators' `out + cost0`."""
        out = self.conv1(x)
        pre = self.conv2(out, residual=postsqu, relu_after=True)          # relu(conv2 [+ postsqu])
        out = self.conv3(pre)
        out = self.conv4(out)
        post = self.conv5(out, residual=presqu if presqu is not None else pre, relu_after=True)
        out = self.conv6(post, residual=out_residual)
        return out, pre, post
