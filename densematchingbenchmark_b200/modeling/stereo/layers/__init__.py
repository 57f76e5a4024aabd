from .basic_layers import (  # noqa: F401
    FusedConvUnit, conv3d_bn, conv3d_bn_relu, deconv3d_bn, deconv3d_bn_relu, fused_plain_conv3d,
)
