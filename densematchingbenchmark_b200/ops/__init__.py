"""Host-side wrappers over the C-ABI kernels (the `dmb.ops` mirror plus the functional layer
the stereo components are built from)."""
from .functional import (  # noqa: F401
    disp_indices, cat_volume, dif_volume, gwc_volume, warp_volume, conv3d_fused, pack_conv_weight,
    upsample_regress, soft_argmin, local_soft_argmin, sga, lga,
)
from .spn import GateRecurrent2dnoind, GateRecurrent2dnoindFunction  # noqa: F401
from .ganet import SGA, LGA  # noqa: F401
