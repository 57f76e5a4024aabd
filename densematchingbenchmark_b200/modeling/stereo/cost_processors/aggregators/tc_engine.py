"""Routing of the PSM/Acf trunk onto the tcgen05 kernels (csrc/conv3d_tc.cu).  Filled in once the
kernels are in the library; until then the trunk runs on the direct kernels."""
from ..... import _cabi as C


def tc_supported(trunk, raw_cost):
    try:
        return bool(C.load().dmb_b200_conv3d_tc_available()) and _shape_ok(trunk, raw_cost)
    except Exception:
        return False


def _shape_ok(trunk, raw_cost):
    return False


def run_trunk_tc(trunk, raw_cost):
    raise NotImplementedError
