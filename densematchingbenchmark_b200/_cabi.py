"""ctypes binding of the C-ABI library (include/dmb_b200.h).

PyTorch supplies device memory (`tensor.data_ptr()`) and the current stream; every kernel is
reached through the plain-C entry points -- the same ones a non-Python host would bind.
There is NO fallback: if `csrc/libdmb_b200.so` is missing or an entry point fails, this raises.
"""
import ctypes
import os
import threading
from ctypes import c_double, c_float, c_int, c_int64, c_longlong, c_void_p, c_char_p, POINTER

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# DMB_B200_LIB: an alternative build of the same library (A/B timing of two builds on one GPU box)
LIB_PATH = os.environ.get("DMB_B200_LIB") or os.path.join(_HERE, "csrc", "libdmb_b200.so")

_P = c_void_p
_I = c_int
_F = c_float
_IP = POINTER(c_int)
_D = c_double
_LL = c_longlong

# name -> argtypes, exactly as declared in include/dmb_b200.h
SIGNATURES = {
    "dmb_b200_cat_volume": [_P, _P, _P, _I, _I, _I, _I, _IP, _I, _P],
    "dmb_b200_dif_volume": [_P, _P, _P, _I, _I, _I, _I, _IP, _I, _P],
    "dmb_b200_gwc_volume": [_P, _P, _P, _I, _I, _I, _I, _I, _IP, _I, _P],
    "dmb_b200_corr1d_volume": [_P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    "dmb_b200_warp_volume": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P],
    "dmb_b200_conv3d_direct": [_P, _P, _P, _P, _P, _I, _I, _I, _IP, _IP, _IP, _I, _I, _I, _I, _P],
    "dmb_b200_upsample_regress": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _F, _F, _P, _P],
    "dmb_b200_soft_argmin": [_P, _P, _I, _I, _I, _I, _F, _I, _F, _F, _P, _P, _P],
    "dmb_b200_local_soft_argmin": [_P, _P, _I, _I, _I, _I, _I, _I, _F, _F, _F, _P],
    "dmb_b200_spn_forward": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "dmb_b200_spn_backward": [_P] * 10 + [_I, _I, _I, _I, _I, _I, _P],
    "dmb_b200_sga": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "dmb_b200_lga": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "dmb_b200_cat_volume_blocked": [_P, _P, _P, _P, _I, _I, _I, _I, _IP, _I, _I, _P],
    "dmb_b200_conv3d_tc": [_P, _P, _I, _P, _F, _P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "dmb_b200_conv2d_tc": [_P, _P, _I, _P, _F, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "dmb_b200_blocked_add": [_P, _P, _P, _P, _P, _P, c_int64, _I, _P],
    "dmb_b200_blocked_dot": [_P, _P, _P, _P, _I, _I, c_int64, _I, _P],
    "dmb_b200_conv3d_tc_head": [_P, _P, _P, _F, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "dmb_b200_head_gather": [_P, _P, _P, _I, _I, _I, _I, _P],
    "dmb_b200_debug_set_trace": [_P],
    "dmb_b200_conv3d_tc_schedule": [_I, _I, _I, _I, _I, _IP],
    "dmb_b200_focal_loss_forward": [_P, _P, _P, _F, _P, _P, _I, _I, _I, _I, _F, _F, _F, _F, _P, _P, _P],
    "dmb_b200_focal_loss_backward": [_P, _P, _P, _F, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _F, _F, _P, _P, _P],
    "dmb_b200_conv3d_tc_pack_weights": [_P, _P, _I, _I, _I, _I, _F, _I, _P],
    "dmb_b200_ncdhw_to_blocked": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "dmb_b200_blocked_to_ncdhw": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "dmb_b200_bn_stats": [_P, _P, _I, _I, _LL, _P],
    "dmb_b200_bn_finalize": [_P, _D, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _I, _P],
    "dmb_b200_bn_apply": [_P, _P, _P, _P, _P, _I, _I, _LL, _I, _P],
    "dmb_b200_bn_backward_reduce": [_P, _P, _P, _P, _P, _P, _I, _I, _LL, _P],
    "dmb_b200_bn_backward_apply": [_P, _P, _P, _P, _P, _P, _P, _D, _P, _P, _I, _I, _LL, _P],
    "dmb_b200_conv3d_wgrad": [_P, _P, _P, _I, _I, _I, _IP, _IP, _I, _I, _P],
    "dmb_b200_conv3d_wgrad_tc": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "dmb_b200_ncdhw_to_blocked_wsplit": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "dmb_b200_upsample_deconv_wgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "dmb_b200_upsample_trilinear_backward": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "dmb_b200_soft_argmin_backward": [_P, _P, _P, _I, _I, _I, _I, _F, _I, _F, _F, _P, _P],
    "dmb_b200_cat_volume_backward": [_P, _P, _P, _I, _I, _I, _I, _IP, _I, _P],
    "dmb_b200_peer_alloc": [POINTER(c_void_p)],
    "dmb_b200_peer_free": [_P],
    "dmb_b200_peer_export": [_P, _P],
    "dmb_b200_peer_import": [_P, POINTER(c_void_p)],
    "dmb_b200_peer_close": [_P],
    "dmb_b200_peer_exchange": [POINTER(c_void_p), _I, _I, _LL, _P, _P, _I, _I, _P],
    "dmb_b200_peer_count": [_P, POINTER(c_longlong)],
    "dmb_b200_peer_sum2_f32": [POINTER(c_void_p), _I, _I, _P, _I, _P, _I, _P, _P, _P],
    "dmb_b200_peer_bn_forward": [POINTER(c_void_p), _I, _I, _P, _P, _F, _F, _F, _P, _P, _P, _P, _P, _I, _P],
    "dmb_b200_dif_volume_backward": [_P, _P, _P, _I, _I, _I, _I, _IP, _I, _P],
}
# entry points that do not return a status code
OTHER = {
    "dmb_b200_abi_version": ([], c_int),
    "dmb_b200_last_error": ([], c_char_p),
    "dmb_b200_launch_count": ([], c_int64),
    "dmb_b200_conv3d_tc_weight_bytes": ([_I, _I, _I, _I], c_int64),
    "dmb_b200_conv3d_tc_available": ([], c_int),
    "dmb_b200_sga_set_bidirectional": ([_I], c_int),
    "dmb_b200_peer_buffer_bytes": ([], c_int64),
    "dmb_b200_conv3d_tc_head_floats": ([_I, _I, _I, _I], c_int64),
}

# debug-only entry points that an older build handed in through DMB_B200_LIB may lack
OPTIONAL = {"dmb_b200_debug_set_trace"}

_lib = None


class DmbB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DmbB200Error(
            "CUDA extension %s not found: build it with `python -m densematchingbenchmark_b200.build` "
            "(there is no CPU or PyTorch fallback for this path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        if name in OPTIONAL and not hasattr(lib, name):
            continue
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    for name, (argtypes, restype) in OTHER.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


# devices of the tensors whose pointers were taken (ptr()) since the last call(): the launch, the
# cudaFuncSetAttribute, the SM count and the TMA descriptors inside the library all act on the calling thread's
# CURRENT device, so call() makes the tensors' device current for the duration of the entry point
_tls = threading.local()


def call(name, *args):
    lib = load()
    devs = getattr(_tls, "devs", None)
    _tls.devs = None
    fn = getattr(lib, name)
    if devs:
        if len(devs) > 1:
            raise DmbB200Error("%s: tensor arguments live on different CUDA devices %s" % (name, sorted(devs)))
        dev = next(iter(devs))
        if dev != torch.cuda.current_device():
            with torch.cuda.device(dev):
                rc = fn(*args)
        else:
            rc = fn(*args)
    else:
        rc = fn(*args)
    if rc != 0:
        msg = lib.dmb_b200_last_error()
        raise DmbB200Error("%s failed (%d): %s" % (name, rc, msg.decode() if msg else "?"))


def launch_count():
    return int(load().dmb_b200_launch_count())


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor must be a contiguous CUDA tensor."""
    if t is None:
        return None
    if not t.is_cuda:
        raise DmbB200Error("dmb_b200 kernels need CUDA tensors (got %s); there is no CPU path" % t.device)
    if not t.is_contiguous():
        raise DmbB200Error("dmb_b200 kernels need contiguous tensors")
    devs = getattr(_tls, "devs", None)
    if devs is None:
        devs = _tls.devs = set()
    devs.add(t.device.index if t.device.index is not None else torch.cuda.current_device())
    return c_void_p(t.data_ptr())


def stream(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def int_array(values):
    arr = (c_int * len(values))(*[int(v) for v in values])
    return arr


def f32(t, name="tensor"):
    """The reference path is float32 end to end (cat_fms.py:32 always allocates fp32)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
