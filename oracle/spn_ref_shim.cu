// TEST INFRASTRUCTURE ONLY.  C-ABI shim around the REFERENCE's own SPN kernels, compiled from where they lie
// (/root/reference/dmb/ops/spn/src/gaterecurrent2dnoind_kernel.cu, never copied into this repo) by
// oracle/build_ref.py into oracle/_ref/libspn_ref.so.  The reference binds these launchers through pybind11 +
// torch::Tensor (gaterecurrent2dnoind_cuda.cpp:7-84); this shim does the same dispatch on raw device pointers so that
// tests/test_gpu_spn_ref.py can pin csrc/scans.cu (and oracle/dmb_oracle.py:spn_scan*) to the reference arithmetic
// on the GPU box, where /root/reference does not exist.
// The reference launches on the legacy default stream and exit()s on a launch error (kernel.cu:542-549).
#include <cuda_runtime.h>

#include "gaterecurrent2dnoind_kernel.h"

extern "C" int spn_ref_forward(int horizontal, int reverse, float* X, float* G1, float* G2, float* G3, float* H, int n,
                               int c, int h, int w) {
    // same dispatch as gaterecurrent2dnoind_forward_cuda (gaterecurrent2dnoind_cuda.cpp:23-40)
    if (horizontal && !reverse) Forward_left_right(n, c, h, w, X, G1, G2, G3, H, horizontal, reverse);
    else if (horizontal && reverse) Forward_right_left(n, c, h, w, X, G1, G2, G3, H, horizontal, reverse);
    else if (!horizontal && !reverse) Forward_top_bottom(n, c, h, w, X, G1, G2, G3, H, horizontal, reverse);
    else Forward_bottom_top(n, c, h, w, X, G1, G2, G3, H, horizontal, reverse);
    return (int)cudaDeviceSynchronize();
}

extern "C" int spn_ref_backward(int horizontal, int reverse, float* H, float* H_diff, float* X, float* G1, float* G2,
                                float* G3, float* X_diff, float* G1_diff, float* G2_diff, float* G3_diff, int n, int c,
                                int h, int w) {
    // same dispatch as gaterecurrent2dnoind_backward_cuda (gaterecurrent2dnoind_cuda.cpp:66-81)
    if (horizontal && !reverse) Backward_left_right(n, c, h, w, X, G1, G2, G3, H, X_diff, G1_diff, G2_diff, G3_diff, H_diff, horizontal, reverse);
    else if (horizontal && reverse) Backward_right_left(n, c, h, w, X, G1, G2, G3, H, X_diff, G1_diff, G2_diff, G3_diff, H_diff, horizontal, reverse);
    else if (!horizontal && !reverse) Backward_top_bottom(n, c, h, w, X, G1, G2, G3, H, X_diff, G1_diff, G2_diff, G3_diff, H_diff, horizontal, reverse);
    else Backward_bottom_top(n, c, h, w, X, G1, G2, G3, H, X_diff, G1_diff, G2_diff, G3_diff, H_diff, horizontal, reverse);
    return (int)cudaDeviceSynchronize();
}
