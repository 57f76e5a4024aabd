// tcgen05 implicit-GEMM trunk -- placeholder entry points (the real kernels replace this file).
#include "common.cuh"
using namespace dmb;
extern "C" int dmb_b200_conv3d_tc(const void*, const void*, const void*, const void*, const float*, const void*,
                                  const void*, void*, void*, float*, int, int, int, const int*, const int*, int, int,
                                  void*) {
    return fail(DMB_ERR_UNSUPPORTED, "conv3d_tc: not built yet");
}
extern "C" int dmb_b200_conv3d_tc_pack_weights(const float*, void*, void*, int, int, void*) {
    return fail(DMB_ERR_UNSUPPORTED, "conv3d_tc_pack_weights: not built yet");
}
extern "C" int64_t dmb_b200_conv3d_tc_weight_bytes(int Cin, int Cout) { return (int64_t)27 * (Cout < 16 ? 16 : Cout) * Cin * 2; }
extern "C" int dmb_b200_conv3d_tc_available(void) { return 0; }
