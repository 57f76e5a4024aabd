// Generic direct 3-D convolution / transposed convolution, fp32 NCDHW, fused
// bias + residual + ReLU epilogue.  This is the any-shape SIMT path (GCNet / StereoNet
// aggregators, odd channel counts, bring-up and cross-checking of the tensor-core trunk);
// the PSMNet/AcfNet trunk shapes go through conv3d_tc.cu on tcgen05.
//
// Replaces the cuDNN calls behind conv3d_bn[_relu] / deconv3d_bn
// (dmb/modeling/stereo/layers/basic_layers.py:68-216) after the caller has folded BatchNorm.
// Oracle: oracle/dmb_oracle.py:conv_unit.
#include "common.cuh"

namespace dmb {

struct ConvParams {
    int B, Cin, Cout;
    int Di, Hi, Wi;
    int Do, Ho, Wo;
    int KD, KH, KW;
    int stride, pad, transposed, relu;
    int ci_tile;
};

// map an output coordinate + tap to an input coordinate; returns false when the tap does not
// contribute (outside the input, or -- transposed -- not on the stride lattice)
__device__ __forceinline__ bool tap_coord(int o, int k, int n_in, int stride, int pad, int transposed, int& i) {
    if (!transposed) {
        i = o * stride - pad + k;
        return i >= 0 && i < n_in;
    }
    const int t = o + pad - k;
    if (t < 0) return false;
    i = t / stride;
    return (t - i * stride) == 0 && i < n_in;
}

// CO_T output channels per thread, one output voxel per thread, 128 threads per CTA.
// Weights of the current input-channel chunk live in shared memory ([ci][tap][co], read as
// broadcast float4), inputs come through the read-only path (every input value is re-used by
// up to 27 taps x CO_T channels out of L1).
template <int CO_T, int KS>  // KS: 3 => 3x3x3 fully unrolled, 0 => runtime kernel size
__global__ void __launch_bounds__(128) conv3d_direct_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias,
                                                            const float* __restrict__ residual, float* __restrict__ y,
                                                            ConvParams p) {
    extern __shared__ __align__(16) float ws[];  // [ci_tile][K3][CO_T]
    const int KD = KS ? KS : p.KD, KH = KS ? KS : p.KH, KW = KS ? KS : p.KW;
    const int K3 = KD * KH * KW;
    const size_t Si = (size_t)p.Di * p.Hi * p.Wi;
    const size_t So = (size_t)p.Do * p.Ho * p.Wo;
    const int b = blockIdx.z;
    const int co0 = blockIdx.y * CO_T;
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = s < So;
    int ow = 0, oh = 0, od = 0;
    if (active) {
        ow = s % p.Wo;
        oh = (s / p.Wo) % p.Ho;
        od = s / ((size_t)p.Wo * p.Ho);
    }

    // per-dimension tap validity and offsets (KS==3: registers; otherwise recomputed in-loop)
    int offd[3], offh[3], offw[3];
    bool vd[3], vh[3], vw[3];
    if (KS == 3) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int i;
            vd[k] = active && tap_coord(od, k, p.Di, p.stride, p.pad, p.transposed, i);
            offd[k] = vd[k] ? i * p.Hi * p.Wi : 0;
            vh[k] = tap_coord(oh, k, p.Hi, p.stride, p.pad, p.transposed, i);
            offh[k] = vh[k] ? i * p.Wi : 0;
            vw[k] = tap_coord(ow, k, p.Wi, p.stride, p.pad, p.transposed, i);
            offw[k] = vw[k] ? i : 0;
        }
    }

    float acc[CO_T];
#pragma unroll
    for (int j = 0; j < CO_T; ++j) acc[j] = 0.f;

    for (int ci0 = 0; ci0 < p.Cin; ci0 += p.ci_tile) {
        const int nci = min(p.ci_tile, p.Cin - ci0);
        __syncthreads();
        for (int i = threadIdx.x; i < nci * K3 * CO_T; i += blockDim.x) {
            const int co = i % CO_T;
            const int tap = (i / CO_T) % K3;
            const int ci = i / (CO_T * K3);
            const int cog = co0 + co;
            ws[i] = (cog < p.Cout) ? __ldg(w + ((size_t)tap * p.Cin + (ci0 + ci)) * p.Cout + cog) : 0.f;
        }
        __syncthreads();
        for (int ci = 0; ci < nci; ++ci) {
            const float* xp = x + ((size_t)b * p.Cin + ci0 + ci) * Si;
            const float* wci = ws + (size_t)ci * K3 * CO_T;
            if (KS == 3) {
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) {
                    if (!vd[kd]) continue;
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        if (!vh[kh]) continue;
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            if (!vw[kw]) continue;
                            const float v = __ldg(xp + offd[kd] + offh[kh] + offw[kw]);
                            const float* wt = wci + ((kd * 3 + kh) * 3 + kw) * CO_T;
                            if (CO_T % 4 == 0) {
#pragma unroll
                                for (int j = 0; j < CO_T / 4; ++j) {
                                    const float4 w4 = reinterpret_cast<const float4*>(wt)[j];
                                    acc[4 * j + 0] = fmaf(v, w4.x, acc[4 * j + 0]);
                                    acc[4 * j + 1] = fmaf(v, w4.y, acc[4 * j + 1]);
                                    acc[4 * j + 2] = fmaf(v, w4.z, acc[4 * j + 2]);
                                    acc[4 * j + 3] = fmaf(v, w4.w, acc[4 * j + 3]);
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < CO_T; ++j) acc[j] = fmaf(v, wt[j], acc[j]);
                            }
                        }
                    }
                }
            } else if (active) {
                for (int kd = 0; kd < KD; ++kd) {
                    int id;
                    if (!tap_coord(od, kd, p.Di, p.stride, p.pad, p.transposed, id)) continue;
                    for (int kh = 0; kh < KH; ++kh) {
                        int ih;
                        if (!tap_coord(oh, kh, p.Hi, p.stride, p.pad, p.transposed, ih)) continue;
                        for (int kw = 0; kw < KW; ++kw) {
                            int iw;
                            if (!tap_coord(ow, kw, p.Wi, p.stride, p.pad, p.transposed, iw)) continue;
                            const float v = __ldg(xp + ((size_t)id * p.Hi + ih) * p.Wi + iw);
                            const float* wt = wci + ((kd * KH + kh) * KW + kw) * CO_T;
#pragma unroll
                            for (int j = 0; j < CO_T; ++j) acc[j] = fmaf(v, wt[j], acc[j]);
                        }
                    }
                }
            }
        }
    }

    if (!active) return;
#pragma unroll
    for (int j = 0; j < CO_T; ++j) {
        const int co = co0 + j;
        if (co < p.Cout) {
            const size_t o = ((size_t)b * p.Cout + co) * So + s;
            float v = acc[j];
            if (bias) v += __ldg(bias + co);
            if (residual) v += __ldg(residual + o);
            if (p.relu) v = fmaxf(v, 0.f);
            y[o] = v;
        }
    }
}

template <int CO_T>
static int launch_direct(const float* x, const float* w, const float* bias, const float* res, float* y, ConvParams p,
                         void* stream) {
    const int K3 = p.KD * p.KH * p.KW;
    int ci_tile = (40 * 1024 / 4) / (K3 * CO_T);
    if (ci_tile < 1) ci_tile = 1;
    if (ci_tile > p.Cin) ci_tile = p.Cin;
    if (ci_tile > 16) ci_tile = 16;
    p.ci_tile = ci_tile;
    const size_t smem = (size_t)ci_tile * K3 * CO_T * 4;
    DMB_REQUIRE(smem <= 200 * 1024, "conv3d_direct: kernel volume %d too large", K3);
    const size_t So = (size_t)p.Do * p.Ho * p.Wo;
    dim3 grid((unsigned)cdiv(So, 128), (unsigned)cdiv(p.Cout, CO_T), p.B);
    DMB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv3d_direct: grid too large");
    const bool k3 = (p.KD == 3 && p.KH == 3 && p.KW == 3);
    if (k3) {
        if (smem > 48 * 1024)
            DMB_CUDA(cudaFuncSetAttribute(conv3d_direct_kernel<CO_T, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv3d_direct_kernel<CO_T, 3><<<grid, 128, smem, as_stream(stream)>>>(x, w, bias, res, y, p);
    } else {
        if (smem > 48 * 1024)
            DMB_CUDA(cudaFuncSetAttribute(conv3d_direct_kernel<CO_T, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv3d_direct_kernel<CO_T, 0><<<grid, 128, smem, as_stream(stream)>>>(x, w, bias, res, y, p);
    }
    return check_launch("conv3d_direct_kernel");
}

}  // namespace dmb

using namespace dmb;

extern "C" int dmb_b200_conv3d_direct(const float* x, const float* w_packed, const float* bias, const float* residual,
                                      float* y, int B, int Cin, int Cout, const int* dims_in, const int* dims_out,
                                      const int* ksize, int stride, int pad, int transposed, int relu, void* stream) {
    DMB_REQUIRE(x && w_packed && y && dims_in && dims_out && ksize, "conv3d_direct: null pointer");
    DMB_REQUIRE(B > 0 && Cin > 0 && Cout > 0, "conv3d_direct: non-positive channel/batch count");
    DMB_REQUIRE(stride >= 1 && pad >= 0, "conv3d_direct: bad stride/pad");
    ConvParams p;
    p.B = B; p.Cin = Cin; p.Cout = Cout;
    p.Di = dims_in[0]; p.Hi = dims_in[1]; p.Wi = dims_in[2];
    p.Do = dims_out[0]; p.Ho = dims_out[1]; p.Wo = dims_out[2];
    p.KD = ksize[0]; p.KH = ksize[1]; p.KW = ksize[2];
    p.stride = stride; p.pad = pad; p.transposed = transposed ? 1 : 0; p.relu = relu ? 1 : 0;
    p.ci_tile = 1;
    DMB_REQUIRE(p.Di > 0 && p.Hi > 0 && p.Wi > 0 && p.Do > 0 && p.Ho > 0 && p.Wo > 0, "conv3d_direct: empty volume");
    DMB_REQUIRE(p.KD > 0 && p.KH > 0 && p.KW > 0, "conv3d_direct: empty kernel");
    DMB_REQUIRE((size_t)p.Di * p.Hi * p.Wi < (1u << 31), "conv3d_direct: input plane exceeds 2^31 elements");
    // output extent must be consistent with the convolution arithmetic
    for (int a = 0; a < 3; ++a) {
        const int ni = dims_in[a], no = dims_out[a], k = ksize[a];
        if (!transposed) {
            const int expect = (ni + 2 * pad - k) / stride + 1;
            DMB_REQUIRE(no == expect, "conv3d_direct: output dim %d is %d, expected %d", a, no, expect);
        } else {
            const int lo = (ni - 1) * stride - 2 * pad + k;   // output_padding 0 .. stride-1 allowed
            DMB_REQUIRE(no >= lo && no < lo + stride, "conv3d_direct: transposed output dim %d is %d, expected %d..%d", a,
                        no, lo, lo + stride - 1);
        }
    }
    if (Cout >= 32) return launch_direct<32>(x, w_packed, bias, residual, y, p, stream);
    if (Cout >= 16) return launch_direct<16>(x, w_packed, bias, residual, y, p, stream);
    if (Cout >= 4) return launch_direct<4>(x, w_packed, bias, residual, y, p, stream);
    return launch_direct<1>(x, w_packed, bias, residual, y, p, stream);
}
