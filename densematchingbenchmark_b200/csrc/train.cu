// Training-mode kernels of the cost-volume path (BASELINE config 5: AcfNet / PSMNet training step).
//
// The reference trains through autograd over cuDNN / ATen ops: Conv3d + BatchNorm3d(batch statistics)
// + ReLU (dmb/modeling/stereo/layers/basic_layers.py:68-216), F.interpolate / ConvTranspose3d upsampling
// (aggregators/PSMNet.py:75-88, AcfNet.py:55-57,81-83), softmax regression
// (disp_predictors/faster_soft_argmin.py:51-75) and the slice-assign volume builder (cat_fms.py:7-48).
// Here every forward piece has a hand-written backward:
//   * batch-norm statistics / apply / backward as two-pass per-channel reductions (double accumulators,
//     so that a multi-GPU run can all-reduce the raw sums between the passes = SyncBN),
//   * conv dgrad = the forward direct kernel run with the roles of the weights swapped (conv3d_direct.cu),
//   * conv wgrad = conv3d_wgrad_k3_kernel below (register-blocked 8x8 outer products per tap),
//   * trilinear / learned-deconv upsampling backward, soft-argmin backward, cat-volume backward.
// fp32 NCDHW throughout: this is the first correct training path (SIMT); the tcgen05 trunk is
// inference-only this round.  Oracle: oracle/dmb_oracle.py (train_* functions, torch autograd).
#include "common.cuh"

namespace dmb {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of two per-thread partials, added atomically (double) to dst_a / dst_b
template <int NT>
__device__ __forceinline__ void block_atomic_add2(float a_f, float b_f, double* dst_a, double* dst_b) {
    __shared__ double sa[NT / 32], sb[NT / 32];
    double a = warp_sum_d((double)a_f), b = warp_sum_d((double)b_f);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        sa[warp] = a;
        sb[warp] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
#pragma unroll
        for (int i = 0; i < NT / 32; ++i) {
            ta += sa[i];
            tb += sb[i];
        }
        atomicAdd(dst_a, ta);
        if (dst_b) atomicAdd(dst_b, tb);
    }
}

constexpr int kRedChunk = 8192;   // elements of one (b, c) plane reduced by one CTA

// ---- batch-norm forward -------------------------------------------------------------------------
// sums[c] += sum z, sums[C + c] += sum z^2 over this CTA's chunk of plane (b, c)
template <bool VEC>
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ z, double* __restrict__ sums, int C,
                                                       size_t S) {
    const int c = blockIdx.y, b = blockIdx.z;
    const float* p = z + ((size_t)b * C + c) * S;
    const size_t s0 = (size_t)blockIdx.x * kRedChunk;
    const size_t s1 = (s0 + kRedChunk < S) ? s0 + kRedChunk : S;
    float a = 0.f, q = 0.f;
    if (VEC) {
        for (size_t s = s0 + 4 * threadIdx.x; s < s1; s += 4 * 256) {   // S % 4 == 0: whole float4s
            const float4 v = ldg_f4(p + s);
            a += (v.x + v.y) + (v.z + v.w);
            q = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, q))));
        }
    } else {
        for (size_t s = s0 + threadIdx.x; s < s1; s += 256) {
            const float v = __ldg(p + s);
            a += v;
            q = fmaf(v, v, q);
        }
    }
    block_atomic_add2<256>(a, q, sums + c, sums + C + c);
}

// mean / invstd / fused scale+shift from the (possibly all-reduced) sums; running statistics updated
// exactly like nn.BatchNorm3d (momentum, unbiased running variance)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ scale,
                                   float* __restrict__ shift, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = sums[c] / count;
    double var = sums[C + c] / count - m * m;
    if (var < 0.0) var = 0.0;
    const double is = 1.0 / sqrt(var + (double)eps);
    const double g = gamma ? (double)gamma[c] : 1.0;
    const double bt = beta ? (double)beta[c] : 0.0;
    mean[c] = (float)m;
    invstd[c] = (float)is;
    scale[c] = (float)(g * is);
    shift[c] = (float)(bt - m * g * is);
    if (running_mean) running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * m);
    if (running_var) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
}

// y = relu?( z * scale[c] + shift[c] + residual )
template <bool VEC>
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ z, const float* __restrict__ scale,
                                                       const float* __restrict__ shift,
                                                       const float* __restrict__ res, float* __restrict__ y, int C,
                                                       size_t S, int relu) {
    const int c = blockIdx.y, b = blockIdx.z;
    const size_t base = ((size_t)b * C + c) * S;
    const float sc = __ldg(scale + c), sh = __ldg(shift + c);
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (VEC) {
        if (4 * i >= S) return;
        float4 v = ldg_f4(z + base + 4 * i);
        v.x = fmaf(v.x, sc, sh); v.y = fmaf(v.y, sc, sh); v.z = fmaf(v.z, sc, sh); v.w = fmaf(v.w, sc, sh);
        if (res) {
            const float4 r = ldg_f4(res + base + 4 * i);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        *reinterpret_cast<float4*>(y + base + 4 * i) = v;
    } else {
        if (i >= S) return;
        float v = fmaf(__ldg(z + base + i), sc, sh);
        if (res) v += __ldg(res + base + i);
        if (relu) v = fmaxf(v, 0.f);
        y[base + i] = v;
    }
}

// ---- batch-norm backward ------------------------------------------------------------------------
// g = dy * (y > 0) when the unit ended in a ReLU (y = forward output), else g = dy.
// sums[c] += sum g ; sums[C + c] += sum g * xhat  (xhat = (z - mean) * invstd; skipped when mean == null)
template <bool VEC>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                            const float* __restrict__ z,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ invstd,
                                                            double* __restrict__ sums, int C, size_t S) {
    const int c = blockIdx.y, b = blockIdx.z;
    const size_t base = ((size_t)b * C + c) * S;
    const size_t s0 = (size_t)blockIdx.x * kRedChunk;
    const size_t s1 = (s0 + kRedChunk < S) ? s0 + kRedChunk : S;
    const float m = mean ? __ldg(mean + c) : 0.f, is = mean ? __ldg(invstd + c) : 0.f;
    float a = 0.f, q = 0.f;
    if (VEC) {
        for (size_t s = s0 + 4 * threadIdx.x; s < s1; s += 4 * 256) {
            float4 g = ldg_f4(dy + base + s);
            if (y) {
                const float4 o = ldg_f4(y + base + s);
                g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f;
                g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
            }
            a += (g.x + g.y) + (g.z + g.w);
            if (mean) {
                const float4 v = ldg_f4(z + base + s);
                q = fmaf(g.x, (v.x - m) * is, q); q = fmaf(g.y, (v.y - m) * is, q);
                q = fmaf(g.z, (v.z - m) * is, q); q = fmaf(g.w, (v.w - m) * is, q);
            }
        }
    } else {
        for (size_t s = s0 + threadIdx.x; s < s1; s += 256) {
            float g = __ldg(dy + base + s);
            if (y) g = __ldg(y + base + s) > 0.f ? g : 0.f;
            a += g;
            if (mean) q = fmaf(g, (__ldg(z + base + s) - m) * is, q);
        }
    }
    block_atomic_add2<256>(a, q, sums + c, mean ? sums + C + c : nullptr);
}

// dz = gamma * invstd * (g - sum_g / N - xhat * sum_g_xhat / N)   (mean != null)
// dz = g                                                            (no batch norm)
// dres (optional) = g : the gradient flowing into the fused residual input
template <bool VEC>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                           const float* __restrict__ z,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ invstd,
                                                           const float* __restrict__ gamma,
                                                           const double* __restrict__ sums, double count,
                                                           float* __restrict__ dz, float* __restrict__ dres, int C,
                                                           size_t S) {
    const int c = blockIdx.y, b = blockIdx.z;
    const size_t base = ((size_t)b * C + c) * S;
    float m = 0.f, is = 0.f, k = 1.f, ma = 0.f, mb = 0.f;
    if (mean) {
        m = __ldg(mean + c);
        is = __ldg(invstd + c);
        k = (gamma ? __ldg(gamma + c) : 1.f) * is;
        ma = (float)(sums[c] / count);
        mb = (float)(sums[C + c] / count);
    }
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (VEC) {
        if (4 * i >= S) return;
        const size_t o = base + 4 * i;
        float4 g = ldg_f4(dy + o);
        if (y) {
            const float4 f = ldg_f4(y + o);
            g.x = f.x > 0.f ? g.x : 0.f; g.y = f.y > 0.f ? g.y : 0.f;
            g.z = f.z > 0.f ? g.z : 0.f; g.w = f.w > 0.f ? g.w : 0.f;
        }
        if (dres) *reinterpret_cast<float4*>(dres + o) = g;
        if (dz) {
            float4 r = g;
            if (mean) {
                const float4 v = ldg_f4(z + o);
                r.x = k * (g.x - ma - (v.x - m) * is * mb); r.y = k * (g.y - ma - (v.y - m) * is * mb);
                r.z = k * (g.z - ma - (v.z - m) * is * mb); r.w = k * (g.w - ma - (v.w - m) * is * mb);
            }
            *reinterpret_cast<float4*>(dz + o) = r;
        }
    } else {
        if (i >= S) return;
        const size_t o = base + i;
        float g = __ldg(dy + o);
        if (y) g = __ldg(y + o) > 0.f ? g : 0.f;
        if (dres) dres[o] = g;
        if (dz) dz[o] = mean ? k * (g - ma - (__ldg(z + o) - m) * is * mb) : g;
    }
}

// ---- conv weight gradient, 3x3x3 --------------------------------------------------------------
// dw[tap][ci][co] += sum_{b, o} dz[b, co, o] * x[b, ci, o*stride - pad + tap]
// One CTA owns a 32(co) x 32(ci) channel tile and walks row segments (b, od, oh, 32 output columns).
// Per segment the dz tile [32 w][32 co] and the nine (kd, kh) input rows [w'][32 ci] are staged with
// 4-byte cp.async copies (transposed to channel-innermost, zero-filled outside the volume), double
// buffered.  432 threads = 27 taps x 4 co-groups x 4 ci-groups hold an 8x8 accumulator block each:
// per output column two 16-byte shared loads per operand feed 64 FMAs.
struct WgradParams {
    int B, Cin, Cout;
    int Di, Hi, Wi, Do, Ho, Wo;
    int pad;
    int nco_t;          // number of 32-wide output-channel tiles
    int nseg_w;         // segments per output row
    long long nseg;     // B * Do * Ho * nseg_w
};

constexpr int WG_THREADS = 448;
constexpr int WG_PITCH = 36;   // floats per staged column: 32 channels + 4 pad (keeps 16-byte alignment)
constexpr int WG_W = 32;

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, bool valid) {
    const int n = valid ? 4 : 0;   // src-size 0: destination is zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(n)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int STRIDE>
__global__ void __launch_bounds__(WG_THREADS, 1) conv3d_wgrad_k3_kernel(const float* __restrict__ x,
                                                                        const float* __restrict__ dz,
                                                                        float* __restrict__ dw, WgradParams p) {
    constexpr int WR = (WG_W - 1) * STRIDE + 3;       // staged input columns per row
    constexpr int XS = 9 * WR * WG_PITCH;             // floats: nine (kd, kh) rows
    constexpr int BUF = XS + WG_W * WG_PITCH;         // + dz tile
    extern __shared__ __align__(16) float sm[];       // [2][BUF]
    const int tid = threadIdx.x;
    const int co0 = (blockIdx.y % p.nco_t) * 32, ci0 = (blockIdx.y / p.nco_t) * 32;

    auto issue = [&](long long seg, float* buf) {
        const int wseg = (int)(seg % p.nseg_w);
        long long r = seg / p.nseg_w;
        const int oh = (int)(r % p.Ho);
        r /= p.Ho;
        const int od = (int)(r % p.Do);
        const int b = (int)(r / p.Do);
        const int w0 = wseg * WG_W;
        for (int e = tid; e < 32 * WG_W; e += WG_THREADS) {
            const int w = e & 31, co = e >> 5;
            const bool ok = (co0 + co < p.Cout) && (w0 + w < p.Wo);
            const float* src = dz + ((((size_t)b * p.Cout + co0 + co) * p.Do + od) * p.Ho + oh) * p.Wo + w0 + w;
            cp_async4(buf + XS + w * WG_PITCH + co, ok ? src : dz, ok);
        }
        const int iw0 = w0 * STRIDE - p.pad;
        for (int e = tid; e < 9 * 32 * WR; e += WG_THREADS) {
            const int j = e % WR;
            const int t = e / WR;
            const int ci = t & 31, rr = t >> 5;
            const int id = od * STRIDE - p.pad + rr / 3, ih = oh * STRIDE - p.pad + rr % 3, iw = iw0 + j;
            const bool ok = (ci0 + ci < p.Cin) && id >= 0 && id < p.Di && ih >= 0 && ih < p.Hi && iw >= 0 && iw < p.Wi;
            const float* src = x + ((((size_t)b * p.Cin + ci0 + ci) * p.Di + id) * p.Hi + ih) * p.Wi + iw;
            cp_async4(buf + (rr * WR + j) * WG_PITCH + ci, ok ? src : x, ok);
        }
    };

    const bool worker = tid < 27 * 16;
    const int tap = worker ? (tid >> 4) : 0, cog = (tid >> 2) & 3, cig = tid & 3;
    const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
    const int x_off = ((kd * 3 + kh) * WR + kw) * WG_PITCH + cig * 8;
    const int d_off = XS + cog * 8;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    long long seg = blockIdx.x;
    int cur = 0;
    if (seg < p.nseg) issue(seg, sm);
    cp_async_commit();
    for (; seg < p.nseg; seg += gridDim.x) {
        const long long nxt = seg + gridDim.x;
        if (nxt < p.nseg) issue(nxt, sm + (cur ^ 1) * BUF);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (worker) {
            const float* xr = sm + cur * BUF + x_off;
            const float* dr = sm + cur * BUF + d_off;
#pragma unroll 2
            for (int w = 0; w < WG_W; ++w) {
                const float4 a0 = *reinterpret_cast<const float4*>(dr + w * WG_PITCH);
                const float4 a1 = *reinterpret_cast<const float4*>(dr + w * WG_PITCH + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(xr + w * STRIDE * WG_PITCH);
                const float4 b1 = *reinterpret_cast<const float4*>(xr + w * STRIDE * WG_PITCH + 4);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
            }
        }
        __syncthreads();
        cur ^= 1;
    }
    cp_async_wait<0>();
    if (!worker) return;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int ci = ci0 + cig * 8 + j;
        if (ci >= p.Cin) continue;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int co = co0 + cog * 8 + i;
            if (co < p.Cout) atomicAdd(dw + ((size_t)tap * p.Cin + ci) * p.Cout + co, acc[i][j]);
        }
    }
}

// weight gradient of AcfNet's learned upsampling ConvTranspose3d(1, 1, 8, stride 4, pad 2):
// dw[k] += sum_i low[i] * dfull[4 i - 2 + k].  Output-stationary: every dfull element is read exactly once
// (coalesced, the 400 MB stream that bounds the kernel); per dimension an output index o pairs with the two
// taps k = r, r + 4 (r = (o + 2) & 3) and the sources i = (o + 2) >> 2, i - 1, so a thread whose (od, oh, ow)
// keep their residues mod 4 owns 8 fixed taps: 8 register accumulators for the whole kernel, one shared-memory
// and one global atomic round at the end.  CTA = 128 (w) x 4 (h) threads, blockIdx.y = od & 3.
__global__ void __launch_bounds__(512) wgrad_up8_kernel(const float* __restrict__ low, const float* __restrict__ dfull,
                                                        float* __restrict__ dw, int B, int Dl, int Hl, int Wl, int D,
                                                        int H, int W) {
    __shared__ float bins[512];
    const int tx = threadIdx.x & 127, ty = threadIdx.x >> 7;
    bins[threadIdx.x] = 0.f;
    __syncthreads();
    const int pd = blockIdx.y;
    const int rd = (pd + 2) & 3, rh = (ty + 2) & 3, rw = (tx + 2) & 3;
    const int wchunks = (W + 127) / 128;
    const int rows = B * Dl * Hl;                                   // D == 4 Dl, H == 4 Hl; one row group = 4 output rows
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    // (the first version decoded a flat 64-bit item index with three divisions PER ELEMENT: ~200 instructions for 9
    // loads and 8 FMAs -- instruction bound at 0.36 TB/s; rows are decoded once now, columns walked with increments)
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int oh4 = row % Hl;
        const int r2 = row / Hl;
        const int od4 = r2 % Dl;
        const int b = r2 / Dl;
        const int od = 4 * od4 + pd, oh = 4 * oh4 + ty;
        const int ida = (od + 2) >> 2, iha = (oh + 2) >> 2;
        const float* grow = dfull + (((size_t)b * D + od) * H + oh) * W;
        const float* lb = low + (size_t)b * Dl * Hl * Wl;
        // the (up to) four source rows of this output row: (ida - a, iha - c), null when outside
        const float* lrow[4];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int id = ida - a, ih = iha - c;
                lrow[a * 2 + c] = (id >= 0 && id < Dl && ih >= 0 && ih < Hl) ? lb + ((size_t)id * Hl + ih) * Wl : nullptr;
            }
        for (int wc = 0; wc < wchunks; ++wc) {
            const int ow = wc * 128 + tx;
            if (ow >= W) break;
            const float g = __ldcs(grow + ow);
            const int iwa = (ow + 2) >> 2;
#pragma unroll
            for (int ac = 0; ac < 4; ++ac) {
                if (!lrow[ac]) continue;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int iw = iwa - e;
                    if (iw < 0 || iw >= Wl) continue;
                    acc[ac * 2 + e] = fmaf(__ldg(lrow[ac] + iw), g, acc[ac * 2 + e]);
                }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int e = 0; e < 2; ++e)
                atomicAdd(&bins[((rd + 4 * a) * 8 + (rh + 4 * c)) * 8 + (rw + 4 * e)], acc[(a * 2 + c) * 2 + e]);
    __syncthreads();
    const float v = bins[threadIdx.x];
    if (v != 0.f) atomicAdd(dw + threadIdx.x, v);
}

// ---- trilinear upsampling backward (align_corners=True) ----------------------------------------
// weight with which output index o reads source index i (the forward's two-tap blend, regress.cu)
__device__ __forceinline__ float tri_weight(int o, int i, float s, int L) {
    const float f = s * (float)o;
    const int i0 = (int)f;
    if (i0 >= L - 1) return i == L - 1 ? 1.f : 0.f;
    const float l1 = f - (float)i0;
    return i == i0 ? 1.f - l1 : (i == i0 + 1 ? l1 : 0.f);
}
__device__ __forceinline__ void tri_range(int i, float s, int O, int& lo, int& hi) {
    if (s <= 0.f) {
        lo = 0;
        hi = O - 1;
        return;
    }
    lo = (int)floorf((float)(i - 1) / s) - 1;
    hi = (int)ceilf((float)(i + 1) / s) + 1;
    lo = lo < 0 ? 0 : lo;
    hi = hi > O - 1 ? O - 1 : hi;
}

// one thread per source voxel gathers its (<= ~10^3) contributing output voxels
__global__ void __launch_bounds__(128) upsample_trilinear_bwd_kernel(const float* __restrict__ dcost,
                                                                     float* __restrict__ dlow, int Dl, int Hl, int Wl,
                                                                     int D, int H, int W) {
    const int wl = blockIdx.x * 32 + (threadIdx.x & 31);
    const int hl = blockIdx.y * 4 + (threadIdx.x >> 5);
    const int b = blockIdx.z / Dl, dl = blockIdx.z % Dl;
    if (wl >= Wl || hl >= Hl) return;
    const float sd = D > 1 ? (float)(Dl - 1) / (float)(D - 1) : 0.f;
    const float sy = H > 1 ? (float)(Hl - 1) / (float)(H - 1) : 0.f;
    const float sx = W > 1 ? (float)(Wl - 1) / (float)(W - 1) : 0.f;
    int d_lo, d_hi, y_lo, y_hi, x_lo, x_hi;
    tri_range(dl, sd, D, d_lo, d_hi);
    tri_range(hl, sy, H, y_lo, y_hi);
    tri_range(wl, sx, W, x_lo, x_hi);
    const float* g = dcost + (size_t)b * D * H * W;
    float acc = 0.f;
    for (int d = d_lo; d <= d_hi; ++d) {
        const float wd = tri_weight(d, dl, sd, Dl);
        if (wd == 0.f) continue;
        for (int yy = y_lo; yy <= y_hi; ++yy) {
            const float wy = tri_weight(yy, hl, sy, Hl) * wd;
            if (wy == 0.f) continue;
            const float* row = g + ((size_t)d * H + yy) * W;
            float racc = 0.f;
            for (int xx = x_lo; xx <= x_hi; ++xx) racc = fmaf(tri_weight(xx, wl, sx, Wl), __ldg(row + xx), racc);
            acc = fmaf(wy, racc, acc);
        }
    }
    dlow[(((size_t)b * Dl + dl) * Hl + hl) * Wl + wl] = acc;
}

// ---- soft-argmin backward ----------------------------------------------------------------------
// disp = sum_d softmax(alpha c)_d v_d  =>  dc_d = alpha p_d (v_d - disp) g ;  normalize == 0: dc_d = alpha v_d g
__global__ void __launch_bounds__(256) soft_argmin_bwd_kernel(const float* __restrict__ cost,
                                                              const float* __restrict__ gdisp,
                                                              float* __restrict__ dcost, int D, size_t HW,
                                                              float alpha, int normalize, float start_disp,
                                                              float disp_step, const float* __restrict__ dvals) {
    const size_t px = (size_t)blockIdx.x * 256 + threadIdx.x;
    const int b = blockIdx.y;
    if (px >= HW) return;
    const float* c = cost + (size_t)b * D * HW + px;
    float* o = dcost + (size_t)b * D * HW + px;
    const float g = __ldg(gdisp + (size_t)b * HW + px);
    if (!normalize) {
        for (int d = 0; d < D; ++d) {
            const float v = dvals ? __ldg(dvals + d) : fmaf((float)d, disp_step, start_disp);
            o[(size_t)d * HW] = alpha * v * g;
        }
        return;
    }
    float m = -INFINITY;
    for (int d = 0; d < D; ++d) m = fmaxf(m, __ldg(c + (size_t)d * HW) * alpha);
    float s = 0.f, t = 0.f;
    for (int d = 0; d < D; ++d) {
        const float e = __expf(__ldg(c + (size_t)d * HW) * alpha - m);
        const float v = dvals ? __ldg(dvals + d) : fmaf((float)d, disp_step, start_disp);
        s += e;
        t = fmaf(e, v, t);
    }
    const float inv = 1.f / s, disp = t * inv, ag = alpha * g * inv;
    for (int d = 0; d < D; ++d) {
        const float e = __expf(__ldg(c + (size_t)d * HW) * alpha - m);
        const float v = dvals ? __ldg(dvals + d) : fmaf((float)d, disp_step, start_disp);
        o[(size_t)d * HW] = ag * e * (v - disp);
    }
}

// ---- cat-volume backward -----------------------------------------------------------------------
// volume[c < C][k][y][x] = left[c][y][x], volume[C + c][k][y][x] = right[c][y][x - i_k] where 0 <= x - i_k < W
// (cat_fms.py:36-44) => dleft[c][y][x] = sum_k valid * dvol[c][k][y][x];
//                       dright[c][y][u] = sum_k [0 <= u + i_k < W] dvol[C + c][k][y][u + i_k]
constexpr int kMaxDispBwd = 256;
struct DispListBwd {
    int d[kMaxDispBwd];
};
// DIF = false: backward of cat_fms (dvol [B,2C,D,H,W]: left half -> dleft, right half -> dright);
// DIF = true : backward of dif_fms (dvol [B,C,D,H,W] = d(ref - tgt): dleft = +sum, dright = -sum over the same channels)
template <bool DIF>
__global__ void __launch_bounds__(128) cat_volume_bwd_kernel(const float* __restrict__ dvol, float* __restrict__ dleft,
                                                             float* __restrict__ dright, int C, int H, int W, int D,
                                                             DispListBwd dl) {
    const int xw = blockIdx.x * 128 + threadIdx.x;
    const int yc = blockIdx.y;              // y + H * c2
    const int y = yc % H, c2 = yc / H;      // c2 in [0, 2C)
    const int b = blockIdx.z;
    if (xw >= W) return;
    const size_t plane = (size_t)H * W;
    const float* v = DIF ? dvol + (((size_t)b * C + (c2 % C)) * D) * plane + (size_t)y * W
                         : dvol + (((size_t)b * 2 * C + c2) * D) * plane + (size_t)y * W;
    float acc = 0.f;
    if (c2 < C) {
        for (int k = 0; k < D; ++k) {
            const int src = xw - dl.d[k];
            if (src >= 0 && src < W) acc += __ldg(v + (size_t)k * plane + xw);
        }
        dleft[(((size_t)b * C + c2) * H + y) * W + xw] = acc;
    } else {
        for (int k = 0; k < D; ++k) {
            const int xo = xw + dl.d[k];
            if (xo >= 0 && xo < W) acc += __ldg(v + (size_t)k * plane + xo);
        }
        dright[(((size_t)b * C + (c2 - C)) * H + y) * W + xw] = DIF ? -acc : acc;
    }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace dmb

using namespace dmb;

extern "C" int dmb_b200_bn_stats(const float* z, double* sums, int B, int C, long long S, void* stream) {
    DMB_REQUIRE(z && sums, "bn_stats: null pointer");
    DMB_REQUIRE(B > 0 && C > 0 && S > 0 && C <= 65535 && B <= 65535, "bn_stats: bad dimensions");
    dim3 grid((unsigned)cdiv(S, kRedChunk), C, B);
    if (S % 4 == 0 && aligned16(z))
        bn_stats_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(z, sums, C, (size_t)S);
    else
        bn_stats_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(z, sums, C, (size_t)S);
    return check_launch("bn_stats_kernel");
}

extern "C" int dmb_b200_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float eps,
                                    float momentum, float* running_mean, float* running_var, float* mean,
                                    float* invstd, float* scale, float* shift, int C, void* stream) {
    DMB_REQUIRE(sums && mean && invstd && scale && shift, "bn_finalize: null pointer");
    DMB_REQUIRE(C > 0 && count > 0.0, "bn_finalize: bad dimensions");
    bn_finalize_kernel<<<(unsigned)cdiv(C, 128), 128, 0, as_stream(stream)>>>(sums, count, gamma, beta, eps, momentum,
                                                                             running_mean, running_var, mean, invstd,
                                                                             scale, shift, C);
    return check_launch("bn_finalize_kernel");
}

extern "C" int dmb_b200_bn_apply(const float* z, const float* scale, const float* shift, const float* residual,
                                 float* y, int B, int C, long long S, int relu, void* stream) {
    DMB_REQUIRE(z && scale && shift && y, "bn_apply: null pointer");
    DMB_REQUIRE(B > 0 && C > 0 && S > 0 && C <= 65535 && B <= 65535, "bn_apply: bad dimensions");
    const bool vec = S % 4 == 0 && aligned16(z) && aligned16(y) && (!residual || aligned16(residual));
    dim3 grid((unsigned)cdiv(vec ? S / 4 : S, 256), C, B);
    if (vec)
        bn_apply_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(z, scale, shift, residual, y, C, (size_t)S, relu);
    else
        bn_apply_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(z, scale, shift, residual, y, C, (size_t)S, relu);
    return check_launch("bn_apply_kernel");
}

extern "C" int dmb_b200_bn_backward_reduce(const float* dy, const float* y_relu, const float* z, const float* mean,
                                           const float* invstd, double* sums, int B, int C, long long S,
                                           void* stream) {
    DMB_REQUIRE(dy && sums, "bn_backward_reduce: null pointer");
    DMB_REQUIRE(!mean || (z && invstd), "bn_backward_reduce: mean given without z / invstd");
    DMB_REQUIRE(B > 0 && C > 0 && S > 0 && C <= 65535 && B <= 65535, "bn_backward_reduce: bad dimensions");
    const bool vec = S % 4 == 0 && aligned16(dy) && (!y_relu || aligned16(y_relu)) && (!z || aligned16(z));
    dim3 grid((unsigned)cdiv(S, kRedChunk), C, B);
    if (vec)
        bn_bwd_reduce_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(dy, y_relu, z, mean, invstd, sums, C, (size_t)S);
    else
        bn_bwd_reduce_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(dy, y_relu, z, mean, invstd, sums, C, (size_t)S);
    return check_launch("bn_bwd_reduce_kernel");
}

extern "C" int dmb_b200_bn_backward_apply(const float* dy, const float* y_relu, const float* z, const float* mean,
                                          const float* invstd, const float* gamma, const double* sums, double count,
                                          float* dz, float* dres, int B, int C, long long S, void* stream) {
    DMB_REQUIRE(dy && (dz || dres), "bn_backward_apply: null pointer");
    DMB_REQUIRE(!mean || (z && invstd && sums && count > 0.0), "bn_backward_apply: incomplete batch-norm state");
    DMB_REQUIRE(B > 0 && C > 0 && S > 0 && C <= 65535 && B <= 65535, "bn_backward_apply: bad dimensions");
    const bool vec = S % 4 == 0 && aligned16(dy) && (!y_relu || aligned16(y_relu)) && (!z || aligned16(z)) &&
                     (!dz || aligned16(dz)) && (!dres || aligned16(dres));
    dim3 grid((unsigned)cdiv(vec ? S / 4 : S, 256), C, B);
    if (vec)
        bn_bwd_apply_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(dy, y_relu, z, mean, invstd, gamma, sums, count,
                                                                       dz, dres, C, (size_t)S);
    else
        bn_bwd_apply_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(dy, y_relu, z, mean, invstd, gamma, sums, count,
                                                                        dz, dres, C, (size_t)S);
    return check_launch("bn_bwd_apply_kernel");
}

template <int STRIDE>
static int launch_wgrad(const float* x, const float* dz, float* dw, WgradParams p, void* stream) {
    constexpr int WR = (WG_W - 1) * STRIDE + 3;
    constexpr size_t smem = 2 * (size_t)(9 * WR * WG_PITCH + WG_W * WG_PITCH) * sizeof(float);
    static_assert(smem <= 227 * 1024, "wgrad staging exceeds shared memory");
    DMB_CUDA(cudaFuncSetAttribute(conv3d_wgrad_k3_kernel<STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = p.nco_t * (int)cdiv(p.Cin, 32);
    long long per_tile = sm_count() / tiles;
    if (per_tile < 1) per_tile = 1;
    if (per_tile > p.nseg) per_tile = p.nseg;
    dim3 grid((unsigned)per_tile, tiles, 1);
    conv3d_wgrad_k3_kernel<STRIDE><<<grid, WG_THREADS, smem, as_stream(stream)>>>(x, dz, dw, p);
    return check_launch("conv3d_wgrad_k3_kernel");
}

// dw_packed [27][Cin][Cout] must be zero-initialised by the caller (CTAs accumulate with atomics)
extern "C" int dmb_b200_conv3d_wgrad(const float* x, const float* dz, float* dw_packed, int B, int Cin, int Cout,
                                     const int* dims_in, const int* dims_out, int stride, int pad, void* stream) {
    DMB_REQUIRE(x && dz && dw_packed && dims_in && dims_out, "conv3d_wgrad: null pointer");
    DMB_REQUIRE(B > 0 && Cin > 0 && Cout > 0, "conv3d_wgrad: non-positive channel/batch count");
    DMB_REQUIRE(stride == 1 || stride == 2, "conv3d_wgrad: stride %d not supported (1 or 2)", stride);
    DMB_REQUIRE(pad >= 0, "conv3d_wgrad: negative padding");
    WgradParams p;
    p.B = B; p.Cin = Cin; p.Cout = Cout;
    p.Di = dims_in[0]; p.Hi = dims_in[1]; p.Wi = dims_in[2];
    p.Do = dims_out[0]; p.Ho = dims_out[1]; p.Wo = dims_out[2];
    p.pad = pad;
    DMB_REQUIRE(p.Di > 0 && p.Hi > 0 && p.Wi > 0 && p.Do > 0 && p.Ho > 0 && p.Wo > 0, "conv3d_wgrad: empty volume");
    for (int a = 0; a < 3; ++a) {
        const int expect = (dims_in[a] + 2 * pad - 3) / stride + 1;
        DMB_REQUIRE(dims_out[a] == expect, "conv3d_wgrad: output dim %d is %d, expected %d", a, dims_out[a], expect);
    }
    p.nco_t = (int)cdiv(Cout, 32);
    p.nseg_w = (int)cdiv(p.Wo, WG_W);
    p.nseg = (long long)B * p.Do * p.Ho * p.nseg_w;
    DMB_REQUIRE(p.nco_t * cdiv(Cin, 32) <= 65535, "conv3d_wgrad: too many channel tiles");
    return stride == 1 ? launch_wgrad<1>(x, dz, dw_packed, p, stream) : launch_wgrad<2>(x, dz, dw_packed, p, stream);
}

// dw [8*8*8] must be zero-initialised by the caller
extern "C" int dmb_b200_upsample_deconv_wgrad(const float* cost_low, const float* dcost, float* dw, int B, int Dl,
                                              int Hl, int Wl, int D, int H, int W, void* stream) {
    DMB_REQUIRE(cost_low && dcost && dw, "upsample_deconv_wgrad: null pointer");
    DMB_REQUIRE(B > 0 && Dl > 0 && Hl > 0 && Wl > 0, "upsample_deconv_wgrad: empty volume");
    DMB_REQUIRE(D == 4 * Dl && H == 4 * Hl && W == 4 * Wl, "upsample_deconv_wgrad: output must be 4x the input");
    DMB_REQUIRE((long long)B * Dl * Hl < (1ll << 31), "upsample_deconv_wgrad: volume too large");
    const long long items = (long long)B * Dl * Hl;  // row groups
    long long ctas = (long long)sm_count();          // x 4 depth phases = 4 CTAs of 512 threads per SM
    if (ctas > items) ctas = items;
    dim3 grid((unsigned)ctas, 4, 1);
    wgrad_up8_kernel<<<grid, 512, 0, as_stream(stream)>>>(cost_low, dcost, dw, B, Dl, Hl, Wl, D, H, W);
    return check_launch("wgrad_up8_kernel");
}

extern "C" int dmb_b200_upsample_trilinear_backward(const float* dcost, float* dcost_low, int B, int Dl, int Hl, int Wl,
                                                    int D, int H, int W, void* stream) {
    DMB_REQUIRE(dcost && dcost_low, "upsample_trilinear_backward: null pointer");
    DMB_REQUIRE(B > 0 && Dl > 0 && Hl > 0 && Wl > 0 && D > 0 && H > 0 && W > 0, "upsample_trilinear_backward: empty volume");
    DMB_REQUIRE((long long)B * Dl <= 65535, "upsample_trilinear_backward: B * Dl too large");
    dim3 grid((unsigned)cdiv(Wl, 32), (unsigned)cdiv(Hl, 4), (unsigned)(B * Dl));
    DMB_REQUIRE(grid.y <= 65535, "upsample_trilinear_backward: Hl too large");
    upsample_trilinear_bwd_kernel<<<grid, 128, 0, as_stream(stream)>>>(dcost, dcost_low, Dl, Hl, Wl, D, H, W);
    return check_launch("upsample_trilinear_bwd_kernel");
}

extern "C" int dmb_b200_soft_argmin_backward(const float* cost, const float* grad_disp, float* dcost, int B, int D, int H,
                                             int W, float alpha, int normalize, float start_disp, float disp_step,
                                             const float* disp_values, void* stream) {
    DMB_REQUIRE(cost && grad_disp && dcost, "soft_argmin_backward: null pointer");
    DMB_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && B <= 65535, "soft_argmin_backward: bad dimensions");
    const size_t HW = (size_t)H * W;
    dim3 grid((unsigned)cdiv(HW, 256), B);
    soft_argmin_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(cost, grad_disp, dcost, D, HW, alpha, normalize,
                                                                start_disp, disp_step, disp_values);
    return check_launch("soft_argmin_bwd_kernel");
}

static int volume_backward(const float* dvol, float* dleft, float* dright, int B, int C, int H, int W,
                           const int* disp_idx_host, int D, bool dif, void* stream) {
    DMB_REQUIRE(dvol && dleft && dright && disp_idx_host, "cat_volume_backward: null pointer");
    DMB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0, "cat_volume_backward: non-positive dimension");
    DMB_REQUIRE(D <= kMaxDispBwd, "cat_volume_backward: at most %d disparity samples", kMaxDispBwd);
    DMB_REQUIRE((long long)H * 2 * C <= 65535 && B <= 65535, "cat_volume_backward: grid dimension too large");
    DispListBwd dl;
    for (int i = 0; i < D; ++i) {
        int d = disp_idx_host[i];
        dl.d[i] = d >= W ? W : (d <= -W ? -W : d);
    }
    dim3 grid((unsigned)cdiv(W, 128), (unsigned)(H * 2 * C), B);
    if (dif) cat_volume_bwd_kernel<true><<<grid, 128, 0, as_stream(stream)>>>(dvol, dleft, dright, C, H, W, D, dl);
    else cat_volume_bwd_kernel<false><<<grid, 128, 0, as_stream(stream)>>>(dvol, dleft, dright, C, H, W, D, dl);
    return check_launch("cat_volume_bwd_kernel");
}

extern "C" int dmb_b200_cat_volume_backward(const float* dvol, float* dleft, float* dright, int B, int C, int H, int W,
                                            const int* disp_idx_host, int D, void* stream) {
    return volume_backward(dvol, dleft, dright, B, C, H, W, disp_idx_host, D, false, stream);
}

extern "C" int dmb_b200_dif_volume_backward(const float* dvol, float* dleft, float* dright, int B, int C, int H, int W,
                                            const int* disp_idx_host, int D, void* stream) {
    return volume_backward(dvol, dleft, dright, B, C, H, W, disp_idx_host, D, true, stream);
}
