// Raw cost-volume builders: concatenation, difference, group-wise correlation, warped
// ("fast_mode") variants, the concatenation volume written straight into the tensor-core trunk's
// blocked 16-bit layout, and the layout converters at the trunk boundary.
//
// Reference behaviour restated by oracle/dmb_oracle.py (cat_volume, dif_volume, gwc_volume,
// warp_volume); reference sources: dmb/modeling/stereo/cost_processors/utils/cat_fms.py:7-82,
// dif_fms.py:7-86, dmb/modeling/stereo/layers/inverse_warp_3d.py:4-52.
//
// All of these are HBM-write-bound: one feature row (<= a few KB) fans out into D rows of the
// volume.  The row is staged ONCE per CTA in shared memory with a 1-D bulk async copy (TMA
// engine, UBLKCP) and re-read from there for every disparity; the volume is written with
// 16-byte streaming stores.
#include <cuda_fp16.h>

#include "common.cuh"

namespace dmb {

constexpr int kMaxDisp = 256;
struct DispList {
    int d[kMaxDisp];
};

// valid output columns for integer disparity d (cat_fms.py:36-44)
__device__ __forceinline__ bool col_valid(int x, int d, int W) { return d >= 0 ? (x >= d) : (x < W + d); }

// ---------------------------------------------------------------------------------------
// cat / dif volume, NCDHW fp32.
// One CTA per source row (b, channel, y).  MODE 0: concat (channel index runs over 2C, the
// CTA copies either the left or the right row); MODE 1: difference (both rows staged).
// ---------------------------------------------------------------------------------------
template <int MODE, bool VEC>
__global__ void __launch_bounds__(256) volume_rows_kernel(const float* __restrict__ left,
                                                          const float* __restrict__ right,
                                                          float* __restrict__ out, int C, int H, int W, int D,
                                                          int k0, int Dtotal, DispList dl) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* row_a = reinterpret_cast<float*>(smem_raw);            // left row (or the only row)
    float* row_b = row_a + ((W + 3) & ~3);                        // right row (MODE 1)
    __shared__ uint64_t bar;

    const int y = blockIdx.x;
    const int c = blockIdx.y;
    const int b = blockIdx.z;
    const int CO = (MODE == 0) ? 2 * C : C;
    const bool is_right = (MODE == 0) && (c >= C);
    const int cs = is_right ? c - C : c;
    const float* src_a = ((MODE == 0 && is_right) ? right : left) + ((size_t)(b * C + cs) * H + y) * W;
    const float* src_b = right + ((size_t)(b * C + cs) * H + y) * W;

    if (VEC) {
        const uint32_t bytes = W * 4;
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, MODE == 1 ? 2 * bytes : bytes);
            bulk_g2s(row_a, src_a, bytes, &bar);
            if (MODE == 1) bulk_g2s(row_b, src_b, bytes, &bar);
        }
        mbar_wait(&bar, 0);
    } else {
        for (int x = threadIdx.x; x < W; x += blockDim.x) {
            row_a[x] = src_a[x];
            if (MODE == 1) row_b[x] = src_b[x];
        }
        __syncthreads();
    }

    float* out_row0 = out + (((size_t)(b * CO + c) * Dtotal + k0) * H + y) * W;
    const size_t kstride = (size_t)H * W;

    if (VEC) {
        // a thread owns 4 consecutive columns and walks the disparities: the left / minuend values stay in registers,
        // only the shifted right row is re-read from shared memory (no index divisions in the loop: the flat-index
        // version spent ~40 instructions per 16-byte store and stalled at 3.9 TB/s)
        for (int x = threadIdx.x << 2; x < W; x += blockDim.x << 2) {
            const bool own_a = (MODE == 1) || !is_right;
            float a[4] = {0.f, 0.f, 0.f, 0.f};
            if (own_a) {
#pragma unroll
                for (int j = 0; j < 4; ++j) a[j] = row_a[x + j];
            }
            float* o = out_row0 + x;
#pragma unroll 4
            for (int k = 0; k < D; ++k, o += kstride) {
                const int d = dl.d[k];
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int xx = x + j;
                    float val = 0.f;
                    if (col_valid(xx, d, W)) {
                        if (MODE == 0)
                            val = is_right ? row_a[xx - d] : a[j];
                        else
                            val = a[j] - row_b[xx - d];
                    }
                    v[j] = val;
                }
                st_cs_f4(o, make_float4(v[0], v[1], v[2], v[3]));
            }
        }
    } else {
        for (int i = threadIdx.x; i < D * W; i += blockDim.x) {
            const int k = i / W;
            const int xx = i - k * W;
            const int d = dl.d[k];
            float val = 0.f;
            if (col_valid(xx, d, W)) {
                if (MODE == 0)
                    val = is_right ? row_a[xx - d] : row_a[xx];
                else
                    val = row_a[xx] - row_b[xx - d];
            }
            out_row0[k * kstride + xx] = val;
        }
    }
}

template <int MODE>
static int launch_volume_rows(const float* left, const float* right, float* out, int B, int C, int H, int W,
                              const int* disp_idx_host, int D, void* stream) {
    DMB_REQUIRE(left && right && out && disp_idx_host, "volume: null pointer");
    DMB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0, "volume: non-positive dimension");
    DMB_REQUIRE(H <= 65535 && (MODE == 0 ? 2 * C : C) <= 65535 && B <= 65535, "volume: grid dimension too large");
    const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(left) | reinterpret_cast<uintptr_t>(right) |
                                       reinterpret_cast<uintptr_t>(out)) % 16 == 0);
    const size_t smem = (size_t)((W + 3) & ~3) * 4 * (MODE == 1 ? 2 : 1);
    DMB_REQUIRE(smem <= 200 * 1024, "volume: feature row too wide (W=%d)", W);
    if (smem > 48 * 1024) {
        DMB_CUDA(cudaFuncSetAttribute(volume_rows_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DMB_CUDA(cudaFuncSetAttribute(volume_rows_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 grid(H, MODE == 0 ? 2 * C : C, B);
    for (int k0 = 0; k0 < D; k0 += kMaxDisp) {
        DispList dl;
        const int n = (D - k0 < kMaxDisp) ? D - k0 : kMaxDisp;
        for (int i = 0; i < n; ++i) {
            dl.d[i] = disp_idx_host[k0 + i];
            // |d| >= W: nothing valid; clamp so that index arithmetic stays in range
            if (dl.d[i] >= W) dl.d[i] = W;
            if (dl.d[i] <= -W) dl.d[i] = -W;
        }
        int vthreads = (int)(((W / 4) + 31) / 32) * 32;          // one thread per 4 columns
        if (vthreads > 256) vthreads = 256;
        if (vec)
            volume_rows_kernel<MODE, true><<<grid, vthreads, smem, as_stream(stream)>>>(left, right, out, C, H, W, n, k0, D, dl);
        else
            volume_rows_kernel<MODE, false><<<grid, 256, smem, as_stream(stream)>>>(left, right, out, C, H, W, n, k0, D, dl);
        int rc = check_launch("volume_rows_kernel");
        if (rc) return rc;
    }
    return DMB_OK;
}

// ---------------------------------------------------------------------------------------
// group-wise correlation, NCDHW fp32.  One CTA per (b, group, y): the group's CPG left and
// right rows are staged in shared memory, every thread produces 4 consecutive columns of one
// disparity row.
// ---------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256) gwc_rows_kernel(const float* __restrict__ left, const float* __restrict__ right,
                                                       float* __restrict__ out, int C, int G, int H, int W, int D,
                                                       int k0, int Dtotal, DispList dl) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int cpg = C / G;
    const int Wp = (W + 3) & ~3;
    float* sl = reinterpret_cast<float*>(smem_raw);   // [cpg][Wp]
    float* sr = sl + cpg * Wp;                        // [cpg][Wp]
    __shared__ uint64_t bar;

    const int y = blockIdx.x, g = blockIdx.y, b = blockIdx.z;
    const size_t plane = (size_t)H * W;
    const float* l0 = left + ((size_t)(b * C + g * cpg) * H + y) * W;
    const float* r0 = right + ((size_t)(b * C + g * cpg) * H + y) * W;

    if (VEC) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, 2u * cpg * W * 4u);
            for (int c = 0; c < cpg; ++c) {
                bulk_g2s(sl + c * Wp, l0 + c * plane, W * 4, &bar);
                bulk_g2s(sr + c * Wp, r0 + c * plane, W * 4, &bar);
            }
        }
        mbar_wait(&bar, 0);
    } else {
        for (int i = threadIdx.x; i < cpg * W; i += blockDim.x) {
            const int c = i / W, x = i - c * W;
            sl[c * Wp + x] = l0[c * plane + x];
            sr[c * Wp + x] = r0[c * plane + x];
        }
        __syncthreads();
    }

    float* out_row0 = out + (((size_t)(b * G + g) * Dtotal + k0) * H + y) * W;
    const float inv = (float)cpg;
    const int W4 = Wp >> 2;
    for (int i = threadIdx.x; i < D * W4; i += blockDim.x) {
        const int k = i / W4;
        const int x = (i - k * W4) << 2;
        const int d = dl.d[k];
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < cpg; ++c) {
            const float* lr = sl + c * Wp + x;
            const float* rr = sr + c * Wp + x - d;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int xx = x + j;
                if (xx < W && col_valid(xx, d, W)) acc[j] = fmaf(lr[j], rr[j], acc[j]);
            }
        }
        float* o = out_row0 + (size_t)k * plane + x;
        if (VEC) {
            st_cs_f4(o, make_float4(acc[0] / inv, acc[1] / inv, acc[2] / inv, acc[3] / inv));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (x + j < W) o[j] = acc[j] / inv;
        }
    }
}

// Fast path for the common disparity list d_k = d_0 + k (unit step): every thread produces a 4 (x) by
// 8 (k) block of outputs, so one staged left vector (4 values) and one 11-wide window of the right row
// feed 32 FMAs -- 12 shared loads per 32 FMAs instead of 8 per 4.  The right rows are staged with PAD
// zeros on both sides: "shifted column outside the image" needs no branch (cat_fms.py:36-44 validity
// is exactly 0 <= x - d < W).
// ALIGNED (d0 % 4 == 0): the right-row window is fetched as three aligned 16-byte loads starting one column earlier
// -- the 11 scalar loads at a 16-byte lane stride were 4-way bank conflicted (44 + 4 shared-memory wavefronts per 32
// FMAs: the kernel sat at the shared-memory bandwidth, 0.45 of the HBM peak; now 12 + 4).
template <int CPG, bool ALIGNED, bool CORR = false>   // channels per group (0 = runtime); CORR: correlation1d_cost epilogue
__global__ void __launch_bounds__(256) gwc_rows_unit_kernel(const float* __restrict__ left, const float* __restrict__ right,
                                                            float* __restrict__ out, int C, int G, int H, int W, int D,
                                                            int d0, int PAD, float scale, float slope, int reverse) {
    // scale: 1 / channels-per-group (GwcNet mean) or 1 (plain correlation sum); slope: negative slope of a leaky ReLU
    // applied to the result (1 = none); reverse: disparity d0 + k is written to output plane D - 1 - k
    // (correlation1d_cost's channel order, correlation1d_cost.py:12-22)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int cpg = CPG ? CPG : C / G;
    const int WR = W + 2 * PAD;
    float* sl = reinterpret_cast<float*>(smem_raw);   // [cpg][W]
    float* sr = sl + cpg * W;                         // [cpg][WR], zero padded
    __shared__ uint64_t bar;
    const int y = blockIdx.x, g = blockIdx.y, b = blockIdx.z;
    const size_t plane = (size_t)H * W;
    const float* l0 = left + ((size_t)(b * C + g * cpg) * H + y) * W;
    const float* r0 = right + ((size_t)(b * C + g * cpg) * H + y) * W;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < cpg * 2 * PAD; i += blockDim.x) {
        const int c = i / (2 * PAD), j = i - c * 2 * PAD;
        sr[c * WR + (j < PAD ? j : W + j)] = 0.f;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, 2u * cpg * W * 4u);
        for (int c = 0; c < cpg; ++c) {
            bulk_g2s(sl + c * W, l0 + c * plane, W * 4, &bar);
            bulk_g2s(sr + c * WR + PAD, r0 + c * plane, W * 4, &bar);
        }
    }
    mbar_wait(&bar, 0);

    float* out_row0 = out + (((size_t)(b * G + g) * D) * H + y) * W;
    const float inv = scale;                      // mean = sum * (1/n): exact for power-of-two group sizes
    const int W4 = W >> 2, KG = (D + 7) >> 3;
    for (int u = threadIdx.x; u < W4 * KG; u += blockDim.x) {
        const int kg = u / W4;
        const int x = (u - kg * W4) << 2;
        const int dk = d0 + kg * 8;                 // disparity of the unit's first k
        float acc[8][4];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[kk][j] = 0.f;
        const float* lp0 = sl + x;
        const float* rp0 = sr + PAD + x - dk - 8;                 // window start: index x - (dk+8) (16-byte aligned if ALIGNED)
#pragma unroll
        for (int c = 0; c < cpg; ++c) {
            const float4 lv = *reinterpret_cast<const float4*>(lp0 + c * W);
            const float l[4] = {lv.x, lv.y, lv.z, lv.w};
            const float* rp = rp0 + c * WR;
            float rw[12];
            if (ALIGNED) {
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float4 t = *reinterpret_cast<const float4*>(rp + 4 * q);
                    rw[4 * q] = t.x; rw[4 * q + 1] = t.y; rw[4 * q + 2] = t.z; rw[4 * q + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int j = 1; j < 12; ++j) rw[j] = rp[j];
            }
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[kk][j] = fmaf(l[j], rw[j - kk + 8], acc[kk][j]);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const int k = kg * 8 + kk;
            if (k < D) {
                float4 o = make_float4(acc[kk][0] * inv, acc[kk][1] * inv, acc[kk][2] * inv, acc[kk][3] * inv);
                if (CORR) {                      // (compile-time: the GwcNet volume pays nothing for it)
                    o.x = o.x > 0.f ? o.x : o.x * slope; o.y = o.y > 0.f ? o.y : o.y * slope;
                    o.z = o.z > 0.f ? o.z : o.z * slope; o.w = o.w > 0.f ? o.w : o.w * slope;
                }
                st_cs_f4(out_row0 + (size_t)((CORR && reverse) ? D - 1 - k : k) * plane + x, o);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// warped ("fast_mode") volumes.  One thread per (b, k, y, x); loops over channels.
// grid_sample arithmetic restated in oracle/dmb_oracle.py:warp_volume.
// ---------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) warp_volume_kernel(const float* __restrict__ left, const float* __restrict__ right,
                                                          const float* __restrict__ disp, float* __restrict__ out,
                                                          int B, int C, int H, int W, int D, float p) {
    const size_t total = (size_t)B * D * H * W;
    const size_t plane = (size_t)H * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = i % W;
        const int y = (i / W) % H;
        const int k = (i / plane) % D;
        const int b = i / (plane * D);
        // same operation order as the reference (inverse_warp_3d.py:36-42, then grid_sample's
        // un-normalisation ((g+1)*size-1)/2); __f*_rn keeps the compiler from contracting to FMA
        const float gx = __fsub_rn((float)x, disp[i]);
        const float nx = __fsub_rn(__fmul_rn(__fdiv_rn(gx, (float)(W - 1)), 2.f), 1.f);
        const float ny = __fsub_rn(__fmul_rn(__fdiv_rn((float)y, (float)(H - 1)), 2.f), 1.f);
        const float nz = __fsub_rn(__fmul_rn(__fdiv_rn((float)k, (float)(D - 1)), 2.f), 1.f);
        const float ix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(nx, 1.f), (float)W), 1.f), 2.f);
        const float iy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(ny, 1.f), (float)H), 1.f), 2.f);
        const float iz = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(nz, 1.f), (float)D), 1.f), 2.f);
        const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
        const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
        const float tx = ix - fx, ty = iy - fy, tz = iz - fz;
        float wz = 0.f;
        if (z0 >= 0 && z0 <= D - 1) wz += 1.f - tz;
        if (z0 + 1 >= 0 && z0 + 1 <= D - 1) wz += tz;
        const bool vx0 = x0 >= 0 && x0 <= W - 1, vx1 = x0 + 1 >= 0 && x0 + 1 <= W - 1;
        const bool vy0 = y0 >= 0 && y0 <= H - 1, vy1 = y0 + 1 >= 0 && y0 + 1 <= H - 1;
        const float w00 = (vx0 && vy0) ? (1.f - tx) * (1.f - ty) * wz : 0.f;
        const float w01 = (vx1 && vy0) ? tx * (1.f - ty) * wz : 0.f;
        const float w10 = (vx0 && vy1) ? (1.f - tx) * ty * wz : 0.f;
        const float w11 = (vx1 && vy1) ? tx * ty * wz : 0.f;
        const int xa = min(max(x0, 0), W - 1), xb = min(max(x0 + 1, 0), W - 1);
        const int ya = min(max(y0, 0), H - 1), yb = min(max(y0 + 1, 0), H - 1);
        float norm_acc = 0.f;
        for (int c = 0; c < C; ++c) {
            const float* rp = right + (size_t)(b * C + c) * plane;
            const float t = w00 * __ldg(rp + ya * W + xa) + w01 * __ldg(rp + ya * W + xb) +
                            w10 * __ldg(rp + yb * W + xa) + w11 * __ldg(rp + yb * W + xb);
            const float l = (t > 0.f) ? __ldg(left + (size_t)(b * C + c) * plane + y * W + x) : 0.f;
            if (MODE == 0) {
                out[(((size_t)(b * 2 * C + c) * D + k) * H + y) * W + x] = l;
                out[(((size_t)(b * 2 * C + C + c) * D + k) * H + y) * W + x] = t;
            } else if (MODE == 1) {
                out[(((size_t)(b * C + c) * D + k) * H + y) * W + x] = l - t;
            } else {
                const float a = fabsf(l - t);
                norm_acc += (p == 1.f) ? a : ((p == 2.f) ? a * a : powf(a, p));
            }
        }
        if (MODE == 2) out[i] = (p == 1.f) ? norm_acc : ((p == 2.f) ? sqrtf(norm_acc) : powf(norm_acc, 1.f / p));
    }
}

// ---------------------------------------------------------------------------------------
// blocked channels-last bf16 (hi[,lo]) layout of the tensor-core trunk: [B][C/8][S][8], S = D*H*W
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack2(float a, float b, int fp16) {
    if (fp16) {
        __half2 v = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&v);
    }
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack2(uint32_t u, int fp16, float& a, float& b) {
    if (fp16) {
        const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&u));
        a = t.x;
        b = t.y;
    } else {
        a = __uint_as_float(u << 16);
        b = __uint_as_float(u & 0xFFFF0000u);
    }
}
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo, int fp16) {
    float h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        h[e] = fp16 ? __half2float(__float2half_rn(v[e])) : __bfloat162float(__float2bfloat16_rn(v[e]));
        l[e] = v[e] - h[e];
    }
    hi = make_uint4(pack2(h[0], h[1], fp16), pack2(h[2], h[3], fp16), pack2(h[4], h[5], fp16), pack2(h[6], h[7], fp16));
    lo = make_uint4(pack2(l[0], l[1], fp16), pack2(l[2], l[3], fp16), pack2(l[4], l[5], fp16), pack2(l[6], l[7], fp16));
}

// x: [B,C,S] fp32 -> hi/lo: [B][C/8][S][8] bf16.  One thread per (b, cb, s): 8 strided-by-S reads
// (coalesced across s) and one 16-byte store per plane.
// wsplit_w > 0: the output rows are W-PARITY-SPLIT -- [..][row][even | odd][W/2][8] instead of [..][row][W][8] -- the
// operand layout of the stride-2 tcgen05 weight gradient (csrc/wgrad_tc.cu); wsplit_w = W (even).
__global__ void __launch_bounds__(256) ncs_to_blocked_kernel(const float* __restrict__ x, uint4* __restrict__ hi,
                                                             uint4* __restrict__ lo, int C, size_t S, size_t total, int fp16,
                                                             int wsplit_w) {
    const int CBS = C / 8;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t s = i % S;
        const size_t r = i / S;
        const int cb = r % CBS;
        const size_t b = r / CBS;
        if (wsplit_w > 0) {                       // output position (row, parity, w2) <- input voxel (row, 2 * w2 + parity)
            const size_t row = s / wsplit_w;
            const int rem = (int)(s - row * wsplit_w), half = wsplit_w / 2;
            const int par = rem / half, w2 = rem - par * half;
            s = row * wsplit_w + 2 * w2 + par;
        }
        const float* src = x + (b * C + (size_t)cb * 8) * S + s;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __ldg(src + e * S);
        uint4 h, l;
        split8(v, h, l, fp16);
        hi[i] = h;
        if (lo) lo[i] = l;
    }
}

__global__ void __launch_bounds__(256) blocked_to_ncs_kernel(const uint4* __restrict__ hi, const uint4* __restrict__ lo,
                                                             float* __restrict__ y, int C, size_t S, size_t total,
                                                             int fp16) {
    const int CBS = C / 8;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t s = i % S;
        const size_t r = i / S;
        const int cb = r % CBS;
        const size_t b = r / CBS;
        const uint4 h = __ldg(hi + i);
        const uint32_t hu[4] = {h.x, h.y, h.z, h.w};
        float v[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) unpack2(hu[k], fp16, v[2 * k], v[2 * k + 1]);
        if (lo) {
            const uint4 l = __ldg(lo + i);
            const uint32_t lu[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float a, c2;
                unpack2(lu[k], fp16, a, c2);
                v[2 * k] += a;
                v[2 * k + 1] += c2;
            }
        }
        float* dst = y + (b * C + (size_t)cb * 8) * S + s;
#pragma unroll
        for (int e = 0; e < 8; ++e) dst[e * S] = v[e];
    }
}

// y = a + b on blocked (hi[, lo]) activations of identical geometry: the skip additions that follow a ReLU and
// therefore cannot ride in the producing convolution's epilogue (GCAggregator, aggregators/GCNet.py:108-116).
__global__ void __launch_bounds__(256) blocked_add_kernel(const uint4* __restrict__ a_hi, const uint4* __restrict__ a_lo,
                                                          const uint4* __restrict__ b_hi, const uint4* __restrict__ b_lo,
                                                          uint4* __restrict__ y_hi, uint4* __restrict__ y_lo, size_t n, int fp16) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
        const uint4* src[4] = {a_hi, a_lo, b_hi, b_lo};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (!src[t]) continue;
            const uint4 q = __ldg(src[t] + i);
            const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float f0, f1;
                unpack2(u[k], fp16, f0, f1);
                v[2 * k] += f0;
                v[2 * k + 1] += f1;
            }
        }
        uint4 h, l;
        split8(v, h, l, fp16);
        y_hi[i] = h;
        if (y_lo) y_lo[i] = l;
    }
}

// y[b][s] = sum_c w[c] * x[b][c][s] for a blocked (hi[, lo]) activation: the 1x1 Conv2d(C, 1, bias=False) that ends a
// confidence head (dmb/modeling/stereo/cmn/cmn.py:31).  One thread per position, C / 8 16-byte loads per plane.
__global__ void __launch_bounds__(256) blocked_dot_kernel(const uint4* __restrict__ hi, const uint4* __restrict__ lo,
                                                          const float* __restrict__ w, float* __restrict__ y, int C, size_t S,
                                                          size_t total, int fp16) {
    extern __shared__ float sw[];
    for (int i = threadIdx.x; i < C; i += blockDim.x) sw[i] = __ldg(w + i);
    __syncthreads();
    const int CBS = C / 8;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / S, s_ = i - b * S;
        float acc = 0.f;
        for (int cb = 0; cb < CBS; ++cb) {
            const size_t o = (b * CBS + cb) * S + s_;
            const uint4 h = __ldg(hi + o);
            const uint32_t hu[4] = {h.x, h.y, h.z, h.w};
            float v[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) unpack2(hu[k], fp16, v[2 * k], v[2 * k + 1]);
            if (lo) {
                const uint4 l = __ldg(lo + o);
                const uint32_t lu[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float a, c2;
                    unpack2(lu[k], fp16, a, c2);
                    v[2 * k] += a;
                    v[2 * k + 1] += c2;
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) acc = fmaf(sw[cb * 8 + e], v[e], acc);
        }
        y[i] = acc;
    }
}

// cat volume straight into the blocked layout: out [B][2C/8][D][H][W][8] from fp32 NCHW features.
// One CTA per (b, 8-channel block, y): the source row's 8 channels are split into 16-bit (hi, lo) ONCE -- the left
// half keeps its own column in registers, the right half stages the converted row in shared memory -- and every
// thread then walks the D disparities of its column with nothing but a validity test, (two shared loads) and two
// 16-byte streaming stores per disparity.  (The first version decoded a flat index with four 64-bit divisions and
// redid the 8 loads + split for every output voxel: ~300 instructions per 32 bytes written, instruction bound at
// 3.4 TB/s where a plain fill reaches 7.2 TB/s on this part -- profiles/README.md.)
__global__ void __launch_bounds__(256) cat_volume_blocked_kernel(const float* __restrict__ left,
                                                                 const float* __restrict__ right,
                                                                 uint4* __restrict__ o_hi, uint4* __restrict__ o_lo,
                                                                 int C, int H, int W, int D, DispList dl, int fp16) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* s_hi = reinterpret_cast<uint4*>(smem_raw);              // [W] converted right row (right half only)
    uint4* s_lo = s_hi + W;
    const int y = blockIdx.x, cb = blockIdx.y, b = blockIdx.z;
    const int CBS = 2 * C / 8, CB_L = C / 8;
    const bool is_r = cb >= CB_L;
    const size_t plane = (size_t)H * W;
    const float* src = (is_r ? right : left) + ((size_t)b * C + (size_t)(is_r ? cb - CB_L : cb) * 8) * plane + (size_t)y * W;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    if (is_r) {                                                    // block-uniform
        for (int x = threadIdx.x; x < W; x += blockDim.x) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __ldg(src + e * plane + x);
            uint4 h, l;
            split8(v, h, l, fp16);
            s_hi[x] = h;
            s_lo[x] = l;
        }
        __syncthreads();
    }
    for (int x0 = 0; x0 < W; x0 += blockDim.x) {
        const int x = x0 + threadIdx.x;
        if (x >= W) break;
        uint4 h = zero, l = zero;
        if (!is_r) {                                               // own column: converted once, kept in registers
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __ldg(src + e * plane + x);
            split8(v, h, l, fp16);
        }
        size_t o = (((size_t)b * CBS + cb) * D) * plane + (size_t)y * W + x;
#pragma unroll 4
        for (int k = 0; k < D; ++k, o += plane) {
            const int d = dl.d[k];
            const bool ok = col_valid(x, d, W);
            uint4 vh = zero, vl = zero;
            if (ok) {
                vh = is_r ? s_hi[x - d] : h;
                vl = is_r ? s_lo[x - d] : l;
            }
            __stcs(o_hi + o, vh);
            if (o_lo) __stcs(o_lo + o, vl);
        }
    }
}

}  // namespace dmb

using namespace dmb;

extern "C" int dmb_b200_cat_volume(const float* left, const float* right, float* out, int B, int C, int H, int W,
                                   const int* disp_idx_host, int D, void* stream) {
    return launch_volume_rows<0>(left, right, out, B, C, H, W, disp_idx_host, D, stream);
}

extern "C" int dmb_b200_dif_volume(const float* left, const float* right, float* out, int B, int C, int H, int W,
                                   const int* disp_idx_host, int D, void* stream) {
    return launch_volume_rows<1>(left, right, out, B, C, H, W, disp_idx_host, D, stream);
}

extern "C" int dmb_b200_gwc_volume(const float* left, const float* right, float* out, int B, int C, int H, int W, int G,
                                   const int* disp_idx_host, int D, void* stream) {
    DMB_REQUIRE(left && right && out && disp_idx_host, "gwc_volume: null pointer");
    DMB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0 && G > 0, "gwc_volume: non-positive dimension");
    DMB_REQUIRE(C % G == 0, "gwc_volume: C=%d not divisible by groups=%d", C, G);
    DMB_REQUIRE(H <= 65535 && G <= 65535 && B <= 65535, "gwc_volume: grid dimension too large");
    const int cpg = C / G;
    const int Wp = (W + 3) & ~3;
    const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(left) | reinterpret_cast<uintptr_t>(right) |
                                       reinterpret_cast<uintptr_t>(out)) % 16 == 0);
    const size_t smem = (size_t)2 * cpg * Wp * 4;
    DMB_REQUIRE(smem <= 200 * 1024, "gwc_volume: group rows do not fit shared memory (cpg=%d, W=%d)", cpg, W);
    if (smem > 48 * 1024) {
        DMB_CUDA(cudaFuncSetAttribute(gwc_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DMB_CUDA(cudaFuncSetAttribute(gwc_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 grid(H, G, B);
    // unit-step disparity list (the shipped configs): register-blocked kernel
    bool unit = vec;
    for (int i = 1; i < D && unit; ++i) unit = disp_idx_host[i] == disp_idx_host[0] + i;
    if (unit) {
        const int d0 = disp_idx_host[0], dlast = d0 + D - 1;
        int maxabs = (d0 < 0 ? -d0 : d0) > (dlast < 0 ? -dlast : dlast) ? (d0 < 0 ? -d0 : d0) : (dlast < 0 ? -dlast : dlast);
        const int PAD = ((maxabs + 8 + 8) + 3) & ~3;          // window reaches 7 before / 3 after the shifted column
        const size_t smem_u = (size_t)cpg * (W + W + 2 * PAD) * 4;
        if (smem_u <= 200 * 1024) {
            // threads: the W/4 x ceil(D/8) register-block units spread evenly over as few sweeps as possible
            const int units = (W / 4) * ((D + 7) / 8), sweeps = (units + 255) / 256;
            int uthreads = (((units + sweeps - 1) / sweeps + 31) / 32) * 32;
            if (uthreads > 256) uthreads = 256;
            const bool aligned = (d0 % 4 == 0);
#define DMB_GWC_UNIT(CPG)                                                                                              \
    do {                                                                                                               \
        if (smem_u > 48 * 1024) {                                                                                      \
            DMB_CUDA(cudaFuncSetAttribute(gwc_rows_unit_kernel<CPG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                          (int)smem_u));                                                               \
            DMB_CUDA(cudaFuncSetAttribute(gwc_rows_unit_kernel<CPG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)smem_u));                                                               \
        }                                                                                                              \
        if (aligned)                                                                                                   \
            gwc_rows_unit_kernel<CPG, true><<<grid, uthreads, smem_u, as_stream(stream)>>>(left, right, out, C, G, H, W, D, d0, PAD, \
                                                                                           1.0f / (float)cpg, 1.0f, 0);          \
        else                                                                                                           \
            gwc_rows_unit_kernel<CPG, false><<<grid, uthreads, smem_u, as_stream(stream)>>>(left, right, out, C, G, H, W, D, d0, PAD, \
                                                                                            1.0f / (float)cpg, 1.0f, 0);         \
    } while (0)
            if (cpg == 8) DMB_GWC_UNIT(8);
            else if (cpg == 4) DMB_GWC_UNIT(4);
            else if (cpg == 16) DMB_GWC_UNIT(16);
            else DMB_GWC_UNIT(0);
#undef DMB_GWC_UNIT
            return check_launch("gwc_rows_unit_kernel");
        }
    }
    for (int k0 = 0; k0 < D; k0 += kMaxDisp) {
        DispList dl;
        const int n = (D - k0 < kMaxDisp) ? D - k0 : kMaxDisp;
        for (int i = 0; i < n; ++i) {
            int d = disp_idx_host[k0 + i];
            dl.d[i] = d >= W ? W : (d <= -W ? -W : d);
        }
        if (vec)
            gwc_rows_kernel<true><<<grid, 256, smem, as_stream(stream)>>>(left, right, out, C, G, H, W, n, k0, D, dl);
        else
            gwc_rows_kernel<false><<<grid, 256, smem, as_stream(stream)>>>(left, right, out, C, G, H, W, n, k0, D, dl);
        int rc = check_launch("gwc_rows_kernel");
        if (rc) return rc;
    }
    return DMB_OK;
}

// correlation1d_cost (dmb/modeling/stereo/cost_processors/utils/correlation1d_cost.py:7-27): the reference calls the
// un-vendored SpatialCorrelationSampler(patch_size = (1, 2 max_disp - 1), kernel 1, stride 1, padding 0, dilation_patch 1)
// -- out[j] = sum_c L[c, y, x] * R[c, y, x + j - (max_disp - 1)], zero outside the image -- keeps the first max_disp
// channels (shifts -(max_disp - 1) .. 0) and applies leaky ReLU(0.1).  I.e. a ONE-group correlation with a plain sum,
// channel j holding disparity max_disp - 1 - j: the GWC register-blocked kernel with scale 1, reversed output planes
// and the leaky ReLU in its store.  out: [B, max_disp, H, W].
extern "C" int dmb_b200_corr1d_volume(const float* left, const float* right, float* out, int B, int C, int H, int W,
                                      int max_disp, float negative_slope, void* stream) {
    DMB_REQUIRE(left && right && out, "corr1d_volume: null pointer");
    DMB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && max_disp > 0, "corr1d_volume: non-positive dimension");
    DMB_REQUIRE(H <= 65535 && B <= 65535, "corr1d_volume: grid dimension too large");
    DMB_REQUIRE(W % 4 == 0 && ((reinterpret_cast<uintptr_t>(left) | reinterpret_cast<uintptr_t>(right) |
                                reinterpret_cast<uintptr_t>(out)) % 16 == 0),
                "corr1d_volume: W must be a multiple of 4 and the tensors 16-byte aligned");
    const int D = max_disp;
    const int PAD = ((D - 1 + 8 + 8) + 3) & ~3;
    const size_t smem_u = (size_t)C * (W + W + 2 * PAD) * 4;
    DMB_REQUIRE(smem_u <= 200 * 1024, "corr1d_volume: %d feature rows of width %d do not fit shared memory", C, W);
    DMB_CUDA(cudaFuncSetAttribute(gwc_rows_unit_kernel<0, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_u));
    const int units = (W / 4) * ((D + 7) / 8), sweeps = (units + 255) / 256;
    int uthreads = (((units + sweeps - 1) / sweeps + 31) / 32) * 32;
    if (uthreads > 256) uthreads = 256;
    dim3 grid(H, 1, B);
    gwc_rows_unit_kernel<0, true, true><<<grid, uthreads, smem_u, as_stream(stream)>>>(left, right, out, C, 1, H, W, D, 0, PAD,
                                                                                        1.0f, negative_slope, 1);
    return check_launch("gwc_rows_unit_kernel<corr1d>");
}

extern "C" int dmb_b200_warp_volume(const float* left, const float* right, const float* disp_sample, float* out, int B,
                                    int C, int H, int W, int D, int mode, float p, void* stream) {
    DMB_REQUIRE(left && right && disp_sample && out, "warp_volume: null pointer");
    DMB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0, "warp_volume: non-positive dimension");
    DMB_REQUIRE(mode >= 0 && mode <= 2, "warp_volume: mode must be 0, 1 or 2");
    DMB_REQUIRE(p > 0.f, "warp_volume: p must be positive");
    const size_t total = (size_t)B * D * H * W;
    const int blocks = (int)((total + 255) / 256 < (size_t)sm_count() * 32 ? (total + 255) / 256 : (size_t)sm_count() * 32);
    if (mode == 0)
        warp_volume_kernel<0><<<blocks, 256, 0, as_stream(stream)>>>(left, right, disp_sample, out, B, C, H, W, D, p);
    else if (mode == 1)
        warp_volume_kernel<1><<<blocks, 256, 0, as_stream(stream)>>>(left, right, disp_sample, out, B, C, H, W, D, p);
    else
        warp_volume_kernel<2><<<blocks, 256, 0, as_stream(stream)>>>(left, right, disp_sample, out, B, C, H, W, D, p);
    return check_launch("warp_volume_kernel");
}

static unsigned grid_for(size_t total) {
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)sm_count() * 16;
    return (unsigned)(blocks > cap ? cap : blocks);
}

extern "C" int dmb_b200_ncdhw_to_blocked(const float* x, void* y_hi, void* y_lo, int B, int C, int D, int H, int W,
                                         int fp16, void* stream) {
    DMB_REQUIRE(x && y_hi, "ncdhw_to_blocked: null pointer");
    DMB_REQUIRE(B > 0 && C > 0 && D > 0 && H > 0 && W > 0, "ncdhw_to_blocked: non-positive dimension");
    DMB_REQUIRE(C % 8 == 0, "ncdhw_to_blocked: C=%d must be a multiple of 8", C);
    const size_t S = (size_t)D * H * W, total = (size_t)B * (C / 8) * S;
    ncs_to_blocked_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(x, (uint4*)y_hi, (uint4*)y_lo, C, S, total, fp16 ? 1 : 0, 0);
    return check_launch("ncs_to_blocked_kernel");
}

extern "C" int dmb_b200_ncdhw_to_blocked_wsplit(const float* x, void* y_hi, void* y_lo, int B, int C, int D, int H, int W,
                                                int fp16, void* stream) {
    DMB_REQUIRE(x && y_hi, "ncdhw_to_blocked_wsplit: null pointer");
    DMB_REQUIRE(B > 0 && C > 0 && D > 0 && H > 0 && W > 0, "ncdhw_to_blocked_wsplit: non-positive dimension");
    DMB_REQUIRE(C % 8 == 0 && W % 2 == 0, "ncdhw_to_blocked_wsplit: C=%d must be a multiple of 8 and W=%d even", C, W);
    const size_t S = (size_t)D * H * W, total = (size_t)B * (C / 8) * S;
    ncs_to_blocked_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(x, (uint4*)y_hi, (uint4*)y_lo, C, S, total, fp16 ? 1 : 0, W);
    return check_launch("ncs_to_blocked_kernel");
}

extern "C" int dmb_b200_blocked_to_ncdhw(const void* x_hi, const void* x_lo, float* y, int B, int C, int D, int H, int W,
                                         int fp16, void* stream) {
    DMB_REQUIRE(x_hi && y, "blocked_to_ncdhw: null pointer");
    DMB_REQUIRE(B > 0 && C > 0 && D > 0 && H > 0 && W > 0, "blocked_to_ncdhw: non-positive dimension");
    DMB_REQUIRE(C % 8 == 0, "blocked_to_ncdhw: C=%d must be a multiple of 8", C);
    const size_t S = (size_t)D * H * W, total = (size_t)B * (C / 8) * S;
    blocked_to_ncs_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>((const uint4*)x_hi, (const uint4*)x_lo, y, C, S,
                                                                         total, fp16 ? 1 : 0);
    return check_launch("blocked_to_ncs_kernel");
}

extern "C" int dmb_b200_blocked_add(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, void* y_hi,
                                    void* y_lo, int64_t n_blocks, int fp16, void* stream) {
    DMB_REQUIRE(a_hi && b_hi && y_hi && n_blocks > 0, "blocked_add: null pointer / empty tensor");
    DMB_REQUIRE((a_lo != nullptr) == (b_lo != nullptr) && (a_lo != nullptr) == (y_lo != nullptr),
                "blocked_add: the three tensors must all be split pairs or all single planes");
    blocked_add_kernel<<<grid_for((size_t)n_blocks), 256, 0, as_stream(stream)>>>((const uint4*)a_hi, (const uint4*)a_lo,
                                                                                 (const uint4*)b_hi, (const uint4*)b_lo, (uint4*)y_hi,
                                                                                 (uint4*)y_lo, (size_t)n_blocks, fp16 ? 1 : 0);
    return check_launch("blocked_add_kernel");
}

extern "C" int dmb_b200_blocked_dot(const void* x_hi, const void* x_lo, const float* w, float* y, int B, int C, int64_t S,
                                    int fp16, void* stream) {
    DMB_REQUIRE(x_hi && w && y, "blocked_dot: null pointer");
    DMB_REQUIRE(B > 0 && C > 0 && S > 0 && C % 8 == 0 && C <= 4096, "blocked_dot: bad dimensions (C %% 8 == 0, C <= 4096)");
    const size_t total = (size_t)B * (size_t)S;
    blocked_dot_kernel<<<grid_for(total), 256, (size_t)C * 4, as_stream(stream)>>>((const uint4*)x_hi, (const uint4*)x_lo, w, y, C,
                                                                                  (size_t)S, total, fp16 ? 1 : 0);
    return check_launch("blocked_dot_kernel");
}

extern "C" int dmb_b200_cat_volume_blocked(const float* left, const float* right, void* out_hi, void* out_lo, int B,
                                           int C, int H, int W, const int* disp_idx_host, int D, int fp16,
                                           void* stream) {
    DMB_REQUIRE(left && right && out_hi && disp_idx_host, "cat_volume_blocked: null pointer");
    DMB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0, "cat_volume_blocked: non-positive dimension");
    DMB_REQUIRE(C % 8 == 0, "cat_volume_blocked: C=%d must be a multiple of 8", C);
    DMB_REQUIRE(D <= kMaxDisp, "cat_volume_blocked: D=%d exceeds %d", D, kMaxDisp);
    DispList dl;
    for (int i = 0; i < D; ++i) {
        int d = disp_idx_host[i];
        dl.d[i] = d >= W ? W : (d <= -W ? -W : d);
    }
    DMB_REQUIRE(H <= 65535 && 2 * C / 8 <= 65535 && B <= 65535, "cat_volume_blocked: grid dimension too large");
    const size_t smem = (size_t)W * 32;
    DMB_REQUIRE(smem <= 200 * 1024, "cat_volume_blocked: feature row too wide (W=%d)", W);
    if (smem > 48 * 1024)
        DMB_CUDA(cudaFuncSetAttribute(cat_volume_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 grid(H, 2 * C / 8, B);
    cat_volume_blocked_kernel<<<grid, 256, smem, as_stream(stream)>>>(left, right, (uint4*)out_hi, (uint4*)out_lo, C, H, W, D,
                                                                     dl, fp16 ? 1 : 0);
    return check_launch("cat_volume_blocked_kernel");
}
