// Cost upsampling (trilinear align_corners=True / AcfNet's learned 8x8x8 stride-4 transposed
// conv) fused with soft-argmin disparity regression, plus the stand-alone predictors.
//
// Reference: F.interpolate in dmb/modeling/stereo/cost_processors/aggregators/PSMNet.py:75-88,
// ConvTranspose3d(1,1,8,4,2) in aggregators/AcfNet.py:55-57,81-83, soft-argmin in
// dmb/modeling/stereo/disp_predictors/{soft_argmin.py:44-75, faster_soft_argmin.py:51-75,
// local_soft_argmin.py:47-105}.  Oracle: oracle/dmb_oracle.py (trilinear_up, acf_aggregator,
// soft_argmin, local_soft_argmin).
//
// The fused kernel reads the LOW-resolution cost (Dl*Hl*Wl floats, L2 resident) and produces
// the disparity map without materialising the D*H*W volume: 8.4 MB of traffic per cost at the
// PSMNet 544x960 config instead of ~1.2 GB (write 401 MB, read 401 MB, softmax write/read).
#include "common.cuh"

namespace dmb {

// running softmax-expectation state: m = running max, s = sum e^(c-m), t = sum e^(c-m) * disp
// e^x for x <= 0 as exp2f(x * log2(e)): the subtraction (c - m) is done first, so the product's rounding
// error is relative to the SHIFTED argument (|x| < ~20 for every term that matters) -- 2-ulp exp2f keeps
// the softmax weights within ~1e-6 relative of expf at a third of its instruction count.
// ex2.approx.ftz is the instruction exp2f() itself is built on (max rel. error 2^-22); the wrapper code that
// exp2f adds for denormal results is not needed: arguments are <= 0 and a flushed e^-88 contributes nothing
__device__ __forceinline__ float exp_neg(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
}

struct SoftState {
    float m, s, t;
    __device__ __forceinline__ void init() { m = -INFINITY; s = 0.f; t = 0.f; }
    __device__ __forceinline__ void push4(const float c[4], const float dv[4]) {
        const float cm = fmaxf(fmaxf(c[0], c[1]), fmaxf(c[2], c[3]));
        if (cm > m) {
            const float f = exp_neg(m - cm);   // m = -inf on the first chunk: f = 0, s = t = 0
            s *= f;
            t *= f;
            m = cm;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float e = exp_neg(c[j] - m);
            s += e;
            t = fmaf(e, dv[j], t);
        }
    }
    __device__ __forceinline__ void push1(float c, float dv) {
        if (c > m) {
            const float f = exp_neg(m - c);
            s *= f;
            t *= f;
            m = c;
        }
        const float e = exp_neg(c - m);
        s += e;
        t = fmaf(e, dv, t);
    }
    __device__ __forceinline__ float result() const { return t / s; }
};

struct RegressParams {
    int B, Dl, Hl, Wl, D, H, W;
    float alpha, start_disp, disp_step;
    int normalize;
};

__device__ __forceinline__ float disp_value(const float* __restrict__ dv, const RegressParams& p, int d) {
    return dv ? __ldg(dv + d) : fmaf((float)d, p.disp_step, p.start_disp);
}

// One thread per output pixel (b, y, x); marches over the output disparities.
// MODE 0: trilinear (align_corners=True).  MODE 1: learned 8^3 stride-4 pad-2 transposed conv.
template <int MODE, bool WRITE_COST, bool REGRESS>
__global__ void __launch_bounds__(256) upsample_regress_kernel(const float* __restrict__ low,
                                                               const float* __restrict__ upw,
                                                               float* __restrict__ cost_out,
                                                               float* __restrict__ disp_out,
                                                               const float* __restrict__ dvals, RegressParams p) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    if (x >= p.W || y >= p.H) return;
    const size_t lplane = (size_t)p.Hl * p.Wl;
    const float* lb = low + (size_t)b * p.Dl * lplane;
    const size_t oplane = (size_t)p.H * p.W;
    float* co = WRITE_COST ? cost_out + (size_t)b * p.D * oplane + (size_t)y * p.W + x : nullptr;

    SoftState st;
    st.init();
    float lin = 0.f;   // normalize == 0: plain weighted sum

    if (MODE == 0) {
        // source coordinates exactly as ATen's area_pixel_compute_source_index (align_corners)
        const float sy = p.H > 1 ? (float)(p.Hl - 1) / (float)(p.H - 1) : 0.f;
        const float sx = p.W > 1 ? (float)(p.Wl - 1) / (float)(p.W - 1) : 0.f;
        const float sd = p.D > 1 ? (float)(p.Dl - 1) / (float)(p.D - 1) : 0.f;
        const float fy = sy * (float)y, fx = sx * (float)x;
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = y0 + (y0 < p.Hl - 1 ? 1 : 0), x1 = x0 + (x0 < p.Wl - 1 ? 1 : 0);
        const float ly1 = fy - (float)y0, ly0 = 1.f - ly1;
        const float lx1 = fx - (float)x0, lx0 = 1.f - lx1;
        const int o00 = y0 * p.Wl + x0, o01 = y0 * p.Wl + x1, o10 = y1 * p.Wl + x0, o11 = y1 * p.Wl + x1;
        auto plane_val = [&](int dl) -> float {
            const float* q = lb + (size_t)dl * lplane;
            return ly0 * (lx0 * __ldg(q + o00) + lx1 * __ldg(q + o01)) +
                   ly1 * (lx0 * __ldg(q + o10) + lx1 * __ldg(q + o11));
        };
        int cur_i = 0;
        float a0 = plane_val(0);
        float a1 = p.Dl > 1 ? plane_val(1) : a0;
        for (int d0 = 0; d0 < p.D; d0 += 4) {
            float c[4], dv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int d = d0 + j;
                if (d < p.D) {
                    const float fd = sd * (float)d;
                    const int i0 = (int)fd;
                    while (cur_i < i0) {   // advance the rolling pair of source planes
                        ++cur_i;
                        a0 = a1;
                        a1 = (cur_i + 1 < p.Dl) ? plane_val(cur_i + 1) : a1;
                    }
                    const float l1 = fd - (float)i0, l0 = 1.f - l1;
                    // i0 == Dl-1: the second tap index equals i0 (t1p = 0) => a1 must equal a0
                    const float hi = (i0 < p.Dl - 1) ? a1 : a0;
                    const float v = l0 * a0 + l1 * hi;
                    if (WRITE_COST) st_cs_f(co + (size_t)d * oplane, v);
                    c[j] = v * p.alpha;
                    dv[j] = REGRESS ? disp_value(dvals, p, d) : 0.f;
                } else {
                    c[j] = -INFINITY;
                    dv[j] = 0.f;
                }
            }
            if (REGRESS) {
                if (p.normalize) {
                    st.push4(c, dv);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (d0 + j < p.D) lin = fmaf(c[j], dv[j], lin);
                }
            }
        }
    } else {
        // out[o] = sum_k in[(o+2-k)/4] * w[k] over k with (o+2-k) % 4 == 0: per dimension the two
        // taps (i_a = (o+2)/4, k = r) and (i_a - 1, k = r+4), r = (o+2) % 4.
        const int ry = (y + 2) & 3, rx = (x + 2) & 3;
        const int ya = (y + 2) >> 2, xa = (x + 2) >> 2;
        int yi[2] = {ya, ya - 1}, xi[2] = {xa, xa - 1};
        const int ky[2] = {ry, ry + 4}, kx[2] = {rx, rx + 4};
        bool cv[4];
        int coff[4];
        float wr[8][4];   // [kd][column]
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
                const int j = a * 2 + c2;
                cv[j] = yi[a] >= 0 && yi[a] < p.Hl && xi[c2] >= 0 && xi[c2] < p.Wl;
                coff[j] = cv[j] ? yi[a] * p.Wl + xi[c2] : 0;
#pragma unroll
                for (int kd = 0; kd < 8; ++kd) wr[kd][j] = cv[j] ? __ldg(upw + (kd * 8 + ky[a]) * 8 + kx[c2]) : 0.f;
            }
        auto col_vals = [&](int dl, float v[4]) {
            const bool ok = dl >= 0 && dl < p.Dl;
            const float* q = lb + (size_t)(ok ? dl : 0) * lplane;
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (ok && cv[j]) ? __ldg(q + coff[j]) : 0.f;
        };
        // D == 4*Dl.  For d = 4q + r': i_a = q + (r'+2)/4, rd = (r'+2) % 4.
        float vm[4], v0[4], vp[4];   // source planes q-1, q, q+1
        col_vals(-1, vm);
        col_vals(0, v0);
        for (int q = 0; q < p.Dl; ++q) {
            col_vals(q + 1, vp);
            float c[4], dv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int rd = (r + 2) & 3;
                const float* va = (r < 2) ? v0 : vp;   // plane i_a
                const float* vb = (r < 2) ? vm : v0;   // plane i_a - 1
                float v = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) v = fmaf(va[j], wr[rd][j], v);
#pragma unroll
                for (int j = 0; j < 4; ++j) v = fmaf(vb[j], wr[rd + 4][j], v);
                const int d = 4 * q + r;
                if (WRITE_COST) st_cs_f(co + (size_t)d * oplane, v);
                c[r] = v * p.alpha;
                dv[r] = REGRESS ? disp_value(dvals, p, d) : 0.f;
            }
            if (REGRESS) {
                if (p.normalize) {
                    st.push4(c, dv);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) lin = fmaf(c[j], dv[j], lin);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                vm[j] = v0[j];
                v0[j] = vp[j];
            }
        }
    }
    if (REGRESS) disp_out[(size_t)b * oplane + (size_t)y * p.W + x] = p.normalize ? st.result() : lin;
}

// Trilinear variant with the thread's 2-D-interpolated source column a[0..Dl) staged in shared memory
// ([Dl][256], each thread touches only its own column: no barrier, no bank conflicts); the depth
// march then costs two shared loads and a blend per output value.
template <bool WRITE_COST, bool REGRESS>
__global__ void __launch_bounds__(256) upsample_trilinear_kernel(const float* __restrict__ low,
                                                                 float* __restrict__ cost_out,
                                                                 float* __restrict__ disp_out,
                                                                 const float* __restrict__ dvals, RegressParams p) {
    extern __shared__ float sa[];
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    if (x >= p.W || y >= p.H) return;
    const size_t lplane = (size_t)p.Hl * p.Wl;
    const float* lb = low + (size_t)b * p.Dl * lplane;
    const size_t oplane = (size_t)p.H * p.W;
    float* co = WRITE_COST ? cost_out + (size_t)b * p.D * oplane + (size_t)y * p.W + x : nullptr;
    float* col = sa + threadIdx.x;

    const float sy = p.H > 1 ? (float)(p.Hl - 1) / (float)(p.H - 1) : 0.f;
    const float sx = p.W > 1 ? (float)(p.Wl - 1) / (float)(p.W - 1) : 0.f;
    const float sd = p.D > 1 ? (float)(p.Dl - 1) / (float)(p.D - 1) : 0.f;
    {
        const float fy = sy * (float)y, fx = sx * (float)x;
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = y0 + (y0 < p.Hl - 1 ? 1 : 0), x1 = x0 + (x0 < p.Wl - 1 ? 1 : 0);
        const float ly1 = fy - (float)y0, ly0 = 1.f - ly1;
        const float lx1 = fx - (float)x0, lx0 = 1.f - lx1;
        const float* q00 = lb + y0 * p.Wl + x0;
        const float* q01 = lb + y0 * p.Wl + x1;
        const float* q10 = lb + y1 * p.Wl + x0;
        const float* q11 = lb + y1 * p.Wl + x1;
#pragma unroll 4
        for (int dl = 0; dl < p.Dl; ++dl) {
            const size_t o = (size_t)dl * lplane;
            col[dl * 256] = ly0 * (lx0 * __ldg(q00 + o) + lx1 * __ldg(q01 + o)) +
                            ly1 * (lx0 * __ldg(q10 + o) + lx1 * __ldg(q11 + o));
        }
    }
    SoftState st;
    st.init();
    float lin = 0.f;
    for (int d0 = 0; d0 < p.D; d0 += 4) {
        float c[4], dv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int d = d0 + j;
            if (d < p.D) {
                const float fd = sd * (float)d;
                const int i0 = (int)fd;
                const int i1 = i0 + (i0 < p.Dl - 1 ? 1 : 0);
                const float l1 = fd - (float)i0, l0 = 1.f - l1;
                const float v = l0 * col[i0 * 256] + l1 * col[i1 * 256];
                if (WRITE_COST) st_cs_f(co + (size_t)d * oplane, v);
                c[j] = v * p.alpha;
                dv[j] = REGRESS ? disp_value(dvals, p, d) : 0.f;
            } else {
                c[j] = -INFINITY;
                dv[j] = 0.f;
            }
        }
        if (REGRESS) {
            if (p.normalize) {
                st.push4(c, dv);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (d0 + j < p.D) lin = fmaf(c[j], dv[j], lin);
            }
        }
    }
    if (REGRESS) disp_out[(size_t)b * oplane + (size_t)y * p.W + x] = p.normalize ? st.result() : lin;
}

// Inference fast path of the above (no cost write, softmax regression): the per-disparity interpolation
// table {byte offset of source plane i0, of i1, l0, l1} and the disparity values are the same for every
// pixel, so each CTA computes them once into shared memory.  The depth march then has no int<->float
// conversions (quarter-rate XU instructions, which -- 3 per value -- bounded the generic kernel) and
// costs two broadcast table loads, two column loads, a blend and the online-softmax update per value.
__global__ void __launch_bounds__(256) upsample_trilinear_regress_kernel(const float* __restrict__ low,
                                                                         float* __restrict__ disp_out,
                                                                         const float* __restrict__ dvals,
                                                                         RegressParams p) {
    extern __shared__ float sa[];
    uint4* tab = reinterpret_cast<uint4*>(sa + (size_t)p.Dl * 256);   // [D]
    float* dvt = reinterpret_cast<float*>(tab + p.D);                 // [D]
    const float sy = p.H > 1 ? (float)(p.Hl - 1) / (float)(p.H - 1) : 0.f;
    const float sx = p.W > 1 ? (float)(p.Wl - 1) / (float)(p.W - 1) : 0.f;
    const float sd = p.D > 1 ? (float)(p.Dl - 1) / (float)(p.D - 1) : 0.f;
    for (int d = threadIdx.x; d < p.D; d += 256) {
        const float fd = sd * (float)d;
        const int i0 = (int)fd;
        const int i1 = i0 + (i0 < p.Dl - 1 ? 1 : 0);
        const float l1 = fd - (float)i0, l0 = 1.f - l1;
        tab[d] = make_uint4((uint32_t)i0 * 1024u, (uint32_t)i1 * 1024u, __float_as_uint(l0), __float_as_uint(l1));
        dvt[d] = disp_value(dvals, p, d);
    }
    const int xr = blockIdx.x * 32 + (threadIdx.x & 31);
    const int yr = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    const bool active = xr < p.W && yr < p.H;
    const int x = min(xr, p.W - 1), y = min(yr, p.H - 1);
    const size_t lplane = (size_t)p.Hl * p.Wl;
    const float* lb = low + (size_t)b * p.Dl * lplane;
    float* col = sa + threadIdx.x;
    {
        const float fy = sy * (float)y, fx = sx * (float)x;
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = y0 + (y0 < p.Hl - 1 ? 1 : 0), x1 = x0 + (x0 < p.Wl - 1 ? 1 : 0);
        const float ly1 = fy - (float)y0, ly0 = 1.f - ly1;
        const float lx1 = fx - (float)x0, lx0 = 1.f - lx1;
        const float* q00 = lb + y0 * p.Wl + x0;
        const float* q01 = lb + y0 * p.Wl + x1;
        const float* q10 = lb + y1 * p.Wl + x0;
        const float* q11 = lb + y1 * p.Wl + x1;
#pragma unroll 4
        for (int dl = 0; dl < p.Dl; ++dl) {
            const size_t o = (size_t)dl * lplane;
            col[dl * 256] = ly0 * (lx0 * __ldg(q00 + o) + lx1 * __ldg(q01 + o)) +
                            ly1 * (lx0 * __ldg(q10 + o) + lx1 * __ldg(q11 + o));
        }
    }
    __syncthreads();
    const unsigned char* colb = reinterpret_cast<const unsigned char*>(col);
    SoftState st;
    st.init();
    int d0 = 0;
    for (; d0 + 4 <= p.D; d0 += 4) {
        float c[4], dv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 e = tab[d0 + j];
            const float a0 = *reinterpret_cast<const float*>(colb + e.x);
            const float a1 = *reinterpret_cast<const float*>(colb + e.y);
            const float v = __uint_as_float(e.z) * a0 + __uint_as_float(e.w) * a1;
            c[j] = v * p.alpha;
            dv[j] = dvt[d0 + j];
        }
        st.push4(c, dv);
    }
    for (; d0 < p.D; ++d0) {
        const uint4 e = tab[d0];
        const float a0 = *reinterpret_cast<const float*>(colb + e.x);
        const float a1 = *reinterpret_cast<const float*>(colb + e.y);
        const float v = __uint_as_float(e.z) * a0 + __uint_as_float(e.w) * a1;
        st.push1(v * p.alpha, dvt[d0]);
    }
    if (active) disp_out[(size_t)b * p.H * p.W + (size_t)y * p.W + x] = st.result();
}

// stand-alone soft-argmin over a materialised cost [B,D,H,W]; one thread per pixel, 8 loads in
// flight per thread, coalesced across x.
template <bool PER_PIXEL>
__global__ void __launch_bounds__(256) soft_argmin_kernel(const float* __restrict__ cost, float* __restrict__ out,
                                                          const float* __restrict__ dvals,
                                                          const float* __restrict__ dsample, RegressParams p) {
    const size_t plane = (size_t)p.H * p.W;
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= plane) return;
    const int b = blockIdx.y;
    const float* cp = cost + (size_t)b * p.D * plane + pix;
    const float* sp = PER_PIXEL ? dsample + (size_t)b * p.D * plane + pix : nullptr;
    SoftState st;
    st.init();
    float lin = 0.f;
    int d = 0;
    for (; d + 8 <= p.D; d += 8) {
        float c[8], dv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = __ldcs(cp + (size_t)(d + j) * plane) * p.alpha;
#pragma unroll
        for (int j = 0; j < 8; ++j) dv[j] = PER_PIXEL ? __ldcs(sp + (size_t)(d + j) * plane) : disp_value(dvals, p, d + j);
        if (p.normalize) {
            st.push4(c, dv);
            st.push4(c + 4, dv + 4);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) lin = fmaf(c[j], dv[j], lin);
        }
    }
    for (; d < p.D; ++d) {
        const float c = __ldcs(cp + (size_t)d * plane) * p.alpha;
        const float dv = PER_PIXEL ? __ldcs(sp + (size_t)d * plane) : disp_value(dvals, p, d);
        if (p.normalize)
            st.push1(c, dv);
        else
            lin = fmaf(c, dv, lin);
    }
    out[(size_t)b * plane + pix] = p.normalize ? st.result() : lin;
}

__global__ void __launch_bounds__(256) local_soft_argmin_kernel(const float* __restrict__ cost, float* __restrict__ out,
                                                                int D, size_t plane, int radius, int rdil, float alpha,
                                                                float start_disp, float dilation) {
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= plane) return;
    const int b = blockIdx.y;
    const float* cp = cost + (size_t)b * D * plane + pix;
    // argmax over D, first maximum wins (torch.argmax)
    float best = -INFINITY;
    int bi = 0;
    int d = 0;
    for (; d + 8 <= D; d += 8) {
        float c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = __ldg(cp + (size_t)(d + j) * plane);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (c[j] > best) {
                best = c[j];
                bi = d + j;
            }
    }
    for (; d < D; ++d) {
        const float c = __ldg(cp + (size_t)d * plane);
        if (c > best) {
            best = c;
            bi = d;
        }
    }
    // window softmax; out-of-range taps: logit -10000*alpha, disparity of the clamped index
    SoftState st;
    st.init();
    for (int r = -radius; r <= radius; ++r) {
        const int idx = bi + r * rdil;
        const bool inside = idx >= 0 && idx <= D - 1;
        const int ic = min(max(idx, 0), D - 1);
        const float g = __ldg(cp + (size_t)ic * plane) * alpha;
        const float logit = inside ? g : (-10000.0f * alpha);
        st.push1(logit, fmaf((float)ic, dilation, start_disp));
    }
    out[(size_t)b * plane + pix] = st.result();
}

}  // namespace dmb

using namespace dmb;

extern "C" int dmb_b200_upsample_regress(const float* cost_low, const float* up_weight, float* cost_out, float* disp_out,
                                         int B, int Dl, int Hl, int Wl, int D, int H, int W, int mode, float alpha,
                                         int normalize, float start_disp, float disp_step, const float* disp_values,
                                         void* stream) {
    DMB_REQUIRE(cost_low, "upsample_regress: null cost_low");
    DMB_REQUIRE(cost_out || disp_out, "upsample_regress: neither cost_out nor disp_out requested");
    DMB_REQUIRE(B > 0 && Dl > 0 && Hl > 0 && Wl > 0 && D > 0 && H > 0 && W > 0, "upsample_regress: non-positive dimension");
    DMB_REQUIRE(mode == 0 || mode == 1, "upsample_regress: mode must be 0 (trilinear) or 1 (deconv 8/4/2)");
    if (mode == 1) {
        DMB_REQUIRE(up_weight, "upsample_regress: mode 1 needs up_weight");
        DMB_REQUIRE(D == 4 * Dl && H == 4 * Hl && W == 4 * Wl, "upsample_regress: mode 1 requires exact x4 upsampling");
    }
    DMB_REQUIRE(B <= 65535 && cdiv(H, 8) <= 65535, "upsample_regress: grid too large");
    RegressParams p{B, Dl, Hl, Wl, D, H, W, alpha, start_disp, disp_step, normalize ? 1 : 0};
    dim3 grid((unsigned)cdiv(W, 32), (unsigned)cdiv(H, 8), B);
    cudaStream_t s = as_stream(stream);
#define DMB_LAUNCH_UP(M, WC, RG) \
    upsample_regress_kernel<M, WC, RG><<<grid, 256, 0, s>>>(cost_low, up_weight, cost_out, disp_out, disp_values, p)
    const size_t tri_smem = (size_t)Dl * 256 * sizeof(float);
    const size_t fast_smem = tri_smem + (size_t)D * 20;
    if (mode == 0 && !cost_out && normalize && fast_smem <= 160 * 1024) {
        if (fast_smem > 48 * 1024)
            DMB_CUDA(cudaFuncSetAttribute(upsample_trilinear_regress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)fast_smem));
        upsample_trilinear_regress_kernel<<<grid, 256, fast_smem, s>>>(cost_low, disp_out, disp_values, p);
        return check_launch("upsample_trilinear_regress_kernel");
    }
    if (mode == 0 && tri_smem <= 160 * 1024) {
#define DMB_LAUNCH_TRI(WC, RG)                                                                                          \
    do {                                                                                                                \
        if (tri_smem > 48 * 1024)                                                                                       \
            DMB_CUDA(cudaFuncSetAttribute(upsample_trilinear_kernel<WC, RG>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)tri_smem));                                                              \
        upsample_trilinear_kernel<WC, RG><<<grid, 256, tri_smem, s>>>(cost_low, cost_out, disp_out, disp_values, p);    \
    } while (0)
        if (cost_out && disp_out) DMB_LAUNCH_TRI(true, true);
        else if (cost_out) DMB_LAUNCH_TRI(true, false);
        else DMB_LAUNCH_TRI(false, true);
#undef DMB_LAUNCH_TRI
        return check_launch("upsample_trilinear_kernel");
    }
    if (mode == 0) {
        if (cost_out && disp_out) DMB_LAUNCH_UP(0, true, true);
        else if (cost_out) DMB_LAUNCH_UP(0, true, false);
        else DMB_LAUNCH_UP(0, false, true);
    } else {
        if (cost_out && disp_out) DMB_LAUNCH_UP(1, true, true);
        else if (cost_out) DMB_LAUNCH_UP(1, true, false);
        else DMB_LAUNCH_UP(1, false, true);
    }
#undef DMB_LAUNCH_UP
    return check_launch("upsample_regress_kernel");
}

extern "C" int dmb_b200_soft_argmin(const float* cost, float* disp_out, int B, int D, int H, int W, float alpha,
                                    int normalize, float start_disp, float disp_step, const float* disp_values,
                                    const float* disp_sample, void* stream) {
    DMB_REQUIRE(cost && disp_out, "soft_argmin: null pointer");
    DMB_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "soft_argmin: non-positive dimension");
    DMB_REQUIRE(!(disp_values && disp_sample), "soft_argmin: disp_values and disp_sample are exclusive");
    DMB_REQUIRE(B <= 65535, "soft_argmin: batch too large");
    RegressParams p{B, 0, 0, 0, D, H, W, alpha, start_disp, disp_step, normalize ? 1 : 0};
    const size_t plane = (size_t)H * W;
    dim3 grid((unsigned)cdiv(plane, 256), B);
    if (disp_sample)
        soft_argmin_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(cost, disp_out, nullptr, disp_sample, p);
    else
        soft_argmin_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(cost, disp_out, disp_values, nullptr, p);
    return check_launch("soft_argmin_kernel");
}

extern "C" int dmb_b200_local_soft_argmin(const float* cost, float* disp_out, int B, int D, int H, int W, int radius,
                                          int radius_dilation, float alpha, float start_disp, float dilation,
                                          void* stream) {
    DMB_REQUIRE(cost && disp_out, "local_soft_argmin: null pointer");
    DMB_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "local_soft_argmin: non-positive dimension");
    DMB_REQUIRE(radius >= 0 && radius_dilation >= 1, "local_soft_argmin: bad radius");
    DMB_REQUIRE(B <= 65535, "local_soft_argmin: batch too large");
    const size_t plane = (size_t)H * W;
    dim3 grid((unsigned)cdiv(plane, 256), B);
    local_soft_argmin_kernel<<<grid, 256, 0, as_stream(stream)>>>(cost, disp_out, D, plane, radius, radius_dilation, alpha,
                                                                  start_disp, dilation);
    return check_launch("local_soft_argmin_kernel");
}
