// Sequential-scan aggregation ops: the SPN 3-neighbour gated scan of dmb.ops (forward and
// backward), and the GANet SGA / LGA layers.
//
// SPN reference: dmb/ops/spn/src/gaterecurrent2dnoind_kernel.cu:130-532 (one launch per column /
// row on the legacy stream); here ONE launch per call, one CTA per (n,c) plane marching over the
// scan axis with the previous line kept in shared memory.
// SGA / LGA have no reference code; semantics = oracle/dmb_oracle.py.
#include "common.cuh"

namespace dmb {

// ---------------------------------------------------------------------------------------
// SPN.  Element (p, t): p = index across the scan, t = index along the scan.
//   horizontal: (h, w) = (p, t);  vertical: (h, w) = (t, p).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ size_t spn_off(int p, int t, int W, int horizontal) {
    return horizontal ? (size_t)p * W + t : (size_t)t * W + p;
}

__global__ void __launch_bounds__(256) spn_forward_kernel(const float* __restrict__ X, const float* __restrict__ G1,
                                                          const float* __restrict__ G2, const float* __restrict__ G3,
                                                          float* __restrict__ Hout, int H, int W, int horizontal,
                                                          int reverse) {
    extern __shared__ float sh[];   // [2][P+2] previous / current line with a zero border
    const int P = horizontal ? H : W;
    const int L = horizontal ? W : H;
    const size_t base = (size_t)blockIdx.x * H * W;
    float* buf0 = sh;
    float* buf1 = sh + (P + 2);
    for (int i = threadIdx.x; i < 2 * (P + 2); i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
    for (int step = 0; step < L; ++step) {
        const int t = reverse ? L - 1 - step : step;
        float* prev = (step & 1) ? buf1 : buf0;
        float* cur = (step & 1) ? buf0 : buf1;
        for (int p = threadIdx.x; p < P; p += blockDim.x) {
            const size_t o = base + spn_off(p, t, W, horizontal);
            const float x = __ldg(X + o);
            float h = x;
            if (step > 0) {
                // gate of a neighbour outside the image reads 0 (get_gate_sf, kernel.cu:79-97)
                const float g1 = (p > 0) ? __ldg(G1 + o) : 0.f;
                const float g2 = __ldg(G2 + o);
                const float g3 = (p < P - 1) ? __ldg(G3 + o) : 0.f;
                h = (1.f - g1 - g2 - g3) * x + (g1 * prev[p] + g2 * prev[p + 1] + g3 * prev[p + 2]);
            }
            cur[p + 1] = h;
            Hout[o] = h;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) spn_backward_kernel(const float* __restrict__ X, const float* __restrict__ G1,
                                                           const float* __restrict__ G2, const float* __restrict__ G3,
                                                           const float* __restrict__ Hout, const float* __restrict__ gout,
                                                           float* __restrict__ gX, float* __restrict__ gG1,
                                                           float* __restrict__ gG2, float* __restrict__ gG3, int H, int W,
                                                           int horizontal, int reverse) {
    // shared: dh of the later line (t+1) multiplied by each of its three gates, zero bordered
    extern __shared__ float sh[];   // [2][3][P+2]
    const int P = horizontal ? H : W;
    const int L = horizontal ? W : H;
    const size_t base = (size_t)blockIdx.x * H * W;
    const int stride = 3 * (P + 2);
    for (int i = threadIdx.x; i < 2 * stride; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
    // walk the scan backwards: the forward's last step first
    for (int step = L - 1; step >= 0; --step) {
        const int t = reverse ? L - 1 - step : step;
        const int tprev = reverse ? t + 1 : t - 1;      // the line the forward read at this step
        float* later = sh + ((step & 1) ? stride : 0);  // products of line step+1
        float* mine = sh + ((step & 1) ? 0 : stride);
        for (int p = threadIdx.x; p < P; p += blockDim.x) {
            const size_t o = base + spn_off(p, t, W, horizontal);
            // dh(t,p) = gout + g1(t+1,p+1) dh(t+1,p+1) + g2(t+1,p) dh(t+1,p) + g3(t+1,p-1) dh(t+1,p-1)
            const float dh = __ldg(gout + o) + later[0 * (P + 2) + p + 2] + later[1 * (P + 2) + p + 1] +
                             later[2 * (P + 2) + p];
            float g1 = 0.f, g2 = 0.f, g3 = 0.f;
            float d1 = 0.f, d2 = 0.f, d3 = 0.f;
            if (step > 0) {
                const float x = __ldg(X + o);
                if (p > 0) {
                    g1 = __ldg(G1 + o);
                    d1 = dh * (__ldg(Hout + base + spn_off(p - 1, tprev, W, horizontal)) - x);
                }
                g2 = __ldg(G2 + o);
                d2 = dh * (__ldg(Hout + base + spn_off(p, tprev, W, horizontal)) - x);
                if (p < P - 1) {
                    g3 = __ldg(G3 + o);
                    d3 = dh * (__ldg(Hout + base + spn_off(p + 1, tprev, W, horizontal)) - x);
                }
            }
            gX[o] = (1.f - g1 - g2 - g3) * dh;
            gG1[o] = d1;
            gG2[o] = d2;
            gG3[o] = d3;
            mine[0 * (P + 2) + p + 1] = g1 * dh;
            mine[1 * (P + 2) + p + 1] = g2 * dh;
            mine[2 * (P + 2) + p + 1] = g3 * dh;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// SGA.  One CTA per group of LPB scan lines of one (b,c); thread = (d, line).
// dir 0: left->right, 1: right->left (lines = rows y, steps along x)
// dir 2: top->bottom, 3: bottom->top (lines = columns x, steps along y)
// The first direction writes `out`, the others max into it.
// ---------------------------------------------------------------------------------------
template <int LPB>
__global__ void __launch_bounds__(1024) sga_dir_kernel(const float* __restrict__ x, const float* __restrict__ guid,
                                                       float* __restrict__ out, int C, int D, int H, int W, int dir,
                                                       int first) {
    extern __shared__ float sh[];   // prev[2][(D+2)*LPB] + partial max [nwarps][LPB]
    const bool horiz = dir < 2;
    const int nlines = horiz ? H : W;
    const int L = horiz ? W : H;
    const int groups = (nlines + LPB - 1) / LPB;
    const int grp = blockIdx.x % groups;
    const int bc = blockIdx.x / groups;
    const int c = bc % C, b = bc / C;
    const int li = threadIdx.x % LPB;
    const int d = threadIdx.x / LPB;
    const int line = grp * LPB + li;
    const bool live = line < nlines && d < D;
    const int nwarps = (blockDim.x + 31) / 32;
    float* prevbuf = sh;                                  // [2][(D+2)][LPB], rows 0 and D+1 stay 0
    float* pmax = sh + 2 * (D + 2) * LPB;                 // [nwarps][LPB]
    for (int i = threadIdx.x; i < 2 * (D + 2) * LPB; i += blockDim.x) prevbuf[i] = 0.f;
    __syncthreads();

    const size_t plane = (size_t)H * W;
    const float* xb = x + ((size_t)(b * C + c) * D + (live ? d : 0)) * plane;
    float* ob = out + ((size_t)(b * C + c) * D + (live ? d : 0)) * plane;
    // guidance [B,4,5,C,H,W]
    const float* gb = guid + (((size_t)(b * 4 + dir) * 5) * C + c) * plane;
    const size_t gk = (size_t)C * plane;
    const bool rev = (dir == 1 || dir == 3);

    for (int step = 0; step < L; ++step) {
        const int t = rev ? L - 1 - step : step;
        float* prev = prevbuf + ((step & 1) ? (D + 2) * LPB : 0);
        float* cur = prevbuf + ((step & 1) ? 0 : (D + 2) * LPB);
        const size_t pos = horiz ? (size_t)line * W + t : (size_t)t * W + line;
        float val = 0.f;
        if (live) {
            float w[5];
            float nrm = 0.f;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                w[k] = __ldg(gb + k * gk + pos);
                nrm += fabsf(w[k]);
            }
            nrm = fmaxf(nrm, 1e-12f);
            val = (w[0] / nrm) * __ldg(xb + pos);
            if (step > 0) {
                // max over d of the previous line: partial maxima were left in pmax
                float mx = pmax[li];
                for (int wi = 1; wi < nwarps; ++wi) mx = fmaxf(mx, pmax[wi * LPB + li]);
                val += (w[1] / nrm) * prev[(d + 1) * LPB + li] + (w[2] / nrm) * prev[d * LPB + li] +
                       (w[3] / nrm) * prev[(d + 2) * LPB + li] + (w[4] / nrm) * mx;
            }
            cur[(d + 1) * LPB + li] = val;
            if (first)
                ob[pos] = val;
            else
                ob[pos] = fmaxf(ob[pos], val);
        }
        __syncthreads();   // everyone has consumed pmax / prev of this step
        // per-line max over d of the line just produced
        float m = live ? val : -INFINITY;
#pragma unroll
        for (int o = 16; o >= LPB; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31) < LPB) pmax[(threadIdx.x >> 5) * LPB + (threadIdx.x & 31)] = m;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// LGA, radius 2 (5x5 window, 75 weights per pixel held in registers).  One thread per pixel,
// marching over d with the three d-planes of the 5x5 neighbourhood kept in registers, so each
// step loads 25 new values and issues 75 FMAs.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) lga_r2_kernel(const float* __restrict__ x, const float* __restrict__ guid,
                                                     float* __restrict__ out, int D, int H, int W) {
    const int px = blockIdx.x * 32 + (threadIdx.x & 31);
    const int py = blockIdx.y * 4 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    if (px >= W || py >= H) return;
    const size_t plane = (size_t)H * W;
    const float* gb = guid + (size_t)b * 75 * plane + (size_t)py * W + px;
    float w[75];
    float nrm = 0.f;
#pragma unroll
    for (int i = 0; i < 75; ++i) {
        w[i] = __ldg(gb + (size_t)i * plane);
        nrm += fabsf(w[i]);
    }
    nrm = fmaxf(nrm, 1e-12f);
#pragma unroll
    for (int i = 0; i < 75; ++i) w[i] = w[i] / nrm;

    int noff[25];
    bool nv[25];
#pragma unroll
    for (int ky = 0; ky < 5; ++ky)
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
            const int yy = py + ky - 2, xx = px + kx - 2;
            nv[ky * 5 + kx] = yy >= 0 && yy < H && xx >= 0 && xx < W;
            noff[ky * 5 + kx] = nv[ky * 5 + kx] ? yy * W + xx : 0;
        }
    const float* xb = x + (size_t)b * D * plane;
    float* ob = out + (size_t)b * D * plane + (size_t)py * W + px;
    float vm[25], v0[25], vp[25];   // planes d-1, d, d+1
#pragma unroll
    for (int i = 0; i < 25; ++i) {
        vm[i] = 0.f;
        v0[i] = nv[i] ? __ldg(xb + noff[i]) : 0.f;
    }
    for (int d = 0; d < D; ++d) {
        const bool has_next = d + 1 < D;
        const float* xn = xb + (size_t)(has_next ? d + 1 : d) * plane;
#pragma unroll
        for (int i = 0; i < 25; ++i) vp[i] = (has_next && nv[i]) ? __ldg(xn + noff[i]) : 0.f;
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 25; ++i) {
            acc = fmaf(w[i], v0[i], acc);
            acc = fmaf(w[25 + i], vm[i], acc);
            acc = fmaf(w[50 + i], vp[i], acc);
        }
        st_cs_f(ob + (size_t)d * plane, acc);
#pragma unroll
        for (int i = 0; i < 25; ++i) {
            vm[i] = v0[i];
            v0[i] = vp[i];
        }
    }
}

// any radius: weights re-read per use (slow path, kept for generality)
__global__ void __launch_bounds__(256) lga_generic_kernel(const float* __restrict__ x, const float* __restrict__ guid,
                                                          float* __restrict__ out, int D, int H, int W, int radius) {
    const int K = 2 * radius + 1;
    const size_t plane = (size_t)H * W;
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= plane) return;
    const int px = pix % W, py = pix / W;
    const int b = blockIdx.y;
    const float* gb = guid + (size_t)b * 3 * K * K * plane + pix;
    float nrm = 0.f;
    for (int i = 0; i < 3 * K * K; ++i) nrm += fabsf(__ldg(gb + (size_t)i * plane));
    nrm = fmaxf(nrm, 1e-12f);
    const float* xb = x + (size_t)b * D * plane;
    for (int d = 0; d < D; ++d) {
        float acc = 0.f;
        for (int t = 0; t < 3; ++t) {
            const int dd = d + (t == 0 ? 0 : (t == 1 ? -1 : 1));
            if (dd < 0 || dd >= D) continue;
            for (int ky = 0; ky < K; ++ky) {
                const int yy = py + ky - radius;
                if (yy < 0 || yy >= H) continue;
                for (int kx = 0; kx < K; ++kx) {
                    const int xx = px + kx - radius;
                    if (xx < 0 || xx >= W) continue;
                    const float wv = __ldg(gb + (size_t)((t * K + ky) * K + kx) * plane) / nrm;
                    acc = fmaf(wv, __ldg(xb + (size_t)dd * plane + (size_t)yy * W + xx), acc);
                }
            }
        }
        out[(size_t)b * D * plane + (size_t)d * plane + pix] = acc;
    }
}

}  // namespace dmb

using namespace dmb;

extern "C" int dmb_b200_spn_forward(const float* X, const float* G1, const float* G2, const float* G3, float* Hout, int N,
                                    int C, int H, int W, int horizontal, int reverse, void* stream) {
    DMB_REQUIRE(X && G1 && G2 && G3 && Hout, "spn_forward: null pointer");
    DMB_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, "spn_forward: non-positive dimension");
    const int P = horizontal ? H : W;
    const size_t smem = (size_t)2 * (P + 2) * 4;
    DMB_REQUIRE(smem <= 48 * 1024, "spn_forward: line of %d elements too long", P);
    spn_forward_kernel<<<N * C, 256, smem, as_stream(stream)>>>(X, G1, G2, G3, Hout, H, W, horizontal ? 1 : 0,
                                                                reverse ? 1 : 0);
    return check_launch("spn_forward_kernel");
}

extern "C" int dmb_b200_spn_backward(const float* X, const float* G1, const float* G2, const float* G3, const float* Hout,
                                     const float* grad_out, float* gX, float* gG1, float* gG2, float* gG3, int N, int C,
                                     int H, int W, int horizontal, int reverse, void* stream) {
    DMB_REQUIRE(X && G1 && G2 && G3 && Hout && grad_out && gX && gG1 && gG2 && gG3, "spn_backward: null pointer");
    DMB_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, "spn_backward: non-positive dimension");
    const int P = horizontal ? H : W;
    const size_t smem = (size_t)2 * 3 * (P + 2) * 4;
    DMB_REQUIRE(smem <= 48 * 1024, "spn_backward: line of %d elements too long", P);
    spn_backward_kernel<<<N * C, 256, smem, as_stream(stream)>>>(X, G1, G2, G3, Hout, grad_out, gX, gG1, gG2, gG3, H, W,
                                                                 horizontal ? 1 : 0, reverse ? 1 : 0);
    return check_launch("spn_backward_kernel");
}

extern "C" int dmb_b200_sga(const float* x, const float* guidance, float* out, int B, int C, int D, int H, int W,
                            void* stream) {
    DMB_REQUIRE(x && guidance && out, "sga: null pointer");
    DMB_REQUIRE(B > 0 && C > 0 && D > 0 && H > 0 && W > 0, "sga: non-positive dimension");
    DMB_REQUIRE(D <= 1024, "sga: D=%d exceeds 1024", D);
    for (int dir = 0; dir < 4; ++dir) {
        const bool horiz = dir < 2;
        // vertical scans: 8 adjacent columns per CTA so that every 32-byte sector is fully used
        int lpb = horiz ? 1 : 8;
        while (lpb > 1 && D * lpb > 1024) lpb >>= 1;
        const int nlines = horiz ? H : W;
        const int groups = (nlines + lpb - 1) / lpb;
        const int threads = ((D * lpb + 31) / 32) * 32;
        const int nwarps = threads / 32;
        const size_t smem = ((size_t)2 * (D + 2) * lpb + (size_t)nwarps * lpb) * 4;
        const unsigned grid = (unsigned)((size_t)B * C * groups);
        const int first = dir == 0 ? 1 : 0;
        cudaStream_t s = as_stream(stream);
        switch (lpb) {
            case 1: sga_dir_kernel<1><<<grid, threads, smem, s>>>(x, guidance, out, C, D, H, W, dir, first); break;
            case 2: sga_dir_kernel<2><<<grid, threads, smem, s>>>(x, guidance, out, C, D, H, W, dir, first); break;
            case 4: sga_dir_kernel<4><<<grid, threads, smem, s>>>(x, guidance, out, C, D, H, W, dir, first); break;
            default: sga_dir_kernel<8><<<grid, threads, smem, s>>>(x, guidance, out, C, D, H, W, dir, first); break;
        }
        int rc = check_launch("sga_dir_kernel");
        if (rc) return rc;
    }
    return DMB_OK;
}

extern "C" int dmb_b200_lga(const float* x, const float* guidance, float* out, int B, int D, int H, int W, int radius,
                            void* stream) {
    DMB_REQUIRE(x && guidance && out, "lga: null pointer");
    DMB_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && radius >= 0, "lga: bad dimension");
    DMB_REQUIRE(B <= 65535, "lga: batch too large");
    if (radius == 2) {
        dim3 grid((unsigned)cdiv(W, 32), (unsigned)cdiv(H, 4), B);
        DMB_REQUIRE(grid.y <= 65535, "lga: image too tall");
        lga_r2_kernel<<<grid, 128, 0, as_stream(stream)>>>(x, guidance, out, D, H, W);
        return check_launch("lga_r2_kernel");
    }
    dim3 grid((unsigned)cdiv((size_t)H * W, 256), B);
    lga_generic_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, guidance, out, D, H, W, radius);
    return check_launch("lga_generic_kernel");
}
