// Sequential-scan aggregation ops: the SPN 3-neighbour gated scan of dmb.ops (forward and
// backward), and the GANet SGA / LGA layers.
//
// SPN reference: dmb/ops/spn/src/gaterecurrent2dnoind_kernel.cu:130-532 (one launch per column /
// row on the legacy stream); here ONE launch per call, one CTA per (n,c) plane marching over the
// scan axis with the previous line kept in shared memory.
// SGA / LGA have no reference code; semantics = oracle/dmb_oracle.py.
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

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
// SGA, tiled kernels (coalesced HBM access in both scan orientations).
//
// Vertical scans (dir 2 top->bottom, 3 bottom->top): a CTA owns 32 adjacent columns of one (b,c);
// lane = column, warp = group of DPT consecutive disparities held in registers.  Neighbouring
// disparities and the max over d cross warps through a tiny double-buffered shared array
// (one barrier per step).  Every global access is a 128-byte row segment.
// ---------------------------------------------------------------------------------------
template <int DPT>
__global__ void __launch_bounds__(256) sga_vertical_kernel(const float* __restrict__ x, const float* __restrict__ guid,
                                                           float* __restrict__ out, int C, int D, int H, int W, int dir,
                                                           int first) {
    __shared__ float lo_s[2][8][32], hi_s[2][8][32], mx_s[2][8][32];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + lane;
    const int c = blockIdx.y, b = blockIdx.z;
    const bool colok = col < W;
    const int dbase = grp * DPT;
    const size_t plane = (size_t)H * W;
    const float* xb = x + ((size_t)(b * C + c) * D + dbase) * plane + (colok ? col : 0);
    float* ob = out + ((size_t)(b * C + c) * D + dbase) * plane + (colok ? col : 0);
    const float* gb = guid + (((size_t)(b * 4 + dir) * 5) * C + c) * plane + (colok ? col : 0);
    const size_t gk = (size_t)C * plane;
    const bool rev = dir == 3;
    float prev[DPT];
#pragma unroll
    for (int j = 0; j < DPT; ++j) prev[j] = 0.f;
    // software pipeline: the loads of step s+1 are issued before step s is computed
    float xn[DPT], on[DPT], wn[5];
    auto fetch = [&](int step) {
        const int y = rev ? H - 1 - step : step;
        const size_t row = (size_t)y * W;
#pragma unroll
        for (int k = 0; k < 5; ++k) wn[k] = colok ? __ldg(gb + k * gk + row) : 0.f;
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            const bool ok = colok && dbase + j < D;
            xn[j] = ok ? __ldg(xb + (size_t)j * plane + row) : 0.f;
            on[j] = (ok && !first) ? ob[(size_t)j * plane + row] : -INFINITY;
        }
    };
    fetch(0);
    for (int step = 0; step < H; ++step) {
        const int y = rev ? H - 1 - step : step;
        const size_t row = (size_t)y * W;
        float xv[DPT], ov[DPT], w[5];
        float nrm = 0.f;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            w[k] = wn[k];
            nrm += fabsf(w[k]);
        }
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            xv[j] = xn[j];
            ov[j] = on[j];
        }
        if (step + 1 < H) fetch(step + 1);
        const float inv = 1.0f / fmaxf(nrm, 1e-12f);
#pragma unroll
        for (int k = 0; k < 5; ++k) w[k] *= inv;
        float below = 0.f, above = 0.f, mx = 0.f;
        if (step > 0) {
            const int pb = (step - 1) & 1;
            below = grp > 0 ? hi_s[pb][grp - 1][lane] : 0.f;                  // A(p-r, dbase-1)
            above = grp < 7 ? lo_s[pb][grp + 1][lane] : 0.f;                  // A(p-r, dbase+DPT)
            mx = mx_s[pb][0][lane];
#pragma unroll
            for (int g = 1; g < 8; ++g) mx = fmaxf(mx, mx_s[pb][g][lane]);
        }
        float cur[DPT];
        float gmax = -INFINITY;
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            const int d = dbase + j;
            float v = 0.f;
            if (d < D) {
                v = w[0] * xv[j];
                if (step > 0) {
                    const float dm = j > 0 ? prev[j - 1] : below;
                    // the value one above the last valid disparity is outside the volume: 0
                    const float dp = (d + 1 < D) ? (j < DPT - 1 ? prev[j + 1] : above) : 0.f;
                    v += w[1] * prev[j] + w[2] * dm + w[3] * dp + w[4] * mx;
                }
                gmax = fmaxf(gmax, v);
                if (colok) ob[(size_t)j * plane + row] = fmaxf(ov[j], v);   // ov = -inf for the first direction
            }
            cur[j] = v;
        }
        const int cb = step & 1;
        lo_s[cb][grp][lane] = cur[0];
        hi_s[cb][grp][lane] = cur[DPT - 1];
        mx_s[cb][grp][lane] = gmax;
#pragma unroll
        for (int j = 0; j < DPT; ++j) prev[j] = cur[j];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// Horizontal scans (dir 0 left->right, 1 right->left): a CTA owns YT=8 rows of one (b,c); warp = row
// (an independent scan line), lane l holds disparities l, l+32, l+64, ... so that d-1 / d+1 are
// lane-neighbours (warp shuffles) and the max over d is a warp reduction: no barrier inside a tile.
// x / out / guidance travel through shared-memory tiles of TT scan steps, loaded and stored as 128-byte
// row segments and transposed on the way in ([y][t][d], odd pitch).
// ---------------------------------------------------------------------------------------
template <int NJ>   // NJ = ceil(D / 32) disparities per lane
__global__ void __launch_bounds__(128) sga_horizontal_kernel(const float* __restrict__ x, const float* __restrict__ guid,
                                                             float* __restrict__ out, int C, int D, int H, int W, int dir,
                                                             int first, int TT) {
    extern __shared__ float sh[];
    constexpr int YT = 4;                                   // rows (= warps) per CTA
    const int Dp = D | 1;                                   // odd pitch: conflict-free transposing stores
    float* xt = sh;                                         // [YT][TT][Dp]
    float* ot = xt + (size_t)YT * TT * Dp;                  // [YT][TT][Dp]
    float* gt = ot + (size_t)YT * TT * Dp;                  // [5][YT][TT]
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    const int y0 = blockIdx.x * YT;
    const int c = blockIdx.y, b = blockIdx.z;
    const size_t plane = (size_t)H * W;
    const float* xb = x + ((size_t)(b * C + c) * D) * plane;
    float* ob = out + ((size_t)(b * C + c) * D) * plane;
    const float* gb = guid + (((size_t)(b * 4 + dir) * 5) * C + c) * plane;
    const size_t gk = (size_t)C * plane;
    const bool rev = dir == 1;
    const int ntiles = (W + TT - 1) / TT;
    // tile transfers: a warp moves one (d, row) segment of TT values per iteration; lanes run along t
    const int tpl = TT < 32 ? TT : 32;                      // active lanes along t (TT is a power of two <= 32)
    const int rows_per_iter = (32 / tpl) * YT;              // (d,row) segments handled per CTA iteration
    const int seg_in_warp = lane / tpl;                     // which of the warp's segments this lane serves
    const int tl = lane % tpl;
    float prev[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) prev[j] = 0.f;
    bool started = false;
    for (int ti = 0; ti < ntiles; ++ti) {
        const int tile = rev ? ntiles - 1 - ti : ti;
        const int t0 = tile * TT;
        const int tx = t0 + tl;
        // batches of 8 (d,row) segments per thread: all global loads of a batch are in flight together
        for (int r0 = wy * (32 / tpl) + seg_in_warp; r0 < D * YT; r0 += 8 * rows_per_iter) {
            float v[8], o[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + u * rows_per_iter;
                const int yy = r % YT, d = r / YT;
                const int y = y0 + yy;
                v[u] = 0.f;
                o[u] = -INFINITY;
                if (r < D * YT && y < H && tx < W) {
                    const size_t a = (size_t)d * plane + (size_t)y * W + tx;
                    v[u] = __ldg(xb + a);
                    if (!first) o[u] = ob[a];
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + u * rows_per_iter;
                if (r < D * YT) {
                    const int yy = r % YT, d = r / YT;
                    const size_t si = ((size_t)yy * TT + tl) * Dp + d;
                    xt[si] = v[u];
                    ot[si] = o[u];
                }
            }
        }
        for (int r = wy * (32 / tpl) + seg_in_warp; r < 5 * YT; r += rows_per_iter) {
            const int yy = r % YT, k = r / YT;
            const int y = y0 + yy;
            gt[(k * YT + yy) * TT + tl] = (y < H && tx < W) ? __ldg(gb + k * gk + (size_t)y * W + tx) : 0.f;
        }
        __syncthreads();
        // ---- scan the TT steps of this warp's row ----
        if (y0 + wy < H) {
            const int nst = min(TT, W - t0);
            for (int s2 = 0; s2 < nst; ++s2) {
                const int t = rev ? nst - 1 - s2 : s2;
                float w[5], nrm = 0.f;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    w[k] = gt[(k * YT + wy) * TT + t];
                    nrm += fabsf(w[k]);
                }
                const float inv = 1.0f / fmaxf(nrm, 1e-12f);
#pragma unroll
                for (int k = 0; k < 5; ++k) w[k] *= inv;
                float mx = -INFINITY;
                if (started) {
#pragma unroll
                    for (int j = 0; j < NJ; ++j)
                        if (j * 32 + lane < D) mx = fmaxf(mx, prev[j]);
                    mx = warp_max(mx);
                }
                const float* xrow = xt + ((size_t)wy * TT + t) * Dp;
                float* orow = ot + ((size_t)wy * TT + t) * Dp;
                float cur[NJ];
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int d = j * 32 + lane;
                    // neighbours in d: lane-1 / lane+1 of the same j, wrapping to j-1 / j+1 at the warp edge
                    float dm = __shfl_up_sync(0xffffffffu, prev[j], 1);
                    const float wrap_dn = __shfl_sync(0xffffffffu, j > 0 ? prev[j > 0 ? j - 1 : 0] : 0.f, 31);
                    if (lane == 0) dm = wrap_dn;
                    float dp = __shfl_down_sync(0xffffffffu, prev[j], 1);
                    const float wrap_up = __shfl_sync(0xffffffffu, j < NJ - 1 ? prev[j < NJ - 1 ? j + 1 : j] : 0.f, 0);
                    if (lane == 31) dp = wrap_up;
                    if (d + 1 >= D) dp = 0.f;
                    float v = 0.f;
                    if (d < D) {
                        v = w[0] * xrow[d];
                        if (started) v += w[1] * prev[j] + w[2] * dm + w[3] * dp + w[4] * mx;
                        orow[d] = fmaxf(orow[d], v);      // -inf for the first direction
                    }
                    cur[j] = v;
                }
#pragma unroll
                for (int j = 0; j < NJ; ++j) prev[j] = cur[j];
                started = true;
            }
        } else {
            started = true;
        }
        __syncthreads();
        for (int r0 = wy * (32 / tpl) + seg_in_warp; r0 < D * YT; r0 += 8 * rows_per_iter) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + u * rows_per_iter;
                const int yy = r % YT, d = r / YT;
                v[u] = (r < D * YT) ? ot[((size_t)yy * TT + tl) * Dp + d] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + u * rows_per_iter;
                const int yy = r % YT, d = r / YT;
                const int y = y0 + yy;
                if (r < D * YT && y < H && tx < W) ob[(size_t)d * plane + (size_t)y * W + tx] = v[u];
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// SGA, register-resident kernels (D <= 128 / 64): a group of G adjacent lanes owns one scan line and each lane
// keeps DPL consecutive disparities of the running aggregate in registers, so the recurrence needs no shared
// memory and no barrier: d-1 / d+1 across the lane boundary are two shuffles, the max over d is a local max
// plus log2(G) xor-shuffles.  All HBM traffic is issued one step (vertical) or one 4-step chunk (horizontal)
// ahead of its use, straight from global memory:
//   vertical   (dir 2/3): lanes run along x -> every load/store instruction covers G row segments of 128/G bytes;
//   horizontal (dir 0/1): lanes run along y, a lane moves 16 bytes (4 scan steps) per disparity and chunk; the
//                         second half of each 32-byte sector is consumed one chunk later out of L1.
// ---------------------------------------------------------------------------------------
// max over the G adjacent lanes of a scan line's lane group as ONE redux.sync instead of log2(G) dependent
// shuffle + max rounds (the max over d sits on the critical path of every scan step): floats are mapped to integers
// of the same order (flip the magnitude bits of negative values), reduced with redux.sync.max.s32 over the group's
// own lane mask, and mapped back.  -inf / finite values only (NaN costs are not ordered by the reference either).
template <int G>
__device__ __forceinline__ float group_max(float v) {
    if (G == 1) return v;
    int i = __float_as_int(v);
    i ^= (i >> 31) & 0x7fffffff;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned mask = (G >= 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (lane & ~(unsigned)(G - 1)));
    i = __reduce_max_sync(mask, i);
    i ^= (i >> 31) & 0x7fffffff;
    return __int_as_float(i);
}

template <int DPL, int G>
__device__ __forceinline__ void sga_lane_step(float (&A)[DPL], const float (&xv)[DPL], float (&w)[5], bool started,
                                              int q, int dbase, int D, float (&cur)[DPL]) {
    float nrm = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) nrm += fabsf(w[k]);
    const float inv = 1.0f / fmaxf(nrm, 1e-12f);
#pragma unroll
    for (int k = 0; k < 5; ++k) w[k] *= inv;
    float below = __shfl_up_sync(0xffffffffu, A[DPL - 1], 1);      // A(p-r, dbase-1)
    float above = __shfl_down_sync(0xffffffffu, A[0], 1);          // A(p-r, dbase+DPL)
    if (q == 0) below = 0.f;
    if (q == G - 1) above = 0.f;
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < DPL; ++j)
        if (dbase + j < D) mx = fmaxf(mx, A[j]);
    mx = group_max<G>(mx);
#pragma unroll
    for (int j = 0; j < DPL; ++j) {
        const int d = dbase + j;
        float v = 0.f;
        if (d < D) {
            v = w[0] * xv[j];
            if (started) {
                const float dm = j > 0 ? A[j > 0 ? j - 1 : 0] : below;
                const float dp = (d + 1 < D) ? (j < DPL - 1 ? A[j < DPL - 1 ? j + 1 : j] : above) : 0.f;
                v += w[1] * A[j] + w[2] * dm + w[3] * dp + w[4] * mx;
            }
        }
        cur[j] = v;
    }
}

template <int DPL, int G>
__global__ void __launch_bounds__(128) sga_v_lanes_kernel(const float* __restrict__ x, const float* __restrict__ guid,
                                                          float* __restrict__ out, int C, int D, int H, int W, int dir,
                                                          int first) {
    constexpr int CPB = 128 / G;                              // columns per CTA
    const int q = threadIdx.x % G;
    const int col = blockIdx.x * CPB + threadIdx.x / G;
    const int c = blockIdx.y, b = blockIdx.z;
    const bool colok = col < W;
    const int dbase = q * DPL;
    const size_t plane = (size_t)H * W;
    const size_t base = ((size_t)(b * C + c) * D + dbase) * plane + (colok ? col : 0);
    const float* xb = x + base;
    float* ob = out + base;
    const float* gb = guid + (((size_t)(b * 4 + dir) * 5) * C + c) * plane + (colok ? col : 0);
    const size_t gk = (size_t)C * plane;
    const bool rev = dir == 3;
    float A[DPL], xn[DPL], on[DPL], wn[5];
#pragma unroll
    for (int j = 0; j < DPL; ++j) A[j] = 0.f;
    auto fetch = [&](int step) {
        const size_t row = (size_t)(rev ? H - 1 - step : step) * W;
#pragma unroll
        for (int k = 0; k < 5; ++k) wn[k] = colok ? __ldg(gb + k * gk + row) : 0.f;
#pragma unroll
        for (int j = 0; j < DPL; ++j) {
            const bool ok = colok && dbase + j < D;
            xn[j] = ok ? __ldcs(xb + (size_t)j * plane + row) : 0.f;
            on[j] = (ok && !first) ? __ldcs(ob + (size_t)j * plane + row) : -INFINITY;
        }
    };
    fetch(0);
    for (int step = 0; step < H; ++step) {
        const size_t row = (size_t)(rev ? H - 1 - step : step) * W;
        float xv[DPL], ov[DPL], w[5], cur[DPL];
#pragma unroll
        for (int k = 0; k < 5; ++k) w[k] = wn[k];
#pragma unroll
        for (int j = 0; j < DPL; ++j) {
            xv[j] = xn[j];
            ov[j] = on[j];
        }
        if (step + 1 < H) fetch(step + 1);
        sga_lane_step<DPL, G>(A, xv, w, step > 0, q, dbase, D, cur);
#pragma unroll
        for (int j = 0; j < DPL; ++j) {
            if (colok && dbase + j < D) ob[(size_t)j * plane + row] = fmaxf(ov[j], cur[j]);   // ov = -inf when first
            A[j] = cur[j];
        }
    }
}

template <int DPL, int G, bool REV>
__global__ void __launch_bounds__(128) sga_h_lanes_kernel(const float* __restrict__ x, const float* __restrict__ guid,
                                                          float* __restrict__ out, int C, int D, int H, int W, int dir,
                                                          int first) {
    constexpr int RPB = 128 / G;                              // rows per CTA
    const int q = threadIdx.x % G;
    const int rowi = blockIdx.x * RPB + threadIdx.x / G;
    const int c = blockIdx.y, b = blockIdx.z;
    const bool rowok = rowi < H;
    const int dbase = q * DPL;
    const size_t plane = (size_t)H * W;
    const size_t base = ((size_t)(b * C + c) * D + dbase) * plane + (size_t)(rowok ? rowi : 0) * W;
    const float* xb = x + base;
    float* ob = out + base;
    const float* gb = guid + (((size_t)(b * 4 + dir) * 5) * C + c) * plane + (size_t)(rowok ? rowi : 0) * W;
    const size_t gk = (size_t)C * plane;
    const int nchunks = W / 4;                                // W % 4 == 0 (checked by the host)
    float A[DPL];
#pragma unroll
    for (int j = 0; j < DPL; ++j) A[j] = 0.f;
    float4 xn[DPL], wn[5];
    auto fetch = [&](int ci) {
        const int t0 = 4 * ci;
#pragma unroll
        for (int k = 0; k < 5; ++k)
            wn[k] = rowok ? __ldg(reinterpret_cast<const float4*>(gb + k * gk + t0)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < DPL; ++j)
            xn[j] = (rowok && dbase + j < D) ? __ldg(reinterpret_cast<const float4*>(xb + (size_t)j * plane + t0))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    fetch(REV ? nchunks - 1 : 0);
    bool started = false;
    for (int i = 0; i < nchunks; ++i) {
        const int t0 = 4 * (REV ? nchunks - 1 - i : i);
        float xc[DPL][4], oc[DPL][4], wc[5][4];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            wc[k][0] = wn[k].x; wc[k][1] = wn[k].y; wc[k][2] = wn[k].z; wc[k][3] = wn[k].w;
        }
#pragma unroll
        for (int j = 0; j < DPL; ++j) {
            xc[j][0] = xn[j].x; xc[j][1] = xn[j].y; xc[j][2] = xn[j].z; xc[j][3] = xn[j].w;
            float4 o = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            if (!first && rowok && dbase + j < D) o = *reinterpret_cast<const float4*>(ob + (size_t)j * plane + t0);
            oc[j][0] = o.x; oc[j][1] = o.y; oc[j][2] = o.z; oc[j][3] = o.w;
        }
        if (i + 1 < nchunks) fetch(REV ? nchunks - 2 - i : i + 1);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            constexpr int dummy = 0;
            (void)dummy;
            const int e = REV ? 3 - s : s;                    // compile-time after unrolling
            float xv[DPL], w[5], cur[DPL];
#pragma unroll
            for (int k = 0; k < 5; ++k) w[k] = wc[k][e];
#pragma unroll
            for (int j = 0; j < DPL; ++j) xv[j] = xc[j][e];
            sga_lane_step<DPL, G>(A, xv, w, started, q, dbase, D, cur);
            started = true;
#pragma unroll
            for (int j = 0; j < DPL; ++j) {
                oc[j][e] = fmaxf(oc[j][e], cur[j]);           // -inf for the first direction
                A[j] = cur[j];
            }
        }
#pragma unroll
        for (int j = 0; j < DPL; ++j)
            if (rowok && dbase + j < D)
                *reinterpret_cast<float4*>(ob + (size_t)j * plane + t0) = make_float4(oc[j][0], oc[j][1], oc[j][2], oc[j][3]);
    }
}

// ---------------------------------------------------------------------------------------
// SGA, BIDIRECTIONAL kernels: both scans of one orientation in one launch.  The first half of a CTA's warps runs the
// forward scan (left->right / top->bottom) of the CTA's scan lines, the second half the backward scan of the SAME
// lines, concurrently.  Phase 1: each direction covers its half of the line; when the launch is the first one to
// write `out` (first = 1) the values are stored plainly, otherwise combined with max.  One __syncthreads.  Phase 2:
// each direction continues through the other half, where `out` now holds the other direction's phase-1 result (and
// whatever was there before), and max-combines.  Two launches (vertical pair, horizontal pair) instead of four,
// each with half the sequential length in flight per line -- and the host runs them over channel groups small
// enough for x and out to stay in L2 between the two launches (dmb_b200_sga below).
// ---------------------------------------------------------------------------------------
// PF: software-pipeline depth in scan steps (vertical) / 4-step chunks (horizontal).  A channel group is only a few
// dozen CTAs, so unlike the all-channel launches there is little thread-level parallelism to hide memory latency
// behind: every operand (x, the guidance taps and, where needed, the previous contents of out) is requested PF steps
// ahead into registers (measured with PF = 1: 2300 cycles per scan step, 0.8 TB/s).
template <int DPL, int G, int PF>
__global__ void __launch_bounds__(256) sga_v_bi_kernel(const float* __restrict__ x, const float* __restrict__ guid,
                                                       float* __restrict__ out, int C, int D, int H, int W, int first) {
    constexpr int CPB = 128 / G;                              // columns per CTA (each direction: 128 threads)
    const int back = threadIdx.x >> 7;                        // 0: top->bottom (dir 2), 1: bottom->top (dir 3)
    const int t = threadIdx.x & 127;
    const int q = t % G;
    const int col = blockIdx.x * CPB + t / G;
    const int c = blockIdx.y, b = blockIdx.z;
    const bool colok = col < W;
    const int dbase = q * DPL;
    const size_t plane = (size_t)H * W;
    const size_t base = ((size_t)(b * C + c) * D + dbase) * plane + (colok ? col : 0);
    const float* xb = x + base;
    float* ob = out + base;
    const float* gb = guid + (((size_t)(b * 4 + 2 + back) * 5) * C + c) * plane + (colok ? col : 0);
    const size_t gk = (size_t)C * plane;
    float A[DPL], xn[PF][DPL], on[PF][DPL], wn[PF][5];
#pragma unroll
    for (int j = 0; j < DPL; ++j) A[j] = 0.f;
    // forward covers rows [0, h1) in phase 1, backward rows [H-1 .. h1] (steps [0, H - h1))
    const int h1 = H / 2;
    const int n1 = back ? H - h1 : h1;
    auto run = [&](int s0, int s1, bool want_out) {
        auto fetch = [&](int u, int step) {                   // u is a compile-time stage index after unrolling
            const size_t row = (size_t)(back ? H - 1 - step : step) * W;
#pragma unroll
            for (int k = 0; k < 5; ++k) wn[u][k] = colok ? __ldg(gb + k * gk + row) : 0.f;
#pragma unroll
            for (int j = 0; j < DPL; ++j) {
                const bool ok = colok && dbase + j < D;
                xn[u][j] = ok ? __ldg(xb + (size_t)j * plane + row) : 0.f;      // (no evict-first: the horizontal pair re-reads x from L2)
                on[u][j] = (ok && want_out) ? ob[(size_t)j * plane + row] : -INFINITY;
            }
        };
#pragma unroll
        for (int u = 0; u < PF; ++u)
            if (s0 + u < s1) fetch(u, s0 + u);
        for (int sb = s0; sb < s1; sb += PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int step = sb + u;
                if (step < s1) {
                    const size_t row = (size_t)(back ? H - 1 - step : step) * W;
                    float xv[DPL], ov[DPL], w[5], cur[DPL];
#pragma unroll
                    for (int k = 0; k < 5; ++k) w[k] = wn[u][k];
#pragma unroll
                    for (int j = 0; j < DPL; ++j) {
                        xv[j] = xn[u][j];
                        ov[j] = on[u][j];
                    }
                    if (step + PF < s1) fetch(u, step + PF);
                    sga_lane_step<DPL, G>(A, xv, w, step > 0, q, dbase, D, cur);
#pragma unroll
                    for (int j = 0; j < DPL; ++j) {
                        if (colok && dbase + j < D) ob[(size_t)j * plane + row] = fmaxf(ov[j], cur[j]);   // ov = -inf: plain store
                        A[j] = cur[j];
                    }
                }
            }
        }
    };
    run(0, n1, !first);
    __syncthreads();                                          // phase-1 stores of both directions visible CTA-wide
    run(n1, H, true);
}

// TPD threads per direction (a multiple of 32): few rows per CTA spread a channel group's rows over many SMs -- the
// horizontal pair works out of L2 and a handful of SMs could not pull the group's slices fast enough
template <int DPL, int G, int TPD, int PF>
__global__ void __launch_bounds__(2 * TPD) sga_h_bi_kernel(const float* __restrict__ x, const float* __restrict__ guid,
                                                           float* __restrict__ out, int C, int D, int H, int W, int first) {
    constexpr int RPB = TPD / G;                              // rows per CTA
    const int back = threadIdx.x / TPD;                       // 0: left->right (dir 0), 1: right->left (dir 1)
    const int t = threadIdx.x % TPD;
    const int q = t % G;
    const int rowi = blockIdx.x * RPB + t / G;
    const int c = blockIdx.y, b = blockIdx.z;
    const bool rowok = rowi < H;
    const int dbase = q * DPL;
    const size_t plane = (size_t)H * W;
    const size_t base = ((size_t)(b * C + c) * D + dbase) * plane + (size_t)(rowok ? rowi : 0) * W;
    const float* xb = x + base;
    float* ob = out + base;
    const float* gb = guid + (((size_t)(b * 4 + back) * 5) * C + c) * plane + (size_t)(rowok ? rowi : 0) * W;
    const size_t gk = (size_t)C * plane;
    const int nchunks = W / 4;                                // W % 4 == 0 (checked by the host)
    float A[DPL];
#pragma unroll
    for (int j = 0; j < DPL; ++j) A[j] = 0.f;
    float4 xn[PF][DPL], on[PF][DPL], wn[PF][5];
    const int c1 = nchunks / 2;
    const int n1 = back ? nchunks - c1 : c1;                  // chunks of phase 1 for this direction
    bool started = false;
    const float4 ninf = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    auto run = [&](int i0, int i1, bool want_out) {
        auto fetch = [&](int u, int i) {                      // i = scan-order chunk index, u = stage (compile time)
            const int t0 = 4 * (back ? nchunks - 1 - i : i);
#pragma unroll
            for (int k = 0; k < 5; ++k)
                wn[u][k] = rowok ? __ldg(reinterpret_cast<const float4*>(gb + k * gk + t0)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < DPL; ++j) {
                const bool ok = rowok && dbase + j < D;
                xn[u][j] = ok ? __ldg(reinterpret_cast<const float4*>(xb + (size_t)j * plane + t0)) : make_float4(0.f, 0.f, 0.f, 0.f);
                on[u][j] = (ok && want_out) ? *reinterpret_cast<const float4*>(ob + (size_t)j * plane + t0) : ninf;
            }
        };
#pragma unroll
        for (int u = 0; u < PF; ++u)
            if (i0 + u < i1) fetch(u, i0 + u);
        for (int ib = i0; ib < i1; ib += PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int i = ib + u;
                if (i < i1) {
                    const int t0 = 4 * (back ? nchunks - 1 - i : i);
                    float xc[DPL][4], oc[DPL][4], wc[5][4];
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        wc[k][0] = wn[u][k].x; wc[k][1] = wn[u][k].y; wc[k][2] = wn[u][k].z; wc[k][3] = wn[u][k].w;
                    }
#pragma unroll
                    for (int j = 0; j < DPL; ++j) {
                        xc[j][0] = xn[u][j].x; xc[j][1] = xn[u][j].y; xc[j][2] = xn[u][j].z; xc[j][3] = xn[u][j].w;
                        oc[j][0] = on[u][j].x; oc[j][1] = on[u][j].y; oc[j][2] = on[u][j].z; oc[j][3] = on[u][j].w;
                    }
                    if (i + PF < i1) fetch(u, i + PF);
#pragma unroll
                    for (int s4 = 0; s4 < 4; ++s4) {
                        float xv[DPL], w[5], cur[DPL];
#pragma unroll
                        for (int k = 0; k < 5; ++k) w[k] = back ? wc[k][3 - s4] : wc[k][s4];
#pragma unroll
                        for (int j = 0; j < DPL; ++j) xv[j] = back ? xc[j][3 - s4] : xc[j][s4];
                        sga_lane_step<DPL, G>(A, xv, w, started, q, dbase, D, cur);
                        started = true;
#pragma unroll
                        for (int j = 0; j < DPL; ++j) {
                            if (back) oc[j][3 - s4] = fmaxf(oc[j][3 - s4], cur[j]);      // -inf: plain store
                            else oc[j][s4] = fmaxf(oc[j][s4], cur[j]);
                            A[j] = cur[j];
                        }
                    }
#pragma unroll
                    for (int j = 0; j < DPL; ++j)
                        if (rowok && dbase + j < D)
                            *reinterpret_cast<float4*>(ob + (size_t)j * plane + t0) = make_float4(oc[j][0], oc[j][1], oc[j][2], oc[j][3]);
                }
            }
        }
    };
    run(0, n1, !first);
    __syncthreads();                                          // phase-1 stores of both directions visible CTA-wide
    run(n1, nchunks, true);
}

// ---------------------------------------------------------------------------------------
// LGA, radius 2 (5x5 window, 75 L1-normalised weights per pixel held in registers).
// Tiled variant: a 32x8 pixel tile per CTA; the NEW depth plane (d+1) of the tile + halo is staged in
// shared memory once per step (double buffered, one barrier per step, the global loads of plane d+2
// are in flight while plane d is computed); the planes d-1, d, d+1 of each thread's 5x5 neighbourhood
// rotate through three register arrays (loop unrolled by 3, no moves).  25 shared loads + 75 FMAs per
// output value.
__global__ void __launch_bounds__(256, 1) lga_r2_tiled_kernel(const float* __restrict__ x, const float* __restrict__ guid,
                                                              float* __restrict__ out, int D, int H, int W) {
    constexpr int TX = 32, TY = 8, PW = TX + 4, PH = TY + 4, NE = PH * PW;   // 432 tile elements
    __shared__ float tile[2][NE];
    const int tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int px = x0 + tx, py = y0 + ty;
    const int b = blockIdx.z;
    const bool inside = px < W && py < H;
    const size_t plane = (size_t)H * W;
    float w[75];
    {
        const float* gb = guid + (size_t)b * 75 * plane + (size_t)(inside ? py : 0) * W + (inside ? px : 0);
        float nrm = 0.f;
#pragma unroll
        for (int i = 0; i < 75; ++i) {
            w[i] = __ldg(gb + (size_t)i * plane);
            nrm += fabsf(w[i]);
        }
        nrm = fmaxf(nrm, 1e-12f);
#pragma unroll
        for (int i = 0; i < 75; ++i) w[i] = w[i] / nrm;
    }
    const float* xb = x + (size_t)b * D * plane;
    float* ob = out + (size_t)b * D * plane + (size_t)py * W + px;
    // this thread stages tile elements e0 = tid and e1 = tid + 256 (if < NE); offsets are loop invariant
    int goff[2];
    bool gok[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int e = tid + k * 256;
        const int r = e / PW, c = e - r * PW;
        const int yy = y0 + r - 2, xx = x0 + c - 2;
        gok[k] = e < NE && yy >= 0 && yy < H && xx >= 0 && xx < W;
        goff[k] = gok[k] ? yy * W + xx : 0;
    }
    auto fetch = [&](int d, float (&r)[2]) {
        const bool dok = d < D;
        const float* q = xb + (size_t)(dok ? d : 0) * plane;
        r[0] = (dok && gok[0]) ? __ldg(q + goff[0]) : 0.f;
        r[1] = (dok && gok[1]) ? __ldg(q + goff[1]) : 0.f;
    };
    auto stage = [&](int slot, const float (&r)[2]) {
        tile[slot][tid] = r[0];
        if (tid + 256 < NE) tile[slot][tid + 256] = r[1];
    };
    const float* tbase0 = &tile[0][ty * PW + tx];
    const float* tbase1 = &tile[1][ty * PW + tx];
    float A[25], Bv[25], Cv[25];
    float pre[2];
    // prologue: plane 0 -> Bv (centre of d=0), A = 0 (plane -1); plane 1 staged in slot 1; plane 2 in flight
    fetch(0, pre);
    stage(0, pre);
    fetch(1, pre);
    __syncthreads();
#pragma unroll
    for (int ky = 0; ky < 5; ++ky)
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
            A[ky * 5 + kx] = 0.f;
            Bv[ky * 5 + kx] = tbase0[ky * PW + kx];
        }
    stage(1, pre);
    fetch(2, pre);
    __syncthreads();
    // step(d, M, Cn, Pl): plane d+1 (slot (d+1)&1) -> Pl; out(d) = w0.Cn + w1.M + w2.Pl; stage plane d+2, prefetch d+3
    auto step = [&](int d, float (&M)[25], float (&Cn)[25], float (&Pl)[25]) {
        const float* tb = ((d + 1) & 1) ? tbase1 : tbase0;
        // six independent accumulation chains (one 75-deep chain left the FMA pipe idle 3 cycles out of 4)
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
#pragma unroll
        for (int ky = 0; ky < 5; ++ky)
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) {
                const int i = ky * 5 + kx;
                Pl[i] = tb[ky * PW + kx];
                if (i & 1) {
                    b0 = fmaf(w[i], Cn[i], b0);
                    b1 = fmaf(w[25 + i], M[i], b1);
                    b2 = fmaf(w[50 + i], Pl[i], b2);
                } else {
                    a0 = fmaf(w[i], Cn[i], a0);
                    a1 = fmaf(w[25 + i], M[i], a1);
                    a2 = fmaf(w[50 + i], Pl[i], a2);
                }
            }
        const float acc = ((a0 + b0) + (a1 + b1)) + (a2 + b2);
        if (inside) st_cs_f(ob + (size_t)d * plane, acc);
        // slot d&1 held plane d (already consumed into registers one step ago): refill it with plane d+2
        stage(d & 1, pre);
        fetch(d + 3, pre);
        __syncthreads();
    };
    int d = 0;
    for (; d + 3 <= D; d += 3) {
        step(d, A, Bv, Cv);
        step(d + 1, Bv, Cv, A);
        step(d + 2, Cv, A, Bv);
    }
    if (d < D) {
        step(d, A, Bv, Cv);
        if (d + 1 < D) step(d + 1, Bv, Cv, A);
    }
}

// LGA, radius 2, accumulator-rotating variant (the production kernel).  Instead of holding the 25-pixel neighbourhood
// of three depth planes in registers (75 + 75 weights = 204 registers, one 8-warp CTA per SM, a barrier per depth step),
// every staged value x(q, d') is used the moment it is read from shared memory for its THREE consumers
//     out(d') += w0(q) x,   out(d'+1) += w1(q) x,   out(d'-1) += w2(q) x
// so a thread keeps the 75 weights and three (x2 for instruction-level parallelism) running sums: ~100 registers, two
// CTAs per SM.  NP depth planes of the tile + halo are staged per barrier pair with cp.async (zero-filled outside the
// volume), double buffered.  25 conflict-free shared loads + 75 FMAs per output value: FMA-issue bound.
template <int NP>
__global__ void __launch_bounds__(256, 2) lga_r2_rot_kernel(const float* __restrict__ x, const float* __restrict__ guid,
                                                            float* __restrict__ out, int D, int H, int W) {
    constexpr int TX = 32, TY = 8, PW = TX + 4, PH = TY + 4, NE = PH * PW;   // 432 tile elements per plane
    constexpr int NS = (NP * NE + 255) / 256;                                 // staged elements per thread and stage
    __shared__ __align__(16) float tile[2][NP * NE];
    const int tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int px = x0 + tx, py = y0 + ty;
    const int b = blockIdx.z;
    const bool inside = px < W && py < H;
    const size_t plane = (size_t)H * W;
    float w[75];
    {
        const float* gb = guid + (size_t)b * 75 * plane + (size_t)(inside ? py : 0) * W + (inside ? px : 0);
        float nrm = 0.f;
#pragma unroll
        for (int i = 0; i < 75; ++i) {
            w[i] = __ldg(gb + (size_t)i * plane);
            nrm += fabsf(w[i]);
        }
        nrm = fmaxf(nrm, 1e-12f);
#pragma unroll
        for (int i = 0; i < 75; ++i) w[i] = w[i] / nrm;
    }
    const float* xb = x + (size_t)b * D * plane;
    float* ob = out + (size_t)b * D * plane + (size_t)py * W + px;
    // loop-invariant staging descriptors: element e = tid + 256 k of a stage = (plane pl, tile position)
    int spl[NS], sgo[NS];                          // plane inside the stage (-1: no element), in-plane global offset (-1: outside)
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const int e = tid + k * 256;
        const int pl = e / NE, idx = e - pl * NE;
        const int r = idx / PW, c = idx - r * PW;
        const int yy = y0 + r - 2, xx = x0 + c - 2;
        spl[k] = e < NP * NE ? pl : -1;
        sgo[k] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? yy * W + xx : -1;
    }
    auto issue = [&](int st) {                     // planes st*NP .. st*NP+NP-1 -> buffer st & 1 (zero-filled outside)
        const int d0 = st * NP;
        float* dst = tile[st & 1];
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            if (spl[k] < 0) continue;
            const int d = d0 + spl[k];
            const bool ok = d < D && sgo[k] >= 0;
            const float* src = ok ? xb + (size_t)d * plane + sgo[k] : xb;
            const uint32_t sz = ok ? 4u : 0u;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst + tid + k * 256)), "l"(src), "r"(sz) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int nst = (D + 1 + NP - 1) / NP;         // planes 0 .. D (plane D is all zeros: it completes out(D-1))
    issue(0);
    issue(1);                                      // (an all-zero stage when nst == 1: harmless)
    float pa = 0.f, pb = 0.f, ca = 0.f, cb = 0.f, na = 0.f, nb = 0.f;   // sums for out(d-1), out(d), out(d+1)
    for (int st = 0; st < nst; ++st) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        const float* tb0 = &tile[st & 1][ty * PW + tx];
#pragma unroll
        for (int pl = 0; pl < NP; ++pl) {
            const float* tb = tb0 + pl * NE;
#pragma unroll
            for (int ky = 0; ky < 5; ++ky)
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) {
                    const int i = ky * 5 + kx;
                    const float v = tb[ky * PW + kx];
                    if (i & 1) {
                        cb = fmaf(w[i], v, cb);
                        nb = fmaf(w[25 + i], v, nb);
                        pb = fmaf(w[50 + i], v, pb);
                    } else {
                        ca = fmaf(w[i], v, ca);
                        na = fmaf(w[25 + i], v, na);
                        pa = fmaf(w[50 + i], v, pa);
                    }
                }
            const int d = st * NP + pl;            // the plane just consumed: out(d-1) is complete
            if (inside && d >= 1 && d <= D) st_cs_f(ob + (size_t)(d - 1) * plane, pa + pb);
            pa = ca; pb = cb; ca = na; cb = nb; na = 0.f; nb = 0.f;
        }
        __syncthreads();                           // everybody is done with buffer st & 1
        issue(st + 2);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// The same kernel with the NP-plane tile (+ halo) of every stage fetched by ONE 4-D TMA box load issued by one thread
// (zero fill outside the image and beyond the last depth plane comes from the tensor map): the cp.async staging above
// costs ~175 integer / address instructions per thread and stage -- 40 % of all instructions executed
// (profiles/r2_ncu_full_ops.txt) -- in a kernel that is issue bound.
__device__ __forceinline__ void tma_load_4d_f32(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

template <int NP>
__global__ void __launch_bounds__(256, 2) lga_r2_tma_kernel(const __grid_constant__ CUtensorMap xmap, const float* __restrict__ guid,
                                                            float* __restrict__ out, int D, int H, int W) {
    // The TMA unit traps (illegal instruction) on a box whose innermost start coordinate is not a multiple of 16 bytes
    // (measured: tools/tma_f32_probe.cu, fp32 c0 = -2 / 30 / 62 fault, 0 / -4 do not), so the halo tile starts at
    // x0 - 4 and is 40 floats wide; the 5 x 5 window of column tx begins at tile column tx + 2.
    constexpr int TX = 32, TY = 8, PW = TX + 8, PH = TY + 4, NE = PH * PW;   // 480 tile elements per plane
    __shared__ __align__(128) float tile[2][NP * NE];
    __shared__ uint64_t full[2];
    const int tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int px = x0 + tx, py = y0 + ty;
    const int b = blockIdx.z;
    const bool inside = px < W && py < H;
    const size_t plane = (size_t)H * W;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    const int nst = (D + 1 + NP - 1) / NP;         // planes 0 .. D (plane D is all zeros: it completes out(D-1))
    const CUtensorMap* const mp = &xmap;
    auto issue = [&, mp](int st) {
        if (tid == 0 && st < nst) {
            mbar_expect_tx(&full[st & 1], NP * NE * 4);
            tma_load_4d_f32(tile[st & 1], mp, &full[st & 1], x0 - 4, y0 - 2, st * NP, b);
        }
    };
    issue(0);
    issue(1);
    float w[75];
    {
        const float* gb = guid + (size_t)b * 75 * plane + (size_t)(inside ? py : 0) * W + (inside ? px : 0);
        float nrm = 0.f;
#pragma unroll
        for (int i = 0; i < 75; ++i) {
            w[i] = __ldg(gb + (size_t)i * plane);
            nrm += fabsf(w[i]);
        }
        nrm = fmaxf(nrm, 1e-12f);
#pragma unroll
        for (int i = 0; i < 75; ++i) w[i] = w[i] / nrm;
    }
    float* ob = out + (size_t)b * D * plane + (size_t)py * W + px;
    float pa = 0.f, pb = 0.f, ca = 0.f, cb = 0.f, na = 0.f, nb = 0.f;   // sums for out(d-1), out(d), out(d+1)
    for (int st = 0; st < nst; ++st) {
        mbar_wait(&full[st & 1], (st >> 1) & 1);
        const float* tb0 = &tile[st & 1][ty * PW + tx + 2];
#pragma unroll
        for (int pl = 0; pl < NP; ++pl) {
            const float* tb = tb0 + pl * NE;
#pragma unroll
            for (int ky = 0; ky < 5; ++ky)
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) {
                    const int i = ky * 5 + kx;
                    const float v = tb[ky * PW + kx];
                    if (i & 1) {
                        cb = fmaf(w[i], v, cb);
                        nb = fmaf(w[25 + i], v, nb);
                        pb = fmaf(w[50 + i], v, pb);
                    } else {
                        ca = fmaf(w[i], v, ca);
                        na = fmaf(w[25 + i], v, na);
                        pa = fmaf(w[50 + i], v, pa);
                    }
                }
            const int d = st * NP + pl;            // the plane just consumed: out(d-1) is complete
            if (inside && d >= 1 && d <= D) st_cs_f(ob + (size_t)(d - 1) * plane, pa + pb);
            pa = ca; pb = cb; ca = na; cb = nb; na = 0.f; nb = 0.f;
        }
        __syncthreads();                           // everybody is done with buffer st & 1
        issue(st + 2);
    }
}

// Two vertically adjacent pixels per thread: the 6 x 5 tile window of the pair is read once and feeds both pixels'
// sums (15 shared loads per output value instead of 25 -- the shared-memory pipe, 128 B per clock and SM, ran level with
// the FMA pipe in the one-pixel kernel: 100 B against 75 FMAs per output).  150 weights + 6 running sums per thread,
// 128 threads per CTA (same 32 x 8 tile, same TMA box), two CTAs per SM.
// MEASURED SLOWER (config 4: 0.504 ms against 0.456 ms for one pixel per thread; with the registers capped at 170 for
// three CTAs per SM: 0.70 ms): 189 registers leave 8 warps per SM, too few to cover the shared-load latency.  Kept
// as an A/B (DMB_B200_LGA_ROT=3); the one-pixel kernel is the default.
template <int NP>
__global__ void __launch_bounds__(128, 2) lga_r2_tma2_kernel(const __grid_constant__ CUtensorMap xmap, const float* __restrict__ guid,
                                                             float* __restrict__ out, int D, int H, int W) {
    constexpr int TX = 32, TY = 8, PW = TX + 8, PH = TY + 4, NE = PH * PW;   // tile starts at x0 - 4 (16-byte aligned box)
    __shared__ __align__(128) float tile[2][NP * NE];
    __shared__ uint64_t full[2];
    const int tid = threadIdx.x;
    const int tx = tid & 31, ty = (tid >> 5) * 2;                            // rows ty, ty + 1 of the tile
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int px = x0 + tx, py = y0 + ty;
    const int b = blockIdx.z;
    const bool in_a = px < W && py < H, in_b = px < W && py + 1 < H;
    const size_t plane = (size_t)H * W;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    const int nst = (D + 1 + NP - 1) / NP;         // planes 0 .. D (plane D is all zeros: it completes out(D-1))
    const CUtensorMap* const mp = &xmap;
    auto issue = [&, mp](int st) {
        if (tid == 0 && st < nst) {
            mbar_expect_tx(&full[st & 1], NP * NE * 4);
            tma_load_4d_f32(tile[st & 1], mp, &full[st & 1], x0 - 4, y0 - 2, st * NP, b);
        }
    };
    issue(0);
    issue(1);
    float wa[75], wb[75];
    {
        const float* ga = guid + (size_t)b * 75 * plane + (size_t)(in_a ? py : 0) * W + (in_a ? px : 0);
        const float* gb = guid + (size_t)b * 75 * plane + (size_t)(in_b ? py + 1 : 0) * W + (in_b ? px : 0);
        float na = 0.f, nb = 0.f;
#pragma unroll
        for (int i = 0; i < 75; ++i) {
            wa[i] = __ldg(ga + (size_t)i * plane);
            wb[i] = __ldg(gb + (size_t)i * plane);
            na += fabsf(wa[i]);
            nb += fabsf(wb[i]);
        }
        na = fmaxf(na, 1e-12f);
        nb = fmaxf(nb, 1e-12f);
#pragma unroll
        for (int i = 0; i < 75; ++i) {
            wa[i] = wa[i] / na;
            wb[i] = wb[i] / nb;
        }
    }
    float* oa = out + (size_t)b * D * plane + (size_t)py * W + px;
    float pa = 0.f, ca = 0.f, na = 0.f, pb = 0.f, cb = 0.f, nb = 0.f;      // sums for out(d-1), out(d), out(d+1) of A, B
    for (int st = 0; st < nst; ++st) {
        mbar_wait(&full[st & 1], (st >> 1) & 1);
        const float* tb0 = &tile[st & 1][ty * PW + tx + 2];
#pragma unroll
        for (int pl = 0; pl < NP; ++pl) {
            const float* tb = tb0 + pl * NE;
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) {
                    const float v = tb[r * PW + kx];
                    if (r < 5) {                   // pixel A: window row ky = r
                        const int i = r * 5 + kx;
                        ca = fmaf(wa[i], v, ca);
                        na = fmaf(wa[25 + i], v, na);
                        pa = fmaf(wa[50 + i], v, pa);
                    }
                    if (r > 0) {                   // pixel B (one row lower): window row ky = r - 1
                        const int i = (r - 1) * 5 + kx;
                        cb = fmaf(wb[i], v, cb);
                        nb = fmaf(wb[25 + i], v, nb);
                        pb = fmaf(wb[50 + i], v, pb);
                    }
                }
            const int d = st * NP + pl;            // the plane just consumed: out(d-1) is complete
            if (d >= 1 && d <= D) {
                if (in_a) st_cs_f(oa + (size_t)(d - 1) * plane, pa);
                if (in_b) st_cs_f(oa + (size_t)(d - 1) * plane + W, pb);
            }
            pa = ca; ca = na; na = 0.f;
            pb = cb; cb = nb; nb = 0.f;
        }
        __syncthreads();                           // everybody is done with buffer st & 1
        issue(st + 2);
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

template <int DPT>
static void launch_sga_v(const float* x, const float* g, float* out, int B, int C, int D, int H, int W, int dir, int first,
                         cudaStream_t s) {
    dim3 grid((unsigned)cdiv(W, 32), C, B);
    sga_vertical_kernel<DPT><<<grid, 256, 0, s>>>(x, g, out, C, D, H, W, dir, first);
}
template <int NJ>
static int launch_sga_h(const float* x, const float* g, float* out, int B, int C, int D, int H, int W, int dir, int first,
                        cudaStream_t s) {
    const int Dp = D | 1;
    int TT = 32;
    auto bytes = [&](int tt) { return ((size_t)2 * 4 * tt * Dp + 5 * 4 * tt) * sizeof(float); };
    while (TT > 4 && bytes(TT) > 72 * 1024) TT >>= 1;         // <= 72 KB per CTA: three CTAs per SM
    const size_t smem = bytes(TT);
    if (smem > 200 * 1024) return 1;
    if (smem > 48 * 1024)
        DMB_CUDA(cudaFuncSetAttribute(sga_horizontal_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)cdiv(H, 4), C, B);
    sga_horizontal_kernel<NJ><<<grid, 128, smem, s>>>(x, g, out, C, D, H, W, dir, first, TT);
    return 0;
}

static int g_sga_bidir = -1;     // -1: read DMB_B200_SGA_BIDIR on first use
extern "C" int dmb_b200_sga_set_bidirectional(int on) {       // selects the SGA schedule; returns the previous setting
    const int prev = g_sga_bidir;
    g_sga_bidir = on < 0 ? -1 : (on ? 1 : 0);
    return prev;
}

extern "C" int dmb_b200_sga(const float* x, const float* guidance, float* out, int B, int C, int D, int H, int W,
                            void* stream) {
    DMB_REQUIRE(x && guidance && out, "sga: null pointer");
    DMB_REQUIRE(B > 0 && C > 0 && D > 0 && H > 0 && W > 0, "sga: non-positive dimension");
    DMB_REQUIRE(D <= 1024, "sga: D=%d exceeds 1024", D);
    DMB_REQUIRE(C <= 65535 && B <= 65535, "sga: grid dimension too large");
    cudaStream_t s = as_stream(stream);
    const bool tiled = D <= 256;
    static int lanes_mode = -1;                    // DMB_B200_SGA_LANES=0: the shared-memory tiled kernels (A/B)
    if (lanes_mode < 0) {
        const char* e = getenv("DMB_B200_SGA_LANES");
        lanes_mode = (e && e[0] == '0') ? 0 : 1;
    }
    const bool aligned16 = (reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(guidance) |
                            reinterpret_cast<uintptr_t>(out)) % 16 == 0;
    // DMB_B200_SGA_BIDIR=1: the bidirectional, L2-blocked launches below.  OFF by default -- measured on B200
    // (profiles/README.md, round 2): they cut the DRAM traffic from 4.8 GB to 1.15 GB (1.15x algorithmic) as designed,
    // but run 4.4 ms against 1.9 ms for the four all-channel launches: with 16-byte-per-lane accesses every load /
    // store request touches 8 (vertical) or 32 (horizontal) distinct lines, the LSU processes about one line per cycle,
    // and a 3-channel group keeps only ~80-100 CTAs busy, so the kernels sit at the L1 tag throughput of those few SMs.
    // The fix is the one the north star names -- TMA-staged tiles so that global traffic bypasses the LSU.
    int bi_mode = g_sga_bidir;
    if (bi_mode < 0) {
        const char* e = getenv("DMB_B200_SGA_BIDIR");
        bi_mode = g_sga_bidir = (e && e[0] == '1') ? 1 : 0;
    }
    if (lanes_mode && bi_mode && D <= 64 && W % 4 == 0 && W >= 8 && H >= 2 && aligned16) {
        // Two bidirectional launches (vertical pair first: its x reads are the coalesced ones, so x comes from HBM
        // once, in full lines; then the horizontal pair) per CHANNEL GROUP, groups sized so that the group's x and
        // out slices (2 * GC * D*H*W*4 bytes) stay in the 126 MB L2 between the two launches: the horizontal pair
        // then reads x and read-modify-writes out in L2, and HBM sees x once, the guidance once and out once.
        const size_t slice = (size_t)D * H * W * 4;
        int GC = (int)((size_t)88 * 1024 * 1024 / (2 * slice));
        if (GC < 1) GC = 1;
        if (GC > C) GC = C;
        const size_t plane = (size_t)H * W;
        for (int b = 0; b < B; ++b) {
            for (int c0 = 0; c0 < C; c0 += GC) {
                const int gc = (C - c0 < GC) ? C - c0 : GC;
                // channel c0 of batch b: x / out advance by whole channel slices; the guidance pointer advances by c0
                // planes inside every (direction, tap) block, its per-tap stride stays C planes
                const float* xg = x + ((size_t)b * C + c0) * D * plane;
                float* og = out + ((size_t)b * C + c0) * D * plane;
                const float* gg = guidance + (size_t)b * 20 * C * plane + (size_t)c0 * plane;
#define DMB_SGA_BI(DPLV, GV, PFV, DPLH, GH, PFH)                                                                              \
    do {                                                                                                                     \
        sga_v_bi_kernel<DPLV, GV, PFV><<<dim3((unsigned)cdiv(W, 128 / GV), gc, 1), 256, 0, s>>>(xg, gg, og, C, D, H, W, 1);   \
        sga_h_bi_kernel<DPLH, GH, 64, PFH><<<dim3((unsigned)cdiv(H, 64 / GH), gc, 1), 128, 0, s>>>(xg, gg, og, C, D, H, W, 0); \
    } while (0)
                // vertical: 8 lanes per column (16 columns = 64-byte row segments per CTA and direction), 3 steps ahead;
                // horizontal: 16 lanes per row (4 disparities each: short per-step chains), 4 rows per CTA and
                // direction, 2-3 chunks (8-12 steps) ahead
                if (D <= 16) DMB_SGA_BI(2, 8, 4, 1, 16, 3);
                else if (D <= 32) DMB_SGA_BI(4, 8, 4, 2, 16, 3);
                else DMB_SGA_BI(8, 8, 3, 4, 16, 2);
#undef DMB_SGA_BI
                int rc = check_launch("sga_bi_kernels");
                if (rc) return rc;
            }
        }
        return DMB_OK;
    }
    for (int dir = 0; dir < 4; ++dir) {
        const int first = dir == 0 ? 1 : 0;
        if (lanes_mode && dir < 2 && D <= 64 && W % 4 == 0 && aligned16) {
            // horizontal, G = 8 lanes per row, DPL = ceil(D / 8) rounded up to 2 / 4 / 8
            dim3 grid((unsigned)cdiv(H, 16), C, B);
#define DMB_SGA_H(DPL)                                                                                              \
    do {                                                                                                            \
        if (dir == 1) sga_h_lanes_kernel<DPL, 8, true><<<grid, 128, 0, s>>>(x, guidance, out, C, D, H, W, dir, first); \
        else sga_h_lanes_kernel<DPL, 8, false><<<grid, 128, 0, s>>>(x, guidance, out, C, D, H, W, dir, first);      \
    } while (0)
            if (D <= 16) DMB_SGA_H(2);
            else if (D <= 32) DMB_SGA_H(4);
            else DMB_SGA_H(8);
#undef DMB_SGA_H
            int rc = check_launch("sga_h_lanes_kernel");
            if (rc) return rc;
            continue;
        }
        if (lanes_mode && dir >= 2 && D <= 128) {
            if (D <= 8) { dim3 grid((unsigned)cdiv(W, 64), C, B); sga_v_lanes_kernel<4, 2><<<grid, 128, 0, s>>>(x, guidance, out, C, D, H, W, dir, first); }
            else if (D <= 16) { dim3 grid((unsigned)cdiv(W, 64), C, B); sga_v_lanes_kernel<8, 2><<<grid, 128, 0, s>>>(x, guidance, out, C, D, H, W, dir, first); }
            else if (D <= 32) { dim3 grid((unsigned)cdiv(W, 64), C, B); sga_v_lanes_kernel<16, 2><<<grid, 128, 0, s>>>(x, guidance, out, C, D, H, W, dir, first); }
            else if (D <= 64) { dim3 grid((unsigned)cdiv(W, 32), C, B); sga_v_lanes_kernel<16, 4><<<grid, 128, 0, s>>>(x, guidance, out, C, D, H, W, dir, first); }
            else { dim3 grid((unsigned)cdiv(W, 32), C, B); sga_v_lanes_kernel<32, 4><<<grid, 128, 0, s>>>(x, guidance, out, C, D, H, W, dir, first); }
            int rc = check_launch("sga_v_lanes_kernel");
            if (rc) return rc;
            continue;
        }
        if (tiled && dir < 2) {
            const int nj = (D + 31) / 32;
            int rc = 0;
            switch (nj) {
                case 1: rc = launch_sga_h<1>(x, guidance, out, B, C, D, H, W, dir, first, s); break;
                case 2: rc = launch_sga_h<2>(x, guidance, out, B, C, D, H, W, dir, first, s); break;
                case 3: rc = launch_sga_h<3>(x, guidance, out, B, C, D, H, W, dir, first, s); break;
                case 4: rc = launch_sga_h<4>(x, guidance, out, B, C, D, H, W, dir, first, s); break;
                case 5: case 6: rc = launch_sga_h<6>(x, guidance, out, B, C, D, H, W, dir, first, s); break;
                default: rc = launch_sga_h<8>(x, guidance, out, B, C, D, H, W, dir, first, s); break;
            }
            if (rc < 0) return rc;
            if (rc == 0) {
                rc = check_launch("sga_horizontal_kernel");
                if (rc) return rc;
                continue;
            }
        } else if (tiled) {
            const int dpt = (D + 7) / 8;
            if (dpt <= 4) launch_sga_v<4>(x, guidance, out, B, C, D, H, W, dir, first, s);
            else if (dpt <= 6) launch_sga_v<6>(x, guidance, out, B, C, D, H, W, dir, first, s);
            else if (dpt <= 8) launch_sga_v<8>(x, guidance, out, B, C, D, H, W, dir, first, s);
            else if (dpt <= 12) launch_sga_v<12>(x, guidance, out, B, C, D, H, W, dir, first, s);
            else if (dpt <= 16) launch_sga_v<16>(x, guidance, out, B, C, D, H, W, dir, first, s);
            else if (dpt <= 24) launch_sga_v<24>(x, guidance, out, B, C, D, H, W, dir, first, s);
            else launch_sga_v<32>(x, guidance, out, B, C, D, H, W, dir, first, s);
            int rc = check_launch("sga_vertical_kernel");
            if (rc) return rc;
            continue;
        }
        // generic fallback (very large D): one CTA per scan line
        const bool horiz = dir < 2;
        int lpb = horiz ? 1 : 8;
        while (lpb > 1 && D * lpb > 1024) lpb >>= 1;
        const int nlines = horiz ? H : W;
        const int groups = (nlines + lpb - 1) / lpb;
        const int threads = ((D * lpb + 31) / 32) * 32;
        const int nwarps = threads / 32;
        const size_t smem = ((size_t)2 * (D + 2) * lpb + (size_t)nwarps * lpb) * 4;
        const unsigned grid = (unsigned)((size_t)B * C * groups);
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
        dim3 grid((unsigned)cdiv(W, 32), (unsigned)cdiv(H, 8), B);
        DMB_REQUIRE(grid.y <= 65535, "lga: image too tall");
        static int rot_mode = -1;                  // DMB_B200_LGA_ROT=0: the register-plane kernel (A/B)
        if (rot_mode < 0) {
            const char* e = getenv("DMB_B200_LGA_ROT");
            // 0: register-plane kernel, 2: cp.async-staged, 3: TMA-staged with two pixels per thread, default 1: TMA-staged
            rot_mode = (e && e[0] >= '0' && e[0] <= '3' && e[0] != '1') ? e[0] - '0' : 1;
        }
        if (rot_mode) {
            // TMA-staged variant when the tensor map can be built (16-byte aligned base and row pitch, sm_100 driver)
            if (rot_mode != 2 && W % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 && tc::device_ok()) {
                CUtensorMap xmap;
                const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
                const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)D * H * W * 4};
                const cuuint32_t box[4] = {40, 12, 4, 1};        // x0 - 4 .. x0 + 35: 16-byte aligned start
                const cuuint32_t estr[4] = {1, 1, 1, 1};
                const CUresult r = tc::encode_fn()(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box,
                                                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r == CUDA_SUCCESS) {
                    if (rot_mode == 3) {           // two pixels per thread (A/B: measured slower)
                        lga_r2_tma2_kernel<4><<<grid, 128, 0, as_stream(stream)>>>(xmap, guidance, out, D, H, W);
                        return check_launch("lga_r2_tma2_kernel");
                    }
                    lga_r2_tma_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(xmap, guidance, out, D, H, W);
                    return check_launch("lga_r2_tma_kernel");
                }
            }
            lga_r2_rot_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(x, guidance, out, D, H, W);
            return check_launch("lga_r2_rot_kernel");
        }
        lga_r2_tiled_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, guidance, out, D, H, W);
        return check_launch("lga_r2_tiled_kernel");
    }
    dim3 grid((unsigned)cdiv((size_t)H * W, 256), B);
    lga_generic_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, guidance, out, D, H, W, radius);
    return check_launch("lga_generic_kernel");
}
