// Round-2 groundwork (NOT part of the library; written after round 1's GPU budget was spent, never run yet):
// a deliberately simple, fully serialised prototype of the tensor-core WEIGHT GRADIENT of a 3x3x3 / stride 1 / pad 1
// convolution on the trunk's blocked 16-bit (hi, lo) layout, self-checked against a float64 host reference:
//     dw[tap][ci][co] = sum_voxels x[ci][v + tap - 1] * g[co][v]            (g = gradient w.r.t. the conv output)
// The contraction runs over VOXELS, so both operands are MN-major views of [C/8][voxel][8 channels] tiles
// (tools/mma_mn_probe.cu checks that reading first).  Per 16-voxel slab of one output row and per tap:
//     A = [x_hi ; x_lo] of the tap-shifted input row   (M = 64: eight 8-channel groups at a uniform pitch)
//     B = g_hi, then g_lo                               (N = 32), both into the tap's accumulator (32 TMEM columns)
// => D[0:32] = x_hi.(g_hi + g_lo),  D[32:64] = x_lo.(g_hi + g_lo);  dw = D[0:32] + D[32:64]  (fp32-grade split product).
// 27 taps x 32 columns exceed TMEM, so taps go in 4 groups of <= 8, each group streaming all slabs again (the real kernel
// will use two CTA populations and M = 128 tap pairs).  No TMA, no pipelining: every slab is
// staged by plain loads, one elected thread issues, everybody waits -- this file is about getting the descriptors,
// the accumulate flags and the TMEM row mapping right, not about speed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/wgrad_tc_proto tools/wgrad_tc_proto.cu
//   tools/_build/wgrad_tc_proto [swap_lbo_sbo=0|1] [m64_lane_map=1|0]
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {   // D f32, A = B = f16, A and B MN-major (bits 15, 16)
    return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

constexpr int C = 32, CB = C / 8, KV = 16;                 // channels, 8-channel blocks, voxels per slab (= MMA K)
constexpr int XV = KV + 2;                                 // staged input voxels per row (one halo voxel each side)
constexpr int X_GROUP = XV * 16, G_GROUP = KV * 16;        // bytes of one 8-channel group: 288, 256
constexpr int X_PLANE = 2 * CB * X_GROUP;                  // [hi|lo][cb] of one (kd, kh) input row: 2304 bytes
constexpr int XS_BYTES = 9 * X_PLANE, GS_BYTES = 2 * CB * G_GROUP;

struct Dims {
    int D, H, W;
};

// x, g: blocked [cb][D][H][W][8] halfs (hi plane, lo plane); dw: [27][C][C] float, zeroed by the host
__global__ void __launch_bounds__(128, 1) wgrad_proto(const __half* __restrict__ x_hi, const __half* __restrict__ x_lo,
                                                      const __half* __restrict__ g_hi, const __half* __restrict__ g_lo,
                                                      float* __restrict__ dw, Dims dm, int swap, int lane_map) {
    __shared__ __align__(1024) unsigned char xs[XS_BYTES];
    __shared__ __align__(1024) unsigned char gs[GS_BYTES];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const size_t plane = (size_t)dm.H * dm.W, vol = (size_t)dm.D * plane;
    const int wchunks = dm.W / KV;
    const int slabs = dm.D * dm.H * wchunks;
    const uint32_t lbo = swap ? X_GROUP : 128u, sbo_x = swap ? 128u : X_GROUP;
    const uint32_t lbo_g = swap ? G_GROUP : 128u, sbo_g = swap ? 128u : G_GROUP;
    uint32_t phase = 0;

    for (int tg = 0; tg < 27; tg += 8) {                   // tap group [tg, tg + 8)
        const int ntap = min(8, 27 - tg);
        int nissued = 0;
        for (int s = blockIdx.x; s < slabs; s += gridDim.x) {
            const int wc = s % wchunks, h = (s / wchunks) % dm.H, d = s / (wchunks * dm.H);
            const int w0 = wc * KV;
            // ---- stage the 9 input rows (zero padded) and the gradient row of this slab
            for (int i = threadIdx.x; i < 9 * 2 * CB * XV; i += blockDim.x) {
                const int v = i % XV, cb = (i / XV) % CB, hl = (i / (XV * CB)) % 2, r = i / (XV * CB * 2);
                const int dd = d + r / 3 - 1, hh = h + r % 3 - 1, ww = w0 + v - 1;
                uint4 val = make_uint4(0, 0, 0, 0);
                if (dd >= 0 && dd < dm.D && hh >= 0 && hh < dm.H && ww >= 0 && ww < dm.W) {
                    const __half* src = (hl ? x_lo : x_hi) + ((size_t)cb * vol + (size_t)dd * plane + (size_t)hh * dm.W + ww) * 8;
                    val = *reinterpret_cast<const uint4*>(src);
                }
                *reinterpret_cast<uint4*>(xs + r * X_PLANE + (hl * CB + cb) * X_GROUP + v * 16) = val;
            }
            for (int i = threadIdx.x; i < 2 * CB * KV; i += blockDim.x) {
                const int v = i % KV, cb = (i / KV) % CB, hl = i / (KV * CB);
                const __half* src = (hl ? g_lo : g_hi) + ((size_t)cb * vol + (size_t)d * plane + (size_t)h * dm.W + w0 + v) * 8;
                *reinterpret_cast<uint4*>(gs + (hl * CB + cb) * G_GROUP + v * 16) = *reinterpret_cast<const uint4*>(src);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (threadIdx.x == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t idesc = make_idesc(64, 32);
                for (int t = 0; t < ntap; ++t) {
                    const int tap = tg + t, r = tap / 3, kw = tap % 3;          // r = kd * 3 + kh
                    const uint64_t ad = make_desc(smem_u32(xs) + r * X_PLANE + kw * 16, lbo, sbo_x);
                    for (int hl = 0; hl < 2; ++hl) {
                        const uint64_t bd = make_desc(smem_u32(gs) + hl * CB * G_GROUP, lbo_g, sbo_g);
                        const uint32_t accumulate = (nissued > 0 || hl > 0) ? 1u : 0u;
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (uint32_t)t * 32),
                                     "l"(ad), "l"(bd), "r"(idesc), "r"(accumulate)
                                     : "memory");
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            ++nissued;
            uint32_t ok = 0;                                   // everybody waits: the tiles are overwritten next
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok)
                             : "r"(smem_u32(&bar)), "r"(phase)
                             : "memory");
            phase ^= 1;
            __syncthreads();
        }
        // ---- epilogue of the group: rows 0..31 = x_hi channels, 32..63 = x_lo channels; M = 64 accumulators sit in
        // TMEM lanes (m % 16) + 32 * (m / 16) (cute tmem_frg_1sm, "half subpartitions") or, lane_map == 0, in lanes m
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (nissued > 0) {
            for (int t = 0; t < ntap; ++t) {
                uint32_t r[32];
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)t * 32;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                int m = -1;                                    // accumulator row held by this TMEM lane
                if (lane_map) {
                    if (lane < 16) m = warp * 16 + lane;
                } else if (warp < 2) {
                    m = warp * 32 + lane;
                }
                if (m >= 0) {
                    const int ci = m % 32;
                    for (int co = 0; co < 32; ++co)
                        atomicAdd(dw + ((size_t)(tg + t) * C + ci) * C + co, __uint_as_float(r[co]));
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
    }
}

int main(int argc, char** argv) {
    const int swap = argc > 1 ? atoi(argv[1]) : 0, lane_map = argc > 2 ? atoi(argv[2]) : 1;
    const Dims dm{4, 6, 32};
    const size_t vol = (size_t)dm.D * dm.H * dm.W;
    std::vector<float> x(C * vol), g(C * vol);
    srand(7);
    for (auto& v : x) v = (float)(rand() % 2001 - 1000) / 997.0f;
    for (auto& v : g) v = (float)(rand() % 2001 - 1000) / 1013.0f;
    // blocked (hi, lo) images
    std::vector<__half> xh(C * vol), xl(C * vol), gh(C * vol), gl(C * vol);
    for (int c = 0; c < C; ++c)
        for (size_t v = 0; v < vol; ++v) {
            const size_t o = ((size_t)(c / 8) * vol + v) * 8 + c % 8;
            const __half hx = __float2half(x[c * vol + v]), hg = __float2half(g[c * vol + v]);
            xh[o] = hx; xl[o] = __float2half(x[c * vol + v] - __half2float(hx));
            gh[o] = hg; gl[o] = __float2half(g[c * vol + v] - __half2float(hg));
        }
    // float64 reference
    std::vector<double> want(27 * C * C, 0.0);
    for (int tap = 0; tap < 27; ++tap) {
        const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
        for (int d = 0; d < dm.D; ++d)
            for (int h = 0; h < dm.H; ++h)
                for (int w = 0; w < dm.W; ++w) {
                    const int dd = d + kd - 1, hh = h + kh - 1, ww = w + kw - 1;
                    if (dd < 0 || dd >= dm.D || hh < 0 || hh >= dm.H || ww < 0 || ww >= dm.W) continue;
                    const size_t vi = ((size_t)dd * dm.H + hh) * dm.W + ww, vo = ((size_t)d * dm.H + h) * dm.W + w;
                    for (int ci = 0; ci < C; ++ci)
                        for (int co = 0; co < C; ++co) want[((size_t)tap * C + ci) * C + co] += (double)x[ci * vol + vi] * g[co * vol + vo];
                }
    }
    __half *d_xh, *d_xl, *d_gh, *d_gl;
    float* d_dw;
    const size_t nb = C * vol * sizeof(__half);
    cudaMalloc(&d_xh, nb); cudaMalloc(&d_xl, nb); cudaMalloc(&d_gh, nb); cudaMalloc(&d_gl, nb);
    cudaMalloc(&d_dw, 27 * C * C * sizeof(float));
    cudaMemcpy(d_xh, xh.data(), nb, cudaMemcpyHostToDevice); cudaMemcpy(d_xl, xl.data(), nb, cudaMemcpyHostToDevice);
    cudaMemcpy(d_gh, gh.data(), nb, cudaMemcpyHostToDevice); cudaMemcpy(d_gl, gl.data(), nb, cudaMemcpyHostToDevice);
    cudaMemset(d_dw, 0, 27 * C * C * sizeof(float));
    wgrad_proto<<<4, 128>>>(d_xh, d_xl, d_gh, d_gl, d_dw, dm, swap, lane_map);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("kernel failed: %s\n", cudaGetErrorString(e));
        return 1;
    }
    std::vector<float> got(27 * C * C);
    cudaMemcpy(got.data(), d_dw, got.size() * sizeof(float), cudaMemcpyDeviceToHost);
    double worst = 0, scale = 0;
    for (size_t i = 0; i < got.size(); ++i) {
        worst = fmax(worst, fabs(got[i] - want[i]));
        scale = fmax(scale, fabs(want[i]));
    }
    printf("swap_lbo_sbo=%d m64_lane_map=%d: max |dw - reference| = %.3e on a scale of %.3e  %s\n", swap, lane_map, worst, scale,
           worst <= 2e-5 * scale ? "PASS (fp32-grade)" : "FAIL");
    return worst <= 2e-5 * scale ? 0 : 2;
}
