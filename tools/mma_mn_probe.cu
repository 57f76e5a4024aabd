// Round-2 groundwork (NOT part of the library, never run in round 1: written after the GPU budget was spent):
// does a tcgen05.mma with MN-major operands read the trunk's blocked layout the way DESIGN.md section 8 item 2
// assumes?  The planned tensor-core weight-gradient kernel contracts over VOXELS (K) with channels as M / N, i.e.
// both operands are "MN-major" views of [C/8][voxel][8 channels] shared-memory tiles:
//     element (channel m, voxel k)  at  (m / 8) * GROUP_PITCH + k * 16 B + (m % 8) * 2 B.
// CUTLASS documents the canonical no-swizzle MN-major layout (cute/atom/mma_traits_sm100.hpp, in 16-byte units) as
//     ((1,n),(8,k)) : ((X,SBO),(1,LBO))  -- 8 consecutive K at a 16-byte pitch, groups of 8 K at LBO, groups of 8 MN
// elements at SBO -- which would make LBO = 128 B and SBO = GROUP_PITCH here.  This probe fills A (M x 16) and
// B (N x 16) with small integers (exact in fp16), issues ONE MMA per candidate (descriptor strides swapped or not,
// major bits set or not, M = 128 and M = 64), reads the accumulator back and reports which candidate reproduces
// D[m][n] = sum_k A[m][k] B[n][k] computed on the host, and which TMEM lane holds row m for M = 64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/mma_mn_probe tools/mma_mn_probe.cu && tools/_build/mma_mn_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// D f32, A = B = f16; bit 15 / 16: A / B is MN-major ("transposed")
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

constexpr int N = 32, K = 16, GROUP_PITCH = 256;      // one K=16 slab: 16 voxels x 16 bytes per 8-channel group

struct Cand {
    int M, a_mn, b_mn;
    uint32_t lbo, sbo;
};

__global__ void __launch_bounds__(128, 1) probe(Cand c, const __half* __restrict__ a_img, const __half* __restrict__ b_img,
                                                float* __restrict__ d_out) {
    __shared__ __align__(1024) unsigned char a_s[16 * GROUP_PITCH];       // up to 128 channels
    __shared__ __align__(1024) unsigned char b_s[4 * GROUP_PITCH];        // 32 channels
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 16 * GROUP_PITCH / 2; i += blockDim.x) reinterpret_cast<__half*>(a_s)[i] = a_img[i];
    for (int i = threadIdx.x; i < 4 * GROUP_PITCH / 2; i += blockDim.x) reinterpret_cast<__half*>(b_s)[i] = b_img[i];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint64_t ad = make_desc(smem_u32(a_s), c.lbo, c.sbo), bd = make_desc(smem_u32(b_s), c.lbo, c.sbo);
        const uint32_t idesc = make_idesc(c.M, N, c.a_mn, c.b_mn);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                     "l"(ad), "l"(bd), "r"(idesc), "r"(0u)
                     : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(smem_u32(&bar))
                     : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // every warp dumps the 32 columns of its 32 TMEM lanes: d_out[lane_global][n]
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int n = 0; n < 32; ++n) d_out[(warp * 32 + lane) * 32 + n] = __uint_as_float(r[n]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
    }
}

static float a_val(int m, int k) { return (float)((m * 5 + k * 3) % 7 - 3); }
static float b_val(int n, int k) { return (float)((n * 3 + k * 7) % 5 - 2); }

int main() {
    // blocked images: element (channel c, voxel k) at (c / 8) * GROUP_PITCH + k * 16 + (c % 8) * 2 bytes
    static __half a_img[16 * GROUP_PITCH / 2], b_img[4 * GROUP_PITCH / 2];
    for (int m = 0; m < 128; ++m)
        for (int k = 0; k < K; ++k) a_img[((m / 8) * GROUP_PITCH + k * 16) / 2 + m % 8] = __float2half(a_val(m, k));
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) b_img[((n / 8) * GROUP_PITCH + k * 16) / 2 + n % 8] = __float2half(b_val(n, k));
    __half *d_a, *d_b;
    float* d_d;
    cudaMalloc(&d_a, sizeof(a_img));
    cudaMalloc(&d_b, sizeof(b_img));
    cudaMalloc(&d_d, 128 * 32 * sizeof(float));
    cudaMemcpy(d_a, a_img, sizeof(a_img), cudaMemcpyHostToDevice);
    cudaMemcpy(d_b, b_img, sizeof(b_img), cudaMemcpyHostToDevice);
    const Cand cands[] = {
        {128, 1, 1, 128, GROUP_PITCH}, {128, 1, 1, GROUP_PITCH, 128},      // MN-major bits, the two stride assignments
        {128, 0, 0, 128, GROUP_PITCH}, {128, 0, 0, GROUP_PITCH, 128},      // control: K-major bits must NOT match
        {64, 1, 1, 128, GROUP_PITCH},  {64, 1, 1, GROUP_PITCH, 128},
    };
    static float h[128 * 32];
    for (const Cand& c : cands) {
        cudaMemset(d_d, 0, sizeof(h));
        probe<<<1, 128>>>(c, d_a, d_b, d_d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("M=%d a_mn=%d b_mn=%d LBO=%u SBO=%u: %s\n", c.M, c.a_mn, c.b_mn, c.lbo, c.sbo, cudaGetErrorString(e));
            return 1;
        }
        cudaMemcpy(h, d_d, sizeof(h), cudaMemcpyDeviceToHost);
        // M = 128: row m in TMEM lane m.  M = 64: try lane = m (dense) and lane = (m % 16) + 32 * (m / 16)
        for (int mapping = 0; mapping < (c.M == 64 ? 2 : 1); ++mapping) {
            double worst = 0;
            for (int m = 0; m < c.M; ++m) {
                const int lane = (c.M == 64 && mapping == 1) ? (m % 16) + 32 * (m / 16) : m;
                for (int n = 0; n < N; ++n) {
                    double want = 0;
                    for (int k = 0; k < K; ++k) want += (double)a_val(m, k) * b_val(n, k);
                    const double err = fabs(want - h[lane * 32 + n]);
                    if (err > worst) worst = err;
                }
            }
            printf("M=%3d majors(A,B)=(%s,%s) LBO=%3u SBO=%3u lanes=%-22s max |D - A.B^T| = %g %s\n", c.M, c.a_mn ? "MN" : "K",
                   c.b_mn ? "MN" : "K", c.lbo, c.sbo, c.M == 64 ? (mapping ? "(m%16)+32*(m/16)" : "m") : "m", worst,
                   worst == 0 ? "<== matches" : "");
        }
    }
    return 0;
}
