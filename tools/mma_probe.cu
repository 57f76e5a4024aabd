// Micro-benchmark: issue cost of one tcgen05.mma (kind::f16, SS operands, K = 16) as a function of N, M and
// the shared-memory layout of the A operand.  Motivation: every kernel of the conv trunk
// runs at ~110 cycles per MMA whatever its N (32..192), i.e. the A-tile fetch, not the tensor pipe, paces it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/mma_probe tools/mma_probe.cu && tools/_build/mma_probe
// Results are timing only (operands are zeros).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
           ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);      // D f32, A = B = f16, K-major
}

struct Cfg {
    int n_mma, M, N, layout;      // layout: 0 none, 2 128B, 4 64B, 6 32B swizzle (A operand only)
    int nacc;                     // independent TMEM accumulators the MMAs rotate over
    uint32_t a_lbo, a_sbo, a_step;
};

__global__ void __launch_bounds__(64, 1) probe(Cfg c, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
        const uint32_t idesc = make_idesc(c.M, c.N);
        const uint64_t bdesc = make_desc(b0, (uint32_t)c.N * 16, 128, 0);
        // descriptors precomputed, 16 MMAs per loop trip: the loop itself must not be what is measured
        uint64_t ad[8];
        for (int j = 0; j < 8; ++j) ad[j] = make_desc(a0 + (uint32_t)j * c.a_step, c.a_lbo, c.a_sbo, (uint32_t)c.layout);
        uint32_t accs[8];
        for (int j = 0; j < 8; ++j) accs[j] = tmem + (uint32_t)(j % c.nacc) * (512 / c.nacc);
        const long long t0 = clock64();
        for (int i = 0; i < c.n_mma; i += 16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(accs[j & 7]),
                    "l"(ad[j & 7]), "l"(bdesc), "r"(idesc), "r"((i + j) >= 8 ? 1u : 0u));
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(smem_u32(&bar))
                : "memory");
        }
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 148 * sizeof(long long));
    const size_t smem = 160 * 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    struct Lay { const char* name; int layout; uint32_t lbo, sbo, step; };
    // none : the trunk's layout (8x16-byte core matrices, row-group pitch 128 B, K-chunk pitch 2560 B, taps = +256 B)
    // 128B/64B/32B swizzle: canonical K-major atoms (8 rows x 128/64/32 B), SBO = 8 rows
    const Lay lays[] = {{"none (trunk)", 0, 2560, 128, 256}, {"none (dense K)", 0, 2048, 128, 4096},
                        {"swizzle 128B", 2, 16, 1024, 16384}, {"swizzle 64B", 4, 16, 512, 8192}, {"swizzle 32B", 6, 16, 256, 4096}};
    const int Ns[] = {16, 32, 64, 96, 128, 192, 256};
    printf("cycles per tcgen05.mma (kind::f16, K=16, SS), 2048 MMAs issued by one thread, median over CTAs\n");
    for (int grid : {1, 148}) {
        for (int nacc : {1, 2, 4, 8}) {
            for (const Lay& l : lays) {
                if (nacc != 2 && l.layout != 0) continue;
                if (grid == 148 && (nacc == 8 || l.layout == 6)) continue;
                printf("grid %3d accumulators %d A-layout %-15s:", grid, nacc, l.name);
                for (int N : Ns) {
                    if (N * nacc > 512) { printf("  N=%-3d    -  ", N); continue; }
                    Cfg c{2048, 128, N, l.layout, nacc, l.lbo, l.sbo, l.step};
                    probe<<<grid, 64, smem>>>(c, d_out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) {
                        printf("  N=%d: %s\n", N, cudaGetErrorString(e));
                        return 1;
                    }
                    long long h[148];
                    cudaMemcpy(h, d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                    printf("  N=%-3d %6.1f", N, (double)h[grid / 2] / 2048.0);
                }
                printf("\n");
            }
        }
    }
    return 0;
}
