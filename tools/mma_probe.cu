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

// The production stride-1 kernel's MMA sequence for one output plane: per (kd, kh, K-step)
// A_hi x [Whi|Wlo] (N=192) then A_lo x Whi (N=96, onto columns 96..191), operand addresses as in the kernel
// (three resident input planes of 2 x 10240 bytes, 110592 bytes of weights), accumulators alternating per plane.
__global__ void __launch_bounds__(192, 1) probe_trunk(int planes, int mode, int commits, int drain, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int done;
    for (int i = threadIdx.x; i < 212 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        done = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(smem_u32(&bar2)));
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
        const uint32_t w0 = smem_u32(smem), pl0 = smem_u32(smem + 110592);
        constexpr uint32_t TAP_BYTES = 4 * 192 * 16, LBO_B = 192 * 16, LBO_A = 2560, PLANE = 10240, STAGE = 20480;
        const uint32_t idesc_main = make_idesc(128, 192), idesc_lo = make_idesc(128, 96);
        const long long t0 = clock64();
        for (int pl = 0; pl < planes; ++pl) {
            const uint32_t acc = tmem + (uint32_t)(pl & 1) * 192;
#pragma unroll
            for (int kd = 0; kd < 3; ++kd) {
                const uint32_t a_stage = pl0 + (uint32_t)((pl + kd) % 5) * STAGE;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const uint64_t bdesc = make_desc(w0 + (kd * 3 + kh) * TAP_BYTES + 2 * kk * LBO_B, LBO_B, 128, 0);
                        const uint32_t a_off = kh * 256 + 2 * kk * LBO_A;
                        const uint64_t ah = make_desc(a_stage + a_off, LBO_A, 128, 0);
                        const uint64_t al = make_desc(a_stage + PLANE + a_off, LBO_A, 128, 0);
                        const uint32_t first = (kd | kh | kk) ? 1u : 0u;
                        if (mode != 2)
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(acc),
                                         "l"(ah), "l"(bdesc), "r"(idesc_main), "r"(first));
                        if (mode != 1)
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(acc + 96),
                                         "l"(al), "l"(bdesc), "r"(idesc_lo), "r"(mode == 2 ? first : 1u));
                    }
                }
            }
            for (int cc = 0; cc < commits; ++cc)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
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
        out[blockIdx.x] = clock64() - t0;
        done = 1;
    } else if (warp >= 2 && drain) {
        // epilogue-like TMEM drain running against the MMAs: each of 4 warps reads `drain` x 32 columns of its
        // lane quarter per iteration, results discarded
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t sink = 0;
        while (!done) {
            for (int j = 0; j < drain; ++j) {
                uint32_t r[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr + 384 + (uint32_t)(j % 4) * 32)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                sink ^= r[0] ^ r[31];
            }
        }
        if (sink == 0x12345678u) out[147] = sink;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

int main() {
    {
        long long* d_o;
        cudaMalloc(&d_o, 148 * sizeof(long long));
        const size_t sm = 212 * 1024;
        cudaFuncSetAttribute(probe_trunk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        const char* names[3] = {"N=192 + N=96 (production pair)", "N=192 only", "N=96 only"};
        for (int grid : {148})
            for (int mode = 0; mode < 3; ++mode)
                for (int commits : {0, 2})
                    for (int drain : {0, 6}) {
                        if (mode != 0 && (commits || drain)) continue;
                        probe_trunk<<<grid, 192, sm>>>(100, mode, commits, drain, d_o);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("probe_trunk: %s\n", cudaGetErrorString(e)); return 1; }
                        long long h[148];
                        cudaMemcpy(h, d_o, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                        printf("trunk sequence, grid %3d, %-30s commits/plane %d, concurrent TMEM drain %d: %6.1f cycles per (kd,kh,K-step) = %5.0f per plane\n",
                               grid, names[mode], commits, drain, (double)h[grid / 2] / (100.0 * 18), (double)h[grid / 2] / 100.0);
                    }
        cudaFree(d_o);
    }

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
