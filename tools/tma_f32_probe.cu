// Probe: which fp32 tiled tensor maps / coordinates the TMA unit accepts (the LGA halo tile: box {36,12,4,1} at x0-2).
// usage: tma_f32_probe <rank 3|4> <box0> <c0> <W>      -- one configuration per process (a fault kills the context)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap m, float* out, int n, int c0, int c1, int c2, int c3) {
    __shared__ __align__(128) float tile[8192];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(n * 4) : "memory");
        if (RANK == 4)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(s32(tile)),
                         "l"(&m), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s32(tile)),
                         "l"(&m), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(s32(&bar)) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char** argv) {
    const int rank = argc > 1 ? atoi(argv[1]) : 4, box0 = argc > 2 ? atoi(argv[2]) : 36, c0 = argc > 3 ? atoi(argv[3]) : -2;
    const int W = argc > 4 ? atoi(argv[4]) : 40, H = 12, D = 9, B = 1;
    const size_t n = (size_t)B * D * H * W;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (float)i;
    float *x, *out;
    cudaMalloc(&x, n * 4); cudaMalloc(&out, 8192 * 4);
    cudaMemcpy(x, h.data(), n * 4, cudaMemcpyHostToDevice);
    EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
    CUtensorMap m;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(rank == 4 ? D : D * B), (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)D * H * W * 4};
    const cuuint32_t box[4] = {(cuuint32_t)box0, 12, 4, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("rank %d box0 %d c0 %d W %d: encode %d", rank, box0, c0, W, (int)r);
    if (r) { printf("\n"); return 1; }
    const int cnt = box0 * 12 * 4;
    if (rank == 4) probe<4><<<1, 128>>>(m, out, cnt, c0, -2, 4, 0); else probe<3><<<1, 128>>>(m, out, cnt, c0, -2, 4, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  run: %s", cudaGetErrorString(e));
    if (!e) {
        std::vector<float> o(cnt); cudaMemcpy(o.data(), out, cnt * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int pl = 0; pl < 4; ++pl) for (int y = 0; y < 12; ++y) for (int xx = 0; xx < box0; ++xx) {
            const int gx = c0 + xx, gy = y - 2, gd = 4 + pl;
            const float want = (gx < 0 || gx >= W || gy < 0 || gy >= H || gd >= D) ? 0.f : (float)((gd * H + gy) * W + gx);
            bad += o[(pl * 12 + y) * box0 + xx] != want;
        }
        printf("  mismatches %d", bad);
    }
    printf("\n");
    return 0;
}
