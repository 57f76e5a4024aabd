// Weight gradient of the trunk's 3x3x3 / stride 1 / pad 1 convolutions on tcgen05 (sm_100a).
//
// Replaces cuDNN's wgrad behind the backward of conv3d_bn[_relu] (dmb/modeling/stereo/layers/basic_layers.py:68-177,
// autograd through nn.Conv3d) in the config-5 training step; oracle: torch autograd of oracle/dmb_oracle.py:conv_unit.
//
//     dw[tap][ci][co] = sum over (b, d, h, w) of  x[b][ci][d+kd-1][h+kh-1][w+kw-1] * g[b][co][d][h][w]
//
// The contraction runs over VOXELS (K of the MMA = 16 consecutive voxels of one row), the channels are the M / N
// dimensions: both operands are MN-major views of the blocked layout [C/8][voxel][8 channels] -- element
// (channel m, voxel k) at (m / 8) * SBO + (k / 8) * 128 B + (k % 8) * 16 B + (m % 8) * 2 B, i.e. LBO = 128 and
// SBO = the pitch between 8-channel groups (verified on the device by tools/mma_mn_probe.cu; a tap is a byte offset).
// Split arithmetic: A = [x_hi ; x_lo] (channels of the hi plane, then of the lo plane), B = [g_hi | g_lo], so one MMA
// yields x_hi.g_hi, x_hi.g_lo, x_lo.g_hi (and the negligible x_lo.g_lo); the epilogue adds the four blocks.
//
// TMEM is what shapes the kernel: 27 taps x (64 x 64) fp32 accumulators = 442 KB do not fit one SM's 256 KB, and a
// tcgen05.mma costs >= 45 cycles however small it is (tools/mma_probe.cu).  So
//   * the three depth taps kd go to three CTA POPULATIONS (blockIdx.y): population kd streams x plane d + kd - 1
//     against g plane d -- one x plane per g plane, no plane ring, and the populations walk the same tiles in the
//     same order so that their loads hit in L2;
//   * inside a population the x tile is laid out [row][hi|lo][channel block][voxel], which makes the 16 channel
//     groups of two ADJACENT rows uniform-strided: one M = 128 MMA covers the taps kh = 0 and kh = 1 (x rows r, r + 1
//     against g row r), one M = 64 MMA the tap kh = 2; N = 64.  Per kw that is a 128-lane and a 64-row accumulator of
//     64 columns: 3 x 128 = 384 of the 512 TMEM columns, resident for the whole kernel (no epilogue overlap needed);
//   * per g row, 16-voxel slab and kw: 2 MMAs of 48 cycles -- 288 cycles per slab and population against ~2400 for one
//     M = 64, N = 32 MMA pair per tap.
// Every CTA adds its partial sums to dw with fp32 atomics at the end (9 taps x 32 x 32 values).
//
// Tiles are TH x TW = 4 x 64 voxels of one depth plane; halos, ragged edges and the planes d = -1 / D are zero-filled
// by the TMA (tensor maps over 8-byte elements: a box row of 66 voxels exceeds the 256-element box limit at 2 bytes).
//
// STRIDE 2 (Hourglass conv1 / conv3, and conv5 / conv6 with the roles of input and output swapped):
//     dw[tap][ca][cg] = sum over the LOW-resolution grid of  a[2d+kd-1][2h+kh-1][2w+kw-1] * g[d][h][w]
// K = 16 consecutive low-resolution voxels pair with every second voxel of `a`, which a 16-byte K pitch cannot
// express, so `a` arrives W-PARITY-SPLIT ([..][H][even | odd][W/2][8], written by the layout conversion that the
// training path runs anyway): kw = 1 reads the even half, kw = 0 / 2 the odd half at offsets -1 / 0.  Rows 2h-1, 2h of
// the tile are adjacent (taps kh = 0, 1: the M = 128 pair), row 2h+1 is the single; 4 x 32 tiles.
#include "common.cuh"
#include "tc_common.cuh"

namespace dmb {
namespace wg {

using namespace dmb::tc;

constexpr int CB = 4;                                // 8-channel blocks per pass (32 channels)
constexpr int NSTAGE = 2;
constexpr int ACC_COLS = 64;                         // [. g_hi | . g_lo]

template <int STRIDE>
struct Geo {
    static constexpr int TH = 4, TW = STRIDE == 1 ? 64 : 32;          // g tile: rows x voxels of one depth plane
    static constexpr int NPAR = STRIDE;                               // w-parity halves of the x tile
    static constexpr int XR = STRIDE == 1 ? TH + 2 : 2 * TH + 1;      // x tile rows incl. halo
    static constexpr int XW = TW + 2;                                 // x tile voxels per row (and parity) incl. halo
    static constexpr uint32_t PW = XW * 16;                           // x: pitch between channel groups: 1056 | 544
    static constexpr uint32_t XROW = 2 * CB * PW;                     // x: one tile row = [hi cb0..3 | lo cb0..3]
    static constexpr uint32_t XPAR = XR * XROW;                       // x: one parity half
    static constexpr uint32_t X_BYTES = NPAR * XPAR;                  // 50688 | 78336
    static constexpr uint32_t PG = TH * TW * 16;                      // g: pitch between channel groups
    static constexpr uint32_t G_BYTES = 2 * CB * PG;                  // 32768 | 16384
    static constexpr uint32_t STAGE_BYTES = X_BYTES + G_BYTES;
    static constexpr uint32_t BAR_OFF = NSTAGE * STAGE_BYTES;
    static constexpr uint32_t SMEM_TOTAL = BAR_OFF + 128;
    static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget exceeded");
    static_assert(XROW % 128 == 0 && (CB * PW) % 128 == 0 && X_BYTES % 128 == 0 && (CB * PG) % 128 == 0,
                  "TMA destinations must be 128-byte aligned");
};

struct Maps {
    CUtensorMap x_hi, x_lo, g_hi, g_lo;
};

struct Params {
    float* dw;                 // [27][Ca][Cg], accumulated with atomics
    int B, D, H, W;
    int tiles_h, tiles_w, n_items;
    int a_cb0, g_cb0;          // first 8-channel block of this pass in x / g
    int Ca, Cg;                // channel counts of dw
    int ca0, cg0;              // channel offsets of this pass in dw
};

// D f32, A = B = 16-bit (fmt 0 = f16, 1 = bf16), both MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t idesc_mn(int m, int n, uint32_t fmt) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void red_add(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

template <bool FP16, int STRIDE>
__global__ void __launch_bounds__(128, 1) wgrad_tc_kernel(const __grid_constant__ Maps maps, const Params p) {
    using G = Geo<STRIDE>;
    constexpr int TH = G::TH, TW = G::TW, XR = G::XR;
    constexpr uint32_t PW = G::PW, XROW = G::XROW, X_BYTES = G::X_BYTES, PG = G::PG, STAGE_BYTES = G::STAGE_BYTES,
                       BAR_OFF = G::BAR_OFF;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + BAR_OFF);     // [NSTAGE]
    uint64_t* empty = full + NSTAGE;                                   // [NSTAGE]
    uint64_t* done = empty + NSTAGE;                                   // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kd = blockIdx.y;                                         // this population's depth tap

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(done, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int my_items = blockIdx.x < p.n_items ? (p.n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            uint32_t n = 0;
            const int per_plane = p.tiles_h * p.tiles_w;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n) {
                const int pl = item / per_plane;                       // b * D + d
                const int r = item - pl * per_plane;
                const int b = pl / p.D, d = pl - b * p.D;
                const int h0 = (r / p.tiles_w) * TH, w0 = (r % p.tiles_w) * TW;
                const uint32_t slot = n % NSTAGE;
                mbar_wait(&empty[slot], ((n / NSTAGE) & 1) ^ 1);
                unsigned char* xs = smem + slot * STAGE_BYTES;
                unsigned char* gs = xs + X_BYTES;
                mbar_expect_tx(&full[slot], STAGE_BYTES);
                // coordinates are in 8-byte elements along w (2 per voxel); out-of-volume parts are zero-filled
                if (STRIDE == 1) {
#pragma unroll 1
                    for (int row = 0; row < XR; ++row) {
                        tma_load_5d(xs + row * XROW, &maps.x_hi, &full[slot], 2 * (w0 - 1), h0 - 1 + row, d + kd - 1, p.a_cb0, b);
                        tma_load_5d(xs + row * XROW + CB * PW, &maps.x_lo, &full[slot], 2 * (w0 - 1), h0 - 1 + row, d + kd - 1, p.a_cb0, b);
                    }
                } else {
                    // W-parity-split input: map row index = 2 * (input row) + parity; tile row `row` = input row 2 h0 - 1 + row,
                    // tile voxel j of either half = half-resolution column w0 - 1 + j
#pragma unroll 1
                    for (int pr = 0; pr < 2 * XR; ++pr) {
                        const int par = pr / XR, row = pr - par * XR;
                        unsigned char* dst = xs + par * G::XPAR + row * XROW;
                        const int hrow = 2 * (2 * h0 - 1 + row) + par;
                        tma_load_5d(dst, &maps.x_hi, &full[slot], 2 * (w0 - 1), hrow, 2 * d + kd - 1, p.a_cb0, b);
                        tma_load_5d(dst + CB * PW, &maps.x_lo, &full[slot], 2 * (w0 - 1), hrow, 2 * d + kd - 1, p.a_cb0, b);
                    }
                }
                tma_load_5d(gs, &maps.g_hi, &full[slot], 2 * w0, h0, d, p.g_cb0, b);
                tma_load_5d(gs + CB * PG, &maps.g_lo, &full[slot], 2 * w0, h0, d, p.g_cb0, b);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        constexpr uint32_t fmt = FP16 ? 0u : 1u;
        constexpr uint32_t idesc_pair = idesc_mn(128, ACC_COLS, fmt);
        constexpr uint32_t idesc_single = idesc_mn(64, ACC_COLS, fmt);
        constexpr uint32_t a_hiw = desc_hi(PW);                        // SBO = pitch between channel groups
        constexpr uint32_t b_hiw = desc_hi(PG);
        const uint32_t smem_addr = smem_u32(smem);
        for (int n = 0; n < my_items; ++n) {
            const uint32_t slot = n % NSTAGE;
            mbar_wait(&full[slot], (n / NSTAGE) & 1);
            tcgen05_fence_after();
            if (elect_one()) {
                const uint32_t xa = smem_addr + slot * STAGE_BYTES, ga = xa + X_BYTES;
                const uint32_t a_lo0 = desc_lo(xa, 128), b_lo0 = desc_lo(ga, 128);   // LBO = 128: next 8 voxels
                const uint32_t acc_flag = n > 0 ? 1u : 0u;
#pragma unroll 1
                for (int r = 0; r < TH; ++r) {
#pragma unroll
                    for (int s = 0; s < TW / 16; ++s) {
                        const uint32_t b_lo = b_lo0 + (((r * TW + s * 16) * 16) >> 4);
                        const uint32_t accumulate = (acc_flag | (uint32_t)(r | s)) ? 1u : 0u;
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            // stride 1: x rows r, r+1, r+2 at voxel offset kw; stride 2: rows 2r, 2r+1, 2r+2 of the even (kw = 1,
                            // offset 1) or odd (kw = 0 / 2, offset 0 / 1) half
                            const uint32_t a_off = STRIDE == 1 ? r * XROW + (s * 16 + kw) * 16
                                                               : (kw == 1 ? 0u : G::XPAR) + 2 * r * XROW + (s * 16 + (kw == 0 ? 0 : 1)) * 16;
                            // x rows r, r+1 (taps kh = 0, 1) -> 128-lane accumulator; x row r+2 (kh = 2) -> 64-row accumulator
                            mma_f16_ss_rt(tmem_base + kw * 2 * ACC_COLS, a_lo0 + (a_off >> 4), a_hiw, b_lo, b_hiw, idesc_pair, accumulate);
                            mma_f16_ss_rt(tmem_base + (kw * 2 + 1) * ACC_COLS, a_lo0 + ((a_off + 2 * XROW) >> 4), a_hiw, b_lo, b_hiw,
                                          idesc_single, accumulate);
                        }
                    }
                }
                commit_one(&empty[slot]);
                if (n == my_items - 1) commit_one(done);
            }
            __syncwarp();
        }
    }

    // ================================ epilogue (all four warps) =========================
    if (my_items > 0) {
        mbar_wait(done, 0);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int kw = 0; kw < 3; ++kw) {
            uint32_t r0[32], r1[32];
            // pair accumulator: TMEM lane m = accumulator row m: kh = m / 64, plane (hi | lo) = (m / 32) % 2, channel = m % 32
            tmem_ld32(taddr + kw * 2 * ACC_COLS, r0);
            tmem_ld32(taddr + kw * 2 * ACC_COLS + 32, r1);
            tmem_ld_wait();
            {
                const int kh = warp >> 1;
                const int tap = (kd * 3 + kh) * 3 + kw;
                float* dst = p.dw + ((size_t)tap * p.Ca + p.ca0 + lane) * p.Cg + p.cg0;
#pragma unroll
                for (int co = 0; co < 32; ++co) red_add(dst + co, __uint_as_float(r0[co]) + __uint_as_float(r1[co]));
            }
            // single accumulator (M = 64): row m sits in TMEM lane (m % 16) + 32 * (m / 16)
            tmem_ld32(taddr + (kw * 2 + 1) * ACC_COLS, r0);
            tmem_ld32(taddr + (kw * 2 + 1) * ACC_COLS + 32, r1);
            tmem_ld_wait();
            if (lane < 16) {
                const int m = warp * 16 + lane;
                const int tap = (kd * 3 + 2) * 3 + kw;
                float* dst = p.dw + ((size_t)tap * p.Ca + p.ca0 + (m & 31)) * p.Cg + p.cg0;
#pragma unroll
                for (int co = 0; co < 32; ++co) red_add(dst + co, __uint_as_float(r0[co]) + __uint_as_float(r1[co]));
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// map over [B][CBS][D][H][W][8 x 16-bit] seen as 8-byte elements: dims (2W, H, D, CBS, B).  wsplit: the tensor is
// W-parity-split ([..][H][2][W/2][8]) and seen as 2H rows of W/2 voxels.
static int make_map(CUtensorMap* map, const void* base, int B, int CBS, int D, int H, int W, int box_w, int box_h,
                    bool wsplit = false) {
    const cuuint64_t rows = wsplit ? 2 * (cuuint64_t)H : (cuuint64_t)H, cols = wsplit ? (cuuint64_t)W / 2 : (cuuint64_t)W;
    const cuuint64_t dims[5] = {cols * 2, rows, (cuuint64_t)D, (cuuint64_t)CBS, (cuuint64_t)B};
    const cuuint64_t strides[4] = {cols * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16,
                                   (cuuint64_t)CBS * D * H * W * 16};
    const cuuint32_t box[5] = {(cuuint32_t)box_w * 2, (cuuint32_t)box_h, 1, CB, 1};
    return encode_typed(map, base, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_INT64);
}

template <bool FP16, int STRIDE>
static int launch(const Maps& maps, Params& p, int Ca, int Cg, void* stream) {
    using G = Geo<STRIDE>;
    p.tiles_h = (int)cdiv(p.H, G::TH);
    p.tiles_w = (int)cdiv(p.W, G::TW);
    p.n_items = p.B * p.D * p.tiles_h * p.tiles_w;
    // three populations (depth taps) of persistent CTAs, one per SM
    int per_pop = sm_count() / 3;
    if (per_pop < 1) per_pop = 1;
    if (per_pop > p.n_items) per_pop = p.n_items;
    const dim3 grid(per_pop, 3);
    DMB_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<FP16, STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_TOTAL));
    for (int ia = 0; ia < Ca / 32; ++ia) {
        for (int ig = 0; ig < Cg / 32; ++ig) {
            p.a_cb0 = ia * CB; p.g_cb0 = ig * CB;
            p.ca0 = ia * 32; p.cg0 = ig * 32;
            wgrad_tc_kernel<FP16, STRIDE><<<grid, 128, G::SMEM_TOTAL, as_stream(stream)>>>(maps, p);
            const int rc = check_launch("wgrad_tc_kernel");
            if (rc) return rc;
        }
    }
    return DMB_OK;
}

}  // namespace wg
}  // namespace dmb

using namespace dmb;
using namespace dmb::wg;

extern "C" int dmb_b200_conv3d_wgrad_tc(const void* a_hi, const void* a_lo, const void* g_hi, const void* g_lo, float* dw,
                                        int B, int Ca, int Cg, int D, int H, int W, int stride, int fp16, void* stream) {
    DMB_REQUIRE(a_hi && a_lo && g_hi && g_lo && dw, "conv3d_wgrad_tc: null pointer (split hi / lo planes are required)");
    DMB_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "conv3d_wgrad_tc: non-positive dimension");
    DMB_REQUIRE(stride == 1 || stride == 2, "conv3d_wgrad_tc: stride %d not supported (1 or 2)", stride);
    DMB_REQUIRE(Ca > 0 && Cg > 0 && Ca % 32 == 0 && Cg % 32 == 0, "conv3d_wgrad_tc: channel counts (%d, %d) must be multiples of 32", Ca, Cg);
    if (!tc::device_ok()) return fail(DMB_ERR_UNSUPPORTED, "conv3d_wgrad_tc: needs an sm_100 device and a TMA-capable driver");
    Maps maps;
    int rc;
    // D, H, W: extents of g (the conv output for a strided conv); `a` has stride times those
    const int Da = stride * D, Ha = stride * H, Wa = stride * W;
    const int xw = stride == 1 ? Geo<1>::XW : Geo<2>::XW, tw = stride == 1 ? Geo<1>::TW : Geo<2>::TW;
    const int th = stride == 1 ? Geo<1>::TH : Geo<2>::TH;
    if ((rc = make_map(&maps.x_hi, a_hi, B, Ca / 8, Da, Ha, Wa, xw, 1, stride == 2))) return rc;
    if ((rc = make_map(&maps.x_lo, a_lo, B, Ca / 8, Da, Ha, Wa, xw, 1, stride == 2))) return rc;
    if ((rc = make_map(&maps.g_hi, g_hi, B, Cg / 8, D, H, W, tw, th))) return rc;
    if ((rc = make_map(&maps.g_lo, g_lo, B, Cg / 8, D, H, W, tw, th))) return rc;
    Params p;
    p.dw = dw;
    p.B = B; p.D = D; p.H = H; p.W = W;
    p.Ca = Ca; p.Cg = Cg;
    if (stride == 1) return fp16 ? launch<true, 1>(maps, p, Ca, Cg, stream) : launch<false, 1>(maps, p, Ca, Cg, stream);
    return fp16 ? launch<true, 2>(maps, p, Ca, Cg, stream) : launch<false, 2>(maps, p, Ca, Cg, stream);
}
