// 3x3x3 convolution of the PSMNet/AcfNet trunk as an implicit GEMM on tcgen05 (sm_100a).
//
// Replaces cuDNN behind conv3d_bn[_relu] (dmb/modeling/stereo/layers/basic_layers.py:68-177) for the
// stride-1 layers of PSMAggregator / AcfAggregator / Hourglass
// (dmb/modeling/stereo/cost_processors/aggregators/PSMNet.py:30-53, utils/hourglass.py:40-52).
// Oracle: oracle/dmb_oracle.py:conv_unit.
//
// Data layout ("blocked channels-last"): activations live in HBM as [B][C/8][D][H][W][8] bf16,
// optionally as a (hi, lo) pair with x ~= hi + lo (hi = bf16(x), lo = bf16(x - hi)), which carries
// ~16 mantissa bits through bf16 tensor-core MMAs with fp32 accumulation ("bf16x3": hi*hi + hi*lo
// + lo*hi).  In this layout the 8 channels of one voxel are 16 contiguous bytes and consecutive
// voxels along W follow at a 16-byte pitch -- exactly the UMMA "no-swizzle K-major" core-matrix
// layout (8 rows x 16 bytes), so ONE TMA box per depth plane (halo tile of (16+2)x(8+2) voxels x
// 32 channels, zero-filled outside the volume) serves all 9 in-plane taps: a tap is just a byte
// offset on the shared-memory descriptor.  Depth planes stream through a 4-deep ring (each plane
// is read from L2 ~1.4x instead of 27x).
//
// The hourglass' stride-2 convolutions (utils/hourglass.py:35-38,45-48) gather the four (h,w)
// parity classes of each input plane with four strided tensor maps; its stride-2 transposed
// convolutions (:53-60) run as 8 output-parity classes, each a 1..8-tap convolution over the input
// grid, accumulated in separate TMEM columns.
//
// One pass = 32 (stride-2: 16) input channels -> 32 output channels (the trunk's 64-channel layers
// run as several passes that accumulate in place).  Per output plane tile (16x8 voxels = M 128):
//   plain : 27 taps x 2 K-steps  MMAs  M128 N32 K16
//   split : per (tap, K-step)   A_hi x [W_hi | W_lo] (N64)  +  A_lo x W_hi (N32, own columns)
// accumulating in TMEM (one short chain per depth tap, double buffered), epilogue = sum of the
// chains + bias + residual + ReLU + 16-bit split + store.
//
// Warp roles (192 threads): warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer (warp 1 also
// owns the TMEM allocation), warps 2..5 = epilogue (TMEM lane quarter = warp_idx % 4).
#include <cuda.h>
#include <algorithm>
#include <stdlib.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace dmb {
namespace tc {

constexpr int TH = 16, TW = 8;              // M tile (in-plane) = 128 rows of the MMA
constexpr int NB = 32;                      // output channels per pass
constexpr int TAPS = 27;
// threads: warp 0 = TMA producer, warp 1 = MMA issuer, then 4 epilogue warps (8 for the transposed
// kernel, whose tiles carry 8 output parity classes = 8x the epilogue work per MMA tile)
// (8 epilogue warps everywhere but the two reference kernels: the epilogue -- TMEM drain, shifted sums, bias /
// residual / ReLU, 16-bit split, stores -- of a 4-warp kernel took longer per tile than the tile's MMAs)
__host__ __device__ constexpr int nthreads_of(int kind, bool head = false) { return (kind >= 2 || head) ? 320 : 192; }
// output channels per pass: 32, except KIND 5 (16: all 64 input channels of the layer fit one pass instead)
__host__ __device__ constexpr int nbo_of(int kind) { return kind == 5 ? 16 : 32; }
// depth-plane ring: the kw-merged kernel's planes are small enough for 6 stages (prefetch across
// work-item boundaries)
__host__ __device__ constexpr int nstage_of(int kind) { return (kind == 4 || kind >= 5) ? 3 : 4; }
// KIND 6, 7, 8: the transposed conv with K = 64 (all input channels of a 64-channel layer in one pass, 32 output
// channels) run as three CLASS-GROUP launches.  Every tap of the 3x3x3 kernel belongs to exactly one of the 8 output
// parity classes (rd, rh, rw), so the 27 x 8 KB of (hi, lo) weights -- too much for shared memory next to the plane
// ring -- split by class: group 0 = classes (1,1,*) (12 taps), group 1 = (0,1,*) + (1,0,*) (12 taps), group 2 =
// (0,0,*) (3 taps).  Each launch writes ITS output voxels exactly once, residual and ReLU included, with the same
// number of MMAs as KIND 2's two input-channel passes -- which write, re-read and re-write the whole output
// (Hourglass conv6: 850 MB moved for 450 MB algorithmic).  The input (1/8 resolution, 50 MB) is re-read from L2.
__host__ __device__ constexpr bool is_t64(int kind) { return kind >= 6; }
__host__ __device__ constexpr bool is_transposed(int kind) { return kind == 2 || kind >= 5; }
__host__ __device__ constexpr int t64_ntaps(int grp) { return grp == 2 ? 3 : 12; }
__host__ __device__ constexpr int t64_nphases(int grp) { return grp == 1 ? 2 : 1; }
// phase pi of group grp handles the two classes (rd, rh, 0) and (rd, rh, 1)
__host__ __device__ constexpr int t64_rd(int grp, int pi) { return grp == 0 ? 1 : (grp == 1 ? pi : 0); }
__host__ __device__ constexpr int t64_rh(int grp, int pi) { return grp == 0 ? 1 : (grp == 1 ? 1 - pi : 0); }
// taps of an output parity r along one axis: r = 0 -> {1}, r = 1 -> {0, 2}
__host__ __device__ constexpr bool t64_uses(int r, int k) { return r == 0 ? k == 1 : k != 1; }
// shared-memory slot of tap (kd, kh, kw) in group grp (-1: not in the group): phases in order, then kd, kh, kw
__host__ __device__ constexpr int t64_slot(int grp, int kd, int kh, int kw) {
    int slot = 0;
    for (int pi = 0; pi < t64_nphases(grp); ++pi)
        for (int d = 0; d < 3; ++d)
            for (int h = 0; h < 3; ++h) {
                if (!t64_uses(t64_rd(grp, pi), d) || !t64_uses(t64_rh(grp, pi), h)) continue;
                if (d == kd && h == kh) return slot + kw;
                slot += 3;
            }
    return -1;
}
constexpr int TMEM_COLS = 512;

// KIND 0: stride-1 conv          M space = output = input grid; halo box 18x10 per plane (reference kernel)
// KIND 3: stride-1 conv, kw taps merged into N (the production kernel for stride-1 layers)
// KIND 1: stride-2 conv (k3,p1)  M space = output grid; four (h,w)-parity sub-tiles per input plane
// KIND 2: stride-2 transposed conv (k3,p1,op1)  M space = INPUT grid; 8 output parity classes,
//         each a small conv over a 17x9 halo box
template <int KIND> struct Geo;
template <> struct Geo<0> {
    static constexpr int CBK = 4;                                  // 8-channel blocks per pass (32 ch)
    static constexpr int PLANE_BYTES = CBK * 18 * 10 * 16;         // 11520
};
template <> struct Geo<1> {
    static constexpr int CBK = 2;                                  // 16 input channels per pass (smem budget)
    // sub-tile (ph,pw): (16+ph) rows x (8+pw) cols, each padded to a multiple of 128 bytes
    static constexpr int SUB_OFF0 = 0, SUB_OFF1 = 4096, SUB_OFF2 = 4096 + 4608, SUB_OFF3 = 4096 + 4608 + 4352;
    static constexpr int PLANE_BYTES = 4096 + 4608 + 4352 + 4992;  // 18048
};
template <> struct Geo<2> {
    static constexpr int CBK = 4;
    static constexpr int PLANE_BYTES = 9856;                       // 4*17*9*16 = 9792, padded to 128
};
// KIND 5: the transposed conv with K = 64: one pass covers ALL input channels of a 64-channel layer and 16 output
// channels (same 110 KB of weights as 32 x 32), so that a 64 -> 32 / 64 -> 64 layer runs as 2 / 4 passes that
// each write their own output channels once -- instead of input-channel passes that read-modify-write the
// output-resolution tensor in place (Hourglass conv6: 200 MB written, re-read and re-written per layer).
template <> struct Geo<5> {
    static constexpr int CBK = 8;
    static constexpr int PLANE_BYTES = 8 * 17 * 9 * 16;            // 19584 = 153 * 128
};
template <> struct Geo<6> : Geo<5> {};
template <> struct Geo<7> : Geo<5> {};
template <> struct Geo<8> : Geo<5> {};
// KIND 3: stride-1 conv with the three kw taps merged into the MMA's N dimension.  The 8 tile columns
// are w0-1 .. w0+6; every MMA multiplies the UNSHIFTED column block with [W(kw=0)|W(kw=1)|W(kw=2)], the
// epilogue adds the three partial results of neighbouring columns (warp shuffles).  6 of 8 columns
// produce outputs, but the A tile is streamed from shared memory 3x less often -- the operand
// bandwidth, not the tensor pipe, bounds these N<=64 MMAs (profiles/README.md).
// Its M tile is 8 rows x 16 columns (two 8-voxel core-matrix groups per row, still a uniform 128-byte
// group pitch because no w halo is stored): 14 of 16 columns carry outputs.
template <> struct Geo<3> {
    static constexpr int CBK = 4;
    static constexpr int PLANE_BYTES = 4 * 6 * 32 * 16;            // 12288: (4+2) rows x 32 columns x 32 channels
};
// KIND 4: stride-2 conv on the kw-merged scheme.  The depth and height strides are taken by the loads
// (plane index 2*od+kd-1; two row-parity boxes per plane through tensor maps with a doubled row pitch --
// full 256-byte rows, unlike KIND 1's 16-byte gathers); along W every input column is computed and the
// epilogue keeps the even centres.  The MMA work doubles, but these layers are small and the 16-byte TMA
// elements of KIND 1 were the bottleneck (measured: 55 us per pass against a 15 us HBM floor).
template <> struct Geo<4> {
    static constexpr int CBK = 4;
    static constexpr int EVEN_BYTES = 4 * 8 * 16 * 16;             // 8192: rows 2*oh
    static constexpr int ODD_BYTES = 4 * 9 * 16 * 16;              // 9216: rows 2*oh-1 .. 2*oh+15
    static constexpr int PLANE_BYTES = EVEN_BYTES + ODD_BYTES;     // 17408
};
constexpr int TH3 = 8, TW3 = 16;            // KIND 4 tile: 8 output rows x 16 input columns
constexpr int TWV = 14;                     // KIND 4: columns of a tile row that are computed with both neighbours
// KIND 3 tile: 4 rows x 32 columns, 30 of which produce outputs (the kw-merged scheme needs one halo column on
// each side INSIDE the M tile).  8 x 16 with 14 valid columns wasted 12.5 % of every MMA plus the ragged last
// tile column (240 = 17.1 x 14); 4 x 32 wastes 6.25 % and tiles the PSMNet grids exactly (136 = 34 x 4,
// 240 = 8 x 30, 120 = 4 x 30, 60 = 2 x 30): 272 instead of 306 tiles at 1/4 resolution, 68 instead of 81 at 1/8,
// 18 instead of 25 at 1/16.  The trunk is POWER bound (tools/tc_clock.py: the SM clock drops to 1.3-1.5 GHz under
// this MMA stream), so MMAs not issued are the only time saved.
constexpr int K3_TH = 4, K3_TW = 32, K3_TWV = 30;
// per-KIND tile steps (M-space rows / output columns a tile advances by)
__host__ __device__ constexpr int th_of(int kind) { return kind == 3 ? K3_TH : (kind == 4 ? TH3 : TH); }
__host__ __device__ constexpr int twstep_of(int kind) { return kind == 3 ? K3_TWV : (kind == 4 ? TWV : TW); }

struct Maps {
    CUtensorMap m[8];   // KIND 0/2: [0]=hi [1]=lo;  KIND 1: [(ph*2+pw)*2 + (0 hi | 1 lo)]
};

struct Params {
    const void* w_blob;        // this pass' packed weights
    const float* bias;         // [32] or null
    const uint4* res_hi;       // residual (blocked, output geometry) or null
    const uint4* res_lo;
    uint4* y_hi;
    uint4* y_lo;
    float* y_f32;              // Cout==1 mode (NCDHW fp32), else null
    const float* res_f32;
    int B;
    int Dm, Hm, Wm;            // M-space grid the tiles / depth segments run over
    int Do, Ho, Wo;            // output grid (indexing of y / residual)
    int in_cb0;                // first 8-channel block of the input tensor used by this pass
    int y_cb0, y_cbs;          // first block / total blocks of y
    int res_cb0, res_cbs;      // same for the residual tensor
    int tiles_h, tiles_w, nseg, seg_len, n_items;
    int balanced, planes_total; // balanced = 1: CTA k of G works planes [k T / G, (k + 1) T / G) of the T = columns x Dm
                                // (tile column, depth plane) sequence, cut into items at column ends (see ItemIter)
    int relu;
    int n_valid_out;           // output channels that exist (1 in y_f32 mode, else 32)
    float acc_scale;           // 2^-k undoing the power-of-two weight pre-scaling (exact)
    int flat;                  // KIND 3 only: 1 = 2-D 3x3 convolution (D == 1): only the kd = 1 taps are issued
    // fused classifier head (KIND 3 only): instead of storing the 32-channel activation a, the epilogue
    // writes its 27 per-tap projections T[tap][voxel] = sum_c a[c] * head_w[tap][c]; the 32->1 3x3x3
    // convolution that follows is then a 27-term gather (head_gather_kernel)
    long long* trace;          // debug timeline (dmb_b200_debug_set_trace) or null: CTA 0 records clock64() per role
    const float* head_w;       // [27][32] fp32 (device) or null
    float* head_t;             // per batch element [9][Do][Ho][Wo] Q planes + [nseg][2][9][Ho][Wo] segment spills, fp32
                               // (null: ordinary layer)
};

// 16-bit element codecs: FP16 = true -> IEEE half (11-bit significand), false -> bfloat16 (8-bit)
template <bool FP16>
__device__ __forceinline__ float round16(float x) {
    return FP16 ? __half2float(__float2half_rn(x)) : __bfloat162float(__float2bfloat16_rn(x));
}
template <bool FP16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    if (FP16) {
        __half2 v = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&v);
    }
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
template <bool FP16>
__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
    const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (FP16) {
            const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&u[i]));
            f[2 * i] = t.x;
            f[2 * i + 1] = t.y;
        } else {
            f[2 * i] = __uint_as_float(u[i] << 16);
            f[2 * i + 1] = __uint_as_float(u[i] & 0xFFFF0000u);
        }
    }
}

template <int KIND, bool SPLIT>
struct Smem {
    using G = Geo<KIND>;
    static constexpr int CBK = G::CBK;
    static constexpr int NKW = (KIND == 3 || KIND == 4) ? 3 : 1;   // kw taps merged into one B block
    static constexpr int NBO = nbo_of(KIND);                       // output channels per pass
    static constexpr int ROWS = (SPLIT ? 2 * NBO : NBO) * NKW;     // B-operand rows per channel block
    static constexpr int TAP_BYTES = CBK * ROWS * 16;              // one B block (a tap, or a (kd,kh) tap row)
    static constexpr int W_BYTES = (is_t64(KIND) ? t64_ntaps(KIND - 6) : TAPS / NKW) * TAP_BYTES;
    static constexpr int PLANE_BYTES = G::PLANE_BYTES;
    static constexpr int STAGE_BYTES = (SPLIT ? 2 : 1) * PLANE_BYTES;
    static constexpr int NST = nstage_of(KIND);
    static constexpr int PLANES_OFF = W_BYTES;
    static constexpr int BAR_OFF = PLANES_OFF + NST * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFF + 256;
    // fused classifier head (KIND 3 head variant): the 32 -> 1 convolution's 27 taps as the N rows of a second B operand
    // ([hi rows | lo rows] x 32 channels, K-major core matrices) and the activated tile as a second A operand
    // (128 voxels x 32 channels, hi plane then lo plane), both written by the CTA's own threads
    static constexpr int HEAD_B_OFF = TOTAL;
    static constexpr int HEAD_B_BYTES = 4 * NB * 16;               // one 16-bit plane of [32 taps][32 channels]: 2048
    static constexpr int HEAD_A_OFF = HEAD_B_OFF + 2 * HEAD_B_BYTES;
    static constexpr int HEAD_A_BYTES = 4 * 128 * 16;              // one 16-bit plane of [128 voxels][32 channels]: 8192
    static constexpr int TOTAL_HEAD = HEAD_A_OFF + (SPLIT ? 2 : 1) * HEAD_A_BYTES;
    static constexpr uint32_t LBO_B = ROWS * 16;
    static constexpr uint32_t SBO_B = 128;
    // The tensor core adds each K=16 partial product into the fp32 accumulator with truncation
    // (measured: error biased toward zero, growing with the chain length).  Accumulation chains are
    // therefore kept short.  KIND 0/1: one accumulator per depth tap kd and, in split mode, a
    // separate one for the small lo*hi term; the epilogue adds them with round-to-nearest.
    static constexpr int KD_COLS = SPLIT ? 64 : 32;                // [hi*Whi | hi*Wlo] of one kd
    static constexpr int LH_COL = 3 * KD_COLS;                     // lo*Whi (split only)
    // KIND 2: one accumulator per output parity class (chains are <= 32 MMAs by construction)
    static constexpr int CLS_COLS = SPLIT ? 2 * NBO : NBO;
    static constexpr int ACC_COLS =
        is_t64(KIND) ? 2 * CLS_COLS
                     : ((KIND == 2 || KIND == 5) ? 4 * CLS_COLS : ((KIND == 3 || KIND == 4) ? ROWS : (SPLIT ? 3 * 64 + 32 : 3 * 32)));
    static constexpr int HEAD_P_COL = 2 * ACC_COLS;                // TMEM columns of the head's 27 (32) tap projections
    static_assert(2 * ACC_COLS <= TMEM_COLS, "accumulators exceed TMEM");
    static_assert(KIND != 3 || HEAD_P_COL + NB <= TMEM_COLS, "no TMEM columns left for the head projections");
    static_assert(TOTAL <= 227 * 1024, "shared memory budget exceeded");
    static_assert(KIND != 3 || TOTAL_HEAD <= 227 * 1024, "shared memory budget exceeded (head variant)");
};

struct Item {
    int b, d0, d1, h0, w0;
};
template <int THSTEP, int TWSTEP>
__device__ __forceinline__ Item decode_item(const Params& p, int item) {
    const int per_b = p.tiles_w * p.tiles_h * p.nseg;
    Item it;
    it.b = item / per_b;
    int r = item - it.b * per_b;
    const int seg = r / (p.tiles_w * p.tiles_h);
    r -= seg * p.tiles_w * p.tiles_h;
    const int th = r / p.tiles_w, tw = r - th * p.tiles_w;
    it.d0 = seg * p.seg_len;
    it.d1 = min(p.Dm, it.d0 + p.seg_len);
    it.h0 = th * THSTEP;
    it.w0 = tw * TWSTEP;
    return it;
}

// Work items of one CTA.  Static schedule (balanced = 0): items blockIdx.x, blockIdx.x + G, ... of the uniform
// (tile column x depth segment) grid.  Balanced schedule: the CTA's contiguous share of the linearised (tile column,
// depth plane) sequence, cut at column ends -- every CTA gets the same number of planes (+-1) in at most a few items,
// where the static schedule's slowest CTA ran ceil(items / G) whole segments (measured on the hourglass' 1/8-resolution
// layers: SMs active 66 % of the kernel, profiles/r2_ncu_full_conv3d_tc_k7.txt).  A share boundary that would leave
// a one-plane item at a column end is moved to the column end.
struct ItemIter {
    int L, L1;
};
__device__ __forceinline__ int snap_boundary(const Params& p, long long k) {
    int L = (int)(k * p.planes_total / gridDim.x);
    const int r = L % p.Dm;
    if (p.Dm >= 4) {
        if (r == 1) L -= 1;
        else if (r == p.Dm - 1) L += 1;
    }
    return L;
}
__device__ __forceinline__ ItemIter items_begin(const Params& p) {
    ItemIter s;
    if (p.balanced) {
        s.L = snap_boundary(p, blockIdx.x);
        s.L1 = blockIdx.x + 1 == gridDim.x ? p.planes_total : snap_boundary(p, (long long)blockIdx.x + 1);
    } else {
        s.L = blockIdx.x;
        s.L1 = p.n_items;
    }
    return s;
}
template <int THSTEP, int TWSTEP>
__device__ __forceinline__ bool items_next(const Params& p, ItemIter& s, Item& it) {
    if (s.L >= s.L1) return false;
    if (p.balanced) {
        const int col = s.L / p.Dm;                // (b, tile row, tile column)
        const int d0 = s.L - col * p.Dm;
        const int d1 = min(p.Dm, d0 + (s.L1 - s.L));
        const int per_b = p.tiles_w * p.tiles_h;
        it.b = col / per_b;
        const int r = col - it.b * per_b;
        const int th = r / p.tiles_w, tw = r - th * p.tiles_w;
        it.d0 = d0;
        it.d1 = d1;
        it.h0 = th * THSTEP;
        it.w0 = tw * TWSTEP;
        s.L += d1 - d0;
    } else {
        it = decode_item<THSTEP, TWSTEP>(p, s.L);
        s.L += gridDim.x;
    }
    return true;
}

// bias + residual + ReLU + 16-bit (hi[,lo]) split + store of one output voxel's 32 channels
template <bool FP16, int NCB = 4>
__device__ __forceinline__ void store_voxel(const Params& p, float (&v)[NB], const float (&bias)[NB], int b, int d, int h,
                                            int w, int cb_off = 0) {
    const size_t plane_sz = (size_t)p.Ho * p.Wo;
    const size_t vox = (size_t)d * plane_sz + (size_t)h * p.Wo + w;
    if (p.y_f32) {
        float o = v[0] + bias[0];
        const size_t oi = (size_t)b * p.Do * plane_sz + vox;
        if (p.res_f32) o += __ldg(p.res_f32 + oi);
        if (p.relu) o = fmaxf(o, 0.f);
        p.y_f32[oi] = o;
        return;
    }
#pragma unroll
    for (int c = 0; c < 8 * NCB; ++c) v[c] += bias[c];
    if (p.res_hi) {
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb) {
            const size_t ri = ((size_t)(b * p.res_cbs + p.res_cb0 + cb_off + cb) * p.Do) * plane_sz + vox;
            float f[8];
            unpack8<FP16>(__ldg(p.res_hi + ri), f);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[cb * 8 + e] += f[e];
            if (p.res_lo) {
                unpack8<FP16>(__ldg(p.res_lo + ri), f);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[cb * 8 + e] += f[e];
            }
        }
    }
    if (p.relu) {
#pragma unroll
        for (int c = 0; c < 8 * NCB; ++c) v[c] = fmaxf(v[c], 0.f);
    }
#pragma unroll
    for (int cb = 0; cb < NCB; ++cb) {
        const size_t yi = ((size_t)(b * p.y_cbs + p.y_cb0 + cb_off + cb) * p.Do) * plane_sz + vox;
        float hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float x = v[cb * 8 + e];
            const float hv = round16<FP16>(x);
            hi[e] = hv;
            lo[e] = x - hv;
        }
        p.y_hi[yi] = make_uint4(pack2<FP16>(hi[0], hi[1]), pack2<FP16>(hi[2], hi[3]), pack2<FP16>(hi[4], hi[5]),
                                pack2<FP16>(hi[6], hi[7]));
        if (p.y_lo)
            p.y_lo[yi] = make_uint4(pack2<FP16>(lo[0], lo[1]), pack2<FP16>(lo[2], lo[3]), pack2<FP16>(lo[4], lo[5]),
                                    pack2<FP16>(lo[6], lo[7]));
    }
}

// fused classifier head.  The 32 -> 1 3x3x3 convolution that follows the 32 -> 32 layer is
//     out(d, h, w) = sum_{kd,kh,kw} P[kd][kh][kw](d + kd - 1, h + kh - 1, w + kw - 1),   P[tap](voxel) = <a(voxel), head_w[tap]>.
// P is itself a small GEMM -- [128 voxels] x [32 channels] x [27 taps] per tile and plane -- and runs on the tensor core:
// the epilogue writes the activated tile (bias, ReLU, 16-bit split) to shared memory as a second A operand, the MMA
// warp multiplies it with the head weights (six N = 32 MMAs into 32 more TMEM columns) in the middle of issuing the
// NEXT plane, and the epilogue reads the 27 projections back.  (Round 1-2 computed P with 864 FMAs per voxel in the
// epilogue, every weight a shared-memory load: the LSU pipe, not the tensor core, bounded the kernel -- 281 us per
// head against 51 us for the same layer without the head.)
// The epilogue thread of a tile position marches along depth, so the kd sum is local to it:
//     Q[kh][kw](d) = P[0][kh][kw](d - 1) + P[1][kh][kw](d) + P[2][kh][kw](d + 1)
// is accumulated in registers across consecutive planes and NINE planes are stored instead of 27 tap planes (the
// gather then reads 9 values per output).  At the two ends of a depth segment the missing neighbour plane belongs
// to another work item: the two boundary projections go to small per-segment "spill" planes that the gather adds.
//   head_t (per batch element): Q [9][D][H][W], then spills [nseg][2][9][H][W]  (0: P[2](d0) -> Q(d0 - 1), 1: P[0](d1 - 1) -> Q(d1))
// K0..K1: the (kh, kw) combinations this warp owns (two warps share a TMEM lane quarter).
__device__ __forceinline__ size_t head_batch_stride(const Params& p) {
    return ((size_t)9 * p.Do + (size_t)18 * p.nseg) * p.Ho * p.Wo;
}
template <int K0, int K1>
__device__ __forceinline__ void head_step(const Params& p, const uint32_t (&pr)[NB], float inv_scale, const Item& it, int d,
                                          int h, int w, float (&qprev)[5], float (&qcur)[5]) {
    const size_t hw_sz = (size_t)p.Ho * p.Wo;
    float* base = p.head_t + (size_t)it.b * head_batch_stride(p) + (size_t)h * p.Wo + w;
    const bool first = d == it.d0;
    float* spill_lo = base + ((size_t)9 * p.Do + (size_t)(it.d0 / p.seg_len) * 18) * hw_sz;
#pragma unroll
    for (int khw = K0; khw < K1; ++khw) {
        // P[kd][khw] of this voxel: columns kd * 9 + khw of the head MMA's accumulator
        const float p0 = __uint_as_float(pr[khw]) * inv_scale;
        const float p1 = __uint_as_float(pr[9 + khw]) * inv_scale;
        const float p2 = __uint_as_float(pr[18 + khw]) * inv_scale;
        const int k = khw - K0;
        if (first) {
            spill_lo[(size_t)khw * hw_sz] = p2;                                  // contribution to Q(d0 - 1)
            qprev[k] = p1;                                                       // Q(d0) so far (P[0](d0 - 1): the neighbour's spill)
        } else {
            base[((size_t)khw * p.Do + (d - 1)) * hw_sz] = qprev[k] + p2;        // Q(d - 1) complete
            qprev[k] = qcur[k] + p1;                                             // Q(d) so far
        }
        qcur[k] = p0;                                                            // towards Q(d + 1)
    }
}
template <int K0, int K1>
__device__ __forceinline__ void head_finish(const Params& p, const Item& it, int h, int w, const float (&qprev)[5],
                                            const float (&qcur)[5]) {
    const size_t hw_sz = (size_t)p.Ho * p.Wo;
    float* base = p.head_t + (size_t)it.b * head_batch_stride(p) + (size_t)h * p.Wo + w;
    float* spill_hi = base + ((size_t)9 * p.Do + (size_t)(it.d0 / p.seg_len) * 18 + 9) * hw_sz;
#pragma unroll
    for (int khw = K0; khw < K1; ++khw) {
        base[((size_t)khw * p.Do + (it.d1 - 1)) * hw_sz] = qprev[khw - K0];      // Q(d1 - 1) without P[2](d1)
        spill_hi[(size_t)khw * hw_sz] = qcur[khw - K0];                          // P[0](d1 - 1): contribution to Q(d1)
    }
}

// out[b,d,h,w] = res + sum_{kh,kw} Qfull[kh][kw][d][h+kh-1][w+kw-1] (zero outside the plane), where Qfull adds the
// neighbouring depth segments' spill planes at the first / last plane of a segment: the 32->1 3x3x3 head
// (aggregators/PSMNet.py:41-52) after head_step / head_finish; every Q element is read once per use (9 per output).
__global__ void __launch_bounds__(256) head_gather_kernel(const float* __restrict__ T, const float* __restrict__ res,
                                                          float* __restrict__ y, int B, int D, int H, int W, int seg_len,
                                                          int nseg) {
    // grid: (plane positions / 256, D, B) -- no 64-bit index divisions (round 2's flat index cost three per thread)
    const int hw = H * W;
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= hw) return;
    const int d = blockIdx.y, b = blockIdx.z;
    const int h = r / W, w = r - h * W;
    const size_t hw_sz = (size_t)hw;
    const float* tb = T + (size_t)b * ((size_t)9 * D + (size_t)18 * nseg) * hw_sz;
    const int seg = d / seg_len;
    const int seg_end = min(D, (seg + 1) * seg_len);
    // spill planes to add: previous segment's "hi" at the first plane, next segment's "lo" at the last plane
    const float* sp_a = (d == seg * seg_len && seg > 0) ? tb + ((size_t)9 * D + (size_t)(seg - 1) * 18 + 9) * hw_sz : nullptr;
    const float* sp_b = (d == seg_end - 1 && seg + 1 < nseg) ? tb + ((size_t)9 * D + (size_t)(seg + 1) * 18) * hw_sz : nullptr;
    const float* qd = tb + (size_t)d * hw_sz + r;          // Q[0][d] at this position; plane khw is khw * D planes further
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int hh = h + kh - 1;
        if (hh < 0 || hh >= H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int ww = w + kw - 1;
            if (ww < 0 || ww >= W) continue;
            const int khw = kh * 3 + kw;
            const int o = (kh - 1) * W + (kw - 1);
            float q = __ldg(qd + (size_t)khw * D * hw_sz + o);
            if (sp_a) q += __ldg(sp_a + (size_t)khw * hw_sz + r + o);
            if (sp_b) q += __ldg(sp_b + (size_t)khw * hw_sz + r + o);
            acc[kh] += q;
        }
    }
    float o = (acc[0] + acc[1]) + acc[2];
    const size_t i = ((size_t)b * D + d) * hw_sz + r;
    if (res) o += __ldg(res + i);
    y[i] = o;
}

// timeline tracing (tools/tc_trace.py): role r of CTA 0 appends clock64() stamps to trace[r * 4096 ...]
__device__ __forceinline__ void trace_stamp(const Params& p, int role, uint32_t& n) {
    if (p.trace && blockIdx.x == 0 && n < 4096) p.trace[role * 4096 + n++] = clock64();
}

template <int KIND, bool SPLIT, bool FP16, bool HEAD>
__global__ void __launch_bounds__(nthreads_of(KIND, HEAD), 1)
conv3d_tc_kernel(const __grid_constant__ Maps maps, const Params p) {
    static_assert(!HEAD || KIND == 3, "the fused classifier head exists for the stride-1 kernel only");
    using S = Smem<KIND, SPLIT>;
    constexpr int CBK = S::CBK;
    constexpr int NSTAGE = S::NST;
    constexpr int EPI_WARPS = nthreads_of(KIND, HEAD) / 32 - 2;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* w_smem = smem;
    unsigned char* planes = smem + S::PLANES_OFF;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);   // [NSTAGE]
    uint64_t* empty = full + NSTAGE;                                    // [NSTAGE]
    uint64_t* tfull = empty + NSTAGE;                                   // [2]
    uint64_t* tempty = tfull + 2;                                       // [2]
    uint64_t* wbar = tempty + 2;                                        // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    uint64_t* a2full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF + 112);   // head: activated tile written (8 warps)
    uint64_t* pfull = a2full + 1;                                               // head: projections complete (commit)
    float head_inv_scale = 1.f;
    if (HEAD) {
        // head weights -> second B operand.  fp16 planes: a power-of-two scale keeps the lo parts normal numbers.
        float* red = reinterpret_cast<float*>(smem + S::BAR_OFF + 192);
        float mx = 0.f;
        for (int i = threadIdx.x; i < TAPS * NB; i += blockDim.x) mx = fmaxf(mx, fabsf(__ldg(p.head_w + i)));
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) red[warp] = mx;
        __syncthreads();
        mx = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
        int e = 0;                                 // scale = 2^e with max|w| * 2^e in [2^9, 2^10)
        if (FP16 && mx > 0.f && mx < 3.0e38f) {
            int ex;
            frexpf(mx, &ex);                       // mx = f * 2^ex, f in [0.5, 1)
            e = max(-40, min(40, 10 - ex));
        }
        const float scale = exp2f((float)e);
        head_inv_scale = exp2f((float)-e);
        unsigned char* b2 = smem + S::HEAD_B_OFF;
        for (int i = threadIdx.x; i < NB * NB; i += blockDim.x) {
            const int n = i >> 5, k = i & 31;      // n: tap (27 real rows, 5 zero rows), k: channel
            const float wv = n < TAPS ? __ldg(p.head_w + n * NB + k) * scale : 0.f;
            const int off = (k >> 3) * (NB * 16) + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2;
            uint16_t hi16, lo16;
            if (FP16) {
                const __half hi = __float2half_rn(wv);
                const __half lo = __float2half_rn(wv - __half2float(hi));
                hi16 = __half_as_ushort(hi);
                lo16 = __half_as_ushort(lo);
            } else {
                __nv_bfloat16 hi, lo;
                split_bf16(wv, hi, lo);
                hi16 = __bfloat16_as_ushort(hi);
                lo16 = __bfloat16_as_ushort(lo);
            }
            *reinterpret_cast<uint16_t*>(b2 + off) = hi16;
            *reinterpret_cast<uint16_t*>(b2 + S::HEAD_B_BYTES + off) = lo16;
        }
        fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core's operand reads
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], EPI_WARPS);
        }
        mbar_init(wbar, 1);
        if (HEAD) {
            mbar_init(a2full, EPI_WARPS);
            mbar_init(pfull, 1);
        }
        fence_mbar_init();
    }
    if (warp == 1) {   // TMEM allocation (whole warp, .sync.aligned)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            mbar_expect_tx(wbar, S::W_BYTES);      // weights: packed in global exactly as they sit in smem
            if (is_t64(KIND)) {                    // only this class group's taps, in slot order
                for (int t = 0; t < TAPS; ++t) {
                    const int slot = t64_slot(KIND - 6, t / 9, (t / 3) % 3, t % 3);
                    if (slot >= 0)
                        bulk_g2s(w_smem + slot * S::TAP_BYTES, reinterpret_cast<const unsigned char*>(p.w_blob) + t * S::TAP_BYTES,
                                 S::TAP_BYTES, wbar);
                }
            } else {
                for (int t = 0; t < TAPS / S::NKW; ++t)
                    bulk_g2s(w_smem + t * S::TAP_BYTES, reinterpret_cast<const unsigned char*>(p.w_blob) + t * S::TAP_BYTES,
                             S::TAP_BYTES, wbar);
            }
            uint32_t n = 0;
            uint32_t ntrace_p = 0;
            const int in_cb0_k4 = p.in_cb0;
            ItemIter iter = items_begin(p);
            Item it;
            while (items_next<th_of(KIND), twstep_of(KIND)>(p, iter, it)) {
                const int nout = it.d1 - it.d0;
                const int nplanes = (KIND == 0 || KIND == 3) ? nout + 2 : ((KIND == 1 || KIND == 4) ? 2 * nout + 1 : nout + 1);   // KIND 2/5: nout + 1
                const int pl0 = (KIND == 0 || KIND == 3) ? it.d0 - 1 : ((KIND == 1 || KIND == 4) ? 2 * it.d0 - 1 : it.d0);
                for (int j = 0; j < nplanes; ++j, ++n) {
                    const int pl = pl0 + j;
                    const uint32_t slot = n % NSTAGE;
                    trace_stamp(p, 2, ntrace_p);                               // [0] plane wanted
                    mbar_wait(&empty[slot], ((n / NSTAGE) & 1) ^ 1);
                    trace_stamp(p, 2, ntrace_p);                               // [1] stage free, TMA issued next
                    unsigned char* dst = planes + slot * S::STAGE_BYTES;
                    if (KIND == 0) {
                        mbar_expect_tx(&full[slot], (SPLIT ? 2 : 1) * CBK * 18 * 10 * 16);
                        tma_load_5d(dst, &maps.m[0], &full[slot], 8 * (it.w0 - 1), it.h0 - 1, pl, p.in_cb0, it.b);
                        if (SPLIT)
                            tma_load_5d(dst + S::PLANE_BYTES, &maps.m[1], &full[slot], 8 * (it.w0 - 1), it.h0 - 1, pl,
                                        p.in_cb0, it.b);
                    } else if (KIND == 3) {
                        mbar_expect_tx(&full[slot], (SPLIT ? 2 : 1) * S::PLANE_BYTES);
                        tma_load_5d(dst, &maps.m[0], &full[slot], 8 * (it.w0 - 1), it.h0 - 1, pl, p.in_cb0, it.b);
                        if (SPLIT)
                            tma_load_5d(dst + S::PLANE_BYTES, &maps.m[1], &full[slot], 8 * (it.w0 - 1), it.h0 - 1, pl,
                                        p.in_cb0, it.b);
                    } else if (KIND == 4) {
                        // m[0]/m[1]: even input rows (hi/lo), m[2]/m[3]: odd input rows; row index = ih >> 1
                        mbar_expect_tx(&full[slot], (SPLIT ? 2 : 1) * S::PLANE_BYTES);
                        tma_load_5d(dst, &maps.m[0], &full[slot], 8 * (it.w0 - 1), it.h0, pl, in_cb0_k4, it.b);
                        tma_load_5d(dst + Geo<4>::EVEN_BYTES, &maps.m[2], &full[slot], 8 * (it.w0 - 1), it.h0 - 1, pl, in_cb0_k4, it.b);
                        if (SPLIT) {
                            tma_load_5d(dst + S::PLANE_BYTES, &maps.m[1], &full[slot], 8 * (it.w0 - 1), it.h0, pl, in_cb0_k4, it.b);
                            tma_load_5d(dst + S::PLANE_BYTES + Geo<4>::EVEN_BYTES, &maps.m[3], &full[slot], 8 * (it.w0 - 1),
                                        it.h0 - 1, pl, in_cb0_k4, it.b);
                        }
                    } else if (is_transposed(KIND)) {
                        mbar_expect_tx(&full[slot], (SPLIT ? 2 : 1) * CBK * 17 * 9 * 16);
                        tma_load_5d(dst, &maps.m[0], &full[slot], 8 * it.w0, it.h0, pl, p.in_cb0, it.b);
                        if (SPLIT)
                            tma_load_5d(dst + S::PLANE_BYTES, &maps.m[1], &full[slot], 8 * it.w0, it.h0, pl, p.in_cb0, it.b);
                    } else {
                        // four parity sub-tiles; odd-parity ones start one index earlier (taps k=0)
                        mbar_expect_tx(&full[slot], (SPLIT ? 2 : 1) * CBK * 16 * (16 * 8 + 16 * 9 + 17 * 8 + 17 * 9));
                        constexpr int offs[4] = {Geo<1>::SUB_OFF0, Geo<1>::SUB_OFF1, Geo<1>::SUB_OFF2, Geo<1>::SUB_OFF3};
#pragma unroll
                        for (int sub = 0; sub < 4; ++sub) {
                            const int ph = sub >> 1, pw = sub & 1;
                            tma_load_5d(dst + offs[sub], &maps.m[sub * 2], &full[slot], 0, it.w0 - pw, it.h0 - ph, pl,
                                        p.in_cb0);
                            if (SPLIT)
                                tma_load_5d(dst + S::PLANE_BYTES + offs[sub], &maps.m[sub * 2 + 1], &full[slot], 0,
                                            it.w0 - pw, it.h0 - ph, pl, p.in_cb0);
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        {
            constexpr uint32_t idesc_main = make_idesc(S::ROWS, FP16 ? 0u : 1u);
            constexpr uint32_t idesc_lo = make_idesc(S::NBO, FP16 ? 0u : 1u);
            mbar_wait(wbar, 0);
            const uint32_t w_addr = smem_u32(w_smem);
            const uint32_t planes_addr = smem_u32(planes);
            const uint32_t b_lo0 = desc_lo(w_addr, S::LBO_B);
            constexpr uint32_t b_hi = desc_hi(S::SBO_B);
            auto bdesc = [&](int tap, int kk) {
                return desc_of(b_lo0 + ((tap * S::TAP_BYTES + 2 * kk * (int)S::LBO_B) >> 4), b_hi);
            };
            uint32_t n_base = 0, t_base = 0;
            uint32_t ntrace = 0;
            // fused head: the projection MMAs of plane t are issued (by the elected lane) in the middle of plane t + 1's
            // issue stream -- by then the epilogue has written plane t's activated tile -- and after the last plane
            uint32_t head_planes = 0;              // planes issued so far (warp-uniform)
            auto issue_head = [&](uint32_t t_head) {
                constexpr uint32_t idesc_head = make_idesc(NB, FP16 ? 0u : 1u);
                constexpr uint32_t LBO_A2 = 128 * 16, LBO_B2 = NB * 16;
                constexpr uint32_t hiw = desc_hi(128);
                const uint32_t a2 = desc_lo(smem_u32(smem + S::HEAD_A_OFF), LBO_A2);
                const uint32_t b2 = desc_lo(smem_u32(smem + S::HEAD_B_OFF), LBO_B2);
                const uint32_t pacc = tmem_base + S::HEAD_P_COL;
                mbar_wait(a2full, t_head & 1);
                tcgen05_fence_after();
                if (SPLIT) {                       // small terms first: lo * Whi, hi * Wlo, then hi * Whi
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const uint32_t ao = (2 * kk * LBO_A2) >> 4, bo = (2 * kk * LBO_B2) >> 4;
                        if (kk == 0)
                            mma_f16_ss<false>(pacc, a2 + (S::HEAD_A_BYTES >> 4) + ao, hiw, b2 + bo, hiw, idesc_head);
                        else
                            mma_f16_ss<true>(pacc, a2 + (S::HEAD_A_BYTES >> 4) + ao, hiw, b2 + bo, hiw, idesc_head);
                    }
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const uint32_t ao = (2 * kk * LBO_A2) >> 4, bo = (2 * kk * LBO_B2) >> 4;
                        mma_f16_ss<true>(pacc, a2 + ao, hiw, b2 + (S::HEAD_B_BYTES >> 4) + bo, hiw, idesc_head);
                    }
                }
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const uint32_t ao = (2 * kk * LBO_A2) >> 4, bo = (2 * kk * LBO_B2) >> 4;
                    if (!SPLIT && kk == 0)
                        mma_f16_ss<false>(pacc, a2 + ao, hiw, b2 + bo, hiw, idesc_head);
                    else
                        mma_f16_ss<true>(pacc, a2 + ao, hiw, b2 + bo, hiw, idesc_head);
                }
                commit_one(pfull);
            };
            ItemIter iter = items_begin(p);
            Item it;
            while (items_next<th_of(KIND), twstep_of(KIND)>(p, iter, it)) {
                const int nout = it.d1 - it.d0;
                if (KIND == 0) {
                    constexpr uint32_t LBO_A = 18 * 10 * 16, SBO_A = 10 * 16;
                    int waited = 0;
                    for (int od = 0; od < nout; ++od) {
                        while (waited < od + 3) {
                            const uint32_t n = n_base + waited;
                            mbar_wait(&full[n % NSTAGE], (n / NSTAGE) & 1);
                            ++waited;
                        }
                        const uint32_t t = t_base + od;
                        const uint32_t buf = t & 1;
                        mbar_wait(&tempty[buf], ((t >> 1) & 1) ^ 1);
                        tcgen05_fence_after();
                        const uint32_t acc0 = tmem_base + buf * S::ACC_COLS;
                        uint32_t first_lo = 1;
#pragma unroll 1
                        for (int kd = 0; kd < 3; ++kd) {
                            const uint32_t a_hi = planes_addr + ((n_base + od + kd) % NSTAGE) * S::STAGE_BYTES;
                            const uint32_t acc = acc0 + kd * S::KD_COLS;
                            const uint32_t a_lo0 = desc_lo(a_hi, LBO_A);
                            const uint32_t b_lo_kd = b_lo0 + ((kd * 9 * S::TAP_BYTES) >> 4);
                            constexpr uint32_t a_hiw = desc_hi(SBO_A);
#pragma unroll
                            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                                for (int kw = 0; kw < 3; ++kw) {
                                    const uint32_t a_off = (kh * 10 + kw) * 16;
#pragma unroll
                                    for (int kk = 0; kk < CBK / 2; ++kk) {
                                        const bool first = (kh == 0 && kw == 0 && kk == 0);
                                        const uint64_t db = desc_of(b_lo_kd + (((kh * 3 + kw) * S::TAP_BYTES + 2 * kk * (int)S::LBO_B) >> 4), b_hi);
                                        tcgen05_mma_bf16(acc, desc_of(a_lo0 + ((a_off + 2 * kk * LBO_A) >> 4), a_hiw), db,
                                                         idesc_main, first ? 0u : 1u);
                                        if (SPLIT) {
                                            tcgen05_mma_bf16(acc0 + S::LH_COL,
                                                             desc_of(a_lo0 + ((S::PLANE_BYTES + a_off + 2 * kk * LBO_A) >> 4), a_hiw),
                                                             db, idesc_lo, (first && first_lo) ? 0u : 1u);
                                        }
                                    }
                                }
                            }
                            first_lo = 0;
                        }
                        tcgen05_commit(&tfull[buf]);
                        tcgen05_commit(&empty[(n_base + od) % NSTAGE]);          // plane d-1 is done
                        if (od == nout - 1) {
                            tcgen05_commit(&empty[(n_base + od + 1) % NSTAGE]);
                            tcgen05_commit(&empty[(n_base + od + 2) % NSTAGE]);
                        }
                    }
                    n_base += nout + 2;
                    t_base += nout;
                } else if (KIND == 3) {
                    constexpr uint32_t LBO_A = (K3_TH + 2) * K3_TW * 16, SBO_A = 8 * 16;   // 6 rows of 512 bytes per channel block
                    constexpr uint32_t idesc_hi3 = make_idesc(3 * NB, FP16 ? 0u : 1u);     // A_lo x [Whi kw0..2]
                    constexpr uint32_t a_hiw = desc_hi(SBO_A);
                    // ring slot of plane (n_base + od), kept as a wrapping counter (no modulo in the loop)
                    uint32_t slot = n_base % NSTAGE, phase = (n_base / NSTAGE) & 1;
                    uint32_t wslot = slot, wphase = phase;                     // next plane to wait for
                    // The barrier waits of plane od+1 (its newest input plane, its accumulator buffer) are taken in the
                    // MIDDLE of issuing plane od -- after the kd = 0, 1 MMAs, before the kd = 2 ones -- so that the issue of
                    // consecutive planes is back to back.  Measured (tools/tc_trace.py): tcgen05.mma issue is synchronous
                    // with execution (issuing a plane's 36 MMAs takes the 2660 cycles they execute in), and the two
                    // mbarrier waits between planes, although always satisfied, cost ~550 cycles of idle tensor pipe.
                    int waited = 0;
                    auto wait_inputs = [&](int upto) {                         // planes [.., upto) of this item present
                        while (waited < upto) {
                            mbar_wait(&full[wslot], wphase);
                            if (++wslot == NSTAGE) { wslot = 0; wphase ^= 1; }
                            ++waited;
                        }
                    };
                    wait_inputs(3);
                    mbar_wait(&tempty[t_base & 1], ((t_base >> 1) & 1) ^ 1);
                    tcgen05_fence_after();
                    for (int od = 0; od < nout; ++od) {
                        const uint32_t t = t_base + od;
                        const uint32_t buf = t & 1;
                        const uint32_t s1 = slot + 1 >= NSTAGE ? slot + 1 - NSTAGE : slot + 1;
                        const uint32_t s2 = slot + 2 >= NSTAGE ? slot + 2 - NSTAGE : slot + 2;
                        const uint32_t acc = tmem_base + buf * S::ACC_COLS;
                        const uint32_t a_lo_kd[3] = {desc_lo(planes_addr + slot * S::STAGE_BYTES, LBO_A),
                                                     desc_lo(planes_addr + s1 * S::STAGE_BYTES, LBO_A),
                                                     desc_lo(planes_addr + s2 * S::STAGE_BYTES, LBO_A)};
                        auto issue_kd = [&](int kd) {
#pragma unroll
                            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                                for (int kk = 0; kk < CBK / 2; ++kk) {
                                    const uint32_t b_lo = b_lo0 + (((kd * 3 + kh) * S::TAP_BYTES + 2 * kk * (int)S::LBO_B) >> 4);
                                    const uint32_t a_off = kh * (K3_TW * 16) + 2 * kk * LBO_A;
                                    if (kd == 0 && kh == 0 && kk == 0)
                                        mma_f16_ss<false>(acc, a_lo_kd[kd] + (a_off >> 4), a_hiw, b_lo, b_hi, idesc_main);
                                    else if (kd == 1 && kh == 0 && kk == 0)    // the first MMA of a flat (2-D) layer
                                        mma_f16_ss_rt(acc, a_lo_kd[kd] + (a_off >> 4), a_hiw, b_lo, b_hi, idesc_main, p.flat ? 0u : 1u);
                                    else
                                        mma_f16_ss<true>(acc, a_lo_kd[kd] + (a_off >> 4), a_hiw, b_lo, b_hi, idesc_main);
                                    if (SPLIT)   // lo*Whi of the three kw lands on the (small) hi*Wlo columns
                                        mma_f16_ss<true>(acc + 3 * NB, a_lo_kd[kd] + ((S::PLANE_BYTES + a_off) >> 4), a_hiw, b_lo, b_hi,
                                                         idesc_hi3);
                                }
                            }
                        };
                        // one elected lane does everything for the plane: the issue stream must not pause (see above)
                        const bool more = od + 1 < nout;
                        if (elect_one()) {
                            trace_stamp(p, 0, ntrace);                         // [0] plane start
                            if (!p.flat) issue_kd(0);                          // flat: a 2-D 3x3 layer = the kd = 1 taps of one plane
                            issue_kd(1);
                            trace_stamp(p, 0, ntrace);                         // [1] kd 0,1 issued
                            if (more) {                                        // barriers of the NEXT plane (already complete
                                                                               // in steady state: one overlapped poll)
                                uint64_t* bt = &tempty[buf ^ 1];
                                const uint32_t pt = (((t + 1) >> 1) & 1) ^ 1;
                                if (!mbar_try_wait2(&full[wslot], wphase, bt, pt)) {
                                    mbar_wait(&full[wslot], wphase);
                                    mbar_wait(bt, pt);
                                }
                                tcgen05_fence_after();
                            }
                            trace_stamp(p, 0, ntrace);                         // [2] next plane's barriers passed
                            if (HEAD && head_planes > 0) issue_head(head_planes - 1);
                            if (!p.flat) issue_kd(2);
                            commit_one(&tfull[buf]);
                            commit_one(&empty[slot]);
                            if (od == nout - 1) {
                                commit_one(&empty[s1]);
                                commit_one(&empty[s2]);
                            }
                            trace_stamp(p, 0, ntrace);                         // [3] plane's MMAs issued
                        }
                        __syncwarp();
                        if (more) {                                            // warp-uniform bookkeeping of the wait above
                            if (++wslot == NSTAGE) { wslot = 0; wphase ^= 1; }
                            ++waited;
                        }
                        if (++slot == NSTAGE) { slot = 0; phase ^= 1; }
                        ++head_planes;
                    }
                    n_base += nout + 2;
                    t_base += nout;
                } else if (KIND == 4) {
                    // plane-driven like KIND 1, one accumulator (all kd) per output like KIND 3
                    constexpr uint32_t idesc_hi3 = make_idesc(3 * NB, FP16 ? 0u : 1u);
                    constexpr uint32_t a_hiw = desc_hi(8 * 16);
                    constexpr uint32_t LBO_E = 8 * 256, LBO_O = 9 * 256;
                    const int nplanes = 2 * nout + 1;
                    auto issue_kd = [&](uint32_t a_stage, uint32_t t, int kd) {      // called by the elected lane only
                        const uint32_t acc = tmem_base + (t & 1) * S::ACC_COLS;
                        const uint32_t e_lo0 = desc_lo(a_stage, LBO_E);
                        const uint32_t o_lo0 = desc_lo(a_stage + Geo<4>::EVEN_BYTES, LBO_O);
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                            for (int kk = 0; kk < CBK / 2; ++kk) {
                                const bool first = (kd == 0 && kh == 0 && kk == 0);
                                const uint32_t b_lo = b_lo0 + (((kd * 3 + kh) * S::TAP_BYTES + 2 * kk * (int)S::LBO_B) >> 4);
                                // kh=1: even rows, local row = oh; kh=0: odd rows, row oh; kh=2: odd rows, row oh+1
                                const uint32_t lo = kh == 1 ? e_lo0 + ((2 * kk * LBO_E) >> 4)
                                                            : o_lo0 + (((kh == 2 ? 256u : 0u) + 2 * kk * LBO_O) >> 4);
                                mma_f16_ss_rt(acc, lo, a_hiw, b_lo, b_hi, idesc_main, first ? 0u : 1u);
                                if (SPLIT)
                                    mma_f16_ss<true>(acc + 3 * NB, lo + (S::PLANE_BYTES >> 4), a_hiw, b_lo, b_hi, idesc_hi3);
                            }
                        }
                    };
                    uint32_t slot = n_base % NSTAGE, phase = (n_base / NSTAGE) & 1;
                    for (int j = 0; j < nplanes; ++j) {
                        mbar_wait(&full[slot], phase);
                        tcgen05_fence_after();
                        const uint32_t a_stage = planes_addr + slot * S::STAGE_BYTES;
                        const bool opens = !(j & 1) && (j >> 1) < nout;          // this plane is kd=0 of a new output
                        if (opens) {
                            const uint32_t t = t_base + (j >> 1);
                            mbar_wait(&tempty[t & 1], ((t >> 1) & 1) ^ 1);
                            tcgen05_fence_after();
                        }
                        if (elect_one()) {
                            if (j & 1) {
                                issue_kd(a_stage, t_base + ((j - 1) >> 1), 1);
                            } else {
                                if (j >= 2) {
                                    const uint32_t t = t_base + (j >> 1) - 1;
                                    issue_kd(a_stage, t, 2);
                                    commit_one(&tfull[t & 1]);
                                }
                                if (opens) issue_kd(a_stage, t_base + (j >> 1), 0);
                            }
                            commit_one(&empty[slot]);
                        }
                        __syncwarp();
                        if (++slot == NSTAGE) { slot = 0; phase ^= 1; }
                    }
                    n_base += nplanes;
                    t_base += nout;
                } else if (KIND == 1) {
                    // plane-driven: input plane 2*od-1+kd feeds chain kd of output od
                    const int nplanes = 2 * nout + 1;
                    auto issue_kd = [&](uint32_t a_stage, uint32_t t, int kd) {
                        const uint32_t acc0 = tmem_base + (t & 1) * S::ACC_COLS;
                        const uint32_t acc = acc0 + kd * S::KD_COLS;
                        const uint32_t a_base = (a_stage & 0x3FFFF) >> 4;
                        const uint32_t b_lo_kd = b_lo0 + ((kd * 9 * S::TAP_BYTES) >> 4);
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                            for (int kw = 0; kw < 3; ++kw) {
                                const bool first = (kh == 0 && kw == 0);
                                const int ph = kh != 1, pw = kw != 1;
                                const int nh = 16 + ph, nw = 8 + pw;
                                const int sub_off = ph ? (pw ? Geo<1>::SUB_OFF3 : Geo<1>::SUB_OFF2)
                                                       : (pw ? Geo<1>::SUB_OFF1 : Geo<1>::SUB_OFF0);
                                const uint32_t lbo = nh * nw * 16, sbo = nw * 16;
                                const uint32_t off = sub_off + ((kh == 2 ? 1 : 0) * nw + (kw == 2 ? 1 : 0)) * 16;
                                const uint32_t lo = (a_base + (off >> 4)) | ((lbo >> 4) << 16);
                                const uint64_t db = desc_of(b_lo_kd + (((kh * 3 + kw) * S::TAP_BYTES) >> 4), b_hi);
                                tcgen05_mma_bf16(acc, desc_of(lo, desc_hi(sbo)), db, idesc_main, first ? 0u : 1u);
                                if (SPLIT)
                                    tcgen05_mma_bf16(acc0 + S::LH_COL, desc_of(lo + (S::PLANE_BYTES >> 4), desc_hi(sbo)), db,
                                                     idesc_lo, (kd == 0 && first) ? 0u : 1u);
                            }
                        }
                    };
                    for (int j = 0; j < nplanes; ++j) {
                        const uint32_t n = n_base + j;
                        const uint32_t slot = n % NSTAGE;
                        mbar_wait(&full[slot], (n / NSTAGE) & 1);
                        tcgen05_fence_after();
                        const uint32_t a_stage = planes_addr + slot * S::STAGE_BYTES;
                        if (j & 1) {
                            issue_kd(a_stage, t_base + ((j - 1) >> 1), 1);
                        } else {
                            if (j >= 2) {
                                const uint32_t t = t_base + (j >> 1) - 1;
                                issue_kd(a_stage, t, 2);
                                tcgen05_commit(&tfull[t & 1]);
                            }
                            if ((j >> 1) < nout) {
                                const uint32_t t = t_base + (j >> 1);
                                mbar_wait(&tempty[t & 1], ((t >> 1) & 1) ^ 1);
                                tcgen05_fence_after();
                                issue_kd(a_stage, t, 0);
                            }
                        }
                        tcgen05_commit(&empty[slot]);
                    }
                    n_base += nplanes;
                    t_base += nout;
                } else if (is_t64(KIND)) {
                    // transposed, K = 64, one class group per launch: per input depth plane NPH phases, each the two
                    // classes (rd, rh, rw = 0 | 1) accumulated in their own TMEM columns (double buffered over phases)
                    constexpr int GRP = KIND - 6;
                    constexpr int NPH = t64_nphases(GRP);
                    constexpr uint32_t LBO_A = 17 * 9 * 16, SBO_A = 9 * 16;
                    constexpr uint32_t a_hiw = desc_hi(SBO_A);
                    auto issue_phase = [&](uint32_t a_same, uint32_t a_next, int rd, int rh, uint32_t accp) {   // elected lane only
#pragma unroll
                        for (int rw = 0; rw < 2; ++rw) {
                            const uint32_t acc = accp + rw * S::CLS_COLS;
                            bool first = true;
#pragma unroll
                            for (int id = 0; id < 2; ++id) {
                                if (id > rd) continue;
                                // rd=0: (kd=1, this plane); rd=1: id0 = (kd=0, next plane), id1 = (kd=2, this plane)
                                const int kd = rd == 0 ? 1 : (id == 0 ? 0 : 2);
                                const uint32_t a_lo0 = desc_lo((rd == 1 && id == 0) ? a_next : a_same, LBO_A);
#pragma unroll
                                for (int ih = 0; ih < 2; ++ih) {
                                    if (ih > rh) continue;
                                    const int kh = rh == 0 ? 1 : (ih == 0 ? 0 : 2);
                                    const int offh = (rh == 1 && ih == 0) ? 1 : 0;
#pragma unroll
                                    for (int iw = 0; iw < 2; ++iw) {
                                        if (iw > rw) continue;
                                        const int kw = rw == 0 ? 1 : (iw == 0 ? 0 : 2);
                                        const int offw = (rw == 1 && iw == 0) ? 1 : 0;
                                        const int wslot_ = t64_slot(GRP, kd, kh, kw);
                                        const uint32_t a_off = (offh * 9 + offw) * 16;
#pragma unroll
                                        for (int kk = 0; kk < CBK / 2; ++kk) {
                                            const uint32_t b_lo = b_lo0 + ((wslot_ * S::TAP_BYTES + 2 * kk * (int)S::LBO_B) >> 4);
                                            mma_f16_ss_rt(acc, a_lo0 + ((a_off + 2 * kk * LBO_A) >> 4), a_hiw, b_lo, b_hi, idesc_main,
                                                          first ? 0u : 1u);
                                            first = false;
                                            if (SPLIT)
                                                mma_f16_ss<true>(acc, a_lo0 + ((S::PLANE_BYTES + a_off + 2 * kk * LBO_A) >> 4), a_hiw, b_lo,
                                                                 b_hi, idesc_lo);
                                        }
                                    }
                                }
                            }
                        }
                    };
                    uint32_t slot = n_base % NSTAGE, phase = (n_base / NSTAGE) & 1;
                    uint32_t wslot = slot, wphase = phase;
                    int waited = 0;
                    for (int od = 0; od < nout; ++od) {
                        while (waited < od + 2) {
                            mbar_wait(&full[wslot], wphase);
                            if (++wslot == NSTAGE) { wslot = 0; wphase ^= 1; }
                            ++waited;
                        }
                        const uint32_t s1 = slot + 1 >= NSTAGE ? slot + 1 - NSTAGE : slot + 1;
                        const uint32_t a_same = planes_addr + slot * S::STAGE_BYTES;
                        const uint32_t a_next = planes_addr + s1 * S::STAGE_BYTES;
#pragma unroll
                        for (int pi = 0; pi < NPH; ++pi) {
                            const uint32_t t = t_base + NPH * od + pi;
                            const uint32_t buf = t & 1;
                            mbar_wait(&tempty[buf], ((t >> 1) & 1) ^ 1);
                            tcgen05_fence_after();
                            if (elect_one()) {
                                issue_phase(a_same, a_next, t64_rd(GRP, pi), t64_rh(GRP, pi), tmem_base + buf * S::ACC_COLS);
                                commit_one(&tfull[buf]);
                                if (pi == NPH - 1) {
                                    commit_one(&empty[slot]);
                                    if (od == nout - 1) commit_one(&empty[s1]);
                                }
                            }
                            __syncwarp();
                        }
                        if (++slot == NSTAGE) { slot = 0; phase ^= 1; }
                    }
                    n_base += nout + 1;
                    t_base += NPH * nout;
                } else {
                    // transposed: per input depth qd two groups (output depth parity rd), 4 classes each
                    constexpr uint32_t LBO_A = 17 * 9 * 16, SBO_A = 9 * 16;
                    constexpr uint32_t a_hiw = desc_hi(SBO_A);
                    auto issue_group = [&](uint32_t a_same, uint32_t a_next, int rd) {        // elected lane only
                        const uint32_t accg = tmem_base + rd * S::ACC_COLS;
#pragma unroll
                        for (int cls = 0; cls < 4; ++cls) {
                            const int rh = cls >> 1, rw = cls & 1;
                            const uint32_t acc = accg + cls * S::CLS_COLS;
                            bool first = true;
#pragma unroll
                            for (int id = 0; id < 2; ++id) {
                                if (id > rd) continue;
                                // rd=0: (kd=1, this plane); rd=1: id0 = (kd=0, next plane), id1 = (kd=2, this plane)
                                const int kd = rd == 0 ? 1 : (id == 0 ? 0 : 2);
                                const uint32_t a_pl = (rd == 1 && id == 0) ? a_next : a_same;
                                const uint32_t a_lo0 = desc_lo(a_pl, LBO_A);
#pragma unroll
                                for (int ih = 0; ih < 2; ++ih) {
                                    if (ih > rh) continue;
                                    const int kh = rh == 0 ? 1 : (ih == 0 ? 0 : 2);
                                    const int offh = (rh == 1 && ih == 0) ? 1 : 0;
#pragma unroll
                                    for (int iw = 0; iw < 2; ++iw) {
                                        if (iw > rw) continue;
                                        const int kw = rw == 0 ? 1 : (iw == 0 ? 0 : 2);
                                        const int offw = (rw == 1 && iw == 0) ? 1 : 0;
                                        const int tap = (kd * 3 + kh) * 3 + kw;
                                        const uint32_t a_off = (offh * 9 + offw) * 16;
#pragma unroll
                                        for (int kk = 0; kk < CBK / 2; ++kk) {
                                            const uint32_t b_lo = b_lo0 + ((tap * S::TAP_BYTES + 2 * kk * (int)S::LBO_B) >> 4);
                                            mma_f16_ss_rt(acc, a_lo0 + ((a_off + 2 * kk * LBO_A) >> 4), a_hiw, b_lo, b_hi, idesc_main,
                                                          first ? 0u : 1u);
                                            first = false;
                                            if (SPLIT)
                                                mma_f16_ss<true>(acc, a_lo0 + ((S::PLANE_BYTES + a_off + 2 * kk * LBO_A) >> 4), a_hiw, b_lo,
                                                                 b_hi, idesc_lo);
                                        }
                                    }
                                }
                            }
                        }
                    };
                    uint32_t slot = n_base % NSTAGE, phase = (n_base / NSTAGE) & 1;
                    uint32_t wslot = slot, wphase = phase;
                    int waited = 0;
                    for (int od = 0; od < nout; ++od) {
                        while (waited < od + 2) {
                            mbar_wait(&full[wslot], wphase);
                            if (++wslot == NSTAGE) { wslot = 0; wphase ^= 1; }
                            ++waited;
                        }
                        const uint32_t s1 = slot + 1 >= NSTAGE ? slot + 1 - NSTAGE : slot + 1;
                        const uint32_t a_same = planes_addr + slot * S::STAGE_BYTES;
                        const uint32_t a_next = planes_addr + s1 * S::STAGE_BYTES;
#pragma unroll 1
                        for (int rd = 0; rd < 2; ++rd) {
                            const uint32_t t = t_base + 2 * od + rd;          // t_base is even: buffer == rd
                            mbar_wait(&tempty[rd], ((t >> 1) & 1) ^ 1);
                            tcgen05_fence_after();
                            if (elect_one()) {
                                if (rd == 0) issue_group(a_same, a_next, 0);
                                else issue_group(a_same, a_next, 1);
                                commit_one(&tfull[rd]);
                                if (rd == 1) {
                                    commit_one(&empty[slot]);
                                    if (od == nout - 1) commit_one(&empty[s1]);
                                }
                            }
                            __syncwarp();
                        }
                        if (++slot == NSTAGE) { slot = 0; phase ^= 1; }
                    }
                    n_base += nout + 1;
                    t_base += 2 * nout;
                }
            }
            if (HEAD && head_planes > 0) {         // the last plane's projections
                if (elect_one()) issue_head(head_planes - 1);
                __syncwarp();
            }
        }
        __syncwarp();
    } else {
        // ================================ epilogue =====================================
        const int q = warp & 3;                    // TMEM lane quarter this warp may access
        const int m = q * 32 + lane;
        // row / column of this thread's accumulator row inside the M tile
        const int hl = KIND == 3 ? (m >> 5) : (KIND == 4 ? (m >> 4) : (m >> 3));
        const int wl = KIND == 3 ? (m & 31) : (KIND == 4 ? (m & 15) : (m & 7));
        // KIND 3/4: two warps per TMEM lane quarter, each owns 16 of the 32 channels (fused head: of the activated tile;
        // the nine (kh, kw) Q sums are then split 5 / 4 between the two)
        constexpr bool HALVES = (KIND == 3 || KIND == 4);
        const int half = HALVES ? ((warp - 2) >> 2) : 0;
        float bias[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            const int cc = half * 16 + c;
            bias[c] = (p.bias && cc < p.n_valid_out && (!HALVES || c < 16)) ? __ldg(p.bias + cc) : 0.f;
        }
        uint32_t t = 0;
        uint32_t ntrace = 0;
        const bool tracer = (warp == 2 && lane == 0);
        ItemIter iter = items_begin(p);
        Item it;
        while (items_next<th_of(KIND), twstep_of(KIND)>(p, iter, it)) {
            const int h = it.h0 + hl;
            // KIND 3/4: tile columns are input columns w0-1 .. w0+14; KIND 4 keeps the even centres (ow = w/2)
            const int win = it.w0 - 1 + wl;
            const int w = KIND == 3 ? win : (KIND == 4 ? (win >> 1) : it.w0 + wl);
            const bool valid = KIND == 3 ? (h < p.Hm && wl >= 1 && wl <= K3_TWV && w < p.Wm)
                             : KIND == 4 ? (h < p.Hm && wl >= 1 && wl <= TWV && (win & 1) == 0 && win < p.Wm)
                                         : (h < p.Hm && w < p.Wm);
            float hq_prev[5], hq_cur[5];                   // fused head: running Q sums of this thread's (kh, kw) share
            for (int d = it.d0; d < it.d1; ++d) {
                if (KIND == 3 || KIND == 4) {
                    const uint32_t buf = t & 1;
                    if (!HEAD && p.res_hi && valid) {          // residual / in-place lines into L1 ahead of the accumulator
                        const size_t plane_sz = (size_t)p.Ho * p.Wo;
                        const size_t vox = (size_t)d * plane_sz + (size_t)h * p.Wo + w;
#pragma unroll
                        for (int cb = 0; cb < 2; ++cb) {
                            const size_t ri = ((size_t)(it.b * p.res_cbs + p.res_cb0 + half * 2 + cb) * p.Do) * plane_sz + vox;
                            prefetch_l1(p.res_hi + ri);
                            if (p.res_lo) prefetch_l1(p.res_lo + ri);
                        }
                    }
                    if (tracer) trace_stamp(p, 1, ntrace);                     // [0] epilogue ready for the plane
                    mbar_wait(&tfull[buf], (t >> 1) & 1);
                    tcgen05_fence_after();
                    if (tracer) trace_stamp(p, 1, ntrace);                     // [1] accumulator complete
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * S::ACC_COLS;
                    float v[NB];
                    // out(w) = D'[w-1][kw=0] + D'[w][kw=1] + D'[w+1][kw=2]; lanes of one tile row are adjacent
                    {
                        uint32_t r0[16], r1[16];
#pragma unroll
                        for (int c = 16; c < NB; ++c) v[c] = 0.f;
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            tmem_ld16(taddr + kw * NB + half * 16, r0);
                            if (SPLIT) tmem_ld16(taddr + 3 * NB + kw * NB + half * 16, r1);
                            tmem_ld_wait();
#pragma unroll
                            for (int c = 0; c < 16; ++c) {
                                float sv = __uint_as_float(r0[c]);
                                if (SPLIT) sv += __uint_as_float(r1[c]);
                                if (kw == 0) v[c] = __shfl_up_sync(0xffffffffu, sv, 1);
                                else if (kw == 1) v[c] += sv;
                                else v[c] = (v[c] + __shfl_down_sync(0xffffffffu, sv, 1)) * p.acc_scale;
                            }
                        }
                    }
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[buf]);
                    if (tracer) trace_stamp(p, 1, ntrace);                     // [2] TMEM drained, buffer released
                    ++t;
                    if constexpr (HEAD) {
                        // activated tile (this warp's 16 channels of its 32 voxels) -> second A operand: K-major core
                        // matrices, row m of channel block cb at cb * 2048 + m * 16; hi plane, then lo plane
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            v[c] += bias[c];
                            if (p.relu) v[c] = fmaxf(v[c], 0.f);
                        }
                        unsigned char* a2 = smem + S::HEAD_A_OFF + (half * 2) * (128 * 16) + m * 16;
#pragma unroll
                        for (int cb = 0; cb < 2; ++cb) {
                            float hi[8], lo[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const float x = v[cb * 8 + e];
                                hi[e] = round16<FP16>(x);
                                lo[e] = x - hi[e];
                            }
                            *reinterpret_cast<uint4*>(a2 + cb * (128 * 16)) =
                                make_uint4(pack2<FP16>(hi[0], hi[1]), pack2<FP16>(hi[2], hi[3]), pack2<FP16>(hi[4], hi[5]), pack2<FP16>(hi[6], hi[7]));
                            if (SPLIT)
                                *reinterpret_cast<uint4*>(a2 + S::HEAD_A_BYTES + cb * (128 * 16)) =
                                    make_uint4(pack2<FP16>(lo[0], lo[1]), pack2<FP16>(lo[2], lo[3]), pack2<FP16>(lo[4], lo[5]), pack2<FP16>(lo[6], lo[7]));
                        }
                        fence_proxy_async();                   // generic-proxy stores -> the tensor core's operand reads
                        __syncwarp();
                        if (lane == 0) mbar_arrive(a2full);
                        // the 27 projections of this thread's voxel, once the MMA warp has run the head MMAs of the plane
                        mbar_wait(pfull, (t - 1) & 1);
                        tcgen05_fence_after();
                        uint32_t pr[NB];
                        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + S::HEAD_P_COL, pr);
                        tmem_ld_wait();
                        tcgen05_fence_before();                // ordered before the next plane's a2full arrival
                        if (valid) {
                            // two warps per TMEM lane quarter: warps 2..5 own (kh, kw) 0..4, warps 6..9 own 5..8
                            if (warp < 6) head_step<0, 5>(p, pr, head_inv_scale, it, d, h, w, hq_prev, hq_cur);
                            else head_step<5, 9>(p, pr, head_inv_scale, it, d, h, w, hq_prev, hq_cur);
                        }
                    } else if (valid) {
                        if (!(p.y_f32 && half)) {              // single-channel output: the first half's warp writes it
                            store_voxel<FP16, 2>(p, v, bias, it.b, d, h, w, half * 2);
                        }
                    }
                    if (tracer) trace_stamp(p, 1, ntrace);                     // [3] stores issued
                } else if (is_t64(KIND)) {
                    constexpr int GRP = KIND - 6;
                    const int rw = (warp - 2) >> 2;            // warps 2..5: class rw = 0; warps 6..9: rw = 1
#pragma unroll
                    for (int pi = 0; pi < t64_nphases(GRP); ++pi, ++t) {
                        const int rd = t64_rd(GRP, pi), rh = t64_rh(GRP, pi);
                        const uint32_t buf = t & 1;
                        if (p.res_hi && valid) {               // residual lines into L1 while the phase's MMAs run
                            const size_t plane_sz = (size_t)p.Ho * p.Wo;
                            const size_t vox = (size_t)(2 * d + rd) * plane_sz + (size_t)(2 * h + rh) * p.Wo + 2 * w + rw;
#pragma unroll
                            for (int cb = 0; cb < 4; ++cb) {
                                const size_t ri = ((size_t)(it.b * p.res_cbs + p.res_cb0 + cb) * p.Do) * plane_sz + vox;
                                prefetch_l1(p.res_hi + ri);
                                if (p.res_lo) prefetch_l1(p.res_lo + ri);
                            }
                        }
                        mbar_wait(&tfull[buf], (t >> 1) & 1);
                        tcgen05_fence_after();
                        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * S::ACC_COLS + rw * S::CLS_COLS;
                        uint32_t r0[32];
                        float v[NB];
                        tmem_ld32(taddr, r0);
                        if (SPLIT) {
                            uint32_t r1[32];
                            tmem_ld32(taddr + 32, r1);
                            tmem_ld_wait();
#pragma unroll
                            for (int c = 0; c < NB; ++c) v[c] = (__uint_as_float(r0[c]) + __uint_as_float(r1[c])) * p.acc_scale;
                        } else {
                            tmem_ld_wait();
#pragma unroll
                            for (int c = 0; c < NB; ++c) v[c] = __uint_as_float(r0[c]) * p.acc_scale;
                        }
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty[buf]);
                        if (valid) store_voxel<FP16, 4>(p, v, bias, it.b, 2 * d + rd, 2 * h + rh, 2 * w + rw);
                    }
                } else if (KIND != 2 && KIND != 5) {
                    const uint32_t buf = t & 1;
                    mbar_wait(&tfull[buf], (t >> 1) & 1);
                    tcgen05_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * S::ACC_COLS;
                    uint32_t r0[32], r1[32];
                    float v[NB];
                    if (SPLIT) {
                        // small terms first: hi*Wlo of the three kd chains + lo*Whi, then the hi*Whi chains
                        float sm[NB];
                        tmem_ld32(taddr + 32, r0);
                        tmem_ld32(taddr + 64 + 32, r1);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < NB; ++c) sm[c] = __uint_as_float(r0[c]) + __uint_as_float(r1[c]);
                        tmem_ld32(taddr + 128 + 32, r0);
                        tmem_ld32(taddr + S::LH_COL, r1);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < NB; ++c) sm[c] += __uint_as_float(r0[c]) + __uint_as_float(r1[c]);
                        tmem_ld32(taddr, r0);
                        tmem_ld32(taddr + 64, r1);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < NB; ++c) v[c] = __uint_as_float(r0[c]) + __uint_as_float(r1[c]);
                        tmem_ld32(taddr + 128, r0);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < NB; ++c) v[c] = ((v[c] + __uint_as_float(r0[c])) + sm[c]) * p.acc_scale;
                    } else {
                        tmem_ld32(taddr, r0);
                        tmem_ld32(taddr + 32, r1);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < NB; ++c) v[c] = __uint_as_float(r0[c]) + __uint_as_float(r1[c]);
                        tmem_ld32(taddr + 64, r0);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < NB; ++c) v[c] = (v[c] + __uint_as_float(r0[c])) * p.acc_scale;
                    }
                    // the accumulator buffer is free as soon as it sits in registers
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[buf]);
                    ++t;
                    if (valid) store_voxel<FP16>(p, v, bias, it.b, d, h, w);
                } else {
#pragma unroll 1
                    for (int rd = 0; rd < 2; ++rd, ++t) {
                        const int cls0 = (warp - 2) >> 2;          // warps 2..5: classes 0,2; warps 6..9: classes 1,3
                        // The residual (or, on an accumulating pass, the output itself) of this warp's two classes is
                        // pulled into L1 while the MMAs of the tile are still running: the epilogue of the transposed
                        // kernels was bound by the DRAM latency of these loads (one exposed round trip per class).
                        if (p.res_hi && valid) {
                            const size_t plane_sz = (size_t)p.Ho * p.Wo;
#pragma unroll
                            for (int ci = 0; ci < 2; ++ci) {
                                const int cls = cls0 + 2 * ci;
                                const size_t vox = (size_t)(2 * d + rd) * plane_sz + (size_t)(2 * h + (cls >> 1)) * p.Wo + 2 * w + (cls & 1);
#pragma unroll
                                for (int cb = 0; cb < (KIND == 5 ? 2 : 4); ++cb) {
                                    const size_t ri = ((size_t)(it.b * p.res_cbs + p.res_cb0 + cb) * p.Do) * plane_sz + vox;
                                    prefetch_l1(p.res_hi + ri);
                                    if (p.res_lo) prefetch_l1(p.res_lo + ri);
                                }
                            }
                        }
                        mbar_wait(&tfull[rd], (t >> 1) & 1);
                        tcgen05_fence_after();
                        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + rd * S::ACC_COLS;
#pragma unroll 1
                        for (int cls = cls0; cls < 4; cls += 2) {
                            uint32_t r0[32];
                            float v[NB];
                            tmem_ld32(taddr + cls * S::CLS_COLS, r0);
                            if (SPLIT && KIND == 5) {
                                // 16 output channels: [hi*Whi + lo*Whi | hi*Wlo] sit in one 32-column load
                                tmem_ld_wait();
#pragma unroll
                                for (int c = 0; c < NB; ++c)
                                    v[c] = c < 16 ? (__uint_as_float(r0[c]) + __uint_as_float(r0[(c & 15) + 16])) * p.acc_scale : 0.f;
                            } else if (SPLIT) {
                                uint32_t r1[32];
                                tmem_ld32(taddr + cls * S::CLS_COLS + 32, r1);
                                tmem_ld_wait();
#pragma unroll
                                for (int c = 0; c < NB; ++c) v[c] = (__uint_as_float(r0[c]) + __uint_as_float(r1[c])) * p.acc_scale;
                            } else {
                                tmem_ld_wait();
#pragma unroll
                                for (int c = 0; c < NB; ++c) v[c] = __uint_as_float(r0[c]) * p.acc_scale;
                            }
                            if (cls >= 2) {
                                tcgen05_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(&tempty[rd]);
                            }
                            if (valid) store_voxel<FP16, (KIND == 5 ? 2 : 4)>(p, v, bias, it.b, 2 * d + rd, 2 * h + (cls >> 1), 2 * w + (cls & 1));
                        }
                    }
                }
            }
            if constexpr (HEAD) {                          // last plane of the segment: Q(d1 - 1) and the upward spill
                if (valid) {
                    if (warp < 6) head_finish<0, 5>(p, it, h, w, hq_prev, hq_cur);
                    else head_finish<5, 9>(p, it, h, w, hq_prev, hq_cur);
                }
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- weight packing ---------------------------------------------------------------------------
// w: [27][Cin][Cout] fp32 -> blobs[(ob*IB + ib)] = [27][cbk][ROWS][8] 16-bit, ROWS = 32 (plain) or 64 (hi rows, lo
// rows); one blob covers 8*cbk input channels x 32 output channels
// nkw = 3 (KIND 3): a block covers one (kd,kh) tap row, rows = [hi kw0|hi kw1|hi kw2|lo kw0|lo kw1|lo kw2]
__global__ void pack_weights_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int Cin, int Cout, int cbk,
                                    int nkw, int split, int fp16, float scale, int nb) {
    const int hi_rows = nb * nkw;
    const int rows = split ? 2 * hi_rows : hi_rows;
    const int nblk = TAPS / nkw;
    const int IB = Cin / (8 * cbk), OB = (Cout + nb - 1) / nb;
    const size_t blob = (size_t)nblk * cbk * rows * 8;
    const size_t total = blob * IB * OB;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int e = r % 8; r /= 8;
        const int row = r % rows; r /= rows;
        const int cb = r % cbk; r /= cbk;
        const int blk = r % nblk; r /= nblk;
        const int ib = r % IB;
        const int ob = r / IB;
        const bool is_lo = row >= hi_rows;
        const int rr = row % hi_rows;
        const int tap = blk * nkw + rr / nb;
        const int co = ob * nb + (rr % nb);
        const int ci = (ib * cbk + cb) * 8 + e;
        const float v = (co < Cout) ? w[((size_t)tap * Cin + ci) * Cout + co] * scale : 0.f;
        if (fp16) {
            const __half hi = __float2half_rn(v);
            const __half lo = __float2half_rn(v - __half2float(hi));
            out[i] = is_lo ? __half_as_ushort(lo) : __half_as_ushort(hi);
        } else {
            __nv_bfloat16 hi, lo;
            split_bf16(v, hi, lo);
            out[i] = is_lo ? __bfloat16_as_ushort(lo) : __bfloat16_as_ushort(hi);
        }
    }
}

// ---- host side -----------------------------------------------------------------------------
// dense map over [B][CBS][D][H][W][8]: dims (w*8, h, d, cb, b), box (bw*8, bh, 1, cbk, 1)
static int make_dense_map(CUtensorMap* map, const void* base, int B, int CBS, int D, int H, int W, int bh, int bw, int cbk,
                          int fp16) {
    const cuuint64_t dims[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)CBS, (cuuint64_t)B};
    const cuuint64_t strides[4] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16,
                                   (cuuint64_t)CBS * D * H * W * 16};
    const cuuint32_t box[5] = {(cuuint32_t)bw * 8, (cuuint32_t)bh, 1, (cuuint32_t)cbk, 1};
    return encode(map, base, dims, strides, box, fp16);
}

// parity map (ph,pw) over one batch element: dims (8, W/2, H/2, D, cb), element (c8, w2, h2, d, cb) lives at
// base + ((cb*D + d)*H + 2*h2+ph)*W*16 + (2*w2+pw)*16
static int make_parity_map(CUtensorMap* map, const void* base_b, int CBS, int D, int H, int W, int ph, int pw, int cbk,
                           int fp16) {
    const unsigned char* base = reinterpret_cast<const unsigned char*>(base_b) + ((size_t)ph * W + pw) * 16;
    const cuuint64_t dims[5] = {8, (cuuint64_t)W / 2, (cuuint64_t)H / 2, (cuuint64_t)D, (cuuint64_t)CBS};
    const cuuint64_t strides[4] = {32, (cuuint64_t)2 * W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
    const cuuint32_t box[5] = {8, (cuuint32_t)(8 + pw), (cuuint32_t)(16 + ph), 1, (cuuint32_t)cbk};
    return encode(map, base, dims, strides, box, fp16);
}

template <int KIND, bool SPLIT, bool FP16>
static int launch_pass(const Maps& maps, const Params& p, int grid, void* stream) {
    const size_t smem = Smem<KIND, SPLIT>::TOTAL;
    DMB_CUDA(cudaFuncSetAttribute(conv3d_tc_kernel<KIND, SPLIT, FP16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3d_tc_kernel<KIND, SPLIT, FP16, false><<<grid, nthreads_of(KIND), smem, as_stream(stream)>>>(maps, p);
    return check_launch("conv3d_tc_kernel");
}

template <bool SPLIT, bool FP16>
static int launch_head(const Maps& maps, const Params& p, int grid, void* stream) {
    const size_t smem = Smem<3, SPLIT>::TOTAL_HEAD;
    DMB_CUDA(cudaFuncSetAttribute(conv3d_tc_kernel<3, SPLIT, FP16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3d_tc_kernel<3, SPLIT, FP16, true><<<grid, nthreads_of(3, true), smem, as_stream(stream)>>>(maps, p);
    return check_launch("conv3d_tc_kernel<head>");
}

template <int KIND>
static int launch_kind(const Maps& maps, const Params& p, int grid, bool split, int fp16, void* stream) {
    if (split)
        return fp16 ? launch_pass<KIND, true, true>(maps, p, grid, stream) : launch_pass<KIND, true, false>(maps, p, grid, stream);
    return fp16 ? launch_pass<KIND, false, true>(maps, p, grid, stream) : launch_pass<KIND, false, false>(maps, p, grid, stream);
}

static int cbk_of(int kind) { return kind == 1 ? Geo<1>::CBK : (kind >= 5 ? Geo<5>::CBK : 4); }
static int nkw_of(int kind) { return (kind == 3 || kind == 4) ? 3 : 1; }

// row-parity map (ph) over the whole batch: dims (w*8, H/2, D, cb, b); row h2 of parity ph is input row 2*h2+ph
static int make_rowparity_map(CUtensorMap* map, const void* base, int B, int CBS, int D, int H, int W, int ph, int rows,
                              int fp16) {
    const unsigned char* b0 = reinterpret_cast<const unsigned char*>(base) + (size_t)ph * W * 16;
    const cuuint64_t dims[5] = {(cuuint64_t)W * 8, (cuuint64_t)H / 2, (cuuint64_t)D, (cuuint64_t)CBS, (cuuint64_t)B};
    const cuuint64_t strides[4] = {(cuuint64_t)2 * W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16,
                                   (cuuint64_t)CBS * D * H * W * 16};
    const cuuint32_t box[5] = {(cuuint32_t)TW3 * 8, (cuuint32_t)rows, 1, 4, 1};
    return encode(map, b0, dims, strides, box, fp16);
}

}  // namespace tc
}  // namespace dmb

using namespace dmb;
using namespace dmb::tc;

extern "C" int dmb_b200_conv3d_tc_available(void) { return device_ok(); }

extern "C" int64_t dmb_b200_conv3d_tc_weight_bytes(int Cin, int Cout, int split, int kind) {
    if (Cin <= 0 || Cout <= 0 || Cin % 32 || kind < 0 || kind > 6) return 0;
    const int cbk = cbk_of(kind), nb = nbo_of(kind);
    if (Cin % (8 * cbk)) return 0;
    const int64_t blob = (int64_t)TAPS * cbk * (split ? 2 * nb : nb) * 16;
    return blob * (Cin / (8 * cbk)) * ((Cout + nb - 1) / nb);
}

extern "C" int dmb_b200_conv3d_tc_pack_weights(const float* w_packed, void* w_blob, int Cin, int Cout, int split,
                                               int fp16, float scale, int kind, void* stream) {
    DMB_REQUIRE(w_packed && w_blob, "conv3d_tc_pack_weights: null pointer");
    DMB_REQUIRE(kind >= 0 && kind <= 6, "conv3d_tc_pack_weights: kind must be 0..6");
    DMB_REQUIRE(kind != 5 || (Cin % 64 == 0 && Cout % 16 == 0), "conv3d_tc_pack_weights: kind 5 needs Cin %% 64 == 0 and Cout %% 16 == 0");
    DMB_REQUIRE(kind != 6 || (Cin == 64 && Cout % 32 == 0), "conv3d_tc_pack_weights: kind 6 needs Cin == 64 and Cout %% 32 == 0");
    DMB_REQUIRE(Cin > 0 && Cin % 32 == 0, "conv3d_tc_pack_weights: Cin=%d must be a multiple of 32", Cin);
    DMB_REQUIRE(Cout > 0 && (Cout % 32 == 0 || Cout < 32), "conv3d_tc_pack_weights: Cout=%d must be <32 or a multiple of 32", Cout);
    DMB_REQUIRE(scale > 0.f, "conv3d_tc_pack_weights: scale must be positive");
    const int64_t n = dmb_b200_conv3d_tc_weight_bytes(Cin, Cout, split, kind) / 2;
    pack_weights_kernel<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(w_packed, (uint16_t*)w_blob, Cin, Cout,
                                                                                cbk_of(kind), nkw_of(kind), split ? 1 : 0, fp16 ? 1 : 0, scale,
                                                                                nbo_of(kind));
    return check_launch("pack_weights_kernel");
}

// Geometry and static schedule of one launch: M-space / output extents, tile counts, depth segmentation.
// Returns the grid size (persistent CTAs).  Pure host arithmetic (dmb_b200_conv3d_tc_schedule exposes it to the
// CPU tests).
static int plan_schedule(Params& p, int kind, int B, int D, int H, int W, bool head = false) {
    p.B = (kind == 1) ? 1 : B;
    const bool s2 = (kind == 1 || kind == 4);
    // KIND 4 tiles W in INPUT columns (every column is computed, even centres are kept)
    p.Dm = s2 ? D / 2 : D; p.Hm = s2 ? H / 2 : H; p.Wm = kind == 1 ? W / 2 : W;
    const bool same = (kind == 0 || kind == 3);
    p.Do = same ? D : (s2 ? D / 2 : 2 * D);
    p.Ho = same ? H : (s2 ? H / 2 : 2 * H);
    p.Wo = same ? W : (s2 ? W / 2 : 2 * W);
    p.tiles_h = (int)cdiv(p.Hm, th_of(kind));
    p.tiles_w = (int)cdiv(p.Wm, twstep_of(kind));
    // depth segments: ~8 work items per persistent CTA; small grids (the 1/8 and 1/16 levels of the
    // hourglass) are cut down to 2-plane segments so that every SM gets work (halo planes are L2 hits)
    const int cols = p.tiles_h * p.tiles_w * p.B;
    static int items_per_sm = -1;     // DMB_B200_TC_ITEMS_PER_SM > 0: the old fixed rule (~that many items per CTA)
    if (items_per_sm < 0) {
        const char* e = getenv("DMB_B200_TC_ITEMS_PER_SM");
        items_per_sm = e ? atoi(e) : 0;
        if (items_per_sm < 0 || items_per_sm > 64) items_per_sm = 0;
    }
    int nseg;
    if (items_per_sm > 0) {
        nseg = (int)cdiv((int64_t)sm_count() * items_per_sm, cols);
    } else {
        // static schedule: the slowest CTA runs ceil(items / SMs) items of seg_len planes each, every item costing
        // about one extra plane (halo loads, pipeline restart).  Take the segment count that minimises that.
        int64_t best = -1;
        nseg = 1;
        for (int c = 1; c <= (p.Dm + 1) / 2; ++c) {
            const int sl = (int)cdiv(p.Dm, c), ns = (int)cdiv(p.Dm, sl);
            const int64_t cost = cdiv((int64_t)cols * ns, sm_count()) * (sl + 1);
            if (best < 0 || cost < best) { best = cost; nseg = ns; }
        }
    }
    if (nseg > p.Dm / 2) nseg = p.Dm / 2;
    if (nseg < 1) nseg = 1;
    p.seg_len = (int)cdiv(p.Dm, nseg);
    p.nseg = (int)cdiv(p.Dm, p.seg_len);
    p.n_items = cols * p.nseg;
    p.balanced = 0;
    p.planes_total = cols * p.Dm;
    // DMB_B200_TC_BALANCED=1 selects the balanced plane schedule (ItemIter).  OFF by default: measured on the PSMNet
    // trunk it is 2 % SLOWER (aggregator 6.98-7.05 ms against 6.89 ms on one box) although the static schedule leaves the
    // SMs of the small hourglass layers idle a third of the time -- the trunk is power bound, the idle SMs' share of the
    // power budget clocks the busy ones higher, and the static schedule's wave-like order shares halo planes in L2.
    static int balanced_on = -1;
    if (balanced_on < 0) {
        const char* e = getenv("DMB_B200_TC_BALANCED");
        balanced_on = (e && e[0] == '1') ? 1 : 0;
    }
    if (balanced_on && !head && (int64_t)cols * p.Dm < ((int64_t)1 << 30)) {
        // (the fused head's spill planes are indexed by the uniform segments: it stays on the static schedule)
        p.balanced = 1;
        const int g = (int)std::min<int64_t>(sm_count(), std::max<int64_t>(1, p.planes_total / 2));
        return g;
    }
    return p.n_items < sm_count() ? p.n_items : sm_count();
}

static long long* g_trace = nullptr;
extern "C" int dmb_b200_debug_set_trace(long long* device_buffer) {   // 3 roles x 4096 stamps, or NULL to stop
    g_trace = device_buffer;
    return DMB_OK;
}

static int conv3d_tc_impl(const void* x_hi, const void* x_lo, int Cin, const void* w_blob, float w_scale,
                          const float* bias, const void* res_hi, const void* res_lo, void* y_hi, void* y_lo,
                          int Cout, float* y_f32, const float* res_f32, int B, int D, int H, int W, int kind,
                          int relu, int fp16, const float* head_w, float* head_t, void* stream, int flat = 0) {
    DMB_REQUIRE(x_hi && w_blob, "conv3d_tc: null input/weights");
    DMB_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "conv3d_tc: non-positive dimension");
    DMB_REQUIRE(kind >= 0 && kind <= 6, "conv3d_tc: kind must be 0/3 (stride 1), 1/4 (stride 2) or 2/5/6 (transposed stride 2)");
    DMB_REQUIRE(kind != 6 || (Cin == 64 && Cout % 32 == 0), "conv3d_tc: kind 6 needs Cin == 64 and Cout %% 32 == 0");
    DMB_REQUIRE(Cin > 0 && Cin % 32 == 0, "conv3d_tc: Cin=%d must be a multiple of 32", Cin);
    DMB_REQUIRE(kind != 5 || (Cin % 64 == 0 && Cout % 32 == 0), "conv3d_tc: kind 5 needs Cin %% 64 == 0 and Cout %% 32 == 0");
    DMB_REQUIRE(w_scale > 0.f, "conv3d_tc: w_scale must be positive");
    const bool scalar_out = (Cout == 1);
    DMB_REQUIRE(scalar_out || (Cout > 0 && Cout % 32 == 0), "conv3d_tc: Cout=%d must be 1 or a multiple of 32", Cout);
    const bool head = head_t != nullptr;
    if (head)
        DMB_REQUIRE(head_w && kind == 3 && Cin == 32 && Cout == 32 && !y_hi && !y_lo && !y_f32 && !res_hi && !res_f32,
                    "conv3d_tc_head: needs a 32->32 stride-1 layer (kind 3), no residual, no activation output");
    else if (scalar_out)
        DMB_REQUIRE(y_f32 && !y_hi, "conv3d_tc: Cout==1 writes y_f32 only");
    else
        DMB_REQUIRE(y_hi && !y_f32 && !res_f32, "conv3d_tc: Cout>=32 writes the blocked 16-bit output");
    DMB_REQUIRE(!res_lo || res_hi, "conv3d_tc: res_lo without res_hi");
    if (kind == 1 || kind == 4)
        DMB_REQUIRE(D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "conv3d_tc: stride-2 needs even input extents");
    if (!device_ok()) return fail(DMB_ERR_UNSUPPORTED, "conv3d_tc: needs an sm_100 device and a TMA-capable driver");
    const bool split = x_lo != nullptr;
    DMB_REQUIRE(head || scalar_out || split == (y_lo != nullptr), "conv3d_tc: x_lo and y_lo must both be given or both be NULL");

    const int cbk = cbk_of(kind), nbo = nbo_of(kind);
    const int IB = Cin / (8 * cbk), OB = scalar_out ? 1 : Cout / nbo;
    const int CBS = Cin / 8;

    DMB_REQUIRE(!flat || (kind == 3 && D == 1 && !head_t), "conv2d_tc: the flat variant is the stride-1 kernel on a single plane");
    Params p;
    p.trace = g_trace;
    p.flat = flat;
    p.head_w = head_w;
    p.head_t = head_t;
    const int grid = plan_schedule(p, kind, B, D, H, W, head_t != nullptr);
    p.n_valid_out = scalar_out ? 1 : nbo;
    p.acc_scale = 1.0f / w_scale;
    const size_t blob = (size_t)TAPS * cbk * (split ? 2 * nbo : nbo) * 16;
    const size_t in_batch_bytes = (size_t)CBS * D * H * W * 16;
    const size_t out_cbs = scalar_out ? 0 : Cout / 8;

    const int nb_outer = (kind == 1) ? B : 1;       // the parity maps cover one batch element each
    for (int bo = 0; bo < nb_outer; ++bo) {
        Maps maps;
        int rc;
        const unsigned char* xh = reinterpret_cast<const unsigned char*>(x_hi) + bo * in_batch_bytes;
        const unsigned char* xl = split ? reinterpret_cast<const unsigned char*>(x_lo) + bo * in_batch_bytes : xh;
        if (kind == 1) {
            for (int sub = 0; sub < 4; ++sub) {
                rc = make_parity_map(&maps.m[sub * 2], xh, CBS, D, H, W, sub >> 1, sub & 1, cbk, fp16);
                if (rc) return rc;
                rc = make_parity_map(&maps.m[sub * 2 + 1], xl, CBS, D, H, W, sub >> 1, sub & 1, cbk, fp16);
                if (rc) return rc;
            }
        } else if (kind == 4) {
            rc = make_rowparity_map(&maps.m[0], xh, B, CBS, D, H, W, 0, 8, fp16);
            if (rc) return rc;
            rc = make_rowparity_map(&maps.m[1], xl, B, CBS, D, H, W, 0, 8, fp16);
            if (rc) return rc;
            rc = make_rowparity_map(&maps.m[2], xh, B, CBS, D, H, W, 1, 9, fp16);
            if (rc) return rc;
            rc = make_rowparity_map(&maps.m[3], xl, B, CBS, D, H, W, 1, 9, fp16);
            if (rc) return rc;
            for (int i = 4; i < 8; ++i) maps.m[i] = maps.m[0];
        } else {
            const int bh = is_transposed(kind) ? 17 : (kind == 3 ? K3_TH + 2 : 18), bw = kind == 0 ? 10 : (kind == 3 ? K3_TW : 9);
            rc = make_dense_map(&maps.m[0], xh, B, CBS, D, H, W, bh, bw, cbk, fp16);
            if (rc) return rc;
            rc = make_dense_map(&maps.m[1], xl, B, CBS, D, H, W, bh, bw, cbk, fp16);
            if (rc) return rc;
            for (int i = 2; i < 8; ++i) maps.m[i] = maps.m[0];
        }
        // per-batch output base offsets when the kernel sees a single batch element
        const size_t out_plane = (size_t)p.Do * p.Ho * p.Wo;
        const size_t yb = (size_t)bo * out_cbs * out_plane;      // in uint4 units (16-byte voxel-block)
        for (int ob = 0; ob < OB; ++ob) {
            for (int ib = 0; ib < IB; ++ib) {
                const bool first = ib == 0, last = ib == IB - 1;
                p.w_blob = reinterpret_cast<const unsigned char*>(w_blob) + (size_t)(ob * IB + ib) * blob;
                p.bias = (first && bias) ? bias + ob * nbo : nullptr;
                p.in_cb0 = ib * cbk;
                p.relu = (last && relu) ? 1 : 0;
                p.y_hi = y_hi ? reinterpret_cast<uint4*>(y_hi) + yb : nullptr;
                p.y_lo = y_lo ? reinterpret_cast<uint4*>(y_lo) + yb : nullptr;
                p.y_cb0 = ob * (nbo / 8);
                p.y_cbs = (int)out_cbs;
                p.y_f32 = y_f32 ? y_f32 + (size_t)bo * out_plane : nullptr;
                p.res_cb0 = ob * (nbo / 8);
                p.res_cbs = p.y_cbs;
                if (first) {                       // the external residual joins on the first pass
                    p.res_hi = res_hi ? reinterpret_cast<const uint4*>(res_hi) + yb : nullptr;
                    p.res_lo = res_lo ? reinterpret_cast<const uint4*>(res_lo) + yb : nullptr;
                    p.res_f32 = res_f32 ? res_f32 + (size_t)bo * out_plane : nullptr;
                } else {                           // later passes accumulate onto the output in place
                    p.res_hi = p.y_hi;
                    p.res_lo = p.y_lo;
                    p.res_f32 = p.y_f32;
                }
                if (head) {
                    rc = split ? (fp16 ? launch_head<true, true>(maps, p, grid, stream)
                                       : launch_head<true, false>(maps, p, grid, stream))
                               : (fp16 ? launch_head<false, true>(maps, p, grid, stream)
                                       : launch_head<false, false>(maps, p, grid, stream));
                } else if (kind == 0) rc = launch_kind<0>(maps, p, grid, split, fp16, stream);
                else if (kind == 1) rc = launch_kind<1>(maps, p, grid, split, fp16, stream);
                else if (kind == 2) rc = launch_kind<2>(maps, p, grid, split, fp16, stream);
                else if (kind == 3) rc = launch_kind<3>(maps, p, grid, split, fp16, stream);
                else if (kind == 5) rc = launch_kind<5>(maps, p, grid, split, fp16, stream);
                else if (kind == 6) {              // three class-group launches, each writing its own output voxels once
                    rc = launch_kind<6>(maps, p, grid, split, fp16, stream);
                    if (!rc) rc = launch_kind<7>(maps, p, grid, split, fp16, stream);
                    if (!rc) rc = launch_kind<8>(maps, p, grid, split, fp16, stream);
                } else rc = launch_kind<4>(maps, p, grid, split, fp16, stream);
                if (rc) return rc;
            }
        }
    }
    return DMB_OK;
}

extern "C" int dmb_b200_conv3d_tc(const void* x_hi, const void* x_lo, int Cin, const void* w_blob, float w_scale,
                                  const float* bias, const void* res_hi, const void* res_lo, void* y_hi, void* y_lo,
                                  int Cout, float* y_f32, const float* res_f32, int B, int D, int H, int W, int kind,
                                  int relu, int fp16, void* stream) {
    return conv3d_tc_impl(x_hi, x_lo, Cin, w_blob, w_scale, bias, res_hi, res_lo, y_hi, y_lo, Cout, y_f32, res_f32, B, D, H,
                          W, kind, relu, fp16, nullptr, nullptr, stream);
}

// 2-D 3x3 / stride 1 / pad 1 convolution (the confidence heads of dmb/modeling/stereo/cmn/cmn.py:29-32:
// conv_bn_relu(192, 64) on a [B,192,H,W] cost volume) as the stride-1 tcgen05 kernel on ONE depth plane: activations
// blocked [B][Cin/8][1][H][W][8], weights packed as a 3-D [27][Cin][Cout] tensor whose kd = 0 / 2 taps are zero
// (pack with kind 3) -- the kernel skips them (9 instead of 27 taps issued per tile).
extern "C" int dmb_b200_conv2d_tc(const void* x_hi, const void* x_lo, int Cin, const void* w_blob, float w_scale,
                                  const float* bias, const void* res_hi, const void* res_lo, void* y_hi, void* y_lo,
                                  int Cout, int B, int H, int W, int relu, int fp16, void* stream) {
    DMB_REQUIRE(Cout > 0 && Cout % 32 == 0, "conv2d_tc: Cout=%d must be a multiple of 32", Cout);
    return conv3d_tc_impl(x_hi, x_lo, Cin, w_blob, w_scale, bias, res_hi, res_lo, y_hi, y_lo, Cout, nullptr, nullptr, B, 1, H,
                          W, 3, relu, fp16, nullptr, nullptr, stream, 1);
}

extern "C" int dmb_b200_conv3d_tc_schedule(int kind, int B, int D, int H, int W, int* out) {
    DMB_REQUIRE(out, "conv3d_tc_schedule: null output");
    DMB_REQUIRE(kind >= 0 && kind <= 6 && B > 0 && D > 0 && H > 0 && W > 0, "conv3d_tc_schedule: bad arguments");
    Params p;
    const int grid = plan_schedule(p, kind, B, D, H, W);
    out[0] = p.tiles_h; out[1] = p.tiles_w; out[2] = p.nseg; out[3] = p.seg_len; out[4] = p.n_items; out[5] = grid;
    out[6] = th_of(kind); out[7] = twstep_of(kind);
    return DMB_OK;
}

extern "C" int dmb_b200_conv3d_tc_head(const void* x_hi, const void* x_lo, const void* w_blob, float w_scale,
                                       const float* bias, const float* head_w, float* head_t, int B, int D, int H, int W,
                                       int relu, int fp16, void* stream) {
    DMB_REQUIRE(head_w && head_t, "conv3d_tc_head: null head weights / tap buffer");
    return conv3d_tc_impl(x_hi, x_lo, 32, w_blob, w_scale, bias, nullptr, nullptr, nullptr, nullptr, 32, nullptr, nullptr, B,
                          D, H, W, 3, relu, fp16, head_w, head_t, stream);
}

extern "C" int64_t dmb_b200_conv3d_tc_head_floats(int B, int D, int H, int W) {
    // floats of the head_t buffer that dmb_b200_conv3d_tc_head writes and dmb_b200_head_gather reads:
    // per batch element 9 Q planes per depth + 2 x 9 spill planes per depth segment of the launch's schedule
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
    Params p;
    plan_schedule(p, 3, B, D, H, W, true);
    return (int64_t)B * ((int64_t)9 * D + (int64_t)18 * p.nseg) * H * W;
}

extern "C" int dmb_b200_head_gather(const float* head_t, const float* res, float* y, int B, int D, int H, int W,
                                    void* stream) {
    DMB_REQUIRE(head_t && y, "head_gather: null pointer");
    DMB_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "head_gather: non-positive dimension");
    Params p;                                      // the same static schedule the head launch used
    plan_schedule(p, 3, B, D, H, W, true);
    DMB_REQUIRE(D <= 65535 && B <= 65535 && (int64_t)H * W < (int64_t)1 << 30, "head_gather: dimension too large");
    const dim3 grid((unsigned)cdiv((int64_t)H * W, 256), (unsigned)D, (unsigned)B);
    head_gather_kernel<<<grid, 256, 0, as_stream(stream)>>>(head_t, res, y, B, D, H, W, p.seg_len, p.nseg);
    return check_launch("head_gather_kernel");
}
