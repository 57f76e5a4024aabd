// Small-vector exchange between the GPUs of one box over peer memory (NVLink / NVSwitch P2P stores), fused into ONE
// kernel per exchange: the synchronised-BatchNorm statistics of the data-parallel training step
// (dmb/apis/train.py:95-97 converts the model to SyncBN; every BatchNorm layer then exchanges 2*C numbers per
// direction).  Roughly 300 such exchanges per step sit on the critical path; as NCCL collectives each costs a kernel
// launch on a side stream, two event hops and ~100 us of host-side bookkeeping -- and torch's own SyncBatchNorm adds a
// host synchronisation per layer (measured: +40 ms on a 56 ms step).  Here every rank writes its vector straight into
// every peer's receive buffer, publishes a sequence number with a system-scope release store, spins (acquire loads) until
// all peers' numbers have arrived in ITS OWN buffer, and reduces / gathers locally in a fixed rank order, so all ranks
// obtain bit-identical results.  No NCCL, no host involvement, the caller's stream.
//
// Buffers are plain cudaMalloc allocations shared through CUDA IPC handles (one process per GPU, like the reference's
// launcher tools/dist_train.sh); the handles travel through torch.distributed once at start-up (utils/dist_utils.py).
// Layout of a rank's buffer: NSLOT x MAXR regions of SLOT_BYTES, then NSLOT x MAXR 64-bit flags.  Exchange number
// `seq` uses slot seq % NSLOT; a slot cannot be overwritten before its reader is done because a peer can only be
// NSLOT exchanges ahead after this rank has itself sent NSLOT - 1 later exchanges, which its stream orders after the
// read (all exchanges of a rank must be issued on one stream, in the same order on every rank).
#include <string.h>

#include "common.cuh"

namespace dmb {
namespace peer {

constexpr int MAXR = 8;                    // ranks of one box
constexpr int NSLOT = 8;
constexpr int SLOT_BYTES = 8192;           // per (slot, source rank): 2048 floats / 1024 doubles
constexpr size_t FLAG_OFF = (size_t)NSLOT * MAXR * SLOT_BYTES;
constexpr size_t SEQ_OFF = FLAG_OFF + (size_t)NSLOT * MAXR * 8;   // this rank's own exchange counter (seq == 0 mode)
constexpr size_t BUF_BYTES = SEQ_OFF + 64;

struct Ptrs {
    unsigned char* p[MAXR];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// ---- the protocol, shared by the kernels below -------------------------------------------------------------------
// seq_arg == 0: the sequence number is the rank's own device-side counter (+1 per exchange, kept in its buffer), so that
// the launch carries no per-call host state and can be captured in a CUDA graph and replayed; every rank issues the
// same exchanges in the same order, so the counters agree.
__device__ __forceinline__ unsigned long long peer_seq(const Ptrs& pp, int rank, unsigned long long seq_arg) {
    const unsigned long long* counter = reinterpret_cast<const unsigned long long*>(pp.p[rank] + SEQ_OFF);
    return seq_arg ? seq_arg : *counter + 1;                   // (launches of one stream are serialised)
}
// this rank's region in rank p's buffer for the exchange's slot
__device__ __forceinline__ uint32_t* peer_out(const Ptrs& pp, int p, int rank, int slot) {
    return reinterpret_cast<uint32_t*>(pp.p[p] + ((size_t)slot * MAXR + rank) * SLOT_BYTES);
}
// rank p's region in my own buffer
__device__ __forceinline__ const unsigned char* peer_in(const Ptrs& pp, int rank, int p, int slot) {
    return pp.p[rank] + ((size_t)slot * MAXR + p) * SLOT_BYTES;
}
// after the payload stores: publish my sequence number to every rank, wait until every rank's payload is in MY buffer
__device__ __forceinline__ void peer_publish_and_wait(const Ptrs& pp, int rank, int world, unsigned long long seq,
                                                      unsigned long long seq_arg, int slot) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        unsigned long long* flag = reinterpret_cast<unsigned long long*>(pp.p[threadIdx.x] + FLAG_OFF) + slot * MAXR + rank;
        st_release_sys(flag, seq);
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(pp.p[rank] + FLAG_OFF) + slot * MAXR + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(mine) < seq) {
            if (clock64() - t0 > 40000000000LL) __trap();      // ~20 s: a peer died; fail loudly instead of hanging the GPU
        }
    }
    __syncthreads();
    // every thread has read the counter (peer_seq) before the barriers above
    if (threadIdx.x == 0 && !seq_arg) *reinterpret_cast<unsigned long long*>(pp.p[rank] + SEQ_OFF) = seq;
}

// mode 0: gather (dst = [world][nwords] 32-bit words), 1: sum of float64 (nwords / 2 values), 2: sum of float32.
// The payload is the concatenation of two source vectors (n0 + n1 = nwords words; src1 may be null with n1 == 0) and
// the result is split the same way over dst0 / dst1 (sums; a gather writes [world][nwords] to dst0).
__global__ void __launch_bounds__(256) peer_exchange_kernel(Ptrs pp, int rank, int world, unsigned long long seq_arg,
                                                            const uint32_t* __restrict__ src0, int n0,
                                                            const uint32_t* __restrict__ src1, uint32_t* __restrict__ dst0,
                                                            uint32_t* __restrict__ dst1, int nwords, int mode) {
    const unsigned long long seq = peer_seq(pp, rank, seq_arg);
    const int slot = (int)(seq % NSLOT);
    // 1. my vector into every rank's receive region (my own included): peer stores over NVLink
    for (int p = 0; p < world; ++p) {
        uint32_t* out = peer_out(pp, p, rank, slot);
        for (int w = threadIdx.x; w < nwords; w += blockDim.x) out[w] = w < n0 ? src0[w] : src1[w - n0];
    }
    // 2. publish, 3. wait until every rank's vector has landed in MY buffer
    peer_publish_and_wait(pp, rank, world, seq, seq_arg, slot);
    // 4. combine out of my own buffer, rank order fixed -> identical bits on every rank
    if (mode == 0) {
        for (int i = threadIdx.x; i < world * nwords; i += blockDim.x) {
            const int p = i / nwords, w = i - p * nwords;
            dst0[i] = reinterpret_cast<const uint32_t*>(peer_in(pp, rank, p, slot))[w];
        }
    } else if (mode == 1) {
        for (int i = threadIdx.x; i < nwords / 2; i += blockDim.x) {
            double acc = 0.0;
            for (int p = 0; p < world; ++p) acc += reinterpret_cast<const double*>(peer_in(pp, rank, p, slot))[i];
            if (2 * i < n0) reinterpret_cast<double*>(dst0)[i] = acc;
            else reinterpret_cast<double*>(dst1)[i - n0 / 2] = acc;
        }
    } else {
        for (int i = threadIdx.x; i < nwords; i += blockDim.x) {
            float acc = 0.f;
            for (int p = 0; p < world; ++p) acc += reinterpret_cast<const float*>(peer_in(pp, rank, p, slot))[i];
            if (i < n0) reinterpret_cast<float*>(dst0)[i] = acc;
            else reinterpret_cast<float*>(dst1)[i - n0] = acc;
        }
    }
}

// Synchronised BatchNorm, forward statistics: exchange of the per-rank (mean, 1/std, count) FUSED with their merge and
// the running-statistics update -- the work of torch's cat + all_gather + batch_norm_gather_stats_with_counts (and
// apex.parallel.SyncBatchNorm's welford_parallel after its all_gather) in one launch.
//   var_p = 1 / invstd_p^2 - eps;  N = sum n_p;  mean = sum n_p mean_p / N;  var = sum n_p (var_p + (mean_p - mean)^2) / N
//   running_mean += momentum (mean - running_mean);  running_var += momentum (var N / (N - 1) - running_var)
__global__ void __launch_bounds__(256) peer_bn_forward_kernel(Ptrs pp, int rank, int world, unsigned long long seq_arg,
                                                              const float* __restrict__ mean, const float* __restrict__ invstd,
                                                              float count, float eps, float momentum,
                                                              float* __restrict__ running_mean, float* __restrict__ running_var,
                                                              float* __restrict__ out_mean, float* __restrict__ out_invstd,
                                                              int* __restrict__ out_counts, int C) {
    const unsigned long long seq = peer_seq(pp, rank, seq_arg);
    const int slot = (int)(seq % NSLOT);
    for (int p = 0; p < world; ++p) {
        float* out = reinterpret_cast<float*>(peer_out(pp, p, rank, slot));
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            out[c] = mean[c];
            out[C + c] = invstd[c];
        }
        if (threadIdx.x == 0) out[2 * C] = count;
    }
    peer_publish_and_wait(pp, rank, world, seq, seq_arg, slot);
    float total = 0.f;
    for (int p = 0; p < world; ++p) total += reinterpret_cast<const float*>(peer_in(pp, rank, p, slot))[2 * C];
    if (threadIdx.x < world && out_counts)
        out_counts[threadIdx.x] = (int)reinterpret_cast<const float*>(peer_in(pp, rank, threadIdx.x, slot))[2 * C];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float m = 0.f;
        for (int p = 0; p < world; ++p) {
            const float* in = reinterpret_cast<const float*>(peer_in(pp, rank, p, slot));
            m += in[2 * C] * in[c];
        }
        m /= total;
        float v = 0.f;
        for (int p = 0; p < world; ++p) {
            const float* in = reinterpret_cast<const float*>(peer_in(pp, rank, p, slot));
            const float is = in[C + c], d = in[c] - m;
            v += in[2 * C] * ((1.f / (is * is) - eps) + d * d);
        }
        v = fmaxf(v / total, 0.f);
        out_mean[c] = m;
        out_invstd[c] = rsqrtf(v + eps);
        if (running_mean) running_mean[c] += momentum * (m - running_mean[c]);
        if (running_var) running_var[c] += momentum * (v * (total / fmaxf(total - 1.f, 1.f)) - running_var[c]);
    }
}

}  // namespace peer
}  // namespace dmb

using namespace dmb;
using namespace dmb::peer;

extern "C" int64_t dmb_b200_peer_buffer_bytes(void) { return (int64_t)BUF_BYTES; }

// Allocates and zeroes this rank's receive buffer (cudaMalloc: CUDA IPC cannot export pooled / VMM allocations).
extern "C" int dmb_b200_peer_alloc(void** ptr) {
    DMB_REQUIRE(ptr, "peer_alloc: null pointer");
    DMB_CUDA(cudaMalloc(ptr, BUF_BYTES));
    DMB_CUDA(cudaMemset(*ptr, 0, BUF_BYTES));
    DMB_CUDA(cudaDeviceSynchronize());
    return DMB_OK;
}
extern "C" int dmb_b200_peer_free(void* ptr) {
    if (ptr) DMB_CUDA(cudaFree(ptr));
    return DMB_OK;
}
// 64-byte CUDA IPC handle of a buffer from dmb_b200_peer_alloc / the mapping of a peer's handle in this process
extern "C" int dmb_b200_peer_export(void* ptr, void* handle64) {
    DMB_REQUIRE(ptr && handle64, "peer_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is expected to be 64 bytes");
    DMB_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), ptr));
    return DMB_OK;
}
extern "C" int dmb_b200_peer_import(const void* handle64, void** ptr) {
    DMB_REQUIRE(ptr && handle64, "peer_import: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    DMB_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DMB_OK;
}
extern "C" int dmb_b200_peer_close(void* ptr) {
    if (ptr) DMB_CUDA(cudaIpcCloseMemHandle(ptr));
    return DMB_OK;
}

static int peer_ptrs(Ptrs& pp, void* const* bufs, int rank, int world, const char* who) {
    DMB_REQUIRE(bufs, "%s: null buffer table", who);
    DMB_REQUIRE(world >= 1 && world <= MAXR && rank >= 0 && rank < world, "%s: rank %d / world %d out of range (max %d)", who, rank, world, MAXR);
    for (int i = 0; i < MAXR; ++i) pp.p[i] = i < world ? reinterpret_cast<unsigned char*>(bufs[i]) : nullptr;
    for (int i = 0; i < world; ++i) DMB_REQUIRE(pp.p[i], "%s: buffer of rank %d is NULL", who, i);
    return DMB_OK;
}

// One exchange.  bufs: HOST array of `world` device pointers (entry `rank` = this rank's own buffer, the others the
// imported peer mappings); seq: 0 = the device-side counter of the buffer (graph-capturable; do not mix with explicit
// numbers on one set of buffers), or 1, 2, 3, ... identical on every rank; src: nbytes (multiple of 4, <= 8192) on this device;
// dst: world * nbytes (mode 0, gather) or nbytes (mode 1: float64 sum, mode 2: float32 sum).
extern "C" int dmb_b200_peer_exchange(void* const* bufs, int rank, int world, long long seq, const void* src, void* dst,
                                      int nbytes, int mode, void* stream) {
    DMB_REQUIRE(src && dst, "peer_exchange: null pointer");
    DMB_REQUIRE(seq >= 0, "peer_exchange: sequence number must be 0 (device-side counter) or 1, 2, 3, ...");
    DMB_REQUIRE(nbytes > 0 && nbytes <= SLOT_BYTES && nbytes % 4 == 0 && (mode != 1 || nbytes % 8 == 0),
                "peer_exchange: %d bytes not supported (multiple of 4, at most %d)", nbytes, SLOT_BYTES);
    DMB_REQUIRE(mode >= 0 && mode <= 2, "peer_exchange: mode must be 0 (gather), 1 (sum f64) or 2 (sum f32)");
    Ptrs pp;
    int rc = peer_ptrs(pp, bufs, rank, world, "peer_exchange");
    if (rc) return rc;
    peer_exchange_kernel<<<1, 256, 0, as_stream(stream)>>>(pp, rank, world, (unsigned long long)seq,
                                                           reinterpret_cast<const uint32_t*>(src), nbytes / 4, nullptr,
                                                           reinterpret_cast<uint32_t*>(dst), nullptr, nbytes / 4, mode);
    return check_launch("peer_exchange_kernel");
}

// Float32 sums of TWO vectors over the ranks in one exchange (the backward statistics of synchronised BatchNorm:
// sum(dy) and sum(dy * (x - mean)), torch.batch_norm_backward_reduce's outputs): dst0[n0], dst1[n1].
extern "C" int dmb_b200_peer_sum2_f32(void* const* bufs, int rank, int world, const float* src0, int n0, const float* src1,
                                      int n1, float* dst0, float* dst1, void* stream) {
    DMB_REQUIRE(src0 && src1 && dst0 && dst1, "peer_sum2_f32: null pointer");
    DMB_REQUIRE(n0 > 0 && n1 > 0 && (n0 + n1) * 4 <= SLOT_BYTES, "peer_sum2_f32: %d + %d values exceed %d bytes", n0, n1, SLOT_BYTES);
    Ptrs pp;
    int rc = peer_ptrs(pp, bufs, rank, world, "peer_sum2_f32");
    if (rc) return rc;
    peer_exchange_kernel<<<1, 256, 0, as_stream(stream)>>>(pp, rank, world, 0ull, reinterpret_cast<const uint32_t*>(src0), n0,
                                                           reinterpret_cast<const uint32_t*>(src1),
                                                           reinterpret_cast<uint32_t*>(dst0), reinterpret_cast<uint32_t*>(dst1),
                                                           n0 + n1, 2);
    return check_launch("peer_exchange_kernel<sum2>");
}

// Forward statistics of synchronised BatchNorm in one launch: this rank's (mean[C], invstd[C], count) -> the statistics
// of the joint batch (out_mean[C], out_invstd[C]), every rank's count (out_counts[world], int32, may be NULL) and the
// running-statistics update (running_mean / running_var may be NULL).  torch.batch_norm_stats produces the inputs,
// torch.batch_norm_elemt consumes the outputs; replaces all_gather + batch_norm_gather_stats_with_counts.
extern "C" int dmb_b200_peer_bn_forward(void* const* bufs, int rank, int world, const float* mean, const float* invstd,
                                        float count, float eps, float momentum, float* running_mean, float* running_var,
                                        float* out_mean, float* out_invstd, int* out_counts, int C, void* stream) {
    DMB_REQUIRE(mean && invstd && out_mean && out_invstd, "peer_bn_forward: null pointer");
    DMB_REQUIRE(C > 0 && (2 * C + 1) * 4 <= SLOT_BYTES, "peer_bn_forward: %d channels exceed the %d-byte slot", C, SLOT_BYTES);
    DMB_REQUIRE(count >= 1.f, "peer_bn_forward: empty batch");
    Ptrs pp;
    int rc = peer_ptrs(pp, bufs, rank, world, "peer_bn_forward");
    if (rc) return rc;
    peer_bn_forward_kernel<<<1, 256, 0, as_stream(stream)>>>(pp, rank, world, 0ull, mean, invstd, count, eps, momentum,
                                                             running_mean, running_var, out_mean, out_invstd, out_counts, C);
    return check_launch("peer_bn_forward_kernel");
}

extern "C" int dmb_b200_peer_count(const void* own_buf, long long* count) {
    DMB_REQUIRE(own_buf && count, "peer_count: null pointer");
    unsigned long long c = 0;
    DMB_CUDA(cudaDeviceSynchronize());
    DMB_CUDA(cudaMemcpy(&c, reinterpret_cast<const unsigned char*>(own_buf) + SEQ_OFF, 8, cudaMemcpyDeviceToHost));
    *count = (long long)c;
    return DMB_OK;
}
