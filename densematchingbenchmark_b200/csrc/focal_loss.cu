// Stereo focal loss on the raw cost volume, forward and backward, one pass over the volume each.
//
// Reference: StereoFocalLoss.loss_per_level, dmb/modeling/stereo/losses/stereo_focal_loss.py:63-101, with the
// ground-truth distribution of LaplaceDisp2Prob, dmb/modeling/stereo/losses/utils/disp2prob.py:29-173:
//     q_d   = softmax_d(-|s_d - gt| / variance) * inner_mask + 1e-40
//     loss  = -sum_pixels outer_mask * sum_d q_d (1 - q_d)^(-coefficient) log_softmax(cost)_d  /  max(#outer_mask, 1)
// The reference materialises five [B,D,H,W] volumes (samples, Laplace logits, gtProb, log-softmax, weight): ~2 GB of
// traffic per 401 MB cost volume.  Here a thread owns a pixel: the Laplace distribution depends only on (gt,
// variance), so its normaliser is computed in registers first; then ONE march over the cost column accumulates the
// online log-sum-exp of the cost together with A = sum q w c and Bs = sum q w, and
//     loss_pixel = -(A - logZ * Bs).
// The backward pass re-reads the cost once and writes d(cost) (and d(variance) for AcfNet's adaptive variance):
//     d loss_pixel / d c_d = softmax(c)_d * Bs - q_d w_d
//     d loss_pixel / d var = -sum_d log p_d * w_d (1 + coefficient q_d / (1 - q_d)) * q~_d (a_d - abar) / var^2,
// a_d = |s_d - gt|, abar = sum q~ a, q~ the softmax before mask / eps.  Oracle: oracle/dmb_oracle.py:stereo_focal_loss.
#include "common.cuh"

namespace dmb {

struct FocalParams {
    int B, D, H, W;
    float lower, upper;       // outer mask: lower < gt < upper            (stereo_focal_loss.py:78-81)
    float inner_end;          // inner mask: gt < start + max_disp - 1     (disp2prob.py:60,126)
    float coefficient;
    float var_scalar;
};

__device__ __forceinline__ float focal_weight(float q, float coefficient) {
    // (1 - q)^(-coefficient); coefficient == 0 -> exactly 1 like torch.pow(x, -0.0)
    return coefficient == 0.f ? 1.f : exp2f(-coefficient * log2f(1.f - q));
}

// Laplace distribution of one pixel: returns 1 / sum_d exp(l_d - lmax) and lmax; optionally abar = sum q~_d a_d
template <bool PER_PIXEL>
__device__ __forceinline__ void laplace_norm(const float* __restrict__ samples, const float* __restrict__ sp, size_t plane,
                                             int D, float g, float inv_var, float& lmax, float& inv_sum, float* abar) {
    float amin = INFINITY;
    for (int d = 0; d < D; ++d) {
        const float s = PER_PIXEL ? __ldg(sp + (size_t)d * plane) : samples[d];
        amin = fminf(amin, fabsf(s - g));
    }
    lmax = -amin * inv_var;
    float sum = 0.f, asum = 0.f;
    for (int d = 0; d < D; ++d) {
        const float s = PER_PIXEL ? __ldg(sp + (size_t)d * plane) : samples[d];
        const float a = fabsf(s - g);
        const float e = expf(-a * inv_var - lmax);
        sum += e;
        asum = fmaf(e, a, asum);
    }
    inv_sum = 1.f / sum;
    if (abar) *abar = asum * inv_sum;
}

template <bool PER_PIXEL>
__global__ void __launch_bounds__(256) focal_loss_fwd_kernel(const float* __restrict__ cost, const float* __restrict__ gt,
                                                             const float* __restrict__ var_map,
                                                             const float* __restrict__ disp_values,
                                                             const float* __restrict__ disp_sample, FocalParams p,
                                                             double* __restrict__ sums, float* __restrict__ stats) {
    extern __shared__ float samples[];                 // [D] shared disparity samples
    __shared__ float red[2][8];
    if (!PER_PIXEL)
        for (int d = threadIdx.x; d < p.D; d += blockDim.x) samples[d] = __ldg(disp_values + d);
    __syncthreads();
    const size_t plane = (size_t)p.H * p.W;
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    float loss = 0.f, valid = 0.f;
    if (pix < plane) {
        const float g0 = __ldg(gt + (size_t)b * plane + pix);
        const bool m_out = g0 > p.lower && g0 < p.upper;
        float logz = 0.f, bs = 0.f;
        if (m_out) {
            const bool m_in = g0 < p.inner_end;        // (g0 > start holds: lower == start)
            const float g = m_in ? g0 : 0.f;
            const float var = var_map ? __ldg(var_map + (size_t)b * plane + pix) : p.var_scalar;
            const float inv_var = 1.f / var;
            const float* sp = PER_PIXEL ? disp_sample + (size_t)b * p.D * plane + pix : nullptr;
            float lmax, inv_sum;
            laplace_norm<PER_PIXEL>(samples, sp, plane, p.D, g, inv_var, lmax, inv_sum, nullptr);
            const float* cp = cost + (size_t)b * p.D * plane + pix;
            float m = -INFINITY, s = 0.f, A = 0.f;
            for (int d0 = 0; d0 < p.D; d0 += 4) {
                float c[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) c[j] = (d0 + j < p.D) ? __ldcs(cp + (size_t)(d0 + j) * plane) : -INFINITY;
                const float cm = fmaxf(fmaxf(c[0], c[1]), fmaxf(c[2], c[3]));
                if (cm > m) {
                    s *= expf(m - cm);
                    m = cm;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (d0 + j < p.D) {
                        const float sd = PER_PIXEL ? __ldg(sp + (size_t)(d0 + j) * plane) : samples[d0 + j];
                        float q = m_in ? expf(-fabsf(sd - g) * inv_var - lmax) * inv_sum : 0.f;
                        q += 1e-40f;
                        const float qw = q * focal_weight(q, p.coefficient);
                        s += expf(c[j] - m);
                        A = fmaf(qw, c[j], A);
                        bs += qw;
                    }
                }
            }
            logz = m + logf(s);
            loss = -(A - logz * bs);
            valid = 1.f;
        }
        if (stats) {
            stats[((size_t)b * 2 + 0) * plane + pix] = logz;
            stats[((size_t)b * 2 + 1) * plane + pix] = bs;
        }
    }
    // block reduction -> two double atomics per CTA
    loss = warp_sum(loss);
    valid = warp_sum(valid);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        red[0][wid] = loss;
        red[1][wid] = valid;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        float l = lane < 8 ? red[0][lane] : 0.f, v = lane < 8 ? red[1][lane] : 0.f;
        l = warp_sum(l);
        v = warp_sum(v);
        if (lane == 0) {
            atomicAdd(sums, (double)l);
            atomicAdd(sums + 1, (double)v);
        }
    }
}

template <bool PER_PIXEL>
__global__ void __launch_bounds__(256) focal_loss_bwd_kernel(const float* __restrict__ cost, const float* __restrict__ gt,
                                                             const float* __restrict__ var_map,
                                                             const float* __restrict__ disp_values,
                                                             const float* __restrict__ disp_sample,
                                                             const float* __restrict__ stats,
                                                             const float* __restrict__ gscale_ptr, FocalParams p,
                                                             float* __restrict__ dcost, float* __restrict__ dvar) {
    extern __shared__ float samples[];
    if (!PER_PIXEL)
        for (int d = threadIdx.x; d < p.D; d += blockDim.x) samples[d] = __ldg(disp_values + d);
    __syncthreads();
    const size_t plane = (size_t)p.H * p.W;
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (pix >= plane) return;
    const float gscale = __ldg(gscale_ptr);            // upstream gradient * level weight / #valid
    const float g0 = __ldg(gt + (size_t)b * plane + pix);
    const bool m_out = g0 > p.lower && g0 < p.upper;
    float* dc = dcost + (size_t)b * p.D * plane + pix;
    if (!m_out) {
        for (int d = 0; d < p.D; ++d) __stcs(dc + (size_t)d * plane, 0.f);
        if (dvar) dvar[(size_t)b * plane + pix] = 0.f;
        return;
    }
    const bool m_in = g0 < p.inner_end;
    const float g = m_in ? g0 : 0.f;
    const float var = var_map ? __ldg(var_map + (size_t)b * plane + pix) : p.var_scalar;
    const float inv_var = 1.f / var;
    const float* sp = PER_PIXEL ? disp_sample + (size_t)b * p.D * plane + pix : nullptr;
    float lmax, inv_sum, abar;
    laplace_norm<PER_PIXEL>(samples, sp, plane, p.D, g, inv_var, lmax, inv_sum, &abar);
    const float logz = __ldg(stats + ((size_t)b * 2 + 0) * plane + pix);
    const float bs = __ldg(stats + ((size_t)b * 2 + 1) * plane + pix);
    const float* cp = cost + (size_t)b * p.D * plane + pix;
    float dv = 0.f;
    for (int d0 = 0; d0 < p.D; d0 += 4) {
        float c[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) c[j] = (d0 + j < p.D) ? __ldcs(cp + (size_t)(d0 + j) * plane) : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (d0 + j < p.D) {
                const float sd = PER_PIXEL ? __ldg(sp + (size_t)(d0 + j) * plane) : samples[d0 + j];
                const float a = fabsf(sd - g);
                const float qs = m_in ? expf(-a * inv_var - lmax) * inv_sum : 0.f;      // softmax before mask / eps
                const float q = qs + 1e-40f;
                const float w = focal_weight(q, p.coefficient);
                const float logp = c[j] - logz;
                __stcs(dc + (size_t)(d0 + j) * plane, gscale * (expf(logp) * bs - q * w));
                // f'(q) = w (1 + coefficient q / (1 - q));  dq/dvar = qs (a - abar) / var^2
                dv = fmaf(logp * w * (1.f + p.coefficient * q / (1.f - q)), qs * (a - abar), dv);
            }
        }
    }
    if (dvar) dvar[(size_t)b * plane + pix] = -gscale * dv * inv_var * inv_var;
}

}  // namespace dmb

using namespace dmb;

static int focal_check(const float* cost, const float* gt, const float* disp_values, const float* disp_sample, int B, int D,
                       int H, int W) {
    DMB_REQUIRE(cost && gt, "focal_loss: null cost / ground truth");
    DMB_REQUIRE((disp_values != nullptr) != (disp_sample != nullptr), "focal_loss: exactly one of disp_values / disp_sample");
    DMB_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "focal_loss: non-positive dimension");
    DMB_REQUIRE(B <= 65535 && D <= 8192, "focal_loss: B or D too large");
    return DMB_OK;
}

extern "C" int dmb_b200_focal_loss_forward(const float* cost, const float* gt, const float* var_map, float var_scalar,
                                           const float* disp_values, const float* disp_sample, int B, int D, int H, int W,
                                           float lower, float upper, float inner_end, float coefficient, double* sums,
                                           float* stats, void* stream) {
    int rc = focal_check(cost, gt, disp_values, disp_sample, B, D, H, W);
    if (rc) return rc;
    DMB_REQUIRE(sums, "focal_loss_forward: null sums");
    DMB_REQUIRE(var_map || var_scalar != 0.f, "focal_loss_forward: zero variance");
    FocalParams p{B, D, H, W, lower, upper, inner_end, coefficient, var_scalar};
    const size_t plane = (size_t)H * W;
    dim3 grid((unsigned)cdiv(plane, 256), B);
    const size_t smem = (size_t)D * sizeof(float);
    if (disp_sample)
        focal_loss_fwd_kernel<true><<<grid, 256, smem, as_stream(stream)>>>(cost, gt, var_map, nullptr, disp_sample, p, sums, stats);
    else
        focal_loss_fwd_kernel<false><<<grid, 256, smem, as_stream(stream)>>>(cost, gt, var_map, disp_values, nullptr, p, sums, stats);
    return check_launch("focal_loss_fwd_kernel");
}

extern "C" int dmb_b200_focal_loss_backward(const float* cost, const float* gt, const float* var_map, float var_scalar,
                                            const float* disp_values, const float* disp_sample, const float* stats,
                                            const float* gscale, int B, int D, int H, int W, float lower, float upper,
                                            float inner_end, float coefficient, float* dcost, float* dvar, void* stream) {
    int rc = focal_check(cost, gt, disp_values, disp_sample, B, D, H, W);
    if (rc) return rc;
    DMB_REQUIRE(stats && gscale && dcost, "focal_loss_backward: null stats / gscale / dcost");
    DMB_REQUIRE(!dvar || var_map, "focal_loss_backward: dvar requested without a variance map");
    FocalParams p{B, D, H, W, lower, upper, inner_end, coefficient, var_scalar};
    const size_t plane = (size_t)H * W;
    dim3 grid((unsigned)cdiv(plane, 256), B);
    const size_t smem = (size_t)D * sizeof(float);
    if (disp_sample)
        focal_loss_bwd_kernel<true><<<grid, 256, smem, as_stream(stream)>>>(cost, gt, var_map, nullptr, disp_sample, stats, gscale, p,
                                                                          dcost, dvar);
    else
        focal_loss_bwd_kernel<false><<<grid, 256, smem, as_stream(stream)>>>(cost, gt, var_map, disp_values, nullptr, stats, gscale, p,
                                                                           dcost, dvar);
    return check_launch("focal_loss_bwd_kernel");
}
