// wtmk_loss.cu — the loss head of the watermark training step ("decoder + loss on device in the
// same graph"), nerf/utils_wtmk_disen.py:592-593 and 636-644:
//
//     pred_rgb = clamp(outputs['image'], 0, 1)                               -> decoder input
//     lossi    = MSELoss(reduction='none')(content_pred_rgb, gt_rgb).mean()
//     lossw    = binary_cross_entropy_with_logits(decoded * 10, message[:, None], reduction='mean')
//     loss     = lambda_w * lossw + lambda_i * lossi
//
// In PyTorch these are ~15 element-wise / reduction launches forward and ~20 backward on tensors of a few thousand
// elements: pure launch latency (~2.5 us each) inside a 1.3 ms step.  Here: one kernel splits the rendered
// [block rays | content rays] image into the clamped decoder input and the content pixels (its backward merges the two
// gradients and applies the clamp mask), one single-CTA kernel evaluates both losses AND their gradients (the loss is a
// fixed function of its inputs, so the backward pass only scales the stored gradients by the incoming scalar - the
// GradScaler's loss scale).  Deterministic: no atomics.
#include "nsig_common.cuh"

namespace nsig {

__global__ void __launch_bounds__(256)
k_split_clamp_fwd(const float* __restrict__ image, uint32_t n_block, uint32_t n_total, float* __restrict__ pred,
                  float* __restrict__ content) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_total) return;
    const float x = image[i];
    if (i < n_block) pred[i] = (x != x) ? x : fminf(fmaxf(x, 0.0f), 1.0f);   // torch.clamp(min=0, max=1) keeps NaN
    else content[i - n_block] = x;
}

__global__ void __launch_bounds__(256)
k_split_clamp_bwd(const float* __restrict__ image, const float* __restrict__ g_pred, const float* __restrict__ g_content,
                  uint32_t n_block, uint32_t n_total, float* __restrict__ g_image) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_total) return;
    float g = 0.0f;
    if (i < n_block) {
        const float x = image[i];
        if (g_pred && x >= 0.0f && x <= 1.0f) g = g_pred[i];   // clamp backward: gradient passes where min <= x <= max
    } else if (g_content) {
        g = g_content[i - n_block];
    }
    g_image[i] = g;
}

constexpr int kLossThreads = 1024;

// out[0] = loss, out[1] = lossi, out[2] = lossw;  g_image[i] = d loss / d image[i], g_logits[j] = d loss / d logits[j]
__global__ void __launch_bounds__(kLossThreads)
k_wtmk_loss_fwd(const float* __restrict__ image, const float* __restrict__ gt, uint32_t n, const float* __restrict__ logits,
                const float* __restrict__ message, uint32_t md, float lambda_w, float lambda_i, float temp,
                float* __restrict__ out, float* __restrict__ g_image, float* __restrict__ g_logits) {
    __shared__ float red[2][kLossThreads / 32];
    const float gi = n ? lambda_i * 2.0f / (float)n : 0.0f;
    float se = 0.0f;
    for (uint32_t i = threadIdx.x; i < n; i += kLossThreads) {
        const float d = image[i] - gt[i];
        se = fmaf(d, d, se);
        g_image[i] = gi * d;
    }
    float bce = 0.0f;
    for (uint32_t j = threadIdx.x; j < md; j += kLossThreads) {
        const float z = logits[j] * temp, y = message[j];
        // max(z,0) - z*y + log(1 + exp(-|z|))   (ATen's binary_cross_entropy_with_logits form)
        bce += fmaxf(z, 0.0f) - z * y + log1pf(expf(-fabsf(z)));
        const float s = 1.0f / (1.0f + expf(-z));
        g_logits[j] = lambda_w * temp * (s - y) / (float)md;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        se += __shfl_xor_sync(NSIG_FULL_MASK, se, o);
        bce += __shfl_xor_sync(NSIG_FULL_MASK, bce, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = se; red[1][warp] = bce; }
    __syncthreads();
    if (warp == 0) {
        se = lane < kLossThreads / 32 ? red[0][lane] : 0.0f;
        bce = lane < kLossThreads / 32 ? red[1][lane] : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            se += __shfl_xor_sync(NSIG_FULL_MASK, se, o);
            bce += __shfl_xor_sync(NSIG_FULL_MASK, bce, o);
        }
        if (lane == 0) {
            const float lossi = n ? se / (float)n : 0.0f, lossw = md ? bce / (float)md : 0.0f;
            out[0] = lambda_w * lossw + lambda_i * lossi;
            out[1] = lossi;
            out[2] = lossw;
        }
    }
}

__global__ void __launch_bounds__(256)
k_wtmk_loss_bwd(const float* __restrict__ g_image, const float* __restrict__ g_logits, uint32_t n, uint32_t md,
                const float* __restrict__ grad_out, float* __restrict__ d_image, float* __restrict__ d_logits) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    const float s = *grad_out;
    if (i < n) d_image[i] = g_image[i] * s;
    else if (i < n + md) d_logits[i - n] = g_logits[i - n] * s;
}

}  // namespace nsig

using namespace nsig;

extern "C" {

int nsig_split_clamp_forward(const float* image, uint32_t n_block, uint32_t n_total, float* pred, float* content,
                             nsig_stream_t stream) {
    if (n_total == 0) return 0;
    if (!image || n_block > n_total || (n_block && !pred) || (n_total > n_block && !content)) return NSIG_EINVAL;
    k_split_clamp_fwd<<<div_up(n_total, 256), 256, 0, (cudaStream_t)stream>>>(image, n_block, n_total, pred, content);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_split_clamp_backward(const float* image, const float* grad_pred, const float* grad_content, uint32_t n_block,
                              uint32_t n_total, float* grad_image, nsig_stream_t stream) {
    if (n_total == 0) return 0;
    if (!image || !grad_image || n_block > n_total) return NSIG_EINVAL;
    k_split_clamp_bwd<<<div_up(n_total, 256), 256, 0, (cudaStream_t)stream>>>(image, grad_pred, grad_content, n_block,
                                                                               n_total, grad_image);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_wtmk_loss_forward(const float* image, const float* gt, uint32_t n, const float* logits, const float* message,
                           uint32_t md, float lambda_w, float lambda_i, float temp, float* out, float* g_image,
                           float* g_logits, nsig_stream_t stream) {
    if (!out || (n && (!image || !gt || !g_image)) || (md && (!logits || !message || !g_logits))) return NSIG_EINVAL;
    k_wtmk_loss_fwd<<<1, kLossThreads, 0, (cudaStream_t)stream>>>(image, gt, n, logits, message, md, lambda_w, lambda_i,
                                                                  temp, out, g_image, g_logits);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_wtmk_loss_backward(const float* g_image, const float* g_logits, uint32_t n, uint32_t md, const float* grad_out,
                            float* d_image, float* d_logits, nsig_stream_t stream) {
    if (n + md == 0) return 0;
    if (!grad_out || (n && (!g_image || !d_image)) || (md && (!g_logits || !d_logits))) return NSIG_EINVAL;
    k_wtmk_loss_bwd<<<div_up(n + md, 256), 256, 0, (cudaStream_t)stream>>>(g_image, g_logits, n, md, grad_out, d_image,
                                                                           d_logits);
    NSIG_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
