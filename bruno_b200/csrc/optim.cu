// Optimiser steps with the semantics of the reference's TF-1.7 optimisers (config_rnn/train.py:121-131), applied to
// one flat fp32 parameter buffer (memory-bound, 128-bit accesses; 20 bytes/parameter for RMSProp):
//   tf.train.RMSPropOptimizer(lr): decay 0.9, momentum 0, epsilon 1e-10 INSIDE the sqrt, rms slot initialised to 1:
//       ms <- decay ms + (1-decay) g^2 ;  p <- p - lr g / sqrt(ms + eps)
//   tf.train.AdamOptimizer(lr, beta1=0.95, beta2=0.9995): lr_t = lr sqrt(1-b2^t)/(1-b1^t);  p <- p - lr_t m/(sqrt(v)+eps)
// `grad_scale` multiplies the gradient first: 1/nr_gpu of the tower average (train.py:103-105) and the student-grad
// schedule (train.py:108-111) are folded in here.
#include "common.cuh"

namespace bruno {

__global__ void __launch_bounds__(256)
rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ ms, size_t n, float lr, float decay,
               float eps, float grad_scale) {
  const size_t i4 = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    float4 pv = *reinterpret_cast<float4*>(p + i4), gv = *reinterpret_cast<const float4*>(g + i4),
           mv = *reinterpret_cast<float4*>(ms + i4);
    float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gg = gp[e] * grad_scale;
      mp[e] = decay * mp[e] + (1.f - decay) * gg * gg;
      pp[e] = pp[e] - lr * gg * rsqrtf(mp[e] + eps);
    }
    *reinterpret_cast<float4*>(p + i4) = pv;
    *reinterpret_cast<float4*>(ms + i4) = mv;
  } else {
    for (size_t i = i4; i < n; ++i) {
      const float gg = g[i] * grad_scale;
      ms[i] = decay * ms[i] + (1.f - decay) * gg * gg;
      p[i] = p[i] - lr * gg * rsqrtf(ms[i] + eps);
    }
  }
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
            float lr_t, float beta1, float beta2, float eps, float grad_scale) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gg = g[i] * grad_scale;
  const float mn = beta1 * m[i] + (1.f - beta1) * gg;
  const float vn = beta2 * v[i] + (1.f - beta2) * gg * gg;
  m[i] = mn;
  v[i] = vn;
  p[i] = p[i] - lr_t * mn / (sqrtf(vn) + eps);
}

// Episodic fine-tuning loss (config_rnn/bn2_omniglot_tp_ft_1s_20w.py:222-226): log_probs [M*B, T] -> per episode m a
// softmax over the B candidate sequences of the LAST element's log-prob; loss = -mean_m log softmax_m[0].  One warp per
// episode (B <= 1024), one block; writes the loss and (optionally) its gradient d loss / d log_probs in the same launch:
//   d loss / d lp[m, b, T-1] = (softmax_m[b] - [b == 0]) / M, zero for the other columns.
__global__ void __launch_bounds__(1024)
episode_softmax_loss_kernel(const float* __restrict__ lp, int M, int B, int T, float* __restrict__ loss,
                            float* __restrict__ dlp) {
  __shared__ float partial[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  float acc = 0.f;
  for (int m = warp; m < M; m += nwarp) {
    const float* row = lp + static_cast<size_t>(m) * B * T + (T - 1);
    float mx = -INFINITY;
    for (int b = lane; b < B; b += 32) mx = fmaxf(mx, row[static_cast<size_t>(b) * T]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
    for (int b = lane; b < B; b += 32) se += expf(row[static_cast<size_t>(b) * T] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    // |log_probs| ~ 1e3: never form lse = mx + log(se) (its rounding at that magnitude would cost 6e-5 relative in the
    // softmax); work with the exact differences row - mx instead
    const float inv_se = 1.0f / se;
    if (lane == 0) acc += (mx - row[0]) + logf(se);     // -log softmax[0]
    if (dlp) {
      float* drow = dlp + static_cast<size_t>(m) * B * T;
      for (int i = lane; i < B * T; i += 32) {
        const int b = i / T, t = i - b * T;
        drow[i] = t == T - 1 ? (expf(row[static_cast<size_t>(b) * T] - mx) * inv_se - (b == 0 ? 1.f : 0.f)) / static_cast<float>(M) : 0.f;
      }
    }
  }
  if (lane == 0) partial[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < nwarp; ++w) t += partial[w];    // fixed order: deterministic
    *loss = t / static_cast<float>(M);
  }
}

}  // namespace bruno

using namespace bruno;

extern "C" {

int bruno_rmsprop_tf17(float* p, const float* g, float* ms, size_t n, float lr, float decay, float eps, float grad_scale,
                       void* stream) {
  BRUNO_REQUIRE(p && g && ms && n > 0, "bruno_rmsprop_tf17: bad arguments");
  BRUNO_REQUIRE(aligned16(p) && aligned16(g) && aligned16(ms), "bruno_rmsprop_tf17: 16-byte alignment required");
  const size_t threads = (n + 3) / 4;
  rmsprop_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, ms, n, lr, decay, eps, grad_scale);
  return check_launch("rmsprop_kernel");
}

int bruno_adam_tf17(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
                    int t, float grad_scale, void* stream) {
  BRUNO_REQUIRE(p && g && m && v && n > 0 && t >= 1, "bruno_adam_tf17: bad arguments");
  const double lr_t = static_cast<double>(lr) * sqrt(1.0 - pow(static_cast<double>(beta2), t)) /
                      (1.0 - pow(static_cast<double>(beta1), t));
  adam_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, n, static_cast<float>(lr_t), beta1, beta2, eps, grad_scale);
  return check_launch("adam_kernel");
}

int bruno_episode_softmax_loss(const float* log_probs, int M, int B, int T, float* loss, float* dlog_probs, void* stream) {
  BRUNO_REQUIRE(log_probs && loss && M >= 1 && B >= 1 && T >= 1, "bruno_episode_softmax_loss: bad arguments");
  const int warps = M < 32 ? M : 32;
  episode_softmax_loss_kernel<<<1, 32 * warps, 0, static_cast<cudaStream_t>(stream)>>>(log_probs, M, B, T, loss, dlog_probs);
  return check_launch("episode_softmax_loss_kernel");
}

}  // extern "C"
