// optim.cu -- AdamW over ONE flat fp32 parameter buffer (the optimizer the reference configures,
// config.yml "optimizer: AdamW", instantiated at matdeeplearn/training/training.py:429-432, applied at
// training.py:49).  torch's fused AdamW hands a single 80k..1M-element tensor to two thread blocks; the
// engine keeps all parameters and gradients in one flat buffer (dist.FlatParameters), so a plain
// grid-wide elementwise kernel is both simpler and an order of magnitude faster, and its
// hyper-parameters live on the device (lr schedulers work without re-capturing the CUDA graph).
//
//   state:  hyper[0..4] = {lr, beta1, beta2, eps, weight_decay} (device), step (device, float count)
//   p <- p * (1 - lr*wd);  m <- b1 m + (1-b1) g;  v <- b2 v + (1-b2) g^2
//   p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)          (torch.optim.AdamW)
#include "common.cuh"

namespace mdl {

__global__ void k_adamw_flat(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, const float* __restrict__ hyper,
                             const float* __restrict__ step, float grad_scale, int64_t n) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const float t = step[0] + 1.0f;  // k_adamw_tick runs after this kernel
  const float bc1 = 1.0f - powf(b1, t), bc2 = 1.0f - powf(b2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    pi -= step_size * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}
__global__ void k_adamw_tick(float* step) { step[0] += 1.0f; }

}  // namespace mdl

using namespace mdl;

extern "C" int mdl_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                              const float* hyper, float* step, float grad_scale, int64_t n, void* stream) {
  MDL_REQUIRE(n >= 0, "adamw_step: bad size");
  if (n == 0) return MDL_OK;
  MDL_REQUIRE(param && grad && exp_avg && exp_avg_sq && hyper && step, "adamw_step: null pointer");
  int grid = (int)std::min<int64_t>(ceil_div<int64_t>(n, 256), (int64_t)kNumSMs * 8);
  k_adamw_flat<<<grid, 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, hyper, step, grad_scale, n);
  MDL_LAUNCHED();
  k_adamw_tick<<<1, 1, 0, as_stream(stream)>>>(step);
  MDL_LAUNCHED();
  return MDL_OK;
}
