// Fused multi-tensor Adam on the flat adapter/router parameter buffer.
//
// The reference trains the adapter and router parameters with torch.optim.Adam(param_group) (AVMOE/AVE/main_trans_v3.py:311-323:
// one learning rate for the adapter stack, default betas/eps, no weight decay, StepLR on top) -- ~1,000 small tensors, i.e.
// ~1,000 tiny launch groups per optimizer step.  Here all trainable tensors are views of ONE fp32 buffer (avmoe_b200/optim.py),
// their gradients views of a second one that the gradient all-reduce reduces in place, and the whole update is a single
// streaming pass: 16 bytes read + 12 bytes written per parameter, HBM-bound.  The 1/world of the gradient average and the
// zeroing of the gradient buffer for the next step are folded into the same pass.
#include "common.cuh"

namespace avmoe {

__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                        long long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                                                        float grad_scale, float bc1, float bc2_sqrt, int zero_grad) {
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float step = lr / bc1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i], gv = reinterpret_cast<float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* pp = &pv.x; float* gg = &gv.x; float* mm = &mv.x; float* vw = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gr = gg[k] * grad_scale;
      if (weight_decay != 0.f) gr += weight_decay * pp[k];
      mm[k] = beta1 * mm[k] + (1.f - beta1) * gr;                 // torch: exp_avg.lerp_(grad, 1 - beta1)
      vw[k] = beta2 * vw[k] + (1.f - beta2) * gr * gr;            // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      const float denom = sqrtf(vw[k]) / bc2_sqrt + eps;
      pp[k] -= step * (mm[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {   // tail (< 4 elements)
    float gr = g[i] * grad_scale;
    if (weight_decay != 0.f) gr += weight_decay * p[i];
    m[i] = beta1 * m[i] + (1.f - beta1) * gr;
    v[i] = beta2 * v[i] + (1.f - beta2) * gr * gr;
    p[i] -= step * (m[i] / (sqrtf(v[i]) / bc2_sqrt + eps));
    if (zero_grad) g[i] = 0.f;
  }
}

}  // namespace avmoe

extern "C" int avmoe_adam_flat(float* p, float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                               float weight_decay, float grad_scale, int64_t step, int zero_grad, avmoe_stream_t stream) {
  using namespace avmoe;
  AVMOE_CHECK_ARG(p && g && m && v && n > 0 && step >= 1, "adam_flat: bad arguments");
  AVMOE_CHECK_ARG(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0, "adam_flat: buffers must be 16-byte aligned");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const long long n4 = n / 4 > 0 ? n / 4 : 1;
  const int threads = 256;
  long long blocks = (n4 + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  adam_flat_kernel<<<(unsigned)blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay,
                                                                                          grad_scale, (float)bc1, (float)sqrt(bc2), zero_grad);
  AVMOE_LAUNCH_CHECK("adam_flat_kernel");
  return 0;
}
