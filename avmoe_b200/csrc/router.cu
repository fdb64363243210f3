// Router of the AVMoE adapter: mean-pooled modalities -> Linear(2C,128) -> ReLU -> Linear(128,32) -> ReLU
// -> Linear(32,E) (+ optional noise) -> softmax -> argmax, all in fp32 with a fixed reduction order so the
// expert selection is reproducible bit-for-bit (AVMOE/AVE/nets/net_trans_v3.py:460-466,472-479; noise:
// AVMOE/AVS/avs_scripts/avs_s4/model/PVT_AVSModel_v2.py:294-296).  One CTA per sample.
// Also: the load-balancing loss quirk  -sum_e log(mean_b p[b,e])  (AVMOE/AVVP/nets/mgn.py:220-224).
#include "common.cuh"

namespace avmoe {

static constexpr int kH1 = 128, kH2 = 32;

// grid BT, block 128.  smem: rin[2C] + h1[128] + h2[32] + logits[E]
__global__ void router_fwd_kernel(const float* __restrict__ px, const float* __restrict__ pv, int chunks_x, int chunks_v, int C, int E,
                                  float inv_n, const float* __restrict__ w0, const float* __restrict__ b0,
                                  const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ w4,
                                  const float* __restrict__ b4, const float* __restrict__ noise, float* __restrict__ probs,
                                  long long* __restrict__ idx, float* __restrict__ h1o, float* __restrict__ h2o, float* __restrict__ rin_o) {
  extern __shared__ float sm[];
  float* rin = sm;
  float* h1 = rin + 2 * C;
  float* h2 = h1 + kH1;
  float* lg = h2 + kH2;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < chunks_x; ++k) s += px[((long long)b * chunks_x + k) * C + c];
    float t = 0.f;
    for (int k = 0; k < chunks_v; ++k) t += pv[((long long)b * chunks_v + k) * C + c];
    rin[c] = s * inv_n;
    rin[C + c] = t * inv_n;
    rin_o[(long long)b * 2 * C + c] = rin[c];
    rin_o[(long long)b * 2 * C + C + c] = rin[C + c];
  }
  __syncthreads();
  for (int j = warp; j < kH1; j += nwarps) {
    const float* wr = w0 + (long long)j * 2 * C;
    float s = 0.f;
    for (int i = lane; i < 2 * C; i += 32) s = fmaf(wr[i], rin[i], s);
    s = warp_sum(s);
    if (lane == 0) { s = fmaxf(s + b0[j], 0.f); h1[j] = s; h1o[(long long)b * kH1 + j] = s; }
  }
  __syncthreads();
  for (int j = warp; j < kH2; j += nwarps) {
    const float* wr = w2 + (long long)j * kH1;
    float s = 0.f;
    for (int i = lane; i < kH1; i += 32) s = fmaf(wr[i], h1[i], s);
    s = warp_sum(s);
    if (lane == 0) { s = fmaxf(s + b2[j], 0.f); h2[j] = s; h2o[(long long)b * kH2 + j] = s; }
  }
  __syncthreads();
  for (int e = warp; e < E; e += nwarps) {
    float s = w4[(long long)e * kH2 + lane] * h2[lane];
    s = warp_sum(s);
    if (lane == 0) lg[e] = s + b4[e] + (noise ? noise[(long long)b * E + e] : 0.f);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float mx = lg[0];
    for (int e = 1; e < E; ++e) mx = fmaxf(mx, lg[e]);
    float sum = 0.f;
    for (int e = 0; e < E; ++e) { lg[e] = expf(lg[e] - mx); sum += lg[e]; }
    int best = 0;
    float bestp = -1.f;
    for (int e = 0; e < E; ++e) {
      const float pr = lg[e] / sum;
      probs[(long long)b * E + e] = pr;
      if (pr > bestp) { bestp = pr; best = e; }   // strict '>' keeps the first maximum, like torch.argmax
    }
    idx[b] = best;
  }
}

// grid BT, block 128: per-sample backward down to d(rin); stores dh1/dh2/dlogits for the weight-gradient kernel
__global__ void router_bwd_sample_kernel(const float* __restrict__ dprobs, const float* __restrict__ probs, const float* __restrict__ h1,
                                         const float* __restrict__ h2, int C, int E, float inv_n, const float* __restrict__ w0,
                                         const float* __restrict__ w2, const float* __restrict__ w4, float* __restrict__ dlg_o,
                                         float* __restrict__ dh2_o, float* __restrict__ dh1_o, float* __restrict__ dpool_x,
                                         float* __restrict__ dpool_v) {
  __shared__ float dlg[AVMOE_MAX_EXPERTS];
  __shared__ float dh2[kH2];
  __shared__ float dh1[kH1];
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    float dot = 0.f;
    for (int e = 0; e < E; ++e) dot += dprobs[(long long)b * E + e] * probs[(long long)b * E + e];
    for (int e = 0; e < E; ++e) {
      const float v = probs[(long long)b * E + e] * (dprobs[(long long)b * E + e] - dot);
      dlg[e] = v;
      dlg_o[(long long)b * E + e] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < kH2) {
    const int j = threadIdx.x;
    float s = 0.f;
    for (int e = 0; e < E; ++e) s = fmaf(dlg[e], w4[(long long)e * kH2 + j], s);
    s = h2[(long long)b * kH2 + j] > 0.f ? s : 0.f;
    dh2[j] = s;
    dh2_o[(long long)b * kH2 + j] = s;
  }
  __syncthreads();
  if (threadIdx.x < kH1) {
    const int j = threadIdx.x;
    float s = 0.f;
    for (int k = 0; k < kH2; ++k) s = fmaf(dh2[k], w2[(long long)k * kH1 + j], s);
    s = h1[(long long)b * kH1 + j] > 0.f ? s : 0.f;
    dh1[j] = s;
    dh1_o[(long long)b * kH1 + j] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < kH1; ++j) s = fmaf(dh1[j], w0[(long long)j * 2 * C + i], s);
    s *= inv_n;
    if (i < C) dpool_x[(long long)b * C + i] = s;
    else dpool_v[(long long)b * C + (i - C)] = s;
  }
}

// dW[j][i] = sum_b dy[b][j] * xin[b][i],  db[j] = sum_b dy[b][j]   (deterministic: fixed 8-way split of b + ordered combine)
// grid (ceil(I/32), J), block (32, 8): x -> input column i, y -> slice of the batch
__global__ void router_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ xin, float* __restrict__ dw,
                                    float* __restrict__ db, int BT, int J, int I) {
  __shared__ float red[8][33];
  const int i = blockIdx.x * 32 + threadIdx.x;
  const int j = blockIdx.y;
  float s = 0.f, sb = 0.f;
  for (int b = threadIdx.y; b < BT; b += 8) {
    const float d = dy[(long long)b * J + j];
    sb += d;
    if (i < I) s = fmaf(d, xin[(long long)b * I + i], s);
  }
  red[threadIdx.y][threadIdx.x] = s;
  if (threadIdx.x == 0) red[threadIdx.y][32] = sb;
  __syncthreads();
  if (threadIdx.y == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    if (i < I) dw[(long long)j * I + i] = t;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      float u = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) u += red[k][32];
      db[j] = u;
    }
  }
}

__global__ void lb_loss_fwd_kernel(const float* __restrict__ probs, float* __restrict__ loss, int BT, int E) {
  __shared__ float pbar[AVMOE_MAX_EXPERTS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < E) {
    float s = 0.f;
    for (int b = lane; b < BT; b += 32) s += probs[(long long)b * E + warp];
    s = warp_sum(s);
    if (lane == 0) pbar[warp] = s / BT;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = 0.f;
    for (int e = 0; e < E; ++e) l -= logf(pbar[e]);
    *loss = l;
  }
}

__global__ void lb_loss_bwd_kernel(const float* __restrict__ probs, const float* __restrict__ dloss, float* __restrict__ dprobs, int BT, int E) {
  __shared__ float coef[AVMOE_MAX_EXPERTS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < E) {
    float s = 0.f;
    for (int b = lane; b < BT; b += 32) s += probs[(long long)b * E + warp];
    s = warp_sum(s);
    if (lane == 0) coef[warp] = -(*dloss) / s;   // -dL / (BT * mean) = -dL / sum
  }
  __syncthreads();
  for (int i = threadIdx.x; i < BT * E; i += blockDim.x) dprobs[i] += coef[i % E];
}

}  // namespace avmoe

using namespace avmoe;

extern "C" int avmoe_router_fwd(const float* pool_x, const float* pool_v, int chunks_x, int chunks_v, int BT, int C, int E, float inv_n,
                                const float* w0, const float* b0, const float* w2, const float* b2, const float* w4, const float* b4,
                                const float* noise, float* probs, int64_t* idx, float* h1, float* h2, float* rin, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(pool_x && pool_v && probs && idx && h1 && h2 && rin && w0 && w2 && w4, "router_fwd: null pointer");
  AVMOE_CHECK_ARG(E >= 1 && E <= AVMOE_MAX_EXPERTS && BT > 0 && C > 0, "router_fwd: bad sizes (E=%d)", E);
  const size_t smem = (2 * (size_t)C + kH1 + kH2 + AVMOE_MAX_EXPERTS) * sizeof(float);
  router_fwd_kernel<<<BT, 512, smem, reinterpret_cast<cudaStream_t>(stream)>>>(pool_x, pool_v, chunks_x, chunks_v, C, E, inv_n, w0, b0, w2, b2,
                                                                                w4, b4, noise, probs, (long long*)idx, h1, h2, rin);
  AVMOE_LAUNCH_CHECK("router_fwd_kernel");
  return 0;
}

extern "C" int avmoe_router_bwd(const float* dprobs, const float* probs, const float* h1, const float* h2, const float* rin, int BT, int C,
                                int E, float inv_n, const float* w0, const float* w2, const float* w4, float* ws /* BT*(E+32+128) */,
                                float* dpool_x, float* dpool_v, float* dw0, float* db0, float* dw2, float* db2, float* dw4, float* db4,
                                avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(dprobs && probs && h1 && h2 && rin && ws && dpool_x && dpool_v && dw0 && db0 && dw2 && db2 && dw4 && db4, "router_bwd: null pointer");
  AVMOE_CHECK_ARG(E >= 1 && E <= AVMOE_MAX_EXPERTS && BT > 0 && C > 0, "router_bwd: bad sizes");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* dlg = ws;
  float* dh2 = dlg + (size_t)BT * E;
  float* dh1 = dh2 + (size_t)BT * kH2;
  router_bwd_sample_kernel<<<BT, 256, 0, st>>>(dprobs, probs, h1, h2, C, E, inv_n, w0, w2, w4, dlg, dh2, dh1, dpool_x, dpool_v);
  AVMOE_LAUNCH_CHECK("router_bwd_sample_kernel");
  const dim3 wblock(32, 8);
  router_wgrad_kernel<<<dim3((2 * C + 31) / 32, kH1), wblock, 0, st>>>(dh1, rin, dw0, db0, BT, kH1, 2 * C);
  router_wgrad_kernel<<<dim3((kH1 + 31) / 32, kH2), wblock, 0, st>>>(dh2, h1, dw2, db2, BT, kH2, kH1);
  router_wgrad_kernel<<<dim3((kH2 + 31) / 32, E), wblock, 0, st>>>(dlg, h2, dw4, db4, BT, E, kH2);
  AVMOE_LAUNCH_CHECK("router_wgrad_kernel");
  return 0;
}

extern "C" int avmoe_lb_loss_fwd(const float* probs, float* loss, int BT, int E, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(probs && loss && E >= 1 && E <= AVMOE_MAX_EXPERTS && BT > 0, "lb_loss_fwd: bad arguments");
  lb_loss_fwd_kernel<<<1, 32 * AVMOE_MAX_EXPERTS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(probs, loss, BT, E);
  AVMOE_LAUNCH_CHECK("lb_loss_fwd_kernel");
  return 0;
}

extern "C" int avmoe_lb_loss_bwd(const float* probs, const float* dloss, float* dprobs, int BT, int E, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(probs && dloss && dprobs && E >= 1 && E <= AVMOE_MAX_EXPERTS && BT > 0, "lb_loss_bwd: bad arguments");
  lb_loss_bwd_kernel<<<1, 32 * AVMOE_MAX_EXPERTS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(probs, dloss, dprobs, BT, E);
  AVMOE_LAUNCH_CHECK("lb_loss_bwd_kernel");
  return 0;
}
