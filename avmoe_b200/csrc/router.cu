// Router of the AVMoE adapter: mean-pooled modalities -> Linear(2C,128) -> ReLU -> Linear(128,32) -> ReLU
// -> Linear(32,E) (+ optional noise) -> softmax -> argmax, all in fp32 with a fixed reduction order so the
// expert selection is reproducible bit-for-bit (AVMOE/AVE/nets/net_trans_v3.py:460-466,472-479; noise:
// AVMOE/AVS/avs_scripts/avs_s4/model/PVT_AVSModel_v2.py:294-296).  The wide first layer is tiled over samples.
// Also: the load-balancing loss quirk  -sum_e log(mean_b p[b,e])  (AVMOE/AVVP/nets/mgn.py:220-224).
#include <atomic>
#include "common.cuh"

namespace avmoe {

static constexpr int kH1 = 128, kH2 = 32;

// ---- forward, layer 1: h1 = relu(rin · W0ᵀ + b0) for a tile of kSB samples x kHB hidden units per CTA.
// One CTA per sample (the first version) made every CTA stream all of W0 (0.4-1.6 MB fp32) from L2 behind dependent
// load->FMA chains; here a CTA reads only its kHB rows of W0 and reuses each weight for kSB samples held in shared memory.
// Each (sample, unit) sum keeps the same order as before -- lane-strided over the inputs, then the warp tree -- so the logits,
// and with them the selected expert, do not depend on the tiling.
// grid (ceil(BT/kSB), kH1/kHB), block 256 (8 warps x 4 units).  smem: rin[kSB][2C]
static constexpr int kSB = 8, kHB = 32;
__global__ void __launch_bounds__(256) router_l1_kernel(const float* __restrict__ px, const float* __restrict__ pv, int chunks_x, int chunks_v, int BT, int C,
                                                        float inv_n, const float* __restrict__ w0, const float* __restrict__ b0,
                                                        float* __restrict__ h1o, float* __restrict__ rin_o) {
  extern __shared__ float rin[];          // [kSB][2C]
  const int b0i = blockIdx.x * kSB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C2 = 2 * C;
  for (int i = threadIdx.x; i < kSB * C2; i += blockDim.x) {
    const int s = i / C2, c2 = i - s * C2;
    const int b = b0i + s;
    float v = 0.f;
    if (b < BT) {
      float t = 0.f;
      if (c2 < C) { for (int k = 0; k < chunks_x; ++k) t += px[((long long)b * chunks_x + k) * C + c2]; }
      else        { for (int k = 0; k < chunks_v; ++k) t += pv[((long long)b * chunks_v + k) * C + (c2 - C)]; }
      v = t * inv_n;
      if (blockIdx.y == 0) rin_o[(long long)b * C2 + c2] = v;
    }
    rin[i] = v;
  }
  __syncthreads();
  constexpr int kU = 4;                    // hidden units per warp
  const int j0 = blockIdx.y * kHB + warp * kU;
  float acc[kU][kSB];
#pragma unroll
  for (int u = 0; u < kU; ++u)
#pragma unroll
    for (int q = 0; q < kSB; ++q) acc[u][q] = 0.f;
  const float* wr = w0 + (long long)j0 * C2;
  for (int i = lane; i < C2; i += 32) {
    float w[kU], r[kSB];
#pragma unroll
    for (int u = 0; u < kU; ++u) w[u] = __ldg(wr + (long long)u * C2 + i);
#pragma unroll
    for (int q = 0; q < kSB; ++q) r[q] = rin[q * C2 + i];
#pragma unroll
    for (int u = 0; u < kU; ++u)
#pragma unroll
      for (int q = 0; q < kSB; ++q) acc[u][q] = fmaf(w[u], r[q], acc[u][q]);
  }
#pragma unroll
  for (int u = 0; u < kU; ++u)
#pragma unroll
    for (int q = 0; q < kSB; ++q) {
      const float t = warp_sum(acc[u][q]);
      if (lane == 0 && b0i + q < BT) h1o[(long long)(b0i + q) * kH1 + j0 + u] = fmaxf(t + b0[j0 + u], 0.f);
    }
}

// ---- forward, layers 2-3 + softmax + argmax: grid BT, block 128.
__global__ void __launch_bounds__(128) router_l23_kernel(int E, const float* __restrict__ h1g, const float* __restrict__ w2, const float* __restrict__ b2,
                                                         const float* __restrict__ w4, const float* __restrict__ b4, const float* __restrict__ noise,
                                                         float* __restrict__ probs, long long* __restrict__ idx, float* __restrict__ h2o) {
  __shared__ float h1[kH1];
  __shared__ float h2[kH2];
  __shared__ float lg[AVMOE_MAX_EXPERTS];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < kH1; i += blockDim.x) h1[i] = h1g[(long long)b * kH1 + i];
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

// ---- backward, per sample down to d(h1): grid BT, block 128; stores dh1/dh2/dlogits for the weight-gradient kernel
__global__ void __launch_bounds__(128) router_bwd_sample_kernel(const float* __restrict__ dprobs, const float* __restrict__ probs, const float* __restrict__ h1,
                                                                const float* __restrict__ h2, int E, const float* __restrict__ w2,
                                                                const float* __restrict__ w4, float* __restrict__ dlg_o, float* __restrict__ dh2_o,
                                                                float* __restrict__ dh1_o) {
  __shared__ float dlg[AVMOE_MAX_EXPERTS];
  __shared__ float dh2[kH2];
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
    dh1_o[(long long)b * kH1 + j] = s;
  }
}

// ---- backward, d(pooled input) = dh1 · W0 / n for a tile of kSB samples x 256 input columns per CTA (W0 read once per sample
// tile, coalesced along the inputs; each column's sum runs over the 128 hidden units in order, as before)
// grid (ceil(BT/kSB), ceil(2C/256)), block 256
__global__ void __launch_bounds__(256) router_bwd_dpool_kernel(const float* __restrict__ dh1g, int BT, int C, float inv_n, const float* __restrict__ w0,
                                                               float* __restrict__ dpool_x, float* __restrict__ dpool_v) {
  __shared__ float dh1[kSB][kH1];
  const int b0i = blockIdx.x * kSB;
  const int C2 = 2 * C;
  for (int i = threadIdx.x; i < kSB * kH1; i += blockDim.x) {
    const int q = i / kH1, j = i - q * kH1;
    dh1[q][j] = b0i + q < BT ? dh1g[(long long)(b0i + q) * kH1 + j] : 0.f;
  }
  __syncthreads();
  const int i = blockIdx.y * blockDim.x + threadIdx.x;
  if (i >= C2) return;
  float acc[kSB];
#pragma unroll
  for (int q = 0; q < kSB; ++q) acc[q] = 0.f;
  constexpr int kJ = 8;                    // weight loads in flight per thread
  for (int j0 = 0; j0 < kH1; j0 += kJ) {
    float w[kJ];
#pragma unroll
    for (int u = 0; u < kJ; ++u) w[u] = __ldg(w0 + (long long)(j0 + u) * C2 + i);
#pragma unroll
    for (int u = 0; u < kJ; ++u)
#pragma unroll
      for (int q = 0; q < kSB; ++q) acc[q] = fmaf(dh1[q][j0 + u], w[u], acc[q]);
  }
#pragma unroll
  for (int q = 0; q < kSB; ++q) {
    const int b = b0i + q;
    if (b < BT) {
      const float v = acc[q] * inv_n;
      if (i < C) dpool_x[(long long)b * C + i] = v;
      else dpool_v[(long long)b * C + (i - C)] = v;
    }
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
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem = (size_t)kSB * 2 * C * sizeof(float);
  AVMOE_CHECK_ARG(smem <= 200 * 1024, "router_fwd: C too large (%d)", C);
  if (smem > 48 * 1024) {
    static std::atomic<int> max_set[64];
    int dev = 0;
    AVMOE_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || max_set[dev].load(std::memory_order_acquire) < (int)smem) {
      AVMOE_CUDA(cudaFuncSetAttribute(router_l1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      if (dev >= 0 && dev < 64) max_set[dev].store(200 * 1024, std::memory_order_release);
    }
  }
  router_l1_kernel<<<dim3((BT + kSB - 1) / kSB, kH1 / kHB), 256, smem, st>>>(pool_x, pool_v, chunks_x, chunks_v, BT, C, inv_n, w0, b0, h1, rin);
  AVMOE_LAUNCH_CHECK("router_l1_kernel");
  router_l23_kernel<<<BT, 128, 0, st>>>(E, h1, w2, b2, w4, b4, noise, probs, (long long*)idx, h2);
  AVMOE_LAUNCH_CHECK("router_l23_kernel");
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
  router_bwd_sample_kernel<<<BT, 128, 0, st>>>(dprobs, probs, h1, h2, E, w2, w4, dlg, dh2, dh1);
  AVMOE_LAUNCH_CHECK("router_bwd_sample_kernel");
  router_bwd_dpool_kernel<<<dim3((BT + kSB - 1) / kSB, (2 * C + 255) / 256), 256, 0, st>>>(dh1, BT, C, inv_n, w0, dpool_x, dpool_v);
  AVMOE_LAUNCH_CHECK("router_bwd_dpool_kernel");
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
