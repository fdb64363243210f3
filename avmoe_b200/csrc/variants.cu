// Single-modal expert pre-transforms of the AVVP and AVS-v1 variants.
//
//  * AVVP (AVMOE/AVVP/nets/mgn.py:132-139): projection-free N x N token self-attention
//        S = x xᵀ,  att = softmax_rows(S),  aux_j = sum_i att[i,j] x_i
//    The two N x N contractions run on the tcgen05 batched GEMM (ops.bgemm); the kernels here are the row softmax
//    (forward) and the fused probability / softmax-backward pass (backward) in between.
//  * AVS-v1 (AVMOE/AVS/avs_scripts/avs_s4/model/PVT_AVSModel_v2.py:212-215): nn.MultiheadAttention(C, 4 heads) called
//    with batch_first=False on [BT,N,C], i.e. attention ALONG THE BT AXIS, batched over the N token positions.
//    in_proj / out_proj are tcgen05 GEMMs; the kernels here are the per-(token position, head) L x L attention core
//    forward and backward (L = BT <= 256), probabilities recomputed in the backward.
#include <type_traits>

#include "block_mma.cuh"
#include "common.cuh"

namespace avmoe {

// ------------------------------------------------------------------------------------------------ AVVP: rows of S
// one warp per row: lse = logsumexp(S[row,:]); P[row,:] = exp(S - lse)
template <typename TP>
__global__ void nxn_softmax_rows_kernel(const float* __restrict__ S, TP* __restrict__ P, float* __restrict__ lse, long long rows, int n) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const float* s = S + row * n;
  float mx = -INFINITY;
  for (int j = lane; j < n; j += 32) mx = fmaxf(mx, s[j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < n; j += 32) sum += expf(s[j] - mx);
  sum = warp_sum(sum);
  const float l = mx + logf(sum);
  if (lane == 0) lse[row] = l;
  TP* p = P + row * n;
  for (int j = lane; j < n; j += 32) p[j] = from_f32<TP>(expf(s[j] - l));
}

// backward rows: att = exp(S - lse); r = sum_j att*datt; ds = att*(datt - r)
template <typename TP>
__global__ void nxn_bwd_rows_kernel(const float* __restrict__ S, const float* __restrict__ dAtt, const float* __restrict__ lse,
                                    TP* __restrict__ att, TP* __restrict__ dS, long long rows, int n) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const float* s = S + row * n;
  const float* da = dAtt + row * n;
  const float l = lse[row];
  float r = 0.f;
  for (int j = lane; j < n; j += 32) r = fmaf(expf(s[j] - l), da[j], r);
  r = warp_sum(r);
  for (int j = lane; j < n; j += 32) {
    const float a = expf(s[j] - l);
    att[row * n + j] = from_f32<TP>(a);
    dS[row * n + j] = from_f32<TP>(a * (da[j] - r));
  }
}

// dst[i] += a[i] + b[i] + c[i]   (b, c optional)
template <typename T>
__global__ void add_into_kernel(T* __restrict__ dst, const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ c, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = to_f32(dst[i]) + to_f32(a[i]);
    if (b) v += to_f32(b[i]);
    if (c) v += to_f32(c[i]);
    dst[i] = from_f32<T>(v);
  }
}

// ------------------------------------------------------------------------------------------------ AVS-v1: attention core
// qkv [L][NB][3C] fp32 (the in_proj GEMM writes fp32), heads H, hd=C/H.  grid (NB, H), block 256.
// q/k/v (and dctx) of one (token position, head) are staged in shared memory as [L][hd+4] when they fit, otherwise the
// contractions read them in place from global memory (row pitch NB*3C).  S lives in shared memory as [L][L+4].
template <typename T>
__global__ void __launch_bounds__(256) mha_attn_fwd_kernel(const float* __restrict__ qkv, T* __restrict__ ctx, int L, int NB, int C, int hd,
                                                           float scale, int staged, const unsigned char* __restrict__ keep, float keep_scale) {
  extern __shared__ float sm[];
  constexpr bool kExact = std::is_same<T, float>::value;
  const int ldh = hd + 4, lds = L + 4;
  float* S = sm;
  const int nb = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float* gq = qkv + (long long)nb * 3 * C + h * hd;
  const int grs = NB * 3 * C;
  const float *q = gq, *k = gq + C, *v = gq + 2 * C;
  int rs = grs;
  if (staged) {
    float* sq = S + L * lds;
    float* sk = sq + L * ldh;
    float* sv = sk + L * ldh;
    for (int l = warp; l < L; l += nwarps)
      for (int d = lane; d < hd; d += 32) {
        sq[l * ldh + d] = gq[(long long)l * grs + d];
        sk[l * ldh + d] = gq[(long long)l * grs + C + d];
        sv[l * ldh + d] = gq[(long long)l * grs + 2 * C + d];
      }
    q = sq; k = sk; v = sv; rs = ldh;
    __syncthreads();
  }
  contract<kExact>(L, L, hd, q, rs, 1, k, 1, rs, [&](int i, int j, float val) { S[i * lds + j] = val * scale; });
  __syncthreads();
  for (int i = warp; i < L; i += nwarps) {
    float mx = -INFINITY;
    for (int j = lane; j < L; j += 32) mx = fmaxf(mx, S[i * lds + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) { const float e = expf(S[i * lds + j] - mx); S[i * lds + j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    if (keep) {    // attention-probability dropout (nn.MultiheadAttention dropout): kept entries scaled by 1/(1-p)
      const unsigned char* kr = keep + (((long long)nb * gridDim.y + h) * L + i) * L;
      for (int j = lane; j < L; j += 32) S[i * lds + j] = kr[j] ? S[i * lds + j] * inv * keep_scale : 0.f;
    } else {
      for (int j = lane; j < L; j += 32) S[i * lds + j] *= inv;
    }
  }
  __syncthreads();
  contract<kExact>(L, hd, L, S, lds, 1, v, rs, 1, [&](int i, int d, float val) {
    ctx[((long long)i * NB + nb) * C + h * hd + d] = from_f32<T>(val);
  });
}

// backward: recompute att; dv = attᵀ dctx; datt = dctx vᵀ; ds = att*(datt - rowsum(att*datt)); dq = scale * ds k; dk = scale * dsᵀ q
template <typename T>
__global__ void __launch_bounds__(256) mha_attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dctx, T* __restrict__ dqkv,
                                                           int L, int NB, int C, int hd, float scale, int staged,
                                                           const unsigned char* __restrict__ keep, float keep_scale) {
  extern __shared__ float sm[];
  constexpr bool kExact = std::is_same<T, float>::value;
  const int ldh = hd + 4, lds = L + 4;
  float* A = sm;                 // att
  float* D = A + L * lds;        // datt -> ds
  const int nb = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float* gq = qkv + (long long)nb * 3 * C + h * hd;
  const float* gg = dctx + (long long)nb * C + h * hd;
  const int grs = NB * 3 * C, ggrs = NB * C;
  const float *q = gq, *k = gq + C, *v = gq + 2 * C, *g = gg;
  int rs = grs, rsg = ggrs;
  if (staged) {
    float* sq = D + L * lds;
    float* sk = sq + L * ldh;
    float* sv = sk + L * ldh;
    float* sg = sv + L * ldh;
    for (int l = warp; l < L; l += nwarps)
      for (int d = lane; d < hd; d += 32) {
        sq[l * ldh + d] = gq[(long long)l * grs + d];
        sk[l * ldh + d] = gq[(long long)l * grs + C + d];
        sv[l * ldh + d] = gq[(long long)l * grs + 2 * C + d];
        sg[l * ldh + d] = gg[(long long)l * ggrs + d];
      }
    q = sq; k = sk; v = sv; g = sg; rs = ldh; rsg = ldh;
    __syncthreads();
  }
  contract<kExact>(L, L, hd, q, rs, 1, k, 1, rs, [&](int i, int j, float val) { A[i * lds + j] = val * scale; });
  contract<kExact>(L, L, hd, g, rsg, 1, v, 1, rs, [&](int i, int j, float val) { D[i * lds + j] = val; });
  __syncthreads();
  for (int i = warp; i < L; i += nwarps) {
    float mx = -INFINITY;
    for (int j = lane; j < L; j += 32) mx = fmaxf(mx, A[i * lds + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) { const float e = expf(A[i * lds + j] - mx); A[i * lds + j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    float r = 0.f;
    const unsigned char* kr = keep ? keep + (((long long)nb * gridDim.y + h) * L + i) * L : nullptr;
    // with dropout the context used att*m (m = keep/(1-p)): the gradient at the un-dropped probabilities is datt*m
    for (int j = lane; j < L; j += 32) {
      const float a = A[i * lds + j] * inv;
      const float m = kr ? (kr[j] ? keep_scale : 0.f) : 1.f;
      const float d = D[i * lds + j] * m;
      A[i * lds + j] = a; D[i * lds + j] = d;
      r = fmaf(a, d, r);
    }
    r = warp_sum(r);
    for (int j = lane; j < L; j += 32) {
      const float a = A[i * lds + j];
      D[i * lds + j] = a * (D[i * lds + j] - r);
      if (kr) A[i * lds + j] = kr[j] ? a * keep_scale : 0.f;      // dv contracts the dropped probabilities
    }
  }
  __syncthreads();
  T* out = dqkv + (long long)nb * 3 * C + h * hd;
  const long long ors = (long long)NB * 3 * C;
  contract<kExact>(L, hd, L, D, lds, 1, k, rs, 1, [&](int i, int d, float val) { out[i * ors + d] = from_f32<T>(val * scale); });
  contract<kExact>(L, hd, L, D, 1, lds, q, rs, 1, [&](int j, int d, float val) { out[j * ors + C + d] = from_f32<T>(val * scale); });
  contract<kExact>(L, hd, L, A, 1, lds, g, rsg, 1, [&](int j, int d, float val) { out[j * ors + 2 * C + d] = from_f32<T>(val); });
}

}  // namespace avmoe

using namespace avmoe;

extern "C" int avmoe_nxn_softmax_rows(const float* S, void* P, float* lse, int64_t rows, int n, int p_dtype, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(S && P && lse && rows > 0 && n > 0, "nxn_softmax_rows: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (p_dtype == AVMOE_F32) nxn_softmax_rows_kernel<float><<<grid, 256, 0, st>>>(S, (float*)P, lse, rows, n);
  else nxn_softmax_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(S, (__nv_bfloat16*)P, lse, rows, n);
  AVMOE_LAUNCH_CHECK("nxn_softmax_rows_kernel");
  return 0;
}

extern "C" int avmoe_nxn_bwd_rows(const float* S, const float* dAtt, const float* lse, void* att, void* dS, int64_t rows, int n, int p_dtype,
                                  avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(S && dAtt && lse && att && dS && rows > 0 && n > 0, "nxn_bwd_rows: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (p_dtype == AVMOE_F32) nxn_bwd_rows_kernel<float><<<grid, 256, 0, st>>>(S, dAtt, lse, (float*)att, (float*)dS, rows, n);
  else nxn_bwd_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(S, dAtt, lse, (__nv_bfloat16*)att, (__nv_bfloat16*)dS, rows, n);
  AVMOE_LAUNCH_CHECK("nxn_bwd_rows_kernel");
  return 0;
}

extern "C" int avmoe_add_into(void* dst, const void* a, const void* b, const void* c, int64_t n, int dtype, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(dst && a && n > 0, "add_into: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (int)min((long long)(8 * num_sms()), (long long)((n + 255) / 256));
  if (dtype == AVMOE_F32) add_into_kernel<float><<<grid, 256, 0, st>>>((float*)dst, (const float*)a, (const float*)b, (const float*)c, n);
  else add_into_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16*)dst, (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (const __nv_bfloat16*)c, n);
  AVMOE_LAUNCH_CHECK("add_into_kernel");
  return 0;
}

extern "C" int avmoe_mha_attn_fwd(const float* qkv, void* ctx, int L, int NB, int C, int heads, int dtype, const unsigned char* keep,
                                  float keep_scale, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(qkv && ctx && L > 0 && NB > 0 && C > 0 && heads > 0 && C % heads == 0, "mha_attn_fwd: bad arguments");
  AVMOE_CHECK_ARG(heads <= 65535, "mha_attn_fwd: grid too large");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int hd = C / heads;
  const size_t base = sizeof(float) * (size_t)L * (L + 4);
  const size_t full = base + sizeof(float) * 3 * (size_t)L * (hd + 4);
  AVMOE_CHECK_ARG(base <= 226 * 1024, "mha_attn_fwd: attended axis BT=%d too long for one CTA", L);
  const int staged = full <= 226 * 1024;
  const size_t smem = staged ? full : base;
  const float scale = 1.f / sqrtf((float)hd);
  dim3 grid(NB, heads);
  if (dtype == AVMOE_F32) {
    AVMOE_CUDA(cudaFuncSetAttribute(mha_attn_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    mha_attn_fwd_kernel<float><<<grid, 256, smem, st>>>(qkv, (float*)ctx, L, NB, C, hd, scale, staged, keep, keep_scale);
  } else {
    AVMOE_CUDA(cudaFuncSetAttribute(mha_attn_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    mha_attn_fwd_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(qkv, (__nv_bfloat16*)ctx, L, NB, C, hd, scale, staged, keep, keep_scale);
  }
  AVMOE_LAUNCH_CHECK("mha_attn_fwd_kernel");
  return 0;
}

extern "C" int avmoe_mha_attn_bwd(const float* qkv, const float* dctx, void* dqkv, int L, int NB, int C, int heads, int dtype,
                                  const unsigned char* keep, float keep_scale, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(qkv && dctx && dqkv && L > 0 && NB > 0 && C > 0 && heads > 0 && C % heads == 0, "mha_attn_bwd: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int hd = C / heads;
  const size_t base = sizeof(float) * 2 * (size_t)L * (L + 4);
  const size_t full = base + sizeof(float) * 4 * (size_t)L * (hd + 4);
  AVMOE_CHECK_ARG(base <= 226 * 1024, "mha_attn_bwd: attended axis BT=%d too long for one CTA", L);
  const int staged = full <= 226 * 1024;
  const size_t smem = staged ? full : base;
  const float scale = 1.f / sqrtf((float)hd);
  dim3 grid(NB, heads);
  if (dtype == AVMOE_F32) {
    AVMOE_CUDA(cudaFuncSetAttribute(mha_attn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    mha_attn_bwd_kernel<float><<<grid, 256, smem, st>>>(qkv, dctx, (float*)dqkv, L, NB, C, hd, scale, staged, keep, keep_scale);
  } else {
    AVMOE_CUDA(cudaFuncSetAttribute(mha_attn_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    mha_attn_bwd_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(qkv, dctx, (__nv_bfloat16*)dqkv, L, NB, C, hd, scale, staged, keep, keep_scale);
  }
  AVMOE_LAUNCH_CHECK("mha_attn_bwd_kernel");
  return 0;
}
