// Softmax / elementwise kernels between the tcgen05 GEMMs of the GEMM-based latent-attention path.
//
// The latent cross-attention of a multimodal expert (AVMOE/AVE/nets/net_trans_v3.py:379-390) is four small
// contractions per direction with T <= 87 latent tokens on one side.  On B200 the batched tcgen05 GEMM runs them at
// multi-TB/s effective bandwidth, so the path is laid out as   GEMM -> softmax kernel -> GEMM   with token-major
// logits  S1ᵀ [BT,Nk,T]  (tokens<-kv, softmax over the Nk axis = per COLUMN) and  S2 [BT,N,T]  (token<-latent, softmax
// over T = per ROW).  These kernels are pure streaming passes over those [rows,T] matrices.
#include "common.cuh"

namespace avmoe {

// Logits produced against a bf16 operand that was split into hi + lo halves (ops.split_operand) arrive as two column
// blocks of the fp32 GEMM output: column t and column off_lo + t.  off_lo == 0 means a single block.
__device__ __forceinline__ float logit(const float* row, int t, int off_lo) { return off_lo ? row[t] + row[off_lo + t] : row[t]; }

// ---- column softmax over the Nk rows of each sample: S [BT][Nk][T] fp32 -> P [BT][Nk][ldp] (first T columns), lse [BT][T]
// grid BT, block 1024 (32 warps): lanes over t (T <= 128 -> up to 4 columns per lane), warps over rows.
template <typename TP>
__global__ void __launch_bounds__(1024) softmax_cols_kernel(const float* __restrict__ S, TP* __restrict__ P, float* __restrict__ lse, int Nk, int T,
                                                            int lds, int ldp, int off_lo, int p_lo) {
  __shared__ float smax[32][128];
  __shared__ float ssum[32][128];
  __shared__ float flse[128];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* s = S + (long long)b * Nk * lds;
  // Each warp walks its rows in groups of kU with all kU row loads issued before any is used: one 128-byte request per warp
  // in flight (the plain loop) leaves a 1024-thread CTA at a few GB/s; kU independent requests per warp hide the latency.
  constexpr int kU = 8, kU2 = 4;   // pass 2 also holds store addresses: 4 rows keep it inside the 64-register budget
  float mx[4], sm[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { mx[k] = -INFINITY; sm[k] = 0.f; }
  const int nk4 = (T + 31) >> 5;      // column blocks of 32 this sample uses (warp-uniform)
  for (int r0 = warp; r0 < Nk; r0 += 32 * kU) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < nk4) {
        const int t = lane + 32 * k;
        float v[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int r = r0 + 32 * u;
          v[u] = (t < T && r < Nk) ? logit(s + (long long)r * lds, t, off_lo) : -INFINITY;
        }
        float m2 = mx[k];
#pragma unroll
        for (int u = 0; u < kU; ++u) m2 = fmaxf(m2, v[u]);
        if (m2 > -INFINITY) {
          float acc = sm[k] * expf(mx[k] - m2);
#pragma unroll
          for (int u = 0; u < kU; ++u) acc += expf(v[u] - m2);     // exp(-inf) = 0 for the padded rows / columns
          sm[k] = acc;
          mx[k] = m2;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { smax[warp][lane + 32 * k] = mx[k]; ssum[warp][lane + 32 * k] = sm[k]; }
  __syncthreads();
  if (threadIdx.x < T) {
    const int t = threadIdx.x;
    float M = -INFINITY;
    for (int w = 0; w < 32; ++w) M = fmaxf(M, smax[w][t]);
    float L = 0.f;
    for (int w = 0; w < 32; ++w) L += ssum[w][t] > 0.f ? ssum[w][t] * expf(smax[w][t] - M) : 0.f;
    const float l = M + logf(L);
    flse[t] = l;
    lse[(long long)b * T + t] = l;
  }
  __syncthreads();
  TP* p = P + (long long)b * Nk * ldp;
  for (int r0 = warp; r0 < Nk; r0 += 32 * kU2) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int t = lane + 32 * k;
      if (t < T) {
        const float l = flse[t];
        float v[kU2];
#pragma unroll
        for (int u = 0; u < kU2; ++u) {
          const int r = r0 + 32 * u;
          v[u] = r < Nk ? logit(s + (long long)r * lds, t, off_lo) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kU2; ++u) {
          const int r = r0 + 32 * u;
          if (r < Nk) {
            const float pv = expf(v[u] - l);
            const TP hi = from_f32<TP>(pv);
            p[(long long)r * ldp + t] = hi;
            if (p_lo) p[(long long)r * ldp + p_lo + t] = from_f32<TP>(pv - to_f32(hi));   // residual: R = (hi + lo)·kv keeps ~16 bits of A1
          }
        }
      }
    }
  }
}

// ---- dS1ᵀ = A1ᵀ * (dA1ᵀ - D[b,t]);  A1ᵀ lives in the first T columns of A12 [BT][Nk][2T], dS1ᵀ is written to the second T
template <typename TP>
__global__ void softmax_cols_bwd_kernel(TP* __restrict__ A12, const float* __restrict__ dA, const float* __restrict__ Dt, long long rows, int Nk, int T,
                                        int ld12, int off2, int ldd, int off_lo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const long long b = row / Nk;
  TP* a = A12 + row * ld12;
  const float* d = dA + row * ldd;
  const float* dt = Dt + b * T;
  for (int t = lane; t < T; t += 32) a[off2 + t] = from_f32<TP>(to_f32(a[t]) * (logit(d, t, off_lo) - dt[t]));
}

// ---- row softmax over T: S [rows][T] fp32 -> A [rows][T]
template <typename TP>
__global__ void softmax_rows_kernel(const float* __restrict__ S, TP* __restrict__ A, long long rows, int T, int lds, int off_lo, int ld) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const float* s = S + row * lds;
  float mx = -INFINITY;
  for (int t = lane; t < T; t += 32) mx = fmaxf(mx, logit(s, t, off_lo));
  mx = warp_max(mx);
  float sum = 0.f;
  for (int t = lane; t < T; t += 32) sum += expf(logit(s, t, off_lo) - mx);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int t = lane; t < ld; t += 32) A[row * ld + t] = from_f32<TP>(t < T ? expf(logit(s, t, off_lo) - mx) * inv : 0.f);   // padding columns are zero
}

// ---- dS = A * (dA - sum_t A*dA)
template <typename TP>
__global__ void softmax_rows_bwd_kernel(const TP* __restrict__ A, const float* __restrict__ dA, TP* __restrict__ dS, long long rows, int T, int ld, int ldd,
                                        int off_lo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const float* d = dA + row * ldd;
  float r = 0.f;
  for (int t = lane; t < T; t += 32) r = fmaf(to_f32(A[row * ld + t]), logit(d, t, off_lo), r);
  r = warp_sum(r);
  for (int t = lane; t < ld; t += 32) dS[row * ld + t] = from_f32<TP>(t < T ? to_f32(A[row * ld + t]) * (logit(d, t, off_lo) - r) : 0.f);
}

// ---- out[b][t][c] = in[b][t][c] (+ in[b][off_lo + t][c]) + tile[t][c]   (fp32; `in` has `rows` rows per sample)
__global__ void tile_add_kernel(const float* __restrict__ in, const float* __restrict__ tile, float* __restrict__ out, int T, int C, int rows, int off_lo,
                                long long total) {
  const long long L = (long long)T * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / L, r = i - b * L;
    const float* src = in + b * rows * C + r;
    float v = src[0] + tile[r];
    if (off_lo) v += src[(long long)off_lo * C];
    out[i] = v;
  }
}

// ---- Dt[b][t] = sum_c dR[b][t][c] * (tokp[b][t][c] - tok[t][c])   (one warp per (b,t))
__global__ void latent_dots_kernel(const float* __restrict__ dR, const float* __restrict__ tokp, const float* __restrict__ tok, float* __restrict__ Dt,
                                   long long rows, int T, int C) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const int t = (int)(row % T);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s = fmaf(dR[row * C + c], tokp[row * C + c] - tok[(long long)t * C + c], s);
  s = warp_sum(s);
  if (lane == 0) Dt[row] = s;
}

// ---- d[b][n][c] += vec[b][c]
template <typename T>
__global__ void add_rowvec_kernel(T* __restrict__ d, const float* __restrict__ vec, int N, int C) {
  const int b = blockIdx.y;
  const long long total = (long long)N * C;
  T* db = d + (long long)b * total;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    db[i] = from_f32<T>(to_f32(db[i]) + vec[(long long)b * C + (int)(i % C)]);
}

// ---- fp32 [G][T][C] -> bf16 [G][2*Tp][C]: rows [0,T) hold hi = bf16(v), rows [Tp,Tp+T) hold lo = bf16(v - hi)
__global__ void cast_split_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long total, int T, int Tp, int C) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long gt = i / C;
    const int t = (int)(gt % T);
    const long long g = gt / T;
    const float v = in[i];
    const __nv_bfloat16 hi = __float2bfloat16(v);
    const long long o = (g * 2 * Tp + t) * C + c;
    out[o] = hi;
    out[o + (long long)Tp * C] = __float2bfloat16(v - __bfloat162float(hi));
  }
}

// ---- out[i] += sum_b in[b][i]
// block (32, 8): x -> element, y -> slice of the batch; the 8 partial sums are combined in a fixed order (deterministic)
__global__ void batch_sum_add_kernel(const float* __restrict__ in, float* __restrict__ out, int BT, long long L) {
  __shared__ float red[8][33];
  const long long i = blockIdx.x * 32LL + threadIdx.x;
  float s = 0.f;
  if (i < L) {
#pragma unroll 4
    for (int b = threadIdx.y; b < BT; b += 8) s += in[(long long)b * L + i];
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && i < L) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    out[i] += t;
  }
}

}  // namespace avmoe

using namespace avmoe;

#define DISPATCH_TP(dt, ...)                                              \
  if ((dt) == AVMOE_F32) { using TP = float; __VA_ARGS__ }                \
  else { using TP = __nv_bfloat16; __VA_ARGS__ }

extern "C" int avmoe_softmax_cols(const float* S, void* P, float* lse, int BT, int Nk, int T, int lds, int ldp, int off_lo, int p_lo, int p_dtype,
                              avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(S && P && lse && BT > 0 && Nk > 0 && T > 0 && T <= 128 && ldp >= T && lds >= T + off_lo && (off_lo == 0 || off_lo >= T) &&
                      (p_lo == 0 || (p_lo >= T && ldp >= p_lo + T)),
                  "softmax_cols: bad arguments (T <= 128)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DISPATCH_TP(p_dtype, softmax_cols_kernel<TP><<<BT, 1024, 0, st>>>(S, (TP*)P, lse, Nk, T, lds, ldp, off_lo, p_lo);)
  AVMOE_LAUNCH_CHECK("softmax_cols_kernel");
  return 0;
}

extern "C" int avmoe_softmax_cols_bwd(void* A12, const float* dA, const float* Dt, int BT, int Nk, int T, int ld12, int off2, int ldd, int off_lo,
                                      int p_dtype, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(A12 && dA && Dt && BT > 0 && Nk > 0 && T > 0 && off2 >= T && ld12 >= off2 + T && ldd >= T + off_lo && (off_lo == 0 || off_lo >= T),
                  "softmax_cols_bwd: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long rows = (long long)BT * Nk;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  DISPATCH_TP(p_dtype, softmax_cols_bwd_kernel<TP><<<grid, 256, 0, st>>>((TP*)A12, dA, Dt, rows, Nk, T, ld12, off2, ldd, off_lo);)
  AVMOE_LAUNCH_CHECK("softmax_cols_bwd_kernel");
  return 0;
}

extern "C" int avmoe_softmax_rows(const float* S, void* A, int64_t rows, int T, int lds, int off_lo, int ld, int p_dtype, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(S && A && rows > 0 && T > 0 && ld >= T && lds >= T + off_lo && (off_lo == 0 || off_lo >= T), "softmax_rows: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  DISPATCH_TP(p_dtype, softmax_rows_kernel<TP><<<grid, 256, 0, st>>>(S, (TP*)A, rows, T, lds, off_lo, ld);)
  AVMOE_LAUNCH_CHECK("softmax_rows_kernel");
  return 0;
}

extern "C" int avmoe_softmax_rows_bwd(const void* A, const float* dA, void* dS, int64_t rows, int T, int ld, int ldd, int off_lo, int p_dtype,
                                      avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(A && dA && dS && rows > 0 && T > 0 && ld >= T && ldd >= T + off_lo && (off_lo == 0 || off_lo >= T), "softmax_rows_bwd: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  DISPATCH_TP(p_dtype, softmax_rows_bwd_kernel<TP><<<grid, 256, 0, st>>>((const TP*)A, dA, (TP*)dS, rows, T, ld, ldd, off_lo);)
  AVMOE_LAUNCH_CHECK("softmax_rows_bwd_kernel");
  return 0;
}

extern "C" int avmoe_tile_add(const float* in, const float* tile, float* out, int BT, int T, int C, int rows, int off_lo, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(in && tile && out && BT > 0 && T > 0 && C > 0 && rows >= T + off_lo && (off_lo == 0 || off_lo >= T), "tile_add: bad arguments");
  const long long total = (long long)BT * T * C;
  const int grid = (int)min((long long)(8 * num_sms()), (total + 255) / 256);
  tile_add_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, tile, out, T, C, rows, off_lo, total);
  AVMOE_LAUNCH_CHECK("tile_add_kernel");
  return 0;
}

extern "C" int avmoe_latent_dots(const float* dR, const float* tokp, const float* tok, float* Dt, int BT, int T, int C, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(dR && tokp && tok && Dt && BT > 0 && T > 0 && C > 0, "latent_dots: bad arguments");
  const long long rows = (long long)BT * T;
  latent_dots_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dR, tokp, tok, Dt, rows, T, C);
  AVMOE_LAUNCH_CHECK("latent_dots_kernel");
  return 0;
}

extern "C" int avmoe_add_rowvec(void* d, const float* vec, int BT, int N, int C, int dtype, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(d && vec && BT > 0 && N > 0 && C > 0, "add_rowvec: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)min((long long)64, ((long long)N * C + 255) / 256), BT);
  if (dtype == AVMOE_F32) add_rowvec_kernel<float><<<grid, 256, 0, st>>>((float*)d, vec, N, C);
  else add_rowvec_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16*)d, vec, N, C);
  AVMOE_LAUNCH_CHECK("add_rowvec_kernel");
  return 0;
}

extern "C" int avmoe_batch_sum_add(const float* in, float* out, int BT, int64_t L, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(in && out && BT > 0 && L > 0, "batch_sum_add: bad arguments");
  AVMOE_CHECK_ARG((L + 31) / 32 < (1ll << 31), "batch_sum_add: too many elements");
  batch_sum_add_kernel<<<(unsigned)((L + 31) / 32), dim3(32, 8), 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, out, BT, L);
  AVMOE_LAUNCH_CHECK("batch_sum_add_kernel");
  return 0;
}

extern "C" int avmoe_cast_split(const float* in, void* out_bf16, int64_t groups, int T, int Tp, int C, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(in && out_bf16 && groups > 0 && T > 0 && Tp >= T && C > 0, "cast_split: bad arguments");
  const long long total = (long long)groups * T * C;
  const int grid = (int)min((long long)(8 * num_sms()), (total + 255) / 256);
  cast_split_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, (__nv_bfloat16*)out_bf16, total, T, Tp, C);
  AVMOE_LAUNCH_CHECK("cast_split_kernel");
  return 0;
}
