// Bandwidth-bound helper kernels of the AVMoE adapter path: layout changes that feed the tensor-core
// GEMMs in K-major form, the router's mean pooling, and the resampler's bias-gradient reductions.
//   transpose / cast   [b][R][C] -> [b][C][R] (optionally fp32 -> bf16): 32x32 smem tiles, coalesced both ways
//   pool               x [BT,N,C] -> per-chunk partial column sums (deterministic two-stage mean,
//                      replaces modal.mean(dim=1), AVMOE/AVE/nets/net_trans_v3.py:474-475)
//   reduce_bias        row / column sums of a gradient tensor (autograd of the conv/linear biases)
#include "common.cuh"
#include "row_pass.cuh"

namespace avmoe {

// ------------------------------------------------------------------------------------------------ transpose
template <typename TIn, typename TOut>
__global__ void transpose_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, int R, int C, long long in_ld, long long in_bs,
                                 long long out_ld, long long out_bs) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const TIn* ip = in + (long long)b * in_bs;
  TOut* op = out + (long long)b * out_bs;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? to_f32(ip[(long long)r * in_ld + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < R) op[(long long)c * out_ld + r] = from_f32<TOut>(tile[threadIdx.x][i]);
  }
}

// bf16 -> bf16 with 64x64 tiles and 4-byte accesses on both sides (full 128-byte rows per warp instruction); needs even R, C,
// pitches and 4-byte aligned bases
__global__ void __launch_bounds__(256) transpose_bf16_x2_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int R, int C,
                                                                long long in_ld, long long in_bs, long long out_ld, long long out_bs) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int b = blockIdx.z;
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const __nv_bfloat16* ip = in + (long long)b * in_bs;
  __nv_bfloat16* op = out + (long long)b * out_bs;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + 2 * tx;
    __nv_bfloat162 v = __floats2bfloat162_rn(0.f, 0.f);
    if (r < R && c < C) v = *reinterpret_cast<const __nv_bfloat162*>(ip + (long long)r * in_ld + c);
    *reinterpret_cast<__nv_bfloat162*>(&tile[ty + 8 * i][2 * tx]) = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int cl = ty + 8 * i;                 // input column -> output row
    const int c = c0 + cl, r = r0 + 2 * tx;    // output columns r, r+1
    if (c < C && r < R) {
      __nv_bfloat162 v;
      v.x = tile[2 * tx][cl];
      v.y = tile[2 * tx + 1][cl];
      *reinterpret_cast<__nv_bfloat162*>(op + (long long)c * out_ld + r) = v;
    }
  }
}

template <typename TIn, typename TOut>
__global__ void cast_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = from_f32<TOut>(to_f32(in[i]));
}

// ------------------------------------------------------------------------------------------------ pool
// grid (chunks, BT); block 256.  part[b][chunk][c] = sum over the chunk's tokens of x[b][n][c]
template <typename T>
__global__ void pool_kernel(const T* __restrict__ x, float* __restrict__ part, int N, int C, int tokens_per_chunk) {
  extern __shared__ float red[];  // [rows_per_iter][C]
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int n0 = chunk * tokens_per_chunk;
  const int n1 = min(N, n0 + tokens_per_chunk);
  const int C4 = C >> 2;
  const int lanes = min(C4, (int)blockDim.x);
  const int rows_per_iter = blockDim.x / lanes;
  const int l = threadIdx.x % lanes, r0 = threadIdx.x / lanes;
  const T* xb = x + (long long)b * N * C;
  if (r0 < rows_per_iter) {
    for (int cg = l; cg < C4; cg += lanes) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int n = n0 + r0; n < n1; n += rows_per_iter) {
        float v[4];
        load4<T>(xb + (long long)n * C + cg * 4, v);
        acc[0] += v[0]; acc[1] += v[1]; acc[2] += v[2]; acc[3] += v[3];
      }
      float* rr = red + (size_t)r0 * C + cg * 4;
      rr[0] = acc[0]; rr[1] = acc[1]; rr[2] = acc[2]; rr[3] = acc[3];
    }
  }
  __syncthreads();
  float* outp = part + ((long long)b * gridDim.x + chunk) * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < rows_per_iter; ++r) s += red[(size_t)r * C + c];
    outp[c] = s;
  }
}

// ------------------------------------------------------------------------------------------------ reduce_bias
// grid (chunks, BT); block 256.  rows_w[n] += sum_c d*colw[c];  cols[c] += sum_n d;  cols_w[c] += sum_n d*roww[n]
template <typename T>
__global__ void reduce_bias_kernel(const T* __restrict__ d, const float* __restrict__ colw, const float* __restrict__ roww,
                                   float* __restrict__ rows_w, float* __restrict__ cols, float* __restrict__ cols_w, int N, int C,
                                   int tokens_per_chunk) {
  extern __shared__ float sm[];  // [2][C] column partials
  float* scol = sm;
  float* scolw = sm + C;
  const int b = blockIdx.y;
  const int n0 = blockIdx.x * tokens_per_chunk;
  const int n1 = min(N, n0 + tokens_per_chunk);
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  const T* db = d + (long long)b * N * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // each warp owns a strided subset of the chunk's tokens; lanes stride the channels
  for (int c0 = lane; c0 < C; c0 += 32) {
    float colacc = 0.f, colwacc = 0.f;
    for (int n = n0 + warp; n < n1; n += nwarps) {
      const float v = to_f32(db[(long long)n * C + c0]);
      colacc += v;
      if (roww) colwacc += v * roww[n];
    }
    atomicAdd(&scol[c0], colacc);
    if (roww) atomicAdd(&scolw[c0], colwacc);
  }
  if (rows_w) {
    for (int n = n0 + warp; n < n1; n += nwarps) {
      float racc = 0.f;
      for (int c = lane; c < C; c += 32) racc += to_f32(db[(long long)n * C + c]) * (colw ? colw[c] : 1.f);
      racc = warp_sum(racc);
      if (lane == 0) atomicAdd(&rows_w[n], racc);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (cols) atomicAdd(&cols[c], scol[c]);
    if (cols_w) atomicAdd(&cols_w[c], scolw[c]);
  }
}

// vectorised single pass (C % 4 == 0): one warp per token row, 4 channels per lane per load; column partials stay in registers
// over the CTA's chunk, the row dot product is reduced by shuffles in the same pass
template <typename T, int kMaxV>
__global__ void reduce_bias_vec_kernel(const T* __restrict__ d, const float* __restrict__ colw, const float* __restrict__ roww,
                                       float* __restrict__ rows_w, float* __restrict__ cols, float* __restrict__ cols_w, int N, int C,
                                       int tokens_per_chunk) {
  extern __shared__ float sm[];  // [2][C] column partials
  float* scol = sm;
  float* scolw = sm + C;
  const int b = blockIdx.y;
  const int n0 = blockIdx.x * tokens_per_chunk;
  const int n1 = min(N, n0 + tokens_per_chunk);
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  const T* db = d + (long long)b * N * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float4 acc[kMaxV], accw[kMaxV];
  ROW_VECS(k, c) acc[k] = accw[k] = f4(0.f);
  // kU rows per warp step, all requested before the first is consumed (narrow rows: one row per warp in flight is latency-bound)
  constexpr int kU = kMaxV <= 2 ? 4 : 1;
  for (int nb = n0 + warp; nb < n1; nb += nwarps * kU) {
    float4 v[kU][kMaxV];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int n = nb + u * nwarps;
      ROW_VECS(k, c) v[u][k] = n < n1 ? ldg_act4<T>(db + (long long)n * C + c) : f4(0.f);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int n = nb + u * nwarps;
      if (n >= n1) break;                                   // warp-uniform
      const float rw = roww ? __ldg(roww + n) : 0.f;
      float racc = 0.f;
      ROW_VECS(k, c) {
        acc[k] = acc[k] + v[u][k];
        accw[k] = accw[k] + rw * v[u][k];
        if (rows_w) racc += hsum(colw ? v[u][k] * ldg4(colw + c) : v[u][k]);
      }
      if (rows_w) {
        racc = warp_sum(racc);
        if (lane == 0) atomicAdd(&rows_w[n], racc);
      }
    }
  }
  ROW_VECS(k, c) {
    atomicAdd(&scol[c], acc[k].x); atomicAdd(&scol[c + 1], acc[k].y); atomicAdd(&scol[c + 2], acc[k].z); atomicAdd(&scol[c + 3], acc[k].w);
    if (roww) {
      atomicAdd(&scolw[c], accw[k].x); atomicAdd(&scolw[c + 1], accw[k].y); atomicAdd(&scolw[c + 2], accw[k].z); atomicAdd(&scolw[c + 3], accw[k].w);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (cols) atomicAdd(&cols[c], scol[c]);
    if (cols_w) atomicAdd(&cols_w[c], scolw[c]);
  }
}

__global__ void rowsum_kernel(const float* __restrict__ w, float* __restrict__ out, int rows, int cols) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += w[(long long)row * cols + c];
  s = warp_sum(s);
  if (lane == 0) out[row] = s;
}

template <typename T>
__global__ void broadcast_rows_kernel(const float* __restrict__ src, T* __restrict__ out, int N, int C) {
  const int b = blockIdx.y;
  const long long total = (long long)N * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    out[(long long)b * total + i] = from_f32<T>(src[(long long)b * C + (int)(i % C)]);
}

static int pick_chunks(int BT, int N, int min_tokens) {
  int chunks = (4 * num_sms() + BT - 1) / BT;
  if (chunks < 1) chunks = 1;
  const int maxc = (N + min_tokens - 1) / min_tokens;
  if (chunks > maxc) chunks = maxc;
  if (chunks > 64) chunks = 64;
  return chunks;
}

}  // namespace avmoe

using namespace avmoe;

extern "C" int avmoe_transpose(const void* in, void* out, int batch, int R, int C, int64_t in_ld, int64_t in_bstride, int64_t out_ld,
                               int64_t out_bstride, int in_dtype, int out_dtype, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(in && out && batch > 0 && R > 0 && C > 0, "transpose: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((C + 31) / 32, (R + 31) / 32, batch), block(32, 8);
  AVMOE_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "transpose: grid too large");
  if (in_dtype == AVMOE_F32 && out_dtype == AVMOE_F32)
    transpose_kernel<float, float><<<grid, block, 0, st>>>((const float*)in, (float*)out, R, C, in_ld, in_bstride, out_ld, out_bstride);
  else if (in_dtype == AVMOE_F32 && out_dtype == AVMOE_BF16)
    transpose_kernel<float, __nv_bfloat16><<<grid, block, 0, st>>>((const float*)in, (__nv_bfloat16*)out, R, C, in_ld, in_bstride, out_ld, out_bstride);
  else if (in_dtype == AVMOE_BF16 && out_dtype == AVMOE_BF16 && R % 2 == 0 && C % 2 == 0 && in_ld % 2 == 0 && in_bstride % 2 == 0 && out_ld % 2 == 0 &&
           out_bstride % 2 == 0 && (uintptr_t)in % 4 == 0 && (uintptr_t)out % 4 == 0 && (R + 63) / 64 <= 65535)
    transpose_bf16_x2_kernel<<<dim3((C + 63) / 64, (R + 63) / 64, batch), block, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, R, C, in_ld, in_bstride,
                                                                                       out_ld, out_bstride);
  else if (in_dtype == AVMOE_BF16 && out_dtype == AVMOE_BF16)
    transpose_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, R, C, in_ld, in_bstride, out_ld, out_bstride);
  else { set_error("transpose: unsupported dtype pair"); return AVMOE_E_UNSUPPORTED; }
  AVMOE_LAUNCH_CHECK("transpose_kernel");
  return 0;
}

extern "C" int avmoe_cast(const void* in, void* out, int64_t n, int in_dtype, int out_dtype, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(in && out && n > 0, "cast: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (int)min((long long)(8 * num_sms()), (long long)((n + 255) / 256));
  if (in_dtype == AVMOE_F32 && out_dtype == AVMOE_BF16) cast_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)in, (__nv_bfloat16*)out, n);
  else if (in_dtype == AVMOE_BF16 && out_dtype == AVMOE_F32) cast_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, (float*)out, n);
  else { set_error("cast: unsupported dtype pair"); return AVMOE_E_UNSUPPORTED; }
  AVMOE_LAUNCH_CHECK("cast_kernel");
  return 0;
}

extern "C" int avmoe_pool_chunks(int BT, int N) { return pick_chunks(BT, N, 32); }

extern "C" int avmoe_pool(const void* x, float* partial, int BT, int N, int C, int chunks, int dtype, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(x && partial && BT > 0 && N > 0 && C > 0 && chunks > 0 && (C % 4) == 0, "pool: bad arguments (C must be a multiple of 4)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int tpc = (N + chunks - 1) / chunks;
  const int C4 = C / 4;
  const int lanes = C4 < 256 ? C4 : 256;
  const int rows = 256 / lanes;
  const size_t smem = (size_t)rows * C * sizeof(float);
  dim3 grid(chunks, BT);
  if (dtype == AVMOE_F32) pool_kernel<float><<<grid, 256, smem, st>>>((const float*)x, partial, N, C, tpc);
  else pool_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>((const __nv_bfloat16*)x, partial, N, C, tpc);
  AVMOE_LAUNCH_CHECK("pool_kernel");
  return 0;
}

extern "C" int avmoe_reduce_bias(const void* d, const float* col_weight, const float* row_weight, float* rows_w, float* cols, float* cols_w,
                                 int BT, int N, int C, int dtype, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(d && BT > 0 && N > 0 && C > 0, "reduce_bias: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (rows_w) AVMOE_CUDA(cudaMemsetAsync(rows_w, 0, sizeof(float) * N, st));
  if (cols) AVMOE_CUDA(cudaMemsetAsync(cols, 0, sizeof(float) * C, st));
  if (cols_w) AVMOE_CUDA(cudaMemsetAsync(cols_w, 0, sizeof(float) * C, st));
  const int chunks = pick_chunks(BT, N, 8);
  const int tpc = (N + chunks - 1) / chunks;
  dim3 grid(chunks, BT);
  const size_t smem = 2 * (size_t)C * sizeof(float);
  if (C % 4 == 0 && C <= 128 * kMaxVLimit) {
    AVMOE_DISPATCH_V(C, {
      if (dtype == AVMOE_F32)
        reduce_bias_vec_kernel<float, kV><<<grid, 256, smem, st>>>((const float*)d, col_weight, row_weight, rows_w, cols, cols_w, N, C, tpc);
      else
        reduce_bias_vec_kernel<__nv_bfloat16, kV><<<grid, 256, smem, st>>>((const __nv_bfloat16*)d, col_weight, row_weight, rows_w, cols, cols_w, N, C, tpc);
    });
  } else if (dtype == AVMOE_F32) reduce_bias_kernel<float><<<grid, 256, smem, st>>>((const float*)d, col_weight, row_weight, rows_w, cols, cols_w, N, C, tpc);
  else reduce_bias_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>((const __nv_bfloat16*)d, col_weight, row_weight, rows_w, cols, cols_w, N, C, tpc);
  AVMOE_LAUNCH_CHECK("reduce_bias_kernel");
  return 0;
}

extern "C" int avmoe_rowsum(const float* w, float* out, int rows, int cols, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(w && out && rows > 0 && cols > 0, "rowsum: bad arguments");
  rowsum_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(w, out, rows, cols);
  AVMOE_LAUNCH_CHECK("rowsum_kernel");
  return 0;
}

extern "C" int avmoe_broadcast_rows(const float* src, void* out, int BT, int N, int C, int dtype, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(src && out && BT > 0 && N > 0 && C > 0, "broadcast_rows: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)min((long long)64, ((long long)N * C + 255) / 256), BT);
  if (dtype == AVMOE_F32) broadcast_rows_kernel<float><<<grid, 256, 0, st>>>(src, (float*)out, N, C);
  else broadcast_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(src, (__nv_bfloat16*)out, N, C);
  AVMOE_LAUNCH_CHECK("broadcast_rows_kernel");
  return 0;
}
