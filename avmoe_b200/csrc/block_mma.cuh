// Block-cooperative small GEMM on the legacy tensor-core path (mma.sync m16n8k8, TF32 operands, fp32 accumulate)
// for the contractions INSIDE the fused, bandwidth-bound adapter kernels (latent attention, bottleneck projections and
// their gradients).  These are tiny per token tile (K or N is T<=87, C/r or a token tile), live entirely in shared
// memory / L1 and never dominate a launch; the dense token-resampler GEMMs run on tcgen05 (gemm.cu).
//
// TF32 keeps bf16 activations exact (8-bit mantissa fits in TF32's 10); fp32 side operands (latent tokens, weights,
// probabilities) are fed as raw fp32 bit patterns, of which the tensor core uses the top 10 mantissa bits — an
// 11-significant-bit truncation, far inside the bf16 tolerance.
#pragma once
#include <stdint.h>

namespace avmoe {

__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}

__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// C(m,n) = sum_k A(m,k) * B(k,n)      A(m,k) = A[m*a_rs + k*a_cs]      B(k,n) = B[k*b_rs + n*b_cs]
// Out-of-range rows / columns / k read as zero.  epi(m, n, value) is invoked once per in-range output element by the
// lane that owns it.  Every thread of the block must call (warp-uniform loops); operands may live in shared or global memory.
template <int WM, int WN, class Epi>
__device__ __forceinline__ void block_mma_tf32(int M, int N, int K, const float* A, int a_rs, int a_cs,
                                               const float* B, int b_rs, int b_cs, Epi&& epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int tiles_m = (M + 16 * WM - 1) / (16 * WM), tiles_n = (N + 8 * WN - 1) / (8 * WN);
  const int kfull = K & ~7;
  for (int tile = warp; tile < tiles_m * tiles_n; tile += nwarps) {
    const int m0 = (tile % tiles_m) * 16 * WM, n0 = (tile / tiles_m) * 8 * WN;
    float acc[WM][WN][4];
#pragma unroll
    for (int i = 0; i < WM; ++i)
#pragma unroll
      for (int j = 0; j < WN; ++j) { acc[i][j][0] = 0.f; acc[i][j][1] = 0.f; acc[i][j][2] = 0.f; acc[i][j][3] = 0.f; }
    // Row / column indices are CLAMPED instead of predicated: out-of-range fragments read valid (duplicate) data and
    // their results are simply never handed to epi.  Only the K tail needs zero fill.
    const float* arow[WM][2];
#pragma unroll
    for (int i = 0; i < WM; ++i) {
      const int r0 = min(m0 + 16 * i + g, M - 1), r1 = min(m0 + 16 * i + g + 8, M - 1);
      arow[i][0] = A + (long long)r0 * a_rs + (long long)t * a_cs;
      arow[i][1] = A + (long long)r1 * a_rs + (long long)t * a_cs;
    }
    const float* bcol[WN];
#pragma unroll
    for (int j = 0; j < WN; ++j) bcol[j] = B + (long long)min(n0 + 8 * j + g, N - 1) * b_cs + (long long)t * b_rs;
    const long long a_k4 = 4ll * a_cs, b_k4 = 4ll * b_rs;
    long long ao = 0, bo = 0;   // 64-bit: global-memory operands can have row pitches of tens of millions of elements
    for (int k0 = 0; k0 < kfull; k0 += 8, ao += 8ll * a_cs, bo += 8ll * b_rs) {
      uint32_t af[WM][4], bf[WN][2];
#pragma unroll
      for (int i = 0; i < WM; ++i) {
        af[i][0] = __float_as_uint(arow[i][0][ao]);
        af[i][1] = __float_as_uint(arow[i][1][ao]);
        af[i][2] = __float_as_uint(arow[i][0][ao + a_k4]);
        af[i][3] = __float_as_uint(arow[i][1][ao + a_k4]);
      }
#pragma unroll
      for (int j = 0; j < WN; ++j) {
        bf[j][0] = __float_as_uint(bcol[j][bo]);
        bf[j][1] = __float_as_uint(bcol[j][bo + b_k4]);
      }
#pragma unroll
      for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j) mma_tf32_16x8x8(acc[i][j], af[i], bf[j]);
    }
    if (kfull < K) {   // K tail: zero-fill k >= K
      const bool va = kfull + t < K, vb = kfull + t + 4 < K;
      uint32_t af[WM][4], bf[WN][2];
#pragma unroll
      for (int i = 0; i < WM; ++i) {
        af[i][0] = __float_as_uint(va ? arow[i][0][ao] : 0.f);
        af[i][1] = __float_as_uint(va ? arow[i][1][ao] : 0.f);
        af[i][2] = __float_as_uint(vb ? arow[i][0][ao + a_k4] : 0.f);
        af[i][3] = __float_as_uint(vb ? arow[i][1][ao + a_k4] : 0.f);
      }
#pragma unroll
      for (int j = 0; j < WN; ++j) {
        bf[j][0] = __float_as_uint(va ? bcol[j][bo] : 0.f);
        bf[j][1] = __float_as_uint(vb ? bcol[j][bo + b_k4] : 0.f);
      }
#pragma unroll
      for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j) mma_tf32_16x8x8(acc[i][j], af[i], bf[j]);
    }
#pragma unroll
    for (int i = 0; i < WM; ++i)
#pragma unroll
      for (int j = 0; j < WN; ++j) {
        const int r0 = m0 + 16 * i + g, r1 = r0 + 8;
        const int c0 = n0 + 8 * j + 2 * t, c1 = c0 + 1;
        if (r0 < M) { if (c0 < N) epi(r0, c0, acc[i][j][0]); if (c1 < N) epi(r0, c1, acc[i][j][1]); }
        if (r1 < M) { if (c0 < N) epi(r1, c0, acc[i][j][2]); if (c1 < N) epi(r1, c1, acc[i][j][3]); }
      }
  }
}

// Strided-destination epilogues: D(m,n) at D[m*d_rs + n*d_cs].  One base address per 16x8 fragment, the four outputs at
// fixed offsets from it — ~6 instructions per fragment instead of per-element index arithmetic.
enum class Out { Store, Add, StorePlusSrc, AtomicAdd };

template <Out kOut>
__device__ __forceinline__ void out_one(float* d, const float* s, float v) {
  if constexpr (kOut == Out::Store) *d = v;
  else if constexpr (kOut == Out::Add) *d += v;
  else if constexpr (kOut == Out::StorePlusSrc) *d = *s + v;
  else atomicAdd(d, v);
}

template <int WM, int WN, Out kOut>
__device__ __forceinline__ void block_mma_tf32_to(int M, int N, int K, const float* A, int a_rs, int a_cs, const float* B, int b_rs, int b_cs,
                                                  float* D, int d_rs, int d_cs, float alpha, const float* S /* same strides as D */) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int tiles_m = (M + 16 * WM - 1) / (16 * WM), tiles_n = (N + 8 * WN - 1) / (8 * WN);
  const int kfull = K & ~7;
  for (int tile = warp; tile < tiles_m * tiles_n; tile += nwarps) {
    const int m0 = (tile % tiles_m) * 16 * WM, n0 = (tile / tiles_m) * 8 * WN;
    float acc[WM][WN][4];
#pragma unroll
    for (int i = 0; i < WM; ++i)
#pragma unroll
      for (int j = 0; j < WN; ++j) { acc[i][j][0] = 0.f; acc[i][j][1] = 0.f; acc[i][j][2] = 0.f; acc[i][j][3] = 0.f; }
    const float* arow[WM][2];
#pragma unroll
    for (int i = 0; i < WM; ++i) {
      const int r0 = min(m0 + 16 * i + g, M - 1), r1 = min(m0 + 16 * i + g + 8, M - 1);
      arow[i][0] = A + (long long)r0 * a_rs + (long long)t * a_cs;
      arow[i][1] = A + (long long)r1 * a_rs + (long long)t * a_cs;
    }
    const float* bcol[WN];
#pragma unroll
    for (int j = 0; j < WN; ++j) bcol[j] = B + (long long)min(n0 + 8 * j + g, N - 1) * b_cs + (long long)t * b_rs;
    const long long a_k4 = 4ll * a_cs, b_k4 = 4ll * b_rs;
    long long ao = 0, bo = 0;
    for (int k0 = 0; k0 < kfull; k0 += 8, ao += 8ll * a_cs, bo += 8ll * b_rs) {
      uint32_t af[WM][4], bf[WN][2];
#pragma unroll
      for (int i = 0; i < WM; ++i) {
        af[i][0] = __float_as_uint(arow[i][0][ao]);
        af[i][1] = __float_as_uint(arow[i][1][ao]);
        af[i][2] = __float_as_uint(arow[i][0][ao + a_k4]);
        af[i][3] = __float_as_uint(arow[i][1][ao + a_k4]);
      }
#pragma unroll
      for (int j = 0; j < WN; ++j) {
        bf[j][0] = __float_as_uint(bcol[j][bo]);
        bf[j][1] = __float_as_uint(bcol[j][bo + b_k4]);
      }
#pragma unroll
      for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j) mma_tf32_16x8x8(acc[i][j], af[i], bf[j]);
    }
    if (kfull < K) {
      const bool va = kfull + t < K, vb = kfull + t + 4 < K;
      uint32_t af[WM][4], bf[WN][2];
#pragma unroll
      for (int i = 0; i < WM; ++i) {
        af[i][0] = __float_as_uint(va ? arow[i][0][ao] : 0.f);
        af[i][1] = __float_as_uint(va ? arow[i][1][ao] : 0.f);
        af[i][2] = __float_as_uint(vb ? arow[i][0][ao + a_k4] : 0.f);
        af[i][3] = __float_as_uint(vb ? arow[i][1][ao + a_k4] : 0.f);
      }
#pragma unroll
      for (int j = 0; j < WN; ++j) {
        bf[j][0] = __float_as_uint(va ? bcol[j][bo] : 0.f);
        bf[j][1] = __float_as_uint(vb ? bcol[j][bo + b_k4] : 0.f);
      }
#pragma unroll
      for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j) mma_tf32_16x8x8(acc[i][j], af[i], bf[j]);
    }
    const bool full = (m0 + 16 * WM <= M) && (n0 + 8 * WN <= N);   // warp-uniform
    const long long base = (long long)(m0 + g) * d_rs + (long long)(n0 + 2 * t) * d_cs;
    const long long r8 = 8ll * d_rs;
#pragma unroll
    for (int i = 0; i < WM; ++i)
#pragma unroll
      for (int j = 0; j < WN; ++j) {
        const long long o = base + 16ll * i * d_rs + 8ll * j * d_cs;
        if (full) {
          out_one<kOut>(D + o, S + o, alpha * acc[i][j][0]);
          out_one<kOut>(D + o + d_cs, S + o + d_cs, alpha * acc[i][j][1]);
          out_one<kOut>(D + o + r8, S + o + r8, alpha * acc[i][j][2]);
          out_one<kOut>(D + o + r8 + d_cs, S + o + r8 + d_cs, alpha * acc[i][j][3]);
        } else {
          const int r0 = m0 + 16 * i + g, r1 = r0 + 8, c0 = n0 + 8 * j + 2 * t, c1 = c0 + 1;
          if (r0 < M && c0 < N) out_one<kOut>(D + o, S + o, alpha * acc[i][j][0]);
          if (r0 < M && c1 < N) out_one<kOut>(D + o + d_cs, S + o + d_cs, alpha * acc[i][j][1]);
          if (r1 < M && c0 < N) out_one<kOut>(D + o + r8, S + o + r8, alpha * acc[i][j][2]);
          if (r1 < M && c1 < N) out_one<kOut>(D + o + r8 + d_cs, S + o + r8 + d_cs, alpha * acc[i][j][3]);
        }
      }
  }
}

}  // namespace avmoe

namespace avmoe {

// Exact fp32 variant with the same contract (used for fp32 activations: the parity mode keeps every contraction in
// full fp32 FMA arithmetic).  Lanes run along m so that A rows hit distinct banks and B is a warp broadcast.
template <class Epi>
__device__ __forceinline__ void block_gemm_fp32(int M, int N, int K, const float* A, int a_rs, int a_cs,
                                                const float* B, int b_rs, int b_cs, Epi&& epi) {
  for (int i = threadIdx.x; i < M * N; i += blockDim.x) {
    const int m = i % M, n = i / M;
    const float* ap = A + (long long)m * a_rs;
    const float* bp = B + (long long)n * b_cs;
    float s = 0.f;
    for (int k = 0; k < K; ++k) s = fmaf(ap[(long long)k * a_cs], bp[(long long)k * b_rs], s);
    epi(m, n, s);
  }
}

// kExact = true: fp32 FMA loops; false: TF32 tensor-core tiles (2x4 m16n8 tiles per warp step)
template <bool kExact, class Epi>
__device__ __forceinline__ void contract(int M, int N, int K, const float* A, int a_rs, int a_cs,
                                         const float* B, int b_rs, int b_cs, Epi&& epi) {
  if constexpr (kExact) block_gemm_fp32(M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, static_cast<Epi&&>(epi));
  else {
    // pick the larger per-warp register tile only when it still gives every warp of the block work
    const int nwarps = blockDim.x >> 5;
    const int mt = (M + 15) >> 4, nt = (N + 7) >> 3;
    if (((mt + 1) >> 1) * ((nt + 3) >> 2) >= nwarps) block_mma_tf32<2, 4>(M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, static_cast<Epi&&>(epi));
    else block_mma_tf32<1, 2>(M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, static_cast<Epi&&>(epi));
  }
}

// D(m,n) (=, +=, = S +, atomic +=) alpha * sum_k A(m,k) B(k,n) with a strided destination (see Out)
template <bool kExact, Out kOut>
__device__ __forceinline__ void contract_to(int M, int N, int K, const float* A, int a_rs, int a_cs, const float* B, int b_rs, int b_cs,
                                            float* D, int d_rs, int d_cs, float alpha = 1.f, const float* S = nullptr) {
  if (S == nullptr) S = D;
  if constexpr (kExact) {
    for (int i = threadIdx.x; i < M * N; i += blockDim.x) {
      const int m = i % M, n = i / M;
      const float* ap = A + (long long)m * a_rs;
      const float* bp = B + (long long)n * b_cs;
      float s = 0.f;
      for (int k = 0; k < K; ++k) s = fmaf(ap[(long long)k * a_cs], bp[(long long)k * b_rs], s);
      const long long o = (long long)m * d_rs + (long long)n * d_cs;
      out_one<kOut>(D + o, S + o, alpha * s);
    }
  } else {
    const int nwarps = blockDim.x >> 5;
    const int mt = (M + 15) >> 4, nt = (N + 7) >> 3;
    if (((mt + 1) >> 1) * ((nt + 3) >> 2) >= nwarps) block_mma_tf32_to<2, 4, kOut>(M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, D, d_rs, d_cs, alpha, S);
    else block_mma_tf32_to<1, 2, kOut>(M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, D, d_rs, d_cs, alpha, S);
  }
}

}  // namespace avmoe
