// Fused expert kernels of the AVMoE adapter (all experts of one MoEAdapter per token tile, fp32 math).
//
//   expert_front_fwd      x tile -> [token<-latent attention | aux add/replace] -> ln_before -> grouped 1x1
//                         down projection -> z (pre-bn1) + bn1 batch sums
//                         (AVMOE/AVE/nets/net_trans_v3.py:386-395, :416-420)
//   bn_moments            sums of a = act(bn1(z)) and a a^T per conv group: bn2's batch statistics follow
//                         algebraically (y = Wu a is linear), so y is never materialised for statistics
//   expert_back_fwd       act -> grouped up projection -> bn2 -> ln_post -> gate -> sum_e probs[b,e] * (.)
//                         (net_trans_v3.py:400-403,430-434,481-486)
//   expert_back_bwd[_stats], expert_front_bwd   hand-derived backward of the above (recompute from z and x)
//
// Token tiles live in shared memory as fp32 [rows][C+1]; activations are read/written in their storage
// dtype with 8/16-byte vectors.  Parameter gradients are accumulated with fp32 atomics, BatchNorm sums
// with fp64 atomics.
#include <type_traits>

#include "block_mma.cuh"
#include "common.cuh"

namespace avmoe {

static constexpr float kLnEps = 1e-5f;
static constexpr double kBnEps = 1e-5;
static constexpr double kBnMomentum = 0.1;
static constexpr int kThreads = 256;

__device__ __forceinline__ int pow2_group(int tn) { return tn < 32 ? tn : 32; }

// sum over groups of `group` consecutive lanes (group is a power of two <= 32)
__device__ __forceinline__ float group_sum(float v, int group) {
  for (int o = group >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__device__ __forceinline__ void load_tile(float* dst, const T* src, int valid_rows, int rows, int C, int ld) {
  const int C4 = C >> 2;
  for (int i = threadIdx.x; i < rows * C4; i += blockDim.x) {
    const int n = i / C4, c4 = i - n * C4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (n < valid_rows) load4<T>(src + (long long)n * C + c4 * 4, v);
    float* d = dst + n * ld + c4 * 4;
    d[0] = v[0]; d[1] = v[1]; d[2] = v[2]; d[3] = v[3];
  }
}

template <typename T>
__device__ __forceinline__ void store_tile(T* dst, const float* src, int valid_rows, int C, int ld) {
  const int C4 = C >> 2;
  for (int i = threadIdx.x; i < valid_rows * C4; i += blockDim.x) {
    const int n = i / C4, c4 = i - n * C4;
    const float* s = src + n * ld + c4 * 4;
    const float v[4] = {s[0], s[1], s[2], s[3]};
    store4<T>(dst + (long long)n * C + c4 * 4, v);
  }
}

// two-pass mean / rstd of every row (one warp per row)
__device__ __forceinline__ void row_stats(const float* buf, int rows, int C, int ld, float* mean, float* rstd) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r = warp; r < rows; r += nwarps) {
    const float* row = buf + r * ld;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += row[c];
    s = warp_sum(s);
    const float m = s / C;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = row[c] - m; ss = fmaf(d, d, ss); }
    ss = warp_sum(ss);
    if (lane == 0) { mean[r] = m; rstd[r] = rsqrtf(ss / C + kLnEps); }
  }
}

__device__ __forceinline__ float block_sum(float v, float* scratch /* >= 32 floats */) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = lane < nwarps ? scratch[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;  // valid in warp 0
}

// ---- token <- latent attention on a tile: probabilities into s2 [rows][T+4], xa = x + g * (A2 tok') into ws ----
template <bool kExact>
__device__ __forceinline__ void latent_tile_fwd(const float* xs, float* ws, float* s2, const float* tp, float g, int rows, int C, int ld, int T) {
  const int lt = T + 4;
  contract<kExact>(rows, T, C, xs, ld, 1, tp, 1, C, [&](int n, int t, float v) { s2[n * lt + t] = v; });
  __syncthreads();
  for (int n = threadIdx.x; n < rows; n += blockDim.x) {
    float* r = s2 + n * lt;
    float mx = r[0];
    for (int t = 1; t < T; ++t) mx = fmaxf(mx, r[t]);
    float sum = 0.f;
    for (int t = 0; t < T; ++t) { const float e = expf(r[t] - mx); r[t] = e; sum += e; }
    const float inv = 1.f / sum;
    for (int t = 0; t < T; ++t) r[t] *= inv;
  }
  __syncthreads();
  contract<kExact>(rows, C, T, s2, lt, 1, tp, C, 1, [&](int n, int c, float v) { ws[n * ld + c] = xs[n * ld + c] + g * v; });
  __syncthreads();
}

// builds ws = the tensor ln_before sees for expert `ex` (see AVMOE_PRE_*); leaves latent probabilities in s2
template <typename T, bool kExact>
__device__ __forceinline__ void pre_transform(const avmoe_expert& ex, const avmoe_expert_args& p, int b, int n0, int valid, int rows,
                                              const float* xs, float* ws, float* s2, int ld) {
  const int C = p.C;
  if (ex.pre_kind == AVMOE_PRE_LATENT) {
    const float* tp = p.tokp + ((long long)b * p.nl + ex.latent_slot) * p.T * C;
    latent_tile_fwd<kExact>(xs, ws, s2, tp, *ex.gate_pre, rows, C, ld, p.T);
  } else if (ex.pre_kind == AVMOE_PRE_NONE) {
    for (int i = threadIdx.x; i < rows * C; i += blockDim.x) { const int n = i / C, c = i - n * C; ws[n * ld + c] = xs[n * ld + c]; }
    __syncthreads();
  } else {
    const T* aux = reinterpret_cast<const T*>(ex.aux) + ((long long)b * p.N + n0) * C;
    load_tile<T>(ws, aux, valid, rows, C, ld);
    __syncthreads();
    if (ex.pre_kind == AVMOE_PRE_AUX_ADD) {
      const float g = *ex.gate_pre;
      for (int i = threadIdx.x; i < rows * C; i += blockDim.x) { const int n = i / C, c = i - n * C; ws[n * ld + c] = xs[n * ld + c] + g * ws[n * ld + c]; }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------ front forward
// grid (ceil(N/TN), BT), block 256
template <typename T>
__global__ void __launch_bounds__(kThreads) expert_front_fwd_kernel(const avmoe_expert_args p, int TN) {
  extern __shared__ float sm[];
  constexpr bool kExact = std::is_same<T, float>::value;
  const int C = p.C, Cr = p.Cr, G = p.G, Cg = C / G, Crg = Cr / G, ld = C + 4, ECr = p.E * Cr, lz = ECr + 1;
  float* xs = sm;                      // [TN][ld]
  float* ws = xs + TN * ld;            // [TN][ld]
  float* s2 = ws + TN * ld;            // [TN][T+4]
  float* zt = s2 + TN * (p.T + 4);     // [TN][E*Cr+1]  z tile of all experts
  float* mean = zt + TN * lz;          // [TN]
  float* rstd = mean + TN;             // [TN]
  const int b = blockIdx.y, n0 = blockIdx.x * TN;
  const int valid = min(TN, p.N - n0);
  const T* xb = reinterpret_cast<const T*>(p.x) + ((long long)b * p.N + n0) * C;
  load_tile<T>(xs, xb, valid, TN, C, ld);
  __syncthreads();
  for (int e = 0; e < p.E; ++e) {
    const avmoe_expert& ex = p.ex[e];
    pre_transform<T, kExact>(ex, p, b, n0, valid, TN, xs, ws, s2, ld);
    if (ex.ln1_w) {
      row_stats(ws, TN, C, ld, mean, rstd);
      __syncthreads();
      for (int i = threadIdx.x; i < TN * C; i += blockDim.x) {
        const int n = i / C, c = i - n * C;
        ws[n * ld + c] = (ws[n * ld + c] - mean[n]) * rstd[n] * __ldg(ex.ln1_w + c) + __ldg(ex.ln1_b + c);
      }
      __syncthreads();
    }
    // grouped down projection: z[n][g*Crg+jl] = sum_cl xn[n][g*Cg+cl] * wd[g*Crg+jl][cl]
    for (int g = 0; g < G; ++g)
      contract<kExact>(TN, Crg, Cg, ws + g * Cg, ld, 1, ex.wd + (long long)g * Crg * Cg, 1, Cg,
                       [&](int n, int jl, float v) { zt[n * lz + e * Cr + g * Crg + jl] = v; });
    __syncthreads();
  }
  // coalesced store of the z tile [valid][E*Cr] and the bn1 batch sums
  float* zg = p.z + ((long long)b * p.N + n0) * ECr;
  for (int i = threadIdx.x; i < valid * ECr; i += blockDim.x) { const int n = i / ECr, q = i - n * ECr; zg[i] = zt[n * lz + q]; }
  if (p.stats1) {
    for (int q = threadIdx.x; q < ECr; q += blockDim.x) {
      float s1 = 0.f, s2v = 0.f;
      for (int n = 0; n < valid; ++n) { const float v = zt[n * lz + q]; s1 += v; s2v = fmaf(v, v, s2v); }
      atomicAdd(&p.stats1[q * 2], (double)s1);
      atomicAdd(&p.stats1[q * 2 + 1], (double)s2v);
    }
  }
}

// ------------------------------------------------------------------------------------------------ bn moments
// persistent: grid = min(tiles, 2*SMs); every thread owns a few (expert, output) accumulators in registers
static constexpr int kMomPerThread = 40;
__global__ void __launch_bounds__(kThreads) bn_moments_kernel(const avmoe_expert_args p, double* m1, double* m2, int TN, long long rows_total) {
  extern __shared__ float sm[];
  const int Cr = p.Cr, G = p.G, Crg = Cr / G, E = p.E;
  const int ECr = E * Cr;
  float* as = sm;  // [TN][ECr+1]
  const int lda = ECr + 1;
  const int n_m1 = ECr, n_m2 = E * G * Crg * Crg;
  const int n_out = n_m1 + n_m2;
  float accv[kMomPerThread];
#pragma unroll
  for (int k = 0; k < kMomPerThread; ++k) accv[k] = 0.f;
  const long long tiles = (rows_total + TN - 1) / TN;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long r0 = tile * TN;
    const int valid = (int)min((long long)TN, rows_total - r0);
    for (int i = threadIdx.x; i < TN * ECr; i += blockDim.x) {
      const int n = i / ECr, q = i - n * ECr;
      float a = 0.f;
      if (n < valid) {
        const int e = q / Cr, j = q - e * Cr;
        float zt = p.z[(r0 + n) * ECr + q];
        zt = zt * p.aff1[(e * Cr + j) * 2] + p.aff1[(e * Cr + j) * 2 + 1];
        a = p.ex[e].relu ? fmaxf(zt, 0.f) : zt;
      }
      as[n * lda + q] = a;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kMomPerThread; ++k) {
      const int o = threadIdx.x + k * kThreads;
      if (o < n_out) {
        float s = 0.f;
        if (o < n_m1) {
          for (int n = 0; n < valid; ++n) s += as[n * lda + o];
        } else {
          const int q = o - n_m1;
          const int e = q / (G * Crg * Crg);
          const int r = q - e * G * Crg * Crg;
          const int g = r / (Crg * Crg);
          const int jk = r - g * Crg * Crg;
          const int j = jk / Crg, kk = jk - j * Crg;
          const int cj = e * Cr + g * Crg + j, ck = e * Cr + g * Crg + kk;
          for (int n = 0; n < valid; ++n) s = fmaf(as[n * lda + cj], as[n * lda + ck], s);
        }
        accv[k] += s;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < kMomPerThread; ++k) {
    const int o = threadIdx.x + k * kThreads;
    if (o < n_out) {
      if (o < n_m1) atomicAdd(&m1[o], (double)accv[k]);
      else atomicAdd(&m2[o - n_m1], (double)accv[k]);
    }
  }
}

// generic (slow) variant when the moment matrix has more entries than kMomPerThread*256: one output per thread, grid-stride
__global__ void bn_moments_big_kernel(const avmoe_expert_args p, double* m1, double* m2, long long rows_total, long long rows_per_cta) {
  const int Cr = p.Cr, G = p.G, Crg = Cr / G, E = p.E, ECr = E * Cr;
  const int n_m1 = ECr, n_out = n_m1 + E * G * Crg * Crg;
  const long long r0 = blockIdx.y * rows_per_cta, r1 = min(rows_total, r0 + rows_per_cta);
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_out; o += gridDim.x * blockDim.x) {
    int cj, ck, e;
    if (o < n_m1) { cj = ck = o; e = o / Cr; }
    else {
      const int q = o - n_m1;
      e = q / (G * Crg * Crg);
      const int r = q - e * G * Crg * Crg;
      const int g = r / (Crg * Crg);
      const int jk = r - g * Crg * Crg;
      cj = e * Cr + g * Crg + jk / Crg; ck = e * Cr + g * Crg + jk % Crg;
    }
    const float sj = p.aff1[cj * 2], hj = p.aff1[cj * 2 + 1], sk = p.aff1[ck * 2], hk = p.aff1[ck * 2 + 1];
    const bool relu = p.ex[e].relu;
    double s = 0.0;
    for (long long r = r0; r < r1; ++r) {
      float aj = p.z[r * ECr + cj] * sj + hj, ak = p.z[r * ECr + ck] * sk + hk;
      if (relu) { aj = fmaxf(aj, 0.f); ak = fmaxf(ak, 0.f); }
      s += (o < n_m1) ? (double)aj : (double)aj * ak;
    }
    if (o < n_m1) atomicAdd(&m1[o], s); else atomicAdd(&m2[o - n_m1], s);
  }
}

// ------------------------------------------------------------------------------------------------ bn finalize
__global__ void bn1_finalize_kernel(const avmoe_bn_args a) {
  const int e = blockIdx.x;
  for (int j = threadIdx.x; j < a.Cr; j += blockDim.x) {
    double mean, var;
    if (a.train) {
      mean = a.stats1[(e * a.Cr + j) * 2] / (double)a.n;
      var = a.stats1[(e * a.Cr + j) * 2 + 1] / (double)a.n - mean * mean;
      if (var < 0) var = 0;
      const double unbiased = var * ((double)a.n / (double)(a.n > 1 ? a.n - 1 : 1));
      a.rm1[e][j] = (float)((1.0 - kBnMomentum) * a.rm1[e][j] + kBnMomentum * mean);
      a.rv1[e][j] = (float)((1.0 - kBnMomentum) * a.rv1[e][j] + kBnMomentum * unbiased);
    } else {
      mean = a.rm1[e][j]; var = a.rv1[e][j];
    }
    const double invstd = 1.0 / sqrt(var + kBnEps);
    const double sc = (double)a.w1[e][j] * invstd;
    a.aff1[(e * a.Cr + j) * 2] = (float)sc;
    a.aff1[(e * a.Cr + j) * 2 + 1] = (float)((double)a.b1[e][j] - mean * sc);
    a.save1[(e * a.Cr + j) * 2] = (float)mean;
    a.save1[(e * a.Cr + j) * 2 + 1] = (float)invstd;
  }
}

__global__ void bn2_finalize_kernel(const avmoe_bn_args a) {
  const int e = blockIdx.x;
  const int Cg = a.C / a.G, Crg = a.Cr / a.G;
  for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < a.C; c += gridDim.y * blockDim.x) {
    double mean, var;
    if (a.train) {
      const int g = c / Cg;
      const float* wr = a.wu[e] + (long long)c * Crg;
      const double* m1 = a.m1 + e * a.Cr + g * Crg;
      const double* m2 = a.m2 + ((long long)(e * a.G + g) * Crg) * Crg;
      double mu = 0.0, ey2 = 0.0;
      for (int j = 0; j < Crg; ++j) {
        mu += (double)wr[j] * m1[j];
        double t = 0.0;
        for (int k = 0; k < Crg; ++k) t += (double)wr[k] * m2[(long long)j * Crg + k];
        ey2 += (double)wr[j] * t;
      }
      mean = mu / (double)a.n;
      var = ey2 / (double)a.n - mean * mean;
      if (var < 0) var = 0;
      const double unbiased = var * ((double)a.n / (double)(a.n > 1 ? a.n - 1 : 1));
      a.rm2[e][c] = (float)((1.0 - kBnMomentum) * a.rm2[e][c] + kBnMomentum * mean);
      a.rv2[e][c] = (float)((1.0 - kBnMomentum) * a.rv2[e][c] + kBnMomentum * unbiased);
    } else {
      mean = a.rm2[e][c]; var = a.rv2[e][c];
    }
    const double invstd = 1.0 / sqrt(var + kBnEps);
    const double sc = (double)a.w2[e][c] * invstd;
    a.aff2[(e * a.C + c) * 2] = (float)sc;
    a.aff2[(e * a.C + c) * 2 + 1] = (float)((double)a.b2[e][c] - mean * sc);
    a.save2[(e * a.C + c) * 2] = (float)mean;
    a.save2[(e * a.C + c) * 2 + 1] = (float)invstd;
  }
}

// ------------------------------------------------------------------------------------------------ back: shared recompute
// as[n][j] = act(bn1(z)) for expert e (invalid rows 0); ys[n][c] = y = Wu a (pre-bn2)
template <bool kExact>
__device__ __forceinline__ void back_recompute(const avmoe_expert_args& p, const avmoe_expert& ex, int e, int b, int n0, int valid, int TN,
                                               float* as, float* ys, int ld) {
  const int C = p.C, Cr = p.Cr, G = p.G, Cg = C / G, Crg = Cr / G, lda = Cr + 4;
  const bool bn = p.bn_mode != AVMOE_BN_OFF;
  for (int i = threadIdx.x; i < TN * Cr; i += blockDim.x) {
    const int n = i / Cr, j = i - n * Cr;
    float a = 0.f;
    if (n < valid) {
      float zt = p.z[(((long long)b * p.N + n0 + n) * p.E + e) * Cr + j];
      if (bn) zt = zt * p.aff1[(e * Cr + j) * 2] + p.aff1[(e * Cr + j) * 2 + 1];
      a = ex.relu ? fmaxf(zt, 0.f) : zt;
    }
    as[n * lda + j] = a;
  }
  __syncthreads();
  for (int g = 0; g < G; ++g)
    contract<kExact>(TN, Cg, Crg, as + g * Crg, lda, 1, ex.wu + (long long)g * Cg * Crg, 1, Crg,
                     [&](int n, int cl, float v) { ys[n * ld + g * Cg + cl] = v; });
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ back forward
template <typename T>
__global__ void __launch_bounds__(kThreads) expert_back_fwd_kernel(const avmoe_expert_args p, int TN) {
  extern __shared__ float sm[];
  constexpr bool kExact = std::is_same<T, float>::value;
  const int C = p.C, Cr = p.Cr, ld = C + 4;
  float* ys = sm;                      // [TN][ld]
  float* acc = ys + TN * ld;           // [TN][ld]
  float* as = acc + TN * ld;           // [TN][Cr+4]
  float* mean = as + TN * (Cr + 4);
  float* rstd = mean + TN;
  const int b = blockIdx.y, n0 = blockIdx.x * TN;
  const int valid = min(TN, p.N - n0);
  const bool bn = p.bn_mode != AVMOE_BN_OFF;
  for (int i = threadIdx.x; i < TN * ld; i += blockDim.x) acc[i] = 0.f;
  for (int e = 0; e < p.E; ++e) {
    const avmoe_expert& ex = p.ex[e];
    back_recompute<kExact>(p, ex, e, b, n0, valid, TN, as, ys, ld);
    if (bn) {
      for (int i = threadIdx.x; i < TN * C; i += blockDim.x) {
        const int n = i / C, c = i - n * C;
        ys[n * ld + c] = ys[n * ld + c] * p.aff2[(e * C + c) * 2] + p.aff2[(e * C + c) * 2 + 1];
      }
      __syncthreads();
    }
    if (ex.ln2_w) { row_stats(ys, TN, C, ld, mean, rstd); __syncthreads(); }
    const float scale = p.probs[(long long)b * p.E + e] * (ex.gate ? *ex.gate : 1.f);
    for (int i = threadIdx.x; i < TN * C; i += blockDim.x) {
      const int n = i / C, c = i - n * C;
      float q = ys[n * ld + c];
      if (ex.ln2_w) q = (q - mean[n]) * rstd[n] * __ldg(ex.ln2_w + c) + __ldg(ex.ln2_b + c);
      acc[n * ld + c] = fmaf(scale, q, acc[n * ld + c]);
    }
    __syncthreads();
  }
  T* ob = reinterpret_cast<T*>(p.out) + ((long long)b * p.N + n0) * C;
  store_tile<T>(ob, acc, valid, C, ld);
}

// ------------------------------------------------------------------------------------------------ back backward
// kStatsOnly: only the bn2 batch reductions (sum dyt, sum dyt*yhat); otherwise the full backward of the back half.
template <typename T, bool kStatsOnly>
__global__ void __launch_bounds__(kThreads) expert_back_bwd_kernel(const avmoe_expert_args p, int TN) {
  extern __shared__ float sm[];
  constexpr bool kExact = std::is_same<T, float>::value;
  const int C = p.C, Cr = p.Cr, G = p.G, Cg = C / G, Crg = Cr / G, ld = C + 4, lda = Cr + 4;
  float* gs0 = sm;                     // [TN][ld]  dout tile (kept)
  float* ys = gs0 + TN * ld;           // [TN][ld]  y (pre-bn2)
  float* gs = ys + TN * ld;            // [TN][ld]  working gradient (dq -> dyt -> dy)
  float* as = gs + TN * ld;            // [TN][Cr+1]
  float* mean = as + TN * lda;
  float* rstd = mean + TN;
  float* m1r = rstd + TN;              // row means of the LN backward
  float* m2r = m1r + TN;
  float* scratch = m2r + TN;           // [32]
  float* dzs = scratch + 32;           // [TN][Cr+4] dzt tile of the current expert
  const int b = blockIdx.y, n0 = blockIdx.x * TN;
  const int valid = min(TN, p.N - n0);
  const int mode = p.bn_mode;
  const double inv_n = 1.0 / ((double)p.BT * p.N);
  const T* gb = reinterpret_cast<const T*>(p.dout) + ((long long)b * p.N + n0) * C;
  load_tile<T>(gs0, gb, valid, TN, C, ld);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int e = 0; e < p.E; ++e) {
    const avmoe_expert& ex = p.ex[e];
    back_recompute<kExact>(p, ex, e, b, n0, valid, TN, as, ys, ld);
    const float* aff2 = p.aff2 ? p.aff2 + (long long)e * C * 2 : nullptr;
    const float* sv2 = p.save2 ? p.save2 + (long long)e * C * 2 : nullptr;
    // ---- ln_post statistics of yt = bn2(y)
    if (ex.ln2_w) {
      if (mode != AVMOE_BN_OFF) {   // row_stats needs yt: materialise into gs temporarily
        for (int i = threadIdx.x; i < TN * C; i += blockDim.x) {
          const int n = i / C, c = i - n * C;
          gs[n * ld + c] = ys[n * ld + c] * aff2[c * 2] + aff2[c * 2 + 1];
        }
        __syncthreads();
        row_stats(gs, TN, C, ld, mean, rstd);
      } else {
        row_stats(ys, TN, C, ld, mean, rstd);
      }
      __syncthreads();
    }
    const float pe = p.probs[(long long)b * p.E + e];
    const float gate = ex.gate ? *ex.gate : 1.f;
    const float sc = pe * gate;
    // ---- per element: q, dq; sum(dout*q) for dprobs/dgate; gs = dq*w (dqh) or dq
    float gq = 0.f;
    for (int i = threadIdx.x; i < TN * C; i += blockDim.x) {
      const int n = i / C, c = i - n * C;
      float yt = ys[n * ld + c];
      if (mode != AVMOE_BN_OFF) yt = yt * aff2[c * 2] + aff2[c * 2 + 1];
      const float g0 = gs0[n * ld + c];
      float q = yt, dqh = g0 * sc;
      if (ex.ln2_w) {
        const float qh = (yt - mean[n]) * rstd[n];
        q = qh * __ldg(ex.ln2_w + c) + __ldg(ex.ln2_b + c);
        dqh *= __ldg(ex.ln2_w + c);
      }
      gq = fmaf(g0, q, gq);
      gs[n * ld + c] = dqh;
    }
    if (!kStatsOnly) {
      const float tot = block_sum(gq, scratch);
      if (threadIdx.x == 0) {
        atomicAdd(&p.dprobs[(long long)b * p.E + e], tot * gate);
        if (ex.g_gate) atomicAdd(ex.g_gate, tot * pe);
      }
    }
    __syncthreads();
    if (ex.ln2_w) {
      // column sums for d ln_post weight/bias (dq = dqh / w is avoided: recompute dq = g0*sc)
      if (!kStatsOnly) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
          float sw = 0.f, sb = 0.f;
          const float a2s = mode != AVMOE_BN_OFF ? aff2[c * 2] : 1.f, a2h = mode != AVMOE_BN_OFF ? aff2[c * 2 + 1] : 0.f;
          for (int n = 0; n < valid; ++n) {
            const float yt = ys[n * ld + c] * a2s + a2h;
            const float qh = (yt - mean[n]) * rstd[n];
            const float dq = gs0[n * ld + c] * sc;
            sw = fmaf(dq, qh, sw);
            sb += dq;
          }
          atomicAdd(ex.g_ln2_w + c, sw);
          atomicAdd(ex.g_ln2_b + c, sb);
        }
      }
      // row means of dqh and dqh*qh
      for (int r = warp; r < TN; r += nwarps) {
        float s1 = 0.f, s2v = 0.f;
        for (int c = lane; c < C; c += 32) {
          float yt = ys[r * ld + c];
          if (mode != AVMOE_BN_OFF) yt = yt * aff2[c * 2] + aff2[c * 2 + 1];
          const float qh = (yt - mean[r]) * rstd[r];
          const float d = gs[r * ld + c];
          s1 += d;
          s2v = fmaf(d, qh, s2v);
        }
        s1 = warp_sum(s1); s2v = warp_sum(s2v);
        if (lane == 0) { m1r[r] = s1 / C; m2r[r] = s2v / C; }
      }
      __syncthreads();
      for (int i = threadIdx.x; i < TN * C; i += blockDim.x) {
        const int n = i / C, c = i - n * C;
        float yt = ys[n * ld + c];
        if (mode != AVMOE_BN_OFF) yt = yt * aff2[c * 2] + aff2[c * 2 + 1];
        const float qh = (yt - mean[n]) * rstd[n];
        gs[n * ld + c] = rstd[n] * (gs[n * ld + c] - m1r[n] - qh * m2r[n]);     // dyt
      }
      __syncthreads();
    }
    // ---- bn2 backward
    if (mode != AVMOE_BN_OFF) {
      if (kStatsOnly || mode == AVMOE_BN_EVAL) {
        double* red = kStatsOnly ? p.red2 : p.red2b;
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
          float s1 = 0.f, s2v = 0.f;
          const float mu = sv2[c * 2], is = sv2[c * 2 + 1];
          for (int n = 0; n < valid; ++n) {
            const float d = gs[n * ld + c];
            s1 += d;
            s2v = fmaf(d, (ys[n * ld + c] - mu) * is, s2v);
          }
          atomicAdd(&red[((long long)e * C + c) * 2], (double)s1);
          atomicAdd(&red[((long long)e * C + c) * 2 + 1], (double)s2v);
        }
        __syncthreads();   // the sums read gs (dyt) before the loop below overwrites it with dy
      }
      if (!kStatsOnly) {
        for (int i = threadIdx.x; i < TN * C; i += blockDim.x) {
          const int n = i / C, c = i - n * C;
          float d = gs[n * ld + c];
          if (mode == AVMOE_BN_TRAIN) {
            const float md = (float)(p.red2[((long long)e * C + c) * 2] * inv_n);
            const float mdy = (float)(p.red2[((long long)e * C + c) * 2 + 1] * inv_n);
            const float yh = (ys[n * ld + c] - sv2[c * 2]) * sv2[c * 2 + 1];
            d = d - md - yh * mdy;
          }
          gs[n * ld + c] = (n < valid) ? d * aff2[c * 2] : 0.f;                   // dy
        }
      }
      __syncthreads();
    }
    if (kStatsOnly) continue;
    // ---- d up_sampler.weight[c][jl] += sum_n dy[n][c] * a[n][g*Crg+jl]      (rows >= valid of gs are zero)
    for (int g = 0; g < G; ++g)
      contract<kExact>(Cg, Crg, TN, gs + g * Cg, 1, ld, as + g * Crg, lda, 1,
                       [&](int cl, int jl, float v) { atomicAdd(ex.g_wu + (long long)(g * Cg + cl) * Crg + jl, v); });
    // ---- da[n][g*Crg+jl] = sum_cl dy[n][g*Cg+cl] * wu[g*Cg+cl][jl]
    for (int g = 0; g < G; ++g)
      contract<kExact>(TN, Crg, Cg, gs + g * Cg, ld, 1, ex.wu + (long long)g * Cg * Crg, Crg, 1,
                       [&](int n, int jl, float v) { dzs[n * lda + g * Crg + jl] = v; });
    __syncthreads();
    // ---- relu mask, store dzt, bn1-backward sums per channel
    for (int i = threadIdx.x; i < valid * Cr; i += blockDim.x) {
      const int n = i / Cr, j = i - n * Cr;
      float d = dzs[n * lda + j];
      if (ex.relu && !(as[n * lda + j] > 0.f)) d = 0.f;
      dzs[n * lda + j] = d;
      p.dzt[(((long long)b * p.N + n0 + n) * p.E + e) * Cr + j] = d;
    }
    __syncthreads();
    if (mode != AVMOE_BN_OFF) {
      for (int j = threadIdx.x; j < Cr; j += blockDim.x) {
        const float mu = p.save1[(e * Cr + j) * 2], is = p.save1[(e * Cr + j) * 2 + 1];
        float s1 = 0.f, s2v = 0.f;
        for (int n = 0; n < valid; ++n) {
          const float d = dzs[n * lda + j];
          s1 += d;
          s2v = fmaf(d, (p.z[(((long long)b * p.N + n0 + n) * p.E + e) * Cr + j] - mu) * is, s2v);
        }
        atomicAdd(&p.red1[((long long)e * Cr + j) * 2], (double)s1);
        atomicAdd(&p.red1[((long long)e * Cr + j) * 2 + 1], (double)s2v);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ front backward
template <typename T>
__global__ void __launch_bounds__(kThreads) expert_front_bwd_kernel(const avmoe_expert_args p, int TN) {
  extern __shared__ float sm[];
  constexpr bool kExact = std::is_same<T, float>::value;
  const int C = p.C, Cr = p.Cr, G = p.G, Cg = C / G, Crg = Cr / G, ld = C + 4, lda = Cr + 4, lt = p.T + 4;
  float* xs = sm;                      // [TN][ld] x tile
  float* ws = xs + TN * ld;            // [TN][ld] xa -> xhat
  float* t1 = ws + TN * ld;            // [TN][ld] dxn -> dxa
  float* dxs = t1 + TN * ld;           // [TN][ld] accumulated dx
  float* s2 = dxs + TN * ld;           // [TN][T+1] token<-latent probabilities
  float* ds2 = s2 + TN * lt;           // [TN][T+1]
  float* dzs = ds2 + TN * lt;          // [TN][Cr+1]
  float* mean = dzs + TN * lda;
  float* rstd = mean + TN;
  float* m1r = rstd + TN;
  float* m2r = m1r + TN;
  float* scratch = m2r + TN;           // [32]
  const int b = blockIdx.y, n0 = blockIdx.x * TN;
  const int valid = min(TN, p.N - n0);
  const int mode = p.bn_mode;
  const double inv_n = 1.0 / ((double)p.BT * p.N);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const T* xb = reinterpret_cast<const T*>(p.x) + ((long long)b * p.N + n0) * C;
  load_tile<T>(xs, xb, valid, TN, C, ld);
  for (int i = threadIdx.x; i < TN * C; i += blockDim.x) {
    const int n = i / C, c = i - n * C;
    dxs[n * ld + c] = p.dpool_x ? p.dpool_x[(long long)b * C + c] : 0.f;
  }
  __syncthreads();
  for (int e = 0; e < p.E; ++e) {
    const avmoe_expert& ex = p.ex[e];
    // ---- dz = bn1 backward of dzt
    for (int i = threadIdx.x; i < TN * Cr; i += blockDim.x) {
      const int n = i / Cr, j = i - n * Cr;
      float d = 0.f;
      if (n < valid) {
        const long long zi = (((long long)b * p.N + n0 + n) * p.E + e) * Cr + j;
        d = p.dzt[zi];
        if (mode != AVMOE_BN_OFF) {
          if (mode == AVMOE_BN_TRAIN) {
            const float zh = (p.z[zi] - p.save1[(e * Cr + j) * 2]) * p.save1[(e * Cr + j) * 2 + 1];
            d = d - (float)(p.red1[((long long)e * Cr + j) * 2] * inv_n) - zh * (float)(p.red1[((long long)e * Cr + j) * 2 + 1] * inv_n);
          }
          d *= p.aff1[(e * Cr + j) * 2];
        }
      }
      dzs[n * lda + j] = d;
    }
    // ---- recompute xa and its LN statistics
    pre_transform<T, kExact>(ex, p, b, n0, valid, TN, xs, ws, s2, ld);
    if (ex.ln1_w) {
      row_stats(ws, TN, C, ld, mean, rstd);
      __syncthreads();
      for (int i = threadIdx.x; i < TN * C; i += blockDim.x) {
        const int n = i / C, c = i - n * C;
        ws[n * ld + c] = (ws[n * ld + c] - mean[n]) * rstd[n];      // xhat
      }
      __syncthreads();
    }
    // ---- d down_sampler.weight[g*Crg+jl][cl] += sum_n dz[n][g*Crg+jl] * xn[n][g*Cg+cl]   (xn staged in t1; dz rows >= valid are 0)
    for (int i = threadIdx.x; i < TN * C; i += blockDim.x) {
      const int n = i / C, c = i - n * C;
      t1[n * ld + c] = ex.ln1_w ? fmaf(ws[n * ld + c], __ldg(ex.ln1_w + c), __ldg(ex.ln1_b + c)) : ws[n * ld + c];
    }
    __syncthreads();
    for (int g = 0; g < G; ++g)
      contract<kExact>(Crg, Cg, TN, dzs + g * Crg, 1, lda, t1 + g * Cg, ld, 1,
                       [&](int jl, int cl, float v) { atomicAdd(ex.g_wd + (long long)(g * Crg + jl) * Cg + cl, v); });
    __syncthreads();
    // ---- dxn[n][g*Cg+cl] = sum_jl dz[n][g*Crg+jl] * wd[g*Crg+jl][cl]
    for (int g = 0; g < G; ++g)
      contract<kExact>(TN, Cg, Crg, dzs + g * Crg, lda, 1, ex.wd + (long long)g * Crg * Cg, Cg, 1,
                       [&](int n, int cl, float v) { t1[n * ld + g * Cg + cl] = v; });
    __syncthreads();
    if (ex.ln1_w) {
      for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float sw = 0.f, sb = 0.f;
        for (int n = 0; n < valid; ++n) { const float d = t1[n * ld + c]; sw = fmaf(d, ws[n * ld + c], sw); sb += d; }
        atomicAdd(ex.g_ln1_w + c, sw);
        atomicAdd(ex.g_ln1_b + c, sb);
      }
      __syncthreads();
      for (int r = warp; r < TN; r += nwarps) {
        float s1 = 0.f, s2v = 0.f;
        for (int c = lane; c < C; c += 32) {
          const float dh = t1[r * ld + c] * __ldg(ex.ln1_w + c);
          s1 += dh;
          s2v = fmaf(dh, ws[r * ld + c], s2v);
        }
        s1 = warp_sum(s1); s2v = warp_sum(s2v);
        if (lane == 0) { m1r[r] = s1 / C; m2r[r] = s2v / C; }
      }
      __syncthreads();
      for (int i = threadIdx.x; i < TN * C; i += blockDim.x) {
        const int n = i / C, c = i - n * C;
        const float dh = t1[n * ld + c] * __ldg(ex.ln1_w + c);
        t1[n * ld + c] = (n < valid) ? rstd[n] * (dh - m1r[n] - ws[n * ld + c] * m2r[n]) : 0.f;    // dxa
      }
      __syncthreads();
    }
    // ---- route dxa through the pre-transform
    if (ex.pre_kind == AVMOE_PRE_AUX_REPLACE) {
      T* ga = reinterpret_cast<T*>(ex.g_aux) + ((long long)b * p.N + n0) * C;
      store_tile<T>(ga, t1, valid, C, ld);
      __syncthreads();
      continue;
    }
    for (int i = threadIdx.x; i < TN * C; i += blockDim.x) { const int n = i / C, c = i - n * C; dxs[n * ld + c] += t1[n * ld + c]; }
    if (ex.pre_kind == AVMOE_PRE_AUX_ADD) {
      const float g = *ex.gate_pre;
      const T* aux = reinterpret_cast<const T*>(ex.aux) + ((long long)b * p.N + n0) * C;
      T* ga = reinterpret_cast<T*>(ex.g_aux) + ((long long)b * p.N + n0) * C;
      float dg = 0.f;
      for (int i = threadIdx.x; i < valid * C; i += blockDim.x) {
        const int n = i / C, c = i - n * C;
        const float d = t1[n * ld + c];
        dg = fmaf(d, to_f32(aux[(long long)n * C + c]), dg);
        ga[(long long)n * C + c] = from_f32<T>(g * d);
      }
      const float tot = block_sum(dg, scratch);
      if (threadIdx.x == 0) atomicAdd(ex.g_gate_pre, tot);
    } else if (ex.pre_kind == AVMOE_PRE_LATENT) {
      const float g = *ex.gate_pre;
      const float* tp = p.tokp + ((long long)b * p.nl + ex.latent_slot) * p.T * C;
      float* dtp = p.dtokp + ((long long)b * p.nl + ex.latent_slot) * p.T * C;
      // u[n][t] = dxa_n . tok'_t
      contract<kExact>(TN, p.T, C, t1, ld, 1, tp, 1, C, [&](int n, int t, float v) { ds2[n * lt + t] = v; });
      __syncthreads();
      // d gate_pre = sum A2*u ; dS2 = A2 * (g*u - sum_t A2*g*u)
      float dg = 0.f;
      for (int n = threadIdx.x; n < TN; n += blockDim.x) {
        float dot = 0.f;
        for (int t = 0; t < p.T; ++t) dot = fmaf(s2[n * lt + t], ds2[n * lt + t], dot);
        if (n < valid) dg += dot;
        for (int t = 0; t < p.T; ++t) ds2[n * lt + t] = (n < valid) ? s2[n * lt + t] * g * (ds2[n * lt + t] - dot) : 0.f;
      }
      const float tot = block_sum(dg, scratch);
      if (threadIdx.x == 0) atomicAdd(ex.g_gate_pre, tot);
      __syncthreads();
      // dx[n][c] += sum_t dS2[n][t] * tok'[t][c]
      contract<kExact>(TN, C, p.T, ds2, lt, 1, tp, C, 1, [&](int n, int c, float v) { dxs[n * ld + c] += v; });
      // d tok'[t][c] += sum_n A2[n][t]*g*dxa[n][c] + dS2[n][t]*x[n][c]       (rows >= valid of t1 / xs are zero)
      contract<kExact>(p.T, C, TN, s2, 1, lt, t1, ld, 1, [&](int t, int c, float v) { atomicAdd(dtp + (long long)t * C + c, g * v); });
      contract<kExact>(p.T, C, TN, ds2, 1, lt, xs, ld, 1, [&](int t, int c, float v) { atomicAdd(dtp + (long long)t * C + c, v); });
    }
    __syncthreads();
  }
  T* dxb = reinterpret_cast<T*>(p.dx) + ((long long)b * p.N + n0) * C;
  store_tile<T>(dxb, dxs, valid, C, ld);
}

// ------------------------------------------------------------------------------------------------ host side
static int pow2_floor(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }
static int pick_tn(int C, int budget_floats) {
  int tn = budget_floats / (C + 4);
  if (tn < 1) tn = 1;
  if (tn > 64) tn = 64;
  return pow2_floor(tn);
}

static int check_common(const avmoe_expert_args* a, const char* who) {
  AVMOE_CHECK_ARG(a != nullptr, "%s: null args", who);
  AVMOE_CHECK_ARG(a->E >= 1 && a->E <= AVMOE_MAX_EXPERTS, "%s: E=%d out of range", who, a->E);
  AVMOE_CHECK_ARG(a->BT > 0 && a->N > 0 && a->C > 0 && a->Cr > 0 && a->G > 0, "%s: bad sizes", who);
  AVMOE_CHECK_ARG(a->C % 4 == 0 && a->C % a->G == 0 && a->Cr % a->G == 0, "%s: C must be a multiple of 4 and of G; Cr a multiple of G", who);
  AVMOE_CHECK_ARG(a->dtype == AVMOE_F32 || a->dtype == AVMOE_BF16, "%s: bad dtype", who);
  AVMOE_CHECK_ARG(a->BT <= 65535, "%s: BT too large for the launch grid", who);
  return 0;
}

template <typename K>
static int set_smem(K kernel, size_t smem, const char* who) {
  AVMOE_CHECK_ARG(smem <= 227 * 1024, "%s: tile does not fit shared memory (%zu bytes)", who, smem);
  AVMOE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return 0;
}

}  // namespace avmoe

using namespace avmoe;

extern "C" int avmoe_expert_front_fwd(const avmoe_expert_args* a, avmoe_stream_t stream) {
  if (int rc = check_common(a, "expert_front_fwd")) return rc;
  AVMOE_CHECK_ARG(a->x && a->z, "expert_front_fwd: null tensor");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int TN = pick_tn(a->C, 12288);
  const size_t smem = sizeof(float) * (2 * (size_t)TN * (a->C + 4) + (size_t)TN * (a->T + 4) + (size_t)TN * (a->E * a->Cr + 1) + 2 * TN);
  dim3 grid((a->N + TN - 1) / TN, a->BT);
  if (a->dtype == AVMOE_F32) {
    if (int rc = set_smem(expert_front_fwd_kernel<float>, smem, "expert_front_fwd")) return rc;
    expert_front_fwd_kernel<float><<<grid, kThreads, smem, st>>>(*a, TN);
  } else {
    if (int rc = set_smem(expert_front_fwd_kernel<__nv_bfloat16>, smem, "expert_front_fwd")) return rc;
    expert_front_fwd_kernel<__nv_bfloat16><<<grid, kThreads, smem, st>>>(*a, TN);
  }
  AVMOE_LAUNCH_CHECK("expert_front_fwd_kernel");
  return 0;
}

extern "C" int avmoe_bn_moments(const avmoe_expert_args* a, double* m1, double* m2, avmoe_stream_t stream) {
  if (int rc = check_common(a, "bn_moments")) return rc;
  AVMOE_CHECK_ARG(a->z && a->aff1 && m1 && m2, "bn_moments: null tensor");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int Crg = a->Cr / a->G;
  const int ECr = a->E * a->Cr;
  const long long n_out = (long long)ECr + (long long)a->E * a->G * Crg * Crg;
  const long long rows = (long long)a->BT * a->N;
  AVMOE_CUDA(cudaMemsetAsync(m1, 0, sizeof(double) * ECr, st));
  AVMOE_CUDA(cudaMemsetAsync(m2, 0, sizeof(double) * (size_t)a->E * a->G * Crg * Crg, st));
  if (n_out <= (long long)kMomPerThread * kThreads) {
    int TN = 8192 / (ECr + 1);
    if (TN > 128) TN = 128;
    if (TN < 1) TN = 1;
    const size_t smem = sizeof(float) * (size_t)TN * (ECr + 1);
    long long tiles = (rows + TN - 1) / TN;
    int grid = (int)(tiles < 2LL * num_sms() ? tiles : 2LL * num_sms());
    if (int rc = set_smem(bn_moments_kernel, smem, "bn_moments")) return rc;
    bn_moments_kernel<<<grid, kThreads, smem, st>>>(*a, m1, m2, TN, rows);
  } else {
    const int gx = (int)((n_out + 255) / 256 < 64 ? (n_out + 255) / 256 : 64);
    int gy = (2 * num_sms() + gx - 1) / gx;
    if (gy > rows) gy = (int)rows;
    const long long rpc = (rows + gy - 1) / gy;
    bn_moments_big_kernel<<<dim3(gx, gy), 256, 0, st>>>(*a, m1, m2, rows, rpc);
  }
  AVMOE_LAUNCH_CHECK("bn_moments_kernel");
  return 0;
}

extern "C" int avmoe_bn1_finalize(const avmoe_bn_args* a, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(a && a->E >= 1 && a->E <= AVMOE_MAX_EXPERTS && a->aff1 && a->save1, "bn1_finalize: bad arguments");
  AVMOE_CHECK_ARG(!a->train || a->stats1, "bn1_finalize: training mode needs the batch sums");
  bn1_finalize_kernel<<<a->E, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*a);
  AVMOE_LAUNCH_CHECK("bn1_finalize_kernel");
  return 0;
}

extern "C" int avmoe_bn2_finalize(const avmoe_bn_args* a, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(a && a->E >= 1 && a->E <= AVMOE_MAX_EXPERTS && a->aff2 && a->save2, "bn2_finalize: bad arguments");
  AVMOE_CHECK_ARG(!a->train || (a->m1 && a->m2), "bn2_finalize: training mode needs the moment sums");
  bn2_finalize_kernel<<<dim3(a->E, (a->C + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*a);
  AVMOE_LAUNCH_CHECK("bn2_finalize_kernel");
  return 0;
}

extern "C" int avmoe_expert_back_fwd(const avmoe_expert_args* a, avmoe_stream_t stream) {
  if (int rc = check_common(a, "expert_back_fwd")) return rc;
  AVMOE_CHECK_ARG(a->z && a->probs && a->out, "expert_back_fwd: null tensor");
  AVMOE_CHECK_ARG(a->bn_mode == AVMOE_BN_OFF || (a->aff1 && a->aff2), "expert_back_fwd: batch-norm affine missing");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int TN = pick_tn(a->C, 12288);
  const size_t smem = sizeof(float) * (2 * (size_t)TN * (a->C + 4) + (size_t)TN * (a->Cr + 4) + 2 * TN);
  dim3 grid((a->N + TN - 1) / TN, a->BT);
  if (a->dtype == AVMOE_F32) {
    if (int rc = set_smem(expert_back_fwd_kernel<float>, smem, "expert_back_fwd")) return rc;
    expert_back_fwd_kernel<float><<<grid, kThreads, smem, st>>>(*a, TN);
  } else {
    if (int rc = set_smem(expert_back_fwd_kernel<__nv_bfloat16>, smem, "expert_back_fwd")) return rc;
    expert_back_fwd_kernel<__nv_bfloat16><<<grid, kThreads, smem, st>>>(*a, TN);
  }
  AVMOE_LAUNCH_CHECK("expert_back_fwd_kernel");
  return 0;
}

template <bool kStatsOnly>
static int launch_back_bwd(const avmoe_expert_args* a, cudaStream_t st, const char* who) {
  if (int rc = check_common(a, who)) return rc;
  AVMOE_CHECK_ARG(a->z && a->probs && a->dout, "%s: null tensor", who);
  AVMOE_CHECK_ARG(a->bn_mode == AVMOE_BN_OFF || (a->aff1 && a->aff2 && a->save1 && a->save2), "%s: batch-norm tensors missing", who);
  if (kStatsOnly) AVMOE_CHECK_ARG(a->bn_mode == AVMOE_BN_TRAIN && a->red2, "%s: only meaningful for bn training", who);
  else {
    AVMOE_CHECK_ARG(a->dzt && a->dprobs, "%s: null output", who);
    AVMOE_CHECK_ARG(a->bn_mode != AVMOE_BN_TRAIN || a->red2, "%s: red2 missing", who);
    AVMOE_CHECK_ARG(a->bn_mode != AVMOE_BN_EVAL || a->red2b, "%s: red2b missing", who);
    AVMOE_CHECK_ARG(a->bn_mode == AVMOE_BN_OFF || a->red1, "%s: red1 missing", who);
  }
  const int TN = pick_tn(a->C, 8192);
  const size_t smem = sizeof(float) * (3 * (size_t)TN * (a->C + 4) + 2 * (size_t)TN * (a->Cr + 4) + 4 * TN + 32);
  dim3 grid((a->N + TN - 1) / TN, a->BT);
  if (a->dtype == AVMOE_F32) {
    if (int rc = set_smem(expert_back_bwd_kernel<float, kStatsOnly>, smem, who)) return rc;
    expert_back_bwd_kernel<float, kStatsOnly><<<grid, kThreads, smem, st>>>(*a, TN);
  } else {
    if (int rc = set_smem(expert_back_bwd_kernel<__nv_bfloat16, kStatsOnly>, smem, who)) return rc;
    expert_back_bwd_kernel<__nv_bfloat16, kStatsOnly><<<grid, kThreads, smem, st>>>(*a, TN);
  }
  AVMOE_LAUNCH_CHECK("expert_back_bwd_kernel");
  return 0;
}

extern "C" int avmoe_expert_back_bwd_stats(const avmoe_expert_args* a, avmoe_stream_t stream) {
  return launch_back_bwd<true>(a, reinterpret_cast<cudaStream_t>(stream), "expert_back_bwd_stats");
}
extern "C" int avmoe_expert_back_bwd(const avmoe_expert_args* a, avmoe_stream_t stream) {
  return launch_back_bwd<false>(a, reinterpret_cast<cudaStream_t>(stream), "expert_back_bwd");
}

extern "C" int avmoe_expert_front_bwd(const avmoe_expert_args* a, avmoe_stream_t stream) {
  if (int rc = check_common(a, "expert_front_bwd")) return rc;
  AVMOE_CHECK_ARG(a->x && a->dzt && a->dx, "expert_front_bwd: null tensor");
  AVMOE_CHECK_ARG(a->bn_mode == AVMOE_BN_OFF || (a->aff1 && a->save1 && a->z && a->red1), "expert_front_bwd: batch-norm tensors missing");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int TN = pick_tn(a->C, 6144);
  const size_t smem = sizeof(float) * (4 * (size_t)TN * (a->C + 4) + 2 * (size_t)TN * (a->T + 4) + (size_t)TN * (a->Cr + 4) + 4 * TN + 32);
  dim3 grid((a->N + TN - 1) / TN, a->BT);
  if (a->dtype == AVMOE_F32) {
    if (int rc = set_smem(expert_front_bwd_kernel<float>, smem, "expert_front_bwd")) return rc;
    expert_front_bwd_kernel<float><<<grid, kThreads, smem, st>>>(*a, TN);
  } else {
    if (int rc = set_smem(expert_front_bwd_kernel<__nv_bfloat16>, smem, "expert_front_bwd")) return rc;
    expert_front_bwd_kernel<__nv_bfloat16><<<grid, kThreads, smem, st>>>(*a, TN);
  }
  AVMOE_LAUNCH_CHECK("expert_front_bwd_kernel");
  return 0;
}
