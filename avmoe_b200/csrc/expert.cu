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
#include <stdlib.h>

#include <type_traits>

#include "block_mma.cuh"
#include "common.cuh"
#include "row_pass.cuh"

namespace avmoe {

// global [rows][C] (storage dtype) -> smem fp32 [rows][ld]; rows >= valid are zero-filled
template <typename T, int kMaxV>
__device__ __forceinline__ void load_tile(float* dst, const T* src, int valid_rows, int rows, int C, int ld) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r = warp; r < rows; r += nwarps) {
    ROW_VECS(k, c) st4(dst + r * ld + c, r < valid_rows ? ldg_act4<T>(src + (long long)r * C + c) : f4(0.f));
  }
}

template <typename T, int kMaxV>
__device__ __forceinline__ void store_tile(T* dst, const float* src, int valid_rows, int C, int ld) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r = warp; r < valid_rows; r += nwarps) {
    ROW_VECS(k, c) st_act4<T>(dst + (long long)r * C + c, ld4(src + r * ld + c));
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

// softmax over the T entries of every row of s2 [rows][lt] (one warp per row, lanes over t)
__device__ __forceinline__ void row_softmax(float* s2, int rows, int T, int lt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r = warp; r < rows; r += nwarps) {
    float* pr = s2 + r * lt;
    float mx = -INFINITY;
    for (int t = lane; t < T; t += 32) mx = fmaxf(mx, pr[t]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int t = lane; t < T; t += 32) { const float e = expf(pr[t] - mx); pr[t] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int t = lane; t < T; t += 32) pr[t] *= inv;
  }
}

// Builds in ws the tensor ln_before sees for expert `ex` (AVMOE_PRE_*), EXCEPT for AVMOE_PRE_NONE where the caller reads xs
// directly.  Latent experts leave the token<-latent probabilities in s2.  Returns the buffer holding xa.
template <typename T, bool kExact, int kMaxV>
__device__ __forceinline__ const float* pre_transform(const avmoe_expert& ex, const avmoe_expert_args& p, int b, int n0, int valid, int rows,
                                                      const float* xs, float* ws, float* s2, int ld) {
  const int C = p.C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  if (ex.pre_kind == AVMOE_PRE_NONE) return xs;
  if (ex.pre_kind == AVMOE_PRE_LATENT) {
    const float* tp = p.tokp + ((long long)b * p.nl + ex.latent_slot) * p.T * C;
    const float g = *ex.gate_pre;
    const int lt = p.T + 4;
    contract_to<kExact, Out::Store>(rows, p.T, C, xs, ld, 1, tp, 1, C, s2, lt, 1);
    __syncthreads();
    row_softmax(s2, rows, p.T, lt);
    __syncthreads();
    contract_to<kExact, Out::StorePlusSrc>(rows, C, p.T, s2, lt, 1, tp, C, 1, ws, ld, 1, g, xs);
    __syncthreads();
    return ws;
  }
  const T* aux = reinterpret_cast<const T*>(ex.aux) + ((long long)b * p.N + n0) * C;
  const float g = ex.pre_kind == AVMOE_PRE_AUX_ADD ? *ex.gate_pre : 0.f;
  for (int r = warp; r < rows; r += nwarps) {
    ROW_VECS(k, c) {
      float4 a = r < valid ? ldg_act4<T>(aux + (long long)r * C + c) : f4(0.f);
      if (ex.pre_kind == AVMOE_PRE_AUX_ADD) a = ld4(xs + r * ld + c) + g * a;
      st4(ws + r * ld + c, a);
    }
  }
  __syncthreads();
  return ws;
}

// ------------------------------------------------------------------------------------------------ front forward
// grid (ceil(N/TN), BT), block 256
template <typename T, int kMaxV>
__global__ void __launch_bounds__(kThreads, 2) expert_front_fwd_kernel(const avmoe_expert_args p, int TN) {
  extern __shared__ float sm[];
  constexpr bool kExact = std::is_same<T, float>::value;
  const int C = p.C, Cr = p.Cr, G = p.G, Cg = C / G, Crg = Cr / G, ld = C + 4, ECr = p.E * Cr, lz = ECr + 1;
  float* xs = sm;                      // [TN][ld]
  float* ws = xs + TN * ld;            // [TN][ld]
  float* s2 = ws + TN * ld;            // [TN][T+4]
  float* zt = s2 + TN * (p.T + 4);     // [TN][E*Cr+1]  z tile of all experts
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int b = blockIdx.y, n0 = blockIdx.x * TN;
  const int valid = min(TN, p.N - n0);
  const T* xb = reinterpret_cast<const T*>(p.x) + ((long long)b * p.N + n0) * C;
  load_tile<T, kMaxV>(xs, xb, valid, TN, C, ld);
  __syncthreads();
  for (int e = 0; e < p.E; ++e) {
    const avmoe_expert& ex = p.ex[e];
    const float* xa = pre_transform<T, kExact, kMaxV>(ex, p, b, n0, valid, TN, xs, ws, s2, ld);
    // ln_before (row pass): ws <- LN(xa) * w + b   (or a plain copy when the variant has no ln_before)
    for (int r = warp; r < TN; r += nwarps) {
      float4 v[kMaxV];
      float sum = 0.f;
      ROW_VECS(k, c) { v[k] = ld4(xa + r * ld + c); sum += hsum(v[k]); }
      if (ex.ln1_w) {
        const float mean = warp_sum(sum) / C;
        float ss = 0.f;
        ROW_VECS(k, c) { const float4 d = v[k] - f4(mean); ss += hsum(d * d); }
        const float rstd = rsqrtf(warp_sum(ss) / C + kLnEps);
        ROW_VECS(k, c) st4(ws + r * ld + c, (rstd * (v[k] - f4(mean))) * ldg4(ex.ln1_w + c) + ldg4(ex.ln1_b + c));
      } else if (xa != ws) {
        ROW_VECS(k, c) st4(ws + r * ld + c, v[k]);
      }
    }
    __syncthreads();
    // grouped down projection: z[n][g*Crg+jl] = sum_cl xn[n][g*Cg+cl] * wd[g*Crg+jl][cl]
    for (int g = 0; g < G; ++g)
      contract_to<kExact, Out::Store>(TN, Crg, Cg, ws + g * Cg, ld, 1, ex.wd + (long long)g * Crg * Cg, 1, Cg,
                                      zt + e * Cr + g * Crg, lz, 1);
    __syncthreads();
  }
  // coalesced store of the z tile [valid][E*Cr] and the bn1 batch sums
  float* zg = p.z + ((long long)b * p.N + n0) * ECr;
  for (int n = warp; n < valid; n += nwarps)
    for (int q = lane; q < ECr; q += 32) zg[(long long)n * ECr + q] = zt[n * lz + q];
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

// Register-blocked form: a thread owns groups of kB consecutive outputs (one a_j broadcast feeds kB products, kB independent
// FMA chains), index arithmetic hoisted out of the tile loop.  Same sums, same fp32-per-tile / fp64-across-tiles accumulation.
template <int kB>
__global__ void __launch_bounds__(kThreads) bn_moments_blk_kernel(const avmoe_expert_args p, double* m1, double* m2, int TN, long long rows_total) {
  extern __shared__ float sm[];
  constexpr int kGroups = kMomPerThread / kB;
  const int Cr = p.Cr, G = p.G, Crg = Cr / G, E = p.E;
  const int ECr = E * Cr;
  float* as = sm;  // [TN][ECr+1]
  const int lda = ECr + 1;
  float* colp = sm + (size_t)TN * lda;   // [3][ECr]: bn1 scale, shift, relu flag per bottleneck column (column q = e*Cr + j)
  for (int q = threadIdx.x; q < ECr; q += blockDim.x) {
    colp[q] = p.aff1[q * 2];
    colp[ECr + q] = p.aff1[q * 2 + 1];
    colp[2 * ECr + q] = p.ex[q / Cr].relu ? 1.f : 0.f;
  }
  __syncthreads();
  const int kq = Crg / kB;
  const int n_m1g = ECr / kB, n_m2g = E * G * Crg * kq;
  const int n_groups = n_m1g + n_m2g;
  float accv[kGroups][kB];
  int cjv[kGroups], ckv[kGroups];      // smem columns: a_j (or -1 for first-moment groups) and the first of the kB partner columns
  long long outv[kGroups];             // linear output index (m1: column, m2: offset by ECr)
#pragma unroll
  for (int t = 0; t < kGroups; ++t) {
#pragma unroll
    for (int b = 0; b < kB; ++b) accv[t][b] = 0.f;
    const int gi = threadIdx.x + t * kThreads;
    cjv[t] = -2; ckv[t] = 0; outv[t] = 0;
    if (gi < n_m1g) {
      cjv[t] = -1; ckv[t] = gi * kB; outv[t] = gi * kB;
    } else if (gi < n_groups) {
      const int q = gi - n_m1g;
      const int e = q / (G * Crg * kq);
      const int r = q - e * G * Crg * kq;
      const int g = r / (Crg * kq);
      const int jk = r - g * Crg * kq;
      const int j = jk / kq, kb0 = (jk - j * kq) * kB;
      cjv[t] = e * Cr + g * Crg + j;
      ckv[t] = e * Cr + g * Crg + kb0;
      outv[t] = (long long)ECr + (((long long)(e * G + g) * Crg + j) * Crg + kb0);
    }
  }
  const long long tiles = (rows_total + TN - 1) / TN;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long r0 = tile * TN;
    const int valid = (int)min((long long)TN, rows_total - r0);
    {   // the tile of z is one contiguous chunk: flat coalesced 16-byte loads (ECr % 4 == 0 for kB = 4, else scalar), up to
        // kDepth requests in flight per thread, (row, column) tracked incrementally -- the fill is latency-bound otherwise
      const float* zt = p.z + r0 * ECr;
      const int total = valid * ECr, cap = TN * ECr;
      constexpr int kV = kB == 4 ? 4 : 1;          // elements per load
      constexpr int kDepth = 8;
      const int stride = kV * kThreads;
      const int sq = stride % ECr, sn = stride / ECr;
      int q = (kV * threadIdx.x) % ECr, n = (kV * threadIdx.x) / ECr;
      for (int i = kV * threadIdx.x; i < cap; i += kDepth * stride) {
        float zv[kDepth][kV];
        int qq[kDepth], nn[kDepth];
#pragma unroll
        for (int u = 0; u < kDepth; ++u) {
          const int ii = i + u * stride;
          qq[u] = q; nn[u] = n;
          if (kV == 4) {
            const float4 t = ii < total ? *reinterpret_cast<const float4*>(zt + ii) : make_float4(0.f, 0.f, 0.f, 0.f);
            zv[u][0] = t.x; zv[u][kV > 1 ? 1 : 0] = t.y; zv[u][kV > 2 ? 2 : 0] = t.z; zv[u][kV > 3 ? 3 : 0] = t.w;
          } else {
            zv[u][0] = ii < total ? zt[ii] : 0.f;
          }
          q += sq; n += sn;
          if (q >= ECr) { q -= ECr; ++n; }
        }
#pragma unroll
        for (int u = 0; u < kDepth; ++u) {
          const int ii = i + u * stride;
          if (ii < cap) {
#pragma unroll
            for (int v = 0; v < kV; ++v) {
              float a = 0.f;
              if (ii < total) {
                a = fmaf(zv[u][v], colp[qq[u] + v], colp[ECr + qq[u] + v]);
                if (colp[2 * ECr + qq[u] + v] != 0.f) a = fmaxf(a, 0.f);
              }
              as[ii + v + nn[u]] = a;      // n * (ECr + 1) + q
            }
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < kGroups; ++t) {
      if (cjv[t] == -2) continue;
      float s[kB];
#pragma unroll
      for (int b = 0; b < kB; ++b) s[b] = 0.f;
      const float* pk = as + ckv[t];
      if (cjv[t] < 0) {
        for (int n = 0; n < valid; ++n) {
#pragma unroll
          for (int b = 0; b < kB; ++b) s[b] += pk[n * lda + b];
        }
      } else {
        const float* pj = as + cjv[t];
        for (int n = 0; n < valid; ++n) {
          const float aj = pj[n * lda];
#pragma unroll
          for (int b = 0; b < kB; ++b) s[b] = fmaf(aj, pk[n * lda + b], s[b]);
        }
      }
#pragma unroll
      for (int b = 0; b < kB; ++b) accv[t][b] += s[b];
    }
    __syncthreads();
  }
#pragma unroll
  for (int t = 0; t < kGroups; ++t) {
    if (cjv[t] == -2) continue;
#pragma unroll
    for (int b = 0; b < kB; ++b) {
      if (cjv[t] < 0) atomicAdd(&m1[outv[t] + b], (double)accv[t][b]);
      else atomicAdd(&m2[outv[t] - ECr + b], (double)accv[t][b]);
    }
  }
}

// Square-blocked form (the default): a work item is a BJ x BJ block of one group's symmetric moment matrix (upper triangle
// only; the mirror is written on output) or BJ first moments.  Per row an item costs two BJ-wide shared loads for BJ*BJ FMAs
// (bn_moments_blk: kB+1 scalar loads for kB FMAs -- shared-load bound), and the tile is a flat copy of z's rows.
template <int BJ>
struct MomVec;
template <>
struct MomVec<4> { using type = float4; };
template <>
struct MomVec<2> { using type = float2; };

template <int BJ, int kItems>
__global__ void __launch_bounds__(kThreads, kItems <= 3 ? 2 : 1) bn_moments_sq_kernel(const avmoe_expert_args p, double* m1, double* m2, int TN, long long rows_total) {
  extern __shared__ __align__(16) float sm[];
  using Vec = typename MomVec<BJ>::type;
  const int Cr = p.Cr, G = p.G, Crg = Cr / G, E = p.E;
  const int ECr = E * Cr;                 // multiple of BJ (launcher)
  float* as = sm;                         // [TN][ECr]
  float* colp = sm + (size_t)TN * ECr;    // [3][ECr]: bn1 scale, shift, relu flag per bottleneck column
  for (int q = threadIdx.x; q < ECr; q += blockDim.x) {
    colp[q] = p.aff1[q * 2];
    colp[ECr + q] = p.aff1[q * 2 + 1];
    colp[2 * ECr + q] = p.ex[q / Cr].relu ? 1.f : 0.f;
  }
  const int nb = Crg / BJ;
  const int per_group = nb * (nb + 1) / 2;
  const int n_sq = E * G * per_group, n_items = n_sq + ECr / BJ;
  float acc[kItems][BJ][BJ];
  int cj[kItems], ck[kItems];             // tile columns of the item's two operands; cj = -1: first moments of ck.., -2: no item
  const int items_pad = kItems == 1 ? min(kThreads, (n_items + 31) & ~31) : kThreads;      // whole warps per row group
  const int rgroups = kItems == 1 ? kThreads / items_pad : 1;
  const int rg = kItems == 1 ? (int)threadIdx.x / items_pad : 0;
#pragma unroll
  for (int t = 0; t < kItems; ++t) {
#pragma unroll
    for (int a = 0; a < BJ; ++a)
#pragma unroll
      for (int b = 0; b < BJ; ++b) acc[t][a][b] = 0.f;
    // few items (narrow bottleneck): the threads left over form extra row groups, each taking every rgroups-th row of the tile
    const int gi = kItems == 1 ? (int)threadIdx.x % items_pad : (int)threadIdx.x + t * kThreads;
    cj[t] = -2; ck[t] = 0;
    if (kItems == 1 && (int)threadIdx.x / items_pad >= rgroups) {
      // idle tail of the block
    } else if (gi < n_sq) {
      const int eg = gi / per_group;
      int r = gi - eg * per_group, jb = 0;
      while (r >= nb - jb) { r -= nb - jb; ++jb; }
      cj[t] = eg * Crg + jb * BJ;          // e*Cr + g*Crg == (e*G + g)*Crg
      ck[t] = eg * Crg + (jb + r) * BJ;
    } else if (gi < n_items) {
      cj[t] = -1; ck[t] = (gi - n_sq) * BJ;
    }
  }
  __syncthreads();
  const long long tiles = (rows_total + TN - 1) / TN;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long r0 = tile * TN;
    const int valid = (int)min((long long)TN, rows_total - r0);
    {   // flat, coalesced BJ-wide copy of the tile with the bn1 affine (+ReLU) applied; kDepth requests in flight per thread
      const float* zt = p.z + r0 * ECr;
      const int total = valid * ECr;
      constexpr int kDepth = 8;
      const int stride = BJ * kThreads;
      const int sq = stride % ECr;
      int q = (BJ * threadIdx.x) % ECr;
      for (int i = BJ * threadIdx.x; i < total; i += kDepth * stride) {
        Vec zv[kDepth];
        int qq[kDepth];
#pragma unroll
        for (int u = 0; u < kDepth; ++u) {
          const int ii = i + u * stride;
          qq[u] = q;
          if (ii < total) zv[u] = *reinterpret_cast<const Vec*>(zt + ii);
          q += sq;
          if (q >= ECr) q -= ECr;
        }
#pragma unroll
        for (int u = 0; u < kDepth; ++u) {
          const int ii = i + u * stride;
          if (ii < total) {
            float* v = reinterpret_cast<float*>(&zv[u]);
#pragma unroll
            for (int b = 0; b < BJ; ++b) {
              float a = fmaf(v[b], colp[qq[u] + b], colp[ECr + qq[u] + b]);
              if (colp[2 * ECr + qq[u] + b] != 0.f) a = fmaxf(a, 0.f);
              v[b] = a;
            }
            *reinterpret_cast<Vec*>(as + ii) = zv[u];
          }
        }
      }
    }
    __syncthreads();
    for (int n = rg; n < valid; n += rgroups) {
      const float* row = as + n * ECr;
#pragma unroll
      for (int t = 0; t < kItems; ++t) {
        if (cj[t] == -2) continue;
        const Vec kv = *reinterpret_cast<const Vec*>(row + ck[t]);
        const float* ak = reinterpret_cast<const float*>(&kv);
        if (cj[t] < 0) {
#pragma unroll
          for (int b = 0; b < BJ; ++b) acc[t][0][b] += ak[b];
        } else {
          const Vec jv = *reinterpret_cast<const Vec*>(row + cj[t]);
          const float* aj = reinterpret_cast<const float*>(&jv);
#pragma unroll
          for (int a = 0; a < BJ; ++a)
#pragma unroll
            for (int b = 0; b < BJ; ++b) acc[t][a][b] = fmaf(aj[a], ak[b], acc[t][a][b]);
        }
      }
    }
    __syncthreads();
  }
  if (kItems == 1 && rgroups > 1 && TN * ECr >= kThreads * BJ * BJ) {
    // the row groups of an item are combined inside the CTA first (fixed order), so each output receives ONE fp64 atomic per CTA:
    // same-address atomics serialise in L2, and 4-8 x as many of them had become the tail of the narrow-bottleneck launches
    float* scr = as;                       // tile storage is free after the last barrier of the tile loop
#pragma unroll
    for (int a = 0; a < BJ; ++a)
#pragma unroll
      for (int b = 0; b < BJ; ++b) scr[(threadIdx.x * BJ + a) * BJ + b] = acc[0][a][b];
    __syncthreads();
    if (rg == 0) {
      for (int g = 1; g < rgroups; ++g)
#pragma unroll
        for (int a = 0; a < BJ; ++a)
#pragma unroll
          for (int b = 0; b < BJ; ++b) acc[0][a][b] += scr[((g * items_pad + (int)threadIdx.x) * BJ + a) * BJ + b];
    } else {
      cj[0] = -2;
    }
  }
#pragma unroll
  for (int t = 0; t < kItems; ++t) {
    if (cj[t] == -2) continue;
    if (cj[t] < 0) {
#pragma unroll
      for (int b = 0; b < BJ; ++b) atomicAdd(&m1[ck[t] + b], (double)acc[t][0][b]);
    } else {
      const int eg = cj[t] / Crg;
      const int j0 = cj[t] - eg * Crg, k0 = ck[t] - eg * Crg;
      double* mm = m2 + (long long)eg * Crg * Crg;
#pragma unroll
      for (int a = 0; a < BJ; ++a)
#pragma unroll
        for (int b = 0; b < BJ; ++b) {
          atomicAdd(&mm[(j0 + a) * Crg + k0 + b], (double)acc[t][a][b]);
          if (j0 != k0) atomicAdd(&mm[(k0 + b) * Crg + j0 + a], (double)acc[t][a][b]);
        }
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
  // one warp per output channel: the quadratic form wᵀ M2 w over the channel's conv group is spread over the lanes
  const int e = blockIdx.x;
  const int Cg = a.C / a.G, Crg = a.Cr / a.G;
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= a.C) return;
  double mean, var;
  if (a.train) {
    const int g = c / Cg;
    const float* wr = a.wu[e] + (long long)c * Crg;
    const double* m1 = a.m1 + e * a.Cr + g * Crg;
    const double* m2 = a.m2 + ((long long)(e * a.G + g) * Crg) * Crg;
    double mu = 0.0, ey2 = 0.0;
    for (int j = lane; j < Crg; j += 32) {
      mu += (double)wr[j] * m1[j];
      double t = 0.0;
      for (int k = 0; k < Crg; ++k) t += (double)wr[k] * m2[(long long)j * Crg + k];
      ey2 += (double)wr[j] * t;
    }
    for (int o = 16; o > 0; o >>= 1) { mu += __shfl_xor_sync(0xffffffffu, mu, o); ey2 += __shfl_xor_sync(0xffffffffu, ey2, o); }
    mean = mu / (double)a.n;
    var = ey2 / (double)a.n - mean * mean;
    if (var < 0) var = 0;
    if (lane == 0) {
      const double unbiased = var * ((double)a.n / (double)(a.n > 1 ? a.n - 1 : 1));
      a.rm2[e][c] = (float)((1.0 - kBnMomentum) * a.rm2[e][c] + kBnMomentum * mean);
      a.rv2[e][c] = (float)((1.0 - kBnMomentum) * a.rv2[e][c] + kBnMomentum * unbiased);
    }
  } else {
    mean = a.rm2[e][c]; var = a.rv2[e][c];
  }
  if (lane == 0) {
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
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const bool bn = p.bn_mode != AVMOE_BN_OFF;
  for (int n = warp; n < TN; n += nwarps) {
    const float* zr = p.z + (((long long)b * p.N + n0 + n) * p.E + e) * Cr;
    for (int j = lane; j < Cr; j += 32) {
      float a = 0.f;
      if (n < valid) {
        float zt = zr[j];
        if (bn) zt = fmaf(zt, __ldg(p.aff1 + (e * Cr + j) * 2), __ldg(p.aff1 + (e * Cr + j) * 2 + 1));
        a = ex.relu ? fmaxf(zt, 0.f) : zt;
      }
      as[n * lda + j] = a;
    }
  }
  __syncthreads();
  for (int g = 0; g < G; ++g)
    contract_to<kExact, Out::Store>(TN, Cg, Crg, as + g * Crg, lda, 1, ex.wu + (long long)g * Cg * Crg, 1, Crg, ys + g * Cg, ld, 1);
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ back forward
template <typename T, int kMaxV>
__global__ void __launch_bounds__(kThreads, 2) expert_back_fwd_kernel(const avmoe_expert_args p, int TN) {
  extern __shared__ float sm[];
  constexpr bool kExact = std::is_same<T, float>::value;
  const int C = p.C, Cr = p.Cr, ld = C + 4;
  float* ys = sm;                      // [TN][ld]
  float* acc = ys + TN * ld;           // [TN][ld]
  float* as = acc + TN * ld;           // [TN][Cr+4]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int b = blockIdx.y, n0 = blockIdx.x * TN;
  const int valid = min(TN, p.N - n0);
  const bool bn = p.bn_mode != AVMOE_BN_OFF;
  for (int e = 0; e < p.E; ++e) {
    const avmoe_expert& ex = p.ex[e];
    back_recompute<kExact>(p, ex, e, b, n0, valid, TN, as, ys, ld);
    const float scale = p.probs[(long long)b * p.E + e] * (ex.gate ? *ex.gate : 1.f);
    // row pass: bn2 affine -> ln_post -> acc += probs*gate*q
    for (int r = warp; r < valid; r += nwarps) {
      float4 v[kMaxV];
      float sum = 0.f;
      ROW_VECS(k, c) {
        v[k] = ld4(ys + r * ld + c);
        if (bn) { float4 sc, sh; ldg_pairs4(p.aff2 + ((long long)e * C + c) * 2, sc, sh); v[k] = v[k] * sc + sh; }
        sum += hsum(v[k]);
      }
      if (ex.ln2_w) {
        const float mean = warp_sum(sum) / C;
        float ss = 0.f;
        ROW_VECS(k, c) { const float4 d = v[k] - f4(mean); ss += hsum(d * d); }
        const float rstd = rsqrtf(warp_sum(ss) / C + kLnEps);
        ROW_VECS(k, c) v[k] = (rstd * (v[k] - f4(mean))) * ldg4(ex.ln2_w + c) + ldg4(ex.ln2_b + c);
      }
      ROW_VECS(k, c) {
        float4 o = scale * v[k];
        if (e > 0) o = o + ld4(acc + r * ld + c);
        st4(acc + r * ld + c, o);
      }
    }
    __syncthreads();
  }
  T* ob = reinterpret_cast<T*>(p.out) + ((long long)b * p.N + n0) * C;
  store_tile<T, kMaxV>(ob, acc, valid, C, ld);
}

// ------------------------------------------------------------------------------------------------ back backward
// kStatsOnly: only the bn2 batch reductions (sum dyt, sum dyt*yhat); otherwise the full backward of the back half.
template <typename T, bool kStatsOnly, int kMaxV>
__global__ void __launch_bounds__(kThreads, 2) expert_back_bwd_kernel(const avmoe_expert_args p, int TN) {
  extern __shared__ float sm[];
  constexpr bool kExact = std::is_same<T, float>::value;
  const int C = p.C, Cr = p.Cr, G = p.G, Cg = C / G, Crg = Cr / G, ld = C + 4, lda = Cr + 4;
  float* gs0 = sm;                     // [TN][ld]  dout tile (kept)
  float* ys = gs0 + TN * ld;           // [TN][ld]  y (pre-bn2)
  float* gs = ys + TN * ld;            // [TN][ld]  dy
  float* as = gs + TN * ld;            // [TN][Cr+4]
  float* dzs = as + TN * lda;          // [TN][Cr+4] dzt tile of the current expert
  float* mds = dzs + TN * lda;         // [2][C] batch means of dyt and dyt*yhat (bn train)
  float* rowm = mds + 2 * C;           // [TN] ln_post row means
  float* rowr = rowm + TN;             // [TN] ln_post row rstd
  float* scratch = rowr + TN;          // [32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int b = blockIdx.y, n0 = blockIdx.x * TN;
  const int valid = min(TN, p.N - n0);
  const int mode = p.bn_mode;
  const double inv_n = 1.0 / ((double)p.BT * p.N);
  const T* gb = reinterpret_cast<const T*>(p.dout) + ((long long)b * p.N + n0) * C;
  load_tile<T, kMaxV>(gs0, gb, valid, TN, C, ld);
  __syncthreads();
  for (int e = 0; e < p.E; ++e) {
    const avmoe_expert& ex = p.ex[e];
    if (!kStatsOnly && mode == AVMOE_BN_TRAIN)
      for (int c = threadIdx.x; c < C; c += blockDim.x) {
        mds[c] = (float)(p.red2[((long long)e * C + c) * 2] * inv_n);
        mds[C + c] = (float)(p.red2[((long long)e * C + c) * 2 + 1] * inv_n);
      }
    back_recompute<kExact>(p, ex, e, b, n0, valid, TN, as, ys, ld);   // (syncs)
    const float pe = p.probs[(long long)b * p.E + e];
    const float gate = ex.gate ? *ex.gate : 1.f;
    const float sc = pe * gate;
    float gq = 0.f;
    // ---- one row pass: yt, ln_post fwd stats, dq, ln_post bwd, (bn2 bwd) -> gs = dy (bn train / off) or dyt (stats, eval)
    for (int r = warp; r < TN; r += nwarps) {
      if (r >= valid) {   // keep the tile zero below the valid rows (it feeds the weight-gradient contraction)
        ROW_VECS(k, c) st4(gs + r * ld + c, f4(0.f));
        if (lane == 0) { rowm[r] = 0.f; rowr[r] = 0.f; }
        continue;
      }
      float4 yt[kMaxV], yh[kMaxV], d[kMaxV];
      float4 s2c[kMaxV];
      float sum = 0.f;
      ROW_VECS(k, c) {
        const float4 y = ld4(ys + r * ld + c);
        yt[k] = y;
        if (mode != AVMOE_BN_OFF) {
          float4 a2s, a2h, mu, is;
          ldg_pairs4(p.aff2 + ((long long)e * C + c) * 2, a2s, a2h);
          ldg_pairs4(p.save2 + ((long long)e * C + c) * 2, mu, is);
          yt[k] = y * a2s + a2h;
          yh[k] = (y - mu) * is;
          s2c[k] = a2s;
        }
        sum += hsum(yt[k]);
      }
      float mean = 0.f, rstd = 1.f;
      if (ex.ln2_w) {
        mean = warp_sum(sum) / C;
        float ss = 0.f;
        ROW_VECS(k, c) { const float4 t = yt[k] - f4(mean); ss += hsum(t * t); }
        rstd = rsqrtf(warp_sum(ss) / C + kLnEps);
        if (lane == 0) { rowm[r] = mean; rowr[r] = rstd; }
      }
      float s1 = 0.f, s2v = 0.f;
      ROW_VECS(k, c) {
        const float4 g0 = ld4(gs0 + r * ld + c);
        const float4 dq = sc * g0;
        if (ex.ln2_w) {
          const float4 w = ldg4(ex.ln2_w + c);
          const float4 qh = rstd * (yt[k] - f4(mean));
          const float4 q = qh * w + ldg4(ex.ln2_b + c);
          gq += hsum(g0 * q);
          d[k] = dq * w;                         // dqh
          s1 += hsum(d[k]);
          s2v += hsum(d[k] * qh);
          yt[k] = qh;                            // keep qh for the LN backward
        } else {
          gq += hsum(g0 * yt[k]);
          d[k] = dq;
        }
      }
      if (ex.ln2_w) {
        const float m1 = warp_sum(s1) / C, m2 = warp_sum(s2v) / C;
        ROW_VECS(k, c) d[k] = rstd * (d[k] - f4(m1) - m2 * yt[k]);      // dyt
      }
      if (mode == AVMOE_BN_TRAIN && !kStatsOnly) {
        ROW_VECS(k, c) d[k] = (d[k] - ld4(mds + c) - yh[k] * ld4(mds + C + c)) * s2c[k];   // dy
      }
      ROW_VECS(k, c) st4(gs + r * ld + c, d[k]);
    }
    __syncthreads();
    // ---- per-channel sums (one thread per channel, rows in order, no shared-memory atomics)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const bool bn = mode != AVMOE_BN_OFF;
      const float a2s = bn ? p.aff2[((long long)e * C + c) * 2] : 1.f, a2h = bn ? p.aff2[((long long)e * C + c) * 2 + 1] : 0.f;
      if (!kStatsOnly && ex.ln2_w) {
        float sw = 0.f, sb = 0.f;
        for (int n = 0; n < valid; ++n) {
          const float qh = (fmaf(ys[n * ld + c], a2s, a2h) - rowm[n]) * rowr[n];
          const float dq = sc * gs0[n * ld + c];
          sw = fmaf(dq, qh, sw);
          sb += dq;
        }
        atomicAdd(ex.g_ln2_w + c, sw);
        atomicAdd(ex.g_ln2_b + c, sb);
      }
      if (bn && (kStatsOnly || mode == AVMOE_BN_EVAL)) {
        const float mu = p.save2[((long long)e * C + c) * 2], is = p.save2[((long long)e * C + c) * 2 + 1];
        float s1 = 0.f, s2v = 0.f;
        for (int n = 0; n < valid; ++n) {
          const float dd = gs[n * ld + c];           // dyt
          s1 += dd;
          s2v = fmaf(dd, (ys[n * ld + c] - mu) * is, s2v);
          if (!kStatsOnly) gs[n * ld + c] = dd * a2s;   // eval mode: dy = scale * dyt
        }
        double* red = kStatsOnly ? p.red2 : p.red2b;
        atomicAdd(&red[((long long)e * C + c) * 2], (double)s1);
        atomicAdd(&red[((long long)e * C + c) * 2 + 1], (double)s2v);
      }
    }
    __syncthreads();
    if (kStatsOnly) continue;
    {
      const float tot = block_sum(gq, scratch);
      if (threadIdx.x == 0) {
        atomicAdd(&p.dprobs[(long long)b * p.E + e], tot * gate);
        if (ex.g_gate) atomicAdd(ex.g_gate, tot * pe);
      }
    }
    // ---- d up_sampler.weight[c][jl] += sum_n dy[n][c] * a[n][g*Crg+jl]      (rows >= valid of gs are zero)
    for (int g = 0; g < G; ++g)
      contract_to<kExact, Out::AtomicAdd>(Cg, Crg, TN, gs + g * Cg, 1, ld, as + g * Crg, lda, 1, ex.g_wu + (long long)g * Cg * Crg, Crg, 1);
    // ---- da[n][g*Crg+jl] = sum_cl dy[n][g*Cg+cl] * wu[g*Cg+cl][jl]
    for (int g = 0; g < G; ++g)
      contract_to<kExact, Out::Store>(TN, Crg, Cg, gs + g * Cg, ld, 1, ex.wu + (long long)g * Cg * Crg, Crg, 1, dzs + g * Crg, lda, 1);
    __syncthreads();
    // ---- relu mask, store dzt (coalesced rows of Cr)
    for (int n = warp; n < valid; n += nwarps) {
      float* dzr = p.dzt + (((long long)b * p.N + n0 + n) * p.E + e) * Cr;
      for (int j = lane; j < Cr; j += 32) {
        float dd = dzs[n * lda + j];
        if (ex.relu && !(as[n * lda + j] > 0.f)) dd = 0.f;
        dzs[n * lda + j] = dd;
        dzr[j] = dd;
      }
    }
    __syncthreads();
    // ---- bn1-backward sums per bottleneck channel
    if (mode != AVMOE_BN_OFF) {
      for (int j = threadIdx.x; j < Cr; j += blockDim.x) {
        const float mu = p.save1[(e * Cr + j) * 2], is = p.save1[(e * Cr + j) * 2 + 1];
        float s1 = 0.f, s2v = 0.f;
        for (int n = 0; n < valid; ++n) {
          const float dd = dzs[n * lda + j];
          s1 += dd;
          s2v = fmaf(dd, (p.z[(((long long)b * p.N + n0 + n) * p.E + e) * Cr + j] - mu) * is, s2v);
        }
        atomicAdd(&p.red1[((long long)e * Cr + j) * 2], (double)s1);
        atomicAdd(&p.red1[((long long)e * Cr + j) * 2 + 1], (double)s2v);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ front backward
template <typename T, int kMaxV>
__global__ void __launch_bounds__(kThreads, 2) expert_front_bwd_kernel(const avmoe_expert_args p, int TN) {
  extern __shared__ float sm[];
  constexpr bool kExact = std::is_same<T, float>::value;
  const int C = p.C, Cr = p.Cr, G = p.G, Cg = C / G, Crg = Cr / G, ld = C + 4, lda = Cr + 4, lt = p.T + 4;
  float* xs = sm;                      // [TN][ld] x tile
  float* ws = xs + TN * ld;            // [TN][ld] xa -> xhat
  float* t1 = ws + TN * ld;            // [TN][ld] xn -> dxn -> dxa
  float* dxs = t1 + TN * ld;           // [TN][ld] accumulated dx
  float* s2 = dxs + TN * ld;           // [TN][T+4] token<-latent probabilities
  float* ds2 = s2 + TN * lt;           // [TN][T+4]
  float* dzs = ds2 + TN * lt;          // [TN][Cr+4]
  float* rstds = dzs + TN * lda;       // [TN]
  float* scratch = rstds + TN;         // [32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int b = blockIdx.y, n0 = blockIdx.x * TN;
  const int valid = min(TN, p.N - n0);
  const int mode = p.bn_mode;
  const double inv_n = 1.0 / ((double)p.BT * p.N);
  const T* xb = reinterpret_cast<const T*>(p.x) + ((long long)b * p.N + n0) * C;
  load_tile<T, kMaxV>(xs, xb, valid, TN, C, ld);
  for (int r = warp; r < TN; r += nwarps) {
    ROW_VECS(k, c) st4(dxs + r * ld + c, p.dpool_x ? ldg4(p.dpool_x + (long long)b * C + c) : f4(0.f));
  }
  __syncthreads();
  for (int e = 0; e < p.E; ++e) {
    const avmoe_expert& ex = p.ex[e];
    // ---- dz = bn1 backward of dzt (rows of Cr, coalesced)
    for (int n = warp; n < TN; n += nwarps) {
      const long long zrow = (((long long)b * p.N + n0 + n) * p.E + e) * Cr;
      for (int j = lane; j < Cr; j += 32) {
        float d = 0.f;
        if (n < valid) {
          d = p.dzt[zrow + j];
          if (mode != AVMOE_BN_OFF) {
            if (mode == AVMOE_BN_TRAIN) {
              const float zh = (p.z[zrow + j] - p.save1[(e * Cr + j) * 2]) * p.save1[(e * Cr + j) * 2 + 1];
              d = d - (float)(p.red1[((long long)e * Cr + j) * 2] * inv_n) - zh * (float)(p.red1[((long long)e * Cr + j) * 2 + 1] * inv_n);
            }
            d *= p.aff1[(e * Cr + j) * 2];
          }
        }
        dzs[n * lda + j] = d;
      }
    }
    // ---- recompute xa; row pass A: LN statistics, ws <- xhat, t1 <- xn
    const float* xa = pre_transform<T, kExact, kMaxV>(ex, p, b, n0, valid, TN, xs, ws, s2, ld);
    __syncthreads();
    for (int r = warp; r < TN; r += nwarps) {
      float4 v[kMaxV];
      float sum = 0.f;
      ROW_VECS(k, c) { v[k] = ld4(xa + r * ld + c); sum += hsum(v[k]); }
      if (ex.ln1_w) {
        const float mean = warp_sum(sum) / C;
        float ss = 0.f;
        ROW_VECS(k, c) { const float4 d = v[k] - f4(mean); ss += hsum(d * d); }
        const float rstd = rsqrtf(warp_sum(ss) / C + kLnEps);
        if (lane == 0) rstds[r] = rstd;
        ROW_VECS(k, c) {
          const float4 xh = rstd * (v[k] - f4(mean));
          st4(ws + r * ld + c, xh);
          st4(t1 + r * ld + c, xh * ldg4(ex.ln1_w + c) + ldg4(ex.ln1_b + c));
        }
      } else {
        ROW_VECS(k, c) { st4(t1 + r * ld + c, v[k]); }
      }
    }
    __syncthreads();
    // ---- d down_sampler.weight[g*Crg+jl][cl] += sum_n dz[n][g*Crg+jl] * xn[n][g*Cg+cl]   (dz rows >= valid are 0)
    for (int g = 0; g < G; ++g)
      contract_to<kExact, Out::AtomicAdd>(Crg, Cg, TN, dzs + g * Crg, 1, lda, t1 + g * Cg, ld, 1, ex.g_wd + (long long)g * Crg * Cg, Cg, 1);
    __syncthreads();
    // ---- dxn[n][g*Cg+cl] = sum_jl dz[n][g*Crg+jl] * wd[g*Crg+jl][cl]
    for (int g = 0; g < G; ++g)
      contract_to<kExact, Out::Store>(TN, Cg, Crg, dzs + g * Crg, lda, 1, ex.wd + (long long)g * Crg * Cg, Cg, 1, t1 + g * Cg, ld, 1);
    __syncthreads();
    // ---- d ln_before weight / bias: column sums of dxn*xhat and dxn (one thread per channel, no atomics in smem)
    if (ex.ln1_w) {
      for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float sw = 0.f, sb = 0.f;
        for (int n = 0; n < valid; ++n) { const float d = t1[n * ld + c]; sw = fmaf(d, ws[n * ld + c], sw); sb += d; }
        atomicAdd(ex.g_ln1_w + c, sw);
        atomicAdd(ex.g_ln1_b + c, sb);
      }
      __syncthreads();
    }
    // ---- row pass B: ln_before backward -> dxa (t1), dx accumulation, aux gradients
    float dg = 0.f;
    const T* auxp = ex.aux ? reinterpret_cast<const T*>(ex.aux) + ((long long)b * p.N + n0) * C : nullptr;
    T* gaux = ex.g_aux ? reinterpret_cast<T*>(ex.g_aux) + ((long long)b * p.N + n0) * C : nullptr;
    const float gpre = ex.gate_pre ? *ex.gate_pre : 0.f;
    for (int r = warp; r < TN; r += nwarps) {
      if (r >= valid) { ROW_VECS(k, c) st4(t1 + r * ld + c, f4(0.f)); continue; }
      float4 d[kMaxV], xh[kMaxV];
      float s1 = 0.f, s2v = 0.f;
      ROW_VECS(k, c) {
        d[k] = ld4(t1 + r * ld + c);
        if (ex.ln1_w) {
          xh[k] = ld4(ws + r * ld + c);
          d[k] = d[k] * ldg4(ex.ln1_w + c);
          s1 += hsum(d[k]);
          s2v += hsum(d[k] * xh[k]);
        }
      }
      if (ex.ln1_w) {
        const float m1 = warp_sum(s1) / C, m2 = warp_sum(s2v) / C, rstd = rstds[r];
        ROW_VECS(k, c) d[k] = rstd * (d[k] - f4(m1) - m2 * xh[k]);          // dxa
      }
      ROW_VECS(k, c) {
        st4(t1 + r * ld + c, d[k]);
        if (ex.pre_kind != AVMOE_PRE_AUX_REPLACE) st4(dxs + r * ld + c, ld4(dxs + r * ld + c) + d[k]);
        if (ex.pre_kind == AVMOE_PRE_AUX_ADD) {
          dg += hsum(d[k] * ldg_act4<T>(auxp + (long long)r * C + c));
          st_act4<T>(gaux + (long long)r * C + c, gpre * d[k]);
        } else if (ex.pre_kind == AVMOE_PRE_AUX_REPLACE) {
          st_act4<T>(gaux + (long long)r * C + c, d[k]);
        }
      }
    }
    __syncthreads();
    if (ex.pre_kind == AVMOE_PRE_LATENT) {
      const float* tp = p.tokp + ((long long)b * p.nl + ex.latent_slot) * p.T * C;
      float* dtp = p.dtokp + ((long long)b * p.nl + ex.latent_slot) * p.T * C;
      // u[n][t] = dxa_n . tok'_t
      contract_to<kExact, Out::Store>(TN, p.T, C, t1, ld, 1, tp, 1, C, ds2, lt, 1);
      __syncthreads();
      // d gate_pre = sum A2*u ; dS2 = A2 * g * (u - sum_t A2*u)      (one warp per row, lanes over t)
      for (int r = warp; r < TN; r += nwarps) {
        float dot = 0.f;
        for (int t = lane; t < p.T; t += 32) dot = fmaf(s2[r * lt + t], ds2[r * lt + t], dot);
        dot = warp_sum(dot);
        if (r < valid && lane == 0) dg += dot;
        for (int t = lane; t < p.T; t += 32) ds2[r * lt + t] = (r < valid) ? s2[r * lt + t] * gpre * (ds2[r * lt + t] - dot) : 0.f;
      }
      __syncthreads();
      // dx[n][c] += sum_t dS2[n][t] * tok'[t][c]
      contract_to<kExact, Out::Add>(TN, C, p.T, ds2, lt, 1, tp, C, 1, dxs, ld, 1);
      // d tok'[t][c] += sum_n A2[n][t]*g*dxa[n][c] + dS2[n][t]*x[n][c]       (rows >= valid of t1 / xs are zero)
      contract_to<kExact, Out::AtomicAdd>(p.T, C, TN, s2, 1, lt, t1, ld, 1, dtp, C, 1, gpre);
      contract_to<kExact, Out::AtomicAdd>(p.T, C, TN, ds2, 1, lt, xs, ld, 1, dtp, C, 1);
    }
    if (ex.g_gate_pre) {
      const float tot = block_sum(dg, scratch);
      if (threadIdx.x == 0) atomicAdd(ex.g_gate_pre, tot);
    }
    __syncthreads();
  }
  T* dxb = reinterpret_cast<T*>(p.dx) + ((long long)b * p.N + n0) * C;
  store_tile<T, kMaxV>(dxb, dxs, valid, C, ld);
}

// ------------------------------------------------------------------------------------------------ host side
static int pow2_floor(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }
// rows per token tile: `nbuf` fp32 [TN][C+4] buffers should stay near 100 KiB so two CTAs share an SM; wide channels
// (stage 3/4) would then get tiles of a few tokens whose per-tile weight traffic dominates, so they may use up to ~190 KiB.
static int pick_tn(int C, int nbuf) {
  static int kb_override = -1;   // diagnostics: AVMOE_TILE_KB overrides the shared-memory budget per CTA
  if (kb_override < 0) { const char* e = getenv("AVMOE_TILE_KB"); kb_override = e ? atoi(e) : 0; }
  int tn = ((kb_override ? kb_override : (nbuf >= 4 ? 64 : 100)) * 1024 / 4) / (nbuf * (C + 4));
  if (tn < 16) tn = (190 * 1024 / 4) / (nbuf * (C + 4));
  if (tn < 1) tn = 1;
  if (tn > 64) tn = 64;
  return pow2_floor(tn);
}

static int check_common(const avmoe_expert_args* a, const char* who) {
  AVMOE_CHECK_ARG(a != nullptr, "%s: null args", who);
  AVMOE_CHECK_ARG(a->E >= 1 && a->E <= AVMOE_MAX_EXPERTS, "%s: E=%d out of range", who, a->E);
  AVMOE_CHECK_ARG(a->BT > 0 && a->N > 0 && a->C > 0 && a->Cr > 0 && a->G > 0, "%s: bad sizes", who);
  AVMOE_CHECK_ARG(a->C % 4 == 0 && a->C % a->G == 0 && a->Cr % a->G == 0, "%s: C must be a multiple of 4 and of G; Cr a multiple of G", who);
  AVMOE_CHECK_ARG(a->C <= 128 * kMaxVLimit, "%s: C=%d exceeds the supported maximum of %d channels", who, a->C, 128 * kMaxVLimit);
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
  const int TN = pick_tn(a->C, 2);
  const size_t smem = sizeof(float) * (2 * (size_t)TN * (a->C + 4) + (size_t)TN * (a->T + 4) + (size_t)TN * (a->E * a->Cr + 1));
  dim3 grid((a->N + TN - 1) / TN, a->BT);
  AVMOE_DISPATCH_V(a->C, {
    if (a->dtype == AVMOE_F32) {
      if (int rc = set_smem(expert_front_fwd_kernel<float, kV>, smem, "expert_front_fwd")) return rc;
      expert_front_fwd_kernel<float, kV><<<grid, kThreads, smem, st>>>(*a, TN);
    } else {
      if (int rc = set_smem(expert_front_fwd_kernel<__nv_bfloat16, kV>, smem, "expert_front_fwd")) return rc;
      expert_front_fwd_kernel<__nv_bfloat16, kV><<<grid, kThreads, smem, st>>>(*a, TN);
    }
  });
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
  // square-blocked kernel: Crg and E*Cr multiples of the block edge, at most 4 items per thread
  const int BJ = (Crg % 4 == 0 && ECr % 4 == 0) ? 4 : ((Crg % 2 == 0 && ECr % 2 == 0) ? 2 : 0);
  long long n_items = 0;
  if (BJ) n_items = (long long)a->E * a->G * ((Crg / BJ) * (Crg / BJ + 1) / 2) + ECr / BJ;
  static const bool sq_on = []() { const char* e = getenv("AVMOE_BN_MOMENTS_SQ"); return !(e && e[0] == '0'); }();
  if (sq_on && BJ && n_items <= 4LL * kThreads && ECr <= 2048) {
    int TN = 8192 / ECr;
    if (TN > 256) TN = 256;
    if (TN < 1) TN = 1;
    const size_t smem = sizeof(float) * ((size_t)TN * ECr + 3 * (size_t)ECr);
    const long long tiles = (rows + TN - 1) / TN;
    const int kI = (int)((n_items + kThreads - 1) / kThreads);
    int occ = 1;
#define AVMOE_MOM_SQ(BJ_, KI_)                                                                                               \
    {                                                                                                                        \
      if (int rc = set_smem(bn_moments_sq_kernel<BJ_, KI_>, smem, "bn_moments")) return rc;                                    \
      AVMOE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bn_moments_sq_kernel<BJ_, KI_>, kThreads, smem));       \
      if (occ < 1) occ = 1;                                                                                                  \
      if (occ > 3) occ = 3;                                                                                                  \
      long long cap = (4LL << 20) / n_out;                 /* bound the fp64 atomics of the final flush */                   \
      if (cap < num_sms()) cap = num_sms();                                                                                  \
      long long g = (long long)occ * num_sms();                                                                              \
      if (g > cap) g = cap;                                                                                                  \
      if (g > tiles) g = tiles;                                                                                              \
      bn_moments_sq_kernel<BJ_, KI_><<<(int)g, kThreads, smem, st>>>(*a, m1, m2, TN, rows);                                   \
    }
    if (BJ == 4) {
      if (kI == 1) AVMOE_MOM_SQ(4, 1) else if (kI == 2) AVMOE_MOM_SQ(4, 2) else if (kI == 3) AVMOE_MOM_SQ(4, 3) else AVMOE_MOM_SQ(4, 4)
    } else {
      if (kI == 1) AVMOE_MOM_SQ(2, 1) else if (kI == 2) AVMOE_MOM_SQ(2, 2) else if (kI == 3) AVMOE_MOM_SQ(2, 3) else AVMOE_MOM_SQ(2, 4)
    }
#undef AVMOE_MOM_SQ
  } else if (n_out <= (long long)kMomPerThread * kThreads) {
    int TN = 8192 / (ECr + 1);
    if (TN > 128) TN = 128;
    if (TN < 1) TN = 1;
    const size_t smem = sizeof(float) * ((size_t)TN * (ECr + 1) + 3 * (size_t)ECr);
    long long tiles = (rows + TN - 1) / TN;
    int grid = (int)(tiles < 2LL * num_sms() ? tiles : 2LL * num_sms());
    if (int rc = set_smem(bn_moments_kernel, smem, "bn_moments")) return rc;
    if (Crg % 4 == 0 && ECr % 4 == 0) {
      if (int rc = set_smem(bn_moments_blk_kernel<4>, smem, "bn_moments")) return rc;
      bn_moments_blk_kernel<4><<<grid, kThreads, smem, st>>>(*a, m1, m2, TN, rows);
    } else if (Crg % 2 == 0 && ECr % 2 == 0) {
      if (int rc = set_smem(bn_moments_blk_kernel<2>, smem, "bn_moments")) return rc;
      bn_moments_blk_kernel<2><<<grid, kThreads, smem, st>>>(*a, m1, m2, TN, rows);
    } else {
      bn_moments_kernel<<<grid, kThreads, smem, st>>>(*a, m1, m2, TN, rows);
    }
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
  bn2_finalize_kernel<<<dim3(a->E, (a->C + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*a);
  AVMOE_LAUNCH_CHECK("bn2_finalize_kernel");
  return 0;
}

extern "C" int avmoe_expert_back_fwd(const avmoe_expert_args* a, avmoe_stream_t stream) {
  if (int rc = check_common(a, "expert_back_fwd")) return rc;
  AVMOE_CHECK_ARG(a->z && a->probs && a->out, "expert_back_fwd: null tensor");
  AVMOE_CHECK_ARG(a->bn_mode == AVMOE_BN_OFF || (a->aff1 && a->aff2), "expert_back_fwd: batch-norm affine missing");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int TN = pick_tn(a->C, 2);
  const size_t smem = sizeof(float) * (2 * (size_t)TN * (a->C + 4) + (size_t)TN * (a->Cr + 4));
  dim3 grid((a->N + TN - 1) / TN, a->BT);
  AVMOE_DISPATCH_V(a->C, {
    if (a->dtype == AVMOE_F32) {
      if (int rc = set_smem(expert_back_fwd_kernel<float, kV>, smem, "expert_back_fwd")) return rc;
      expert_back_fwd_kernel<float, kV><<<grid, kThreads, smem, st>>>(*a, TN);
    } else {
      if (int rc = set_smem(expert_back_fwd_kernel<__nv_bfloat16, kV>, smem, "expert_back_fwd")) return rc;
      expert_back_fwd_kernel<__nv_bfloat16, kV><<<grid, kThreads, smem, st>>>(*a, TN);
    }
  });
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
  const int TN = pick_tn(a->C, 3);
  const size_t smem = sizeof(float) * (3 * (size_t)TN * (a->C + 4) + 2 * (size_t)TN * (a->Cr + 4) + 2 * (size_t)a->C + 2 * TN + 32);
  dim3 grid((a->N + TN - 1) / TN, a->BT);
  AVMOE_DISPATCH_V(a->C, {
    if (a->dtype == AVMOE_F32) {
      if (int rc = set_smem(expert_back_bwd_kernel<float, kStatsOnly, kV>, smem, who)) return rc;
      expert_back_bwd_kernel<float, kStatsOnly, kV><<<grid, kThreads, smem, st>>>(*a, TN);
    } else {
      if (int rc = set_smem(expert_back_bwd_kernel<__nv_bfloat16, kStatsOnly, kV>, smem, who)) return rc;
      expert_back_bwd_kernel<__nv_bfloat16, kStatsOnly, kV><<<grid, kThreads, smem, st>>>(*a, TN);
    }
  });
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
  const int TN = pick_tn(a->C, 4);
  const size_t smem = sizeof(float) * (4 * (size_t)TN * (a->C + 4) + 2 * (size_t)TN * (a->T + 4) + (size_t)TN * (a->Cr + 4) + TN + 32);
  dim3 grid((a->N + TN - 1) / TN, a->BT);
  AVMOE_DISPATCH_V(a->C, {
    if (a->dtype == AVMOE_F32) {
      if (int rc = set_smem(expert_front_bwd_kernel<float, kV>, smem, "expert_front_bwd")) return rc;
      expert_front_bwd_kernel<float, kV><<<grid, kThreads, smem, st>>>(*a, TN);
    } else {
      if (int rc = set_smem(expert_front_bwd_kernel<__nv_bfloat16, kV>, smem, "expert_front_bwd")) return rc;
      expert_front_bwd_kernel<__nv_bfloat16, kV><<<grid, kThreads, smem, st>>>(*a, TN);
    }
  });
  AVMOE_LAUNCH_CHECK("expert_front_bwd_kernel");
  return 0;
}
