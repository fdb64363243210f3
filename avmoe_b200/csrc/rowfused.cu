// Row-fused expert passes for bf16 token streams (AVMOE/AVE/nets/net_trans_v3.py:379-434, 481-486; call-site add :703-720).
//
// One warp owns a tile of 16 token rows.  All per-row work of an adapter's experts -- latent-token attention of the row against
// the T <= 96 updated latent tokens, gate_pre residual, ln_before, grouped down projection (front pass); bn1 affine, activation,
// grouped up projection, bn2, ln_post, gate, probability-weighted combine over experts, residual add (back pass) -- is done on
// that tile without leaving the SM: contractions on warp-level bf16 tensor-core tiles (csrc/mma_tile.cuh) against packed weight
// operands that live in L1/L2, row statistics by quad shuffles, column (over-rows) gradient sums by an 8-lane reduce-scatter
// into per-CTA shared-memory accumulators.  HBM sees each row's inputs and outputs once per pass; what the GEMM-based path
// (csrc/expert_rows.cu + avmoe_gemm) materialised between kernels -- xl, logits, A2, x_res, a, yhat -- is recomputed instead.
// The passes are HBM-bound by construction; the dense contractions over ALL rows (weight gradients, d tok')
// stay on the tcgen05 GEMM, fed by the operands the backward passes emit.
#include "mma_tile.cuh"
#include "row_pass.cuh"

namespace avmoe {
namespace rf {

static constexpr int kMaxC = 1536;

// ------------------------------------------------------------------------------------------------ packed B operands
__global__ void rf_pack_b_kernel(const float* __restrict__ src, long long bstride, long long stride_n, long long stride_k, int nb, int Nn,
                                 int K, int groups, int group_local_k, const float* __restrict__ scale_n, int sns,
                                 const float* __restrict__ scale_k, int sks, int hi_lo, uint4* __restrict__ dst, int NP, int KS) {
  const long long total = (long long)nb * NP * KS * 32;
  const int ng = Nn / groups, kg = K / groups;
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const int lane = (int)(o & 31);
    long long q = o >> 5;
    const int s = (int)(q % KS); q /= KS;
    const int p = (int)(q % NP);
    const int b = (int)(q / NP);
    const int g = lane >> 2, t = lane & 3;
    uint32_t regs[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int half = r & 1;
      const int n = hi_lo ? p * 8 + g : (2 * p + (r >> 1)) * 8 + g;
      float v[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int k = 16 * s + 8 * half + 2 * t + i;
        float w = 0.f;
        if (n < Nn && k < K) {
          int ns = n, ks = k;
          bool on = true;
          if (groups > 1) {
            const int gn = n / ng, gk = k / kg;
            on = gn == gk;
            if (group_local_k) ks = k - gk * kg; else ns = n - gn * ng;
          }
          if (on) {
            w = __ldg(src + b * bstride + ns * stride_n + ks * stride_k);
            if (scale_n) w *= __ldg(scale_n + ((long long)b * Nn + n) * sns);
            if (scale_k) w *= __ldg(scale_k + ((long long)b * K + k) * sks);
          }
        }
        if (hi_lo && (r >> 1)) w = w - __bfloat162float(__float2bfloat16_rn(w));   // residual of the bf16 rounding
        v[i] = w;
      }
      regs[r] = pack_bf16x2(v[0], v[1]);
    }
    dst[o] = make_uint4(regs[0], regs[1], regs[2], regs[3]);
  }
}

// ------------------------------------------------------------------------------------------------ tile <-> global helpers
// 16-row x C bf16 tile, rows contiguous in global memory (row pitch C elements); shared-memory row pitch = 2C + 16 bytes so
// that ldmatrix row fetches and accumulator-layout 4-byte accesses are bank-conflict free (C % 16 == 0).
__device__ __forceinline__ void tile_load_async(unsigned char* tile, int pitch, const __nv_bfloat16* gsrc, long long r0, long long rows, int C,
                                                int lane) {
  const int chunks = C >> 3;                      // 16-byte chunks per row
  const int nrow = (int)min(16LL, rows - r0);
  for (int i = lane; i < nrow * chunks; i += 32) {
    const int r = i / chunks, ch = i - r * chunks;
    cp_async16(tile + r * pitch + ch * 16, gsrc + (r0 + r) * C + ch * 8);
  }
}
// coalesced store of a staged tile; ADD: fp32 add of a residual row read from `gres` first
template <bool ADD>
__device__ __forceinline__ void tile_store(const unsigned char* tile, int pitch, __nv_bfloat16* gdst, long long ld, const __nv_bfloat16* gres,
                                           long long r0, long long rows, int C, int lane) {
  const int chunks = C >> 3;
  const int nrow = (int)min(16LL, rows - r0);
  for (int i = lane; i < nrow * chunks; i += 32) {
    const int r = i / chunks, ch = i - r * chunks;
    uint4 v = *reinterpret_cast<const uint4*>(tile + r * pitch + ch * 16);
    if (ADD) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(gres + (r0 + r) * C + ch * 8));
      uint32_t* vv = reinterpret_cast<uint32_t*>(&v);
      const uint32_t* qq = reinterpret_cast<const uint32_t*>(&q);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_bf16x2(vv[j]), b = unpack_bf16x2(qq[j]);
        vv[j] = pack_bf16x2(a.x + b.x, a.y + b.y);
      }
    }
    *reinterpret_cast<uint4*>(gdst + (r0 + r) * ld + ch * 8) = v;
  }
}
__device__ __forceinline__ uint32_t& tile_at(unsigned char* tile, int pitch, int row, int col) {
  return *reinterpret_cast<uint32_t*>(tile + row * pitch + col * 2);
}

// ------------------------------------------------------------------------------------------------ back pass (forward and backward)
// k-steps of the up projection that can be non-zero for the n-tile pair starting at column c0 (block-diagonal weights)
__device__ __forceinline__ void up_ksteps(int c0, int C, int Cr, int G, int& lo, int& hi) {
  const int cg = C / G, crg = Cr / G;
  const int g0 = c0 / cg, g1 = (c0 + 15) / cg;
  lo = (g0 * crg) >> 4;
  hi = min((Cr + 15) >> 4, ((g1 + 1) * crg + 15) >> 4);
}

template <int E, int CRK, bool BWD>
__global__ void __launch_bounds__(256) rf_back_kernel(const avmoe_rf_back_args p, long long rows, int tiles_per_cta) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int C = p.C, Cr = p.Cr, N = p.N;
  const int pitch = C * 2 + 16;
  const int NP = C >> 4;
  const float invC = 1.f / C;
  // shared memory: [BWD: per-CTA column accumulators E x 4 x C fp32 + E gate sums] [per warp: 1 (fwd) or 3 (bwd) tiles]
  float* colacc = reinterpret_cast<float*>(smem);
  const size_t acc_bytes = BWD ? ((size_t)E * 4 * C + 32) * sizeof(float) : 0;
  unsigned char* wbase = smem + acc_bytes + (size_t)warp * (BWD ? 3 : 1) * 16 * pitch;
  unsigned char* t_stage = wbase;                         // fwd: out tile; bwd: dy2 tile
  unsigned char* t_dout = wbase + (BWD ? 16 * pitch : 0);
  unsigned char* t_yh = wbase + (BWD ? 32 * pitch : 0);
  if (BWD) {
    for (int i = threadIdx.x; i < E * 4 * C + 32; i += blockDim.x) colacc[i] = 0.f;
    __syncthreads();
  }
  float* gateacc = colacc + (size_t)E * 4 * C;
  const uint4* wup = static_cast<const uint4*>(p.wu_pack);
  const size_t wu_stride = (size_t)NP * CRK * 32;        // uint4 per expert
  const long long ntiles = (rows + 15) >> 4;
  const long long tile0 = (long long)blockIdx.x * tiles_per_cta;
  const long long tile1 = min(ntiles, tile0 + tiles_per_cta);
  float gate_sum[E];
#pragma unroll
  for (int e = 0; e < E; ++e) gate_sum[e] = 0.f;

  for (long long tile = tile0 + warp; tile < tile1; tile += nwarps) {
    const long long r0 = tile << 4;
    const long long rowA = r0 + g, rowB = r0 + g + 8;
    const bool vA = rowA < rows, vB = rowB < rows;
    if (BWD) {
      tile_load_async(t_dout, pitch, static_cast<const __nv_bfloat16*>(p.dout), r0, rows, C, lane);
      cp_async_commit();
    }
    // ---- a = act(bn1(z)) as A fragments (K = bottleneck channels, zero padded to 16*CRK)
    uint32_t af[E][CRK][4];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const bool relu = p.ex[e].relu != 0;
#pragma unroll
      for (int s = 0; s < CRK; ++s) {
#pragma unroll
        for (int kh = 0; kh < 2; ++kh) {
          const int k0 = 16 * s + 8 * kh + 2 * t;
          float2 zA = make_float2(0.f, 0.f), zB = make_float2(0.f, 0.f);
          if (k0 < Cr) {
            if (vA) zA = __ldg(reinterpret_cast<const float2*>(p.z + (rowA * E + e) * Cr + k0));
            if (vB) zB = __ldg(reinterpret_cast<const float2*>(p.z + (rowB * E + e) * Cr + k0));
            if (p.aff1) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(p.aff1 + ((size_t)e * Cr + k0) * 2));   // sc0 sh0 sc1 sh1
              zA.x = zA.x * a.x + a.y; zA.y = zA.y * a.z + a.w;
              zB.x = zB.x * a.x + a.y; zB.y = zB.y * a.z + a.w;
            }
            if (relu) { zA.x = fmaxf(zA.x, 0.f); zA.y = fmaxf(zA.y, 0.f); zB.x = fmaxf(zB.x, 0.f); zB.y = fmaxf(zB.y, 0.f); }
            if (!vA) zA = make_float2(0.f, 0.f);
            if (!vB) zB = make_float2(0.f, 0.f);
          }
          af[e][s][2 * kh] = pack_bf16x2(zA.x, zA.y);
          af[e][s][2 * kh + 1] = pack_bf16x2(zB.x, zB.y);
          if (BWD && k0 < p.Crp) {     // a is also an operand of the weight-gradient GEMM (padding columns are zero)
            __nv_bfloat16* act = static_cast<__nv_bfloat16*>(p.act) + (size_t)e * rows * p.Crp;
            if (vA) *reinterpret_cast<uint32_t*>(act + rowA * p.Crp + k0) = af[e][s][2 * kh];
            if (vB) *reinterpret_cast<uint32_t*>(act + rowB * p.Crp + k0) = af[e][s][2 * kh + 1];
          }
        }
      }
      if (BWD) {   // zero the padding columns beyond 16*CRK (Crp is a multiple of 32)
        __nv_bfloat16* act = static_cast<__nv_bfloat16*>(p.act) + (size_t)e * rows * p.Crp;
        for (int k0 = 16 * CRK + 2 * t; k0 < p.Crp; k0 += 8) {
          if (vA) *reinterpret_cast<uint32_t*>(act + rowA * p.Crp + k0) = 0u;
          if (vB) *reinterpret_cast<uint32_t*>(act + rowB * p.Crp + k0) = 0u;
        }
      }
    }
    // ---- per-row combine weights
    const long long bA = vA ? rowA / N : 0, bB = vB ? rowB / N : 0;
    float scA[E], scB[E], gate[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      gate[e] = p.ex[e].gate ? __ldg(p.ex[e].gate) : 1.f;
      scA[e] = vA ? __ldg(p.probs + bA * E + e) * gate[e] : 0.f;
      scB[e] = vB ? __ldg(p.probs + bB * E + e) * gate[e] : 0.f;
    }
    // yhat (normalised up projection) of an n-tile pair: acc[h][0..3], h = n-tile within the pair
    auto up_pair = [&](int e, int pr, float (&acc)[2][4]) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[h][i] = 0.f;
      int lo, hi;
      up_ksteps(pr << 4, C, Cr, p.G, lo, hi);
      const uint4* w = wup + e * wu_stride;
#pragma unroll
      for (int s = 0; s < CRK; ++s) {
        if (s >= lo && s < hi) {
          const uint4 b = ldg_b(w, pr, CRK, s, lane);
          mma16816(acc[0], af[e][s], b.x, b.y);
          mma16816(acc[1], af[e][s], b.z, b.w);
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float2 yb = __ldg(reinterpret_cast<const float2*>(p.ybi + (size_t)e * C + (pr << 4) + h * 8 + 2 * t));
        acc[h][0] += yb.x; acc[h][1] += yb.y; acc[h][2] += yb.x; acc[h][3] += yb.y;
      }
    };
    // bn2 affine: the tensor ln_post sees
    auto bn2_aff = [&](int e, int c, float (&v)[4]) {
      if (p.bn2_w[e]) {
        const float2 w = __ldg(reinterpret_cast<const float2*>(p.bn2_w[e] + c)), b = __ldg(reinterpret_cast<const float2*>(p.bn2_b[e] + c));
        v[0] = v[0] * w.x + b.x; v[1] = v[1] * w.y + b.y; v[2] = v[2] * w.x + b.x; v[3] = v[3] * w.y + b.y;
      }
    };
    // ---- pass 1: ln_post row statistics (one pass, fp32 sums)
    float mean[E][2], rstd[E][2];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      mean[e][0] = mean[e][1] = 0.f;
      rstd[e][0] = rstd[e][1] = 1.f;
      if (p.ex[e].ln2_w) {
        float s1A = 0.f, s2A = 0.f, s1B = 0.f, s2B = 0.f;
        for (int pr = 0; pr < NP; ++pr) {
          float acc[2][4];
          up_pair(e, pr, acc);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            bn2_aff(e, (pr << 4) + h * 8 + 2 * t, acc[h]);
            s1A += acc[h][0] + acc[h][1]; s2A += acc[h][0] * acc[h][0] + acc[h][1] * acc[h][1];
            s1B += acc[h][2] + acc[h][3]; s2B += acc[h][2] * acc[h][2] + acc[h][3] * acc[h][3];
          }
        }
        s1A = quad_sum(s1A); s2A = quad_sum(s2A); s1B = quad_sum(s1B); s2B = quad_sum(s2B);
        mean[e][0] = s1A * invC; mean[e][1] = s1B * invC;
        rstd[e][0] = rsqrtf(fmaxf(s2A * invC - mean[e][0] * mean[e][0], 0.f) + kLnEps);
        rstd[e][1] = rsqrtf(fmaxf(s2B * invC - mean[e][1] * mean[e][1], 0.f) + kLnEps);
      }
    }
    if (!BWD) {
      // ---- pass 2: out = (res +) sum_e probs*gate * ln_post(bn2(yhat_e))
      for (int pr = 0; pr < NP; ++pr) {
        float o[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int i = 0; i < 4; ++i) o[h][i] = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          float acc[2][4];
          up_pair(e, pr, acc);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = (pr << 4) + h * 8 + 2 * t;
            bn2_aff(e, c, acc[h]);
            if (p.ex[e].ln2_w) {
              const float2 w = __ldg(reinterpret_cast<const float2*>(p.ex[e].ln2_w + c)), b = __ldg(reinterpret_cast<const float2*>(p.ex[e].ln2_b + c));
              acc[h][0] = (acc[h][0] - mean[e][0]) * rstd[e][0] * w.x + b.x;
              acc[h][1] = (acc[h][1] - mean[e][0]) * rstd[e][0] * w.y + b.y;
              acc[h][2] = (acc[h][2] - mean[e][1]) * rstd[e][1] * w.x + b.x;
              acc[h][3] = (acc[h][3] - mean[e][1]) * rstd[e][1] * w.y + b.y;
            }
            o[h][0] += scA[e] * acc[h][0]; o[h][1] += scA[e] * acc[h][1];
            o[h][2] += scB[e] * acc[h][2]; o[h][3] += scB[e] * acc[h][3];
          }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = (pr << 4) + h * 8 + 2 * t;
          tile_at(t_stage, pitch, g, c) = pack_bf16x2(o[h][0], o[h][1]);
          tile_at(t_stage, pitch, g + 8, c) = pack_bf16x2(o[h][2], o[h][3]);
        }
      }
      __syncwarp();
      if (p.res) tile_store<true>(t_stage, pitch, static_cast<__nv_bfloat16*>(p.out), C, static_cast<const __nv_bfloat16*>(p.res), r0, rows, C, lane);
      else tile_store<false>(t_stage, pitch, static_cast<__nv_bfloat16*>(p.out), C, nullptr, r0, rows, C, lane);
      __syncwarp();
    } else {
      cp_async_wait<0>();
      __syncwarp();
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const bool has_ln = p.ex[e].ln2_w != nullptr;
        const bool has_bn = p.bn2_w[e] != nullptr;
        float* acc_e = colacc + (size_t)e * 4 * C;       // [4][C]: d ln2_w, d ln2_b, sum dy2, sum dy2*yhat
        // ---- pass 2: <dout, o_e> per row (d probs, d gate), ln_post backward row sums, d ln_post weight/bias column sums
        float dotA = 0.f, dotB = 0.f, maA = 0.f, maB = 0.f, mbA = 0.f, mbB = 0.f;
        for (int pr = 0; pr < NP; ++pr) {
          float acc[2][4];
          up_pair(e, pr, acc);
          float cs[8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = (pr << 4) + h * 8 + 2 * t;
            bn2_aff(e, c, acc[h]);
            const float2 dA = unpack_bf16x2(tile_at(t_dout, pitch, g, c)), dB = unpack_bf16x2(tile_at(t_dout, pitch, g + 8, c));
            const float d[4] = {dA.x, dA.y, dB.x, dB.y};
            float2 w = make_float2(1.f, 1.f), b = make_float2(0.f, 0.f);
            if (has_ln) { w = __ldg(reinterpret_cast<const float2*>(p.ex[e].ln2_w + c)); b = __ldg(reinterpret_cast<const float2*>(p.ex[e].ln2_b + c)); }
            float xh[4];
            xh[0] = (acc[h][0] - mean[e][0]) * rstd[e][0]; xh[1] = (acc[h][1] - mean[e][0]) * rstd[e][0];
            xh[2] = (acc[h][2] - mean[e][1]) * rstd[e][1]; xh[3] = (acc[h][3] - mean[e][1]) * rstd[e][1];
            dotA += d[0] * (xh[0] * w.x + b.x) + d[1] * (xh[1] * w.y + b.y);
            dotB += d[2] * (xh[2] * w.x + b.x) + d[3] * (xh[3] * w.y + b.y);
            const float t0 = scA[e] * d[0], t1 = scA[e] * d[1], t2 = scB[e] * d[2], t3 = scB[e] * d[3];   // d o_e
            maA += t0 * w.x + t1 * w.y; maB += t2 * w.x + t3 * w.y;
            mbA += t0 * w.x * xh[0] + t1 * w.y * xh[1]; mbB += t2 * w.x * xh[2] + t3 * w.y * xh[3];
            cs[4 * h + 0] = t0 * xh[0] + t2 * xh[2]; cs[4 * h + 1] = t1 * xh[1] + t3 * xh[3];    // d ln2_w columns c, c+1
            cs[4 * h + 2] = t0 + t2; cs[4 * h + 3] = t1 + t3;                                      // d ln2_b
          }
          if (has_ln) {
            // cs index i -> (h = i>>2, quantity q = (i>>1)&1, column offset i&1)
            const float s = reduce_scatter8(cs, lane);
            const int i = rs_index(lane);
            atomicAdd(acc_e + ((i >> 1) & 1) * C + (pr << 4) + (i >> 2) * 8 + 2 * t + (i & 1), s);
          }
        }
        dotA = quad_sum(dotA); dotB = quad_sum(dotB);
        maA = quad_sum(maA) * invC; maB = quad_sum(maB) * invC; mbA = quad_sum(mbA) * invC; mbB = quad_sum(mbB) * invC;
        // d probs[b,e] += gate * <dout, o_e>; d gate += probs * <dout, o_e>
        if (!vA) dotA = 0.f;
        if (!vB) dotB = 0.f;
        if (t == 0 && p.ex[e].gate) gate_sum[e] += __ldg(p.probs + bA * E + e) * dotA + __ldg(p.probs + bB * E + e) * dotB;
        const long long b_first = r0 / N, b_last = min(r0 + 15, rows - 1) / N;
        if (b_first == b_last) {          // the usual case: the tile lies inside one sample -> one atomic per tile
          float s = t == 0 ? dotA + dotB : 0.f;
          s += __shfl_xor_sync(0xffffffffu, s, 4); s += __shfl_xor_sync(0xffffffffu, s, 8); s += __shfl_xor_sync(0xffffffffu, s, 16);
          if (lane == 0) atomicAdd(p.dprobs + b_first * E + e, gate[e] * s);
        } else if (t == 0) {
          if (vA) atomicAdd(p.dprobs + bA * E + e, gate[e] * dotA);
          if (vB) atomicAdd(p.dprobs + bB * E + e, gate[e] * dotB);
        }
        // ---- pass 3: dy2 = gradient at the bn2 output; yhat; bn2 batch reductions
        __nv_bfloat16* yd = static_cast<__nv_bfloat16*>(p.yd) + (size_t)e * rows * 2 * C;
        for (int pr = 0; pr < NP; ++pr) {
          float acc[2][4];
          up_pair(e, pr, acc);
          float cs[8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = (pr << 4) + h * 8 + 2 * t;
            const float yh[4] = {acc[h][0], acc[h][1], acc[h][2], acc[h][3]};
            tile_at(t_yh, pitch, g, c) = pack_bf16x2(yh[0], yh[1]);
            tile_at(t_yh, pitch, g + 8, c) = pack_bf16x2(yh[2], yh[3]);
            bn2_aff(e, c, acc[h]);
            const float2 dA = unpack_bf16x2(tile_at(t_dout, pitch, g, c)), dB = unpack_bf16x2(tile_at(t_dout, pitch, g + 8, c));
            float dy[4] = {scA[e] * dA.x, scA[e] * dA.y, scB[e] * dB.x, scB[e] * dB.y};
            if (has_ln) {
              const float2 w = __ldg(reinterpret_cast<const float2*>(p.ex[e].ln2_w + c));
              const float xh0 = (acc[h][0] - mean[e][0]) * rstd[e][0], xh1 = (acc[h][1] - mean[e][0]) * rstd[e][0];
              const float xh2 = (acc[h][2] - mean[e][1]) * rstd[e][1], xh3 = (acc[h][3] - mean[e][1]) * rstd[e][1];
              dy[0] = rstd[e][0] * (dy[0] * w.x - maA - mbA * xh0);
              dy[1] = rstd[e][0] * (dy[1] * w.y - maA - mbA * xh1);
              dy[2] = rstd[e][1] * (dy[2] * w.x - maB - mbB * xh2);
              dy[3] = rstd[e][1] * (dy[3] * w.y - maB - mbB * xh3);
            }
            tile_at(t_stage, pitch, g, c) = pack_bf16x2(dy[0], dy[1]);
            tile_at(t_stage, pitch, g + 8, c) = pack_bf16x2(dy[2], dy[3]);
            cs[4 * h + 0] = dy[0] + dy[2]; cs[4 * h + 1] = dy[1] + dy[3];
            cs[4 * h + 2] = dy[0] * yh[0] + dy[2] * yh[2]; cs[4 * h + 3] = dy[1] * yh[1] + dy[3] * yh[3];
          }
          if (has_bn) {
            const float s = reduce_scatter8(cs, lane);
            const int i = rs_index(lane);
            atomicAdd(acc_e + (2 + ((i >> 1) & 1)) * C + (pr << 4) + (i >> 2) * 8 + 2 * t + (i & 1), s);
          }
        }
        __syncwarp();
        tile_store<false>(t_yh, pitch, yd, 2 * C, nullptr, r0, rows, C, lane);
        tile_store<false>(t_stage, pitch, yd + C, 2 * C, nullptr, r0, rows, C, lane);
        __syncwarp();
      }
    }
  }
  if (BWD) {
    // ---- flush the per-CTA column accumulators and gate sums
#pragma unroll
    for (int e = 0; e < E; ++e) {
      float s = gate_sum[e];
      s += __shfl_xor_sync(0xffffffffu, s, 4); s += __shfl_xor_sync(0xffffffffu, s, 8); s += __shfl_xor_sync(0xffffffffu, s, 16);
      if (lane == 0 && p.ex[e].gate) atomicAdd(gateacc + e, s);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < E * C; i += blockDim.x) {
      const int e = i / C, c = i - e * C;
      const float* a = colacc + (size_t)e * 4 * C;
      if (p.ex[e].ln2_w) { atomicAdd(p.ex[e].g_ln2_w + c, a[c]); atomicAdd(p.ex[e].g_ln2_b + c, a[C + c]); }
      if (p.bn2_w[e]) {
        atomicAdd(p.red2 + ((size_t)e * C + c) * 2, (double)a[2 * C + c]);
        atomicAdd(p.red2 + ((size_t)e * C + c) * 2 + 1, (double)a[3 * C + c]);
      }
    }
    if (threadIdx.x < E && p.ex[threadIdx.x].gate) atomicAdd(p.ex[threadIdx.x].g_gate, gateacc[threadIdx.x]);
  }
}

static int back_check(const avmoe_rf_back_args* a, bool bwd, const char* who) {
  AVMOE_CHECK_ARG(a != nullptr, "%s: null args", who);
  AVMOE_CHECK_ARG(avmoe_rf_supported(a->C, a->Cr, a->E, 0), "%s: geometry not supported by the row-fused passes (C=%d Cr=%d E=%d)", who, a->C,
                  a->Cr, a->E);
  AVMOE_CHECK_ARG(a->z && a->wu_pack && a->ybi && a->probs, "%s: null tensor", who);
  if (bwd) AVMOE_CHECK_ARG(a->dout && a->yd && a->act && a->dprobs, "%s: null backward tensor", who);
  else AVMOE_CHECK_ARG(a->out != nullptr, "%s: null out", who);
  return 0;
}

template <int E, int CRK>
static int launch_back(const avmoe_rf_back_args* a, bool bwd, cudaStream_t st) {
  const long long rows = (long long)a->BT * a->N;
  const long long ntiles = (rows + 15) / 16;
  const int pitch = a->C * 2 + 16;
  const size_t per_warp = (size_t)(bwd ? 3 : 1) * 16 * pitch;
  const size_t acc_bytes = bwd ? ((size_t)E * 4 * a->C + 32) * sizeof(float) : 0;
  int warps = 8;
  while (warps > 1 && acc_bytes + warps * per_warp > 100 * 1024) warps >>= 1;    // two CTAs per SM where possible
  const size_t smem = acc_bytes + warps * per_warp;
  AVMOE_CHECK_ARG(smem <= 227 * 1024, "rf_back: tile does not fit shared memory");
  // persistent-style grid: a few CTAs per SM, each owning a contiguous range of tiles (one flush of the column sums per CTA)
  long long ctas = (long long)num_sms() * (smem > 64 * 1024 ? 2 : 4);
  long long tpc = (ntiles + ctas - 1) / ctas;
  tpc = (tpc + warps - 1) / warps * warps;
  if (tpc < warps) tpc = warps;
  const unsigned grid = (unsigned)((ntiles + tpc - 1) / tpc);
  if (bwd) {
    AVMOE_CUDA(cudaFuncSetAttribute(rf_back_kernel<E, CRK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    rf_back_kernel<E, CRK, true><<<grid, warps * 32, smem, st>>>(*a, rows, (int)tpc);
  } else {
    AVMOE_CUDA(cudaFuncSetAttribute(rf_back_kernel<E, CRK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    rf_back_kernel<E, CRK, false><<<grid, warps * 32, smem, st>>>(*a, rows, (int)tpc);
  }
  AVMOE_LAUNCH_CHECK("rf_back_kernel");
  return 0;
}

template <int E>
static int dispatch_back_crk(const avmoe_rf_back_args* a, bool bwd, cudaStream_t st) {
  const int crk = (a->Cr + 15) / 16;
  switch (crk) {
    case 1: return launch_back<E, 1>(a, bwd, st);
    case 2: return launch_back<E, 2>(a, bwd, st);
    case 3: return launch_back<E, 3>(a, bwd, st);
    case 6: return launch_back<E, 6>(a, bwd, st);
    default: break;
  }
  set_error("rf_back: unsupported bottleneck width %d", a->Cr);
  return AVMOE_E_BADARG;
}

static int dispatch_back(const avmoe_rf_back_args* a, bool bwd, cudaStream_t st) {
  const int rc = back_check(a, bwd, bwd ? "rf_back_bwd" : "rf_back_fwd");
  if (rc) return rc;
  switch (a->E) {
    case 2: return dispatch_back_crk<2>(a, bwd, st);
    case 3: return dispatch_back_crk<3>(a, bwd, st);
    case 4: return dispatch_back_crk<4>(a, bwd, st);
    default: break;
  }
  set_error("rf_back: unsupported expert count %d", a->E);
  return AVMOE_E_BADARG;
}

}  // namespace rf
}  // namespace avmoe

extern "C" int avmoe_rf_supported(int C, int Cr, int E, int T) {
  const int crk = (Cr + 15) / 16;
  const bool crk_ok = crk == 1 || crk == 2 || crk == 3 || crk == 6;
  return (C % 16 == 0 && C <= avmoe::rf::kMaxC && Cr % 2 == 0 && crk_ok && E >= 2 && E <= 4 && T <= 96) ? 1 : 0;
}

extern "C" int64_t avmoe_rf_pack_b_bytes(int nb, int Nn, int K, int hi_lo) {
  const int64_t NP = hi_lo ? (Nn + 7) / 8 : (Nn + 15) / 16, KS = (K + 15) / 16;
  return (int64_t)nb * NP * KS * 32 * 16;
}

extern "C" int avmoe_rf_pack_b(const float* src, int64_t bstride, int64_t stride_n, int64_t stride_k, int nb, int Nn, int K, int groups,
                               int group_local_k, const float* scale_n, int scale_n_stride, const float* scale_k, int scale_k_stride,
                               int hi_lo, void* dst, avmoe_stream_t stream) {
  using namespace avmoe;
  AVMOE_CHECK_ARG(src && dst && nb > 0 && Nn > 0 && K > 0 && groups >= 1, "rf_pack_b: bad arguments");
  AVMOE_CHECK_ARG(Nn % groups == 0 && K % groups == 0, "rf_pack_b: groups must divide both extents");
  const int NP = hi_lo ? (Nn + 7) / 8 : (Nn + 15) / 16, KS = (K + 15) / 16;
  const long long total = (long long)nb * NP * KS * 32;
  const int threads = 256;
  const unsigned grid = (unsigned)std::min<long long>((total + threads - 1) / threads, 4096);
  rf::rf_pack_b_kernel<<<grid, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, bstride, stride_n, stride_k, nb, Nn, K, groups,
                                                                                 group_local_k, scale_n, scale_n_stride, scale_k,
                                                                                 scale_k_stride, hi_lo, static_cast<uint4*>(dst), NP, KS);
  AVMOE_LAUNCH_CHECK("rf_pack_b_kernel");
  return 0;
}

extern "C" int avmoe_rf_back_fwd(const avmoe_rf_back_args* a, avmoe_stream_t stream) {
  return avmoe::rf::dispatch_back(a, false, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int avmoe_rf_back_bwd(const avmoe_rf_back_args* a, avmoe_stream_t stream) {
  return avmoe::rf::dispatch_back(a, true, reinterpret_cast<cudaStream_t>(stream));
}
