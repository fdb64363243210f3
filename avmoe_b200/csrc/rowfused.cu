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

static bool rf_capable(int C, int Cr, int E, int T);

namespace avmoe {
namespace rf {

static constexpr int kMaxC = 1536;
static constexpr int kFastC = 192;

// ------------------------------------------------------------------------------------------------ packed B operands
__global__ void rf_pack_b_kernel(const float* __restrict__ src, long long bstride, long long stride_n, long long stride_k, int nb, int Nn,
                                 int K, int groups, int group_local_k, const float* __restrict__ scale_n, int sns,
                                 const float* __restrict__ scale_k, int sks, int hi_lo, uint4* __restrict__ dst, int NP, int KS) {
  // NP comes from the padded n extent (Nn_pad): columns n >= Nn are zero
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
  const uint32_t mg = magic_of(chunks);
  const int nrow = (int)min(16LL, rows - r0);
  for (int i = lane; i < nrow * chunks; i += 32) {
    const int r = (int)magic_div(i, mg), ch = i - r * chunks;
    cp_async16(tile + r * pitch + ch * 16, gsrc + (r0 + r) * C + ch * 8);
  }
  for (int i = nrow * chunks + lane; i < 16 * chunks; i += 32) {     // rows beyond the tensor: zeros, so no pass needs to mask them
    const int r = (int)magic_div(i, mg), ch = i - r * chunks;
    *reinterpret_cast<uint4*>(tile + r * pitch + ch * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
}
// coalesced store of a staged tile; ADD: fp32 add of a residual row read from `gres` first
template <bool ADD>
__device__ __forceinline__ void tile_store(const unsigned char* tile, int pitch, __nv_bfloat16* gdst, long long ld, const __nv_bfloat16* gres,
                                           long long r0, long long rows, int C, int lane) {
  const int chunks = C >> 3;
  const uint32_t mg = magic_of(chunks);
  const int nrow = (int)min(16LL, rows - r0);
  for (int i = lane; i < nrow * chunks; i += 32) {
    const int r = (int)magic_div(i, mg), ch = i - r * chunks;
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
// the two tile rows a lane touches in accumulator layout (g and g+8) as 32-bit words: element pair (row, columns c, c+1) = A[c >> 1]
struct TileRows {
  uint32_t* A;
  uint32_t* B;
  __device__ __forceinline__ TileRows(const unsigned char* tile, int pitch, int g)
      : A(reinterpret_cast<uint32_t*>(const_cast<unsigned char*>(tile) + g * pitch)),
        B(reinterpret_cast<uint32_t*>(const_cast<unsigned char*>(tile) + (g + 8) * pitch)) {}
};

// Per-warp staging block for one n-tile pair (16 rows x 16 bf16 columns): values leave the tensor-core accumulator layout as 4-byte
// words, a __syncwarp later every lane moves one 16-byte chunk, so each row's 32 bytes reach global memory as ONE full sector
// (4-byte stores straight from the accumulator layout half-fill two sectors per row and were measured ~40 % slower end to end).
struct PairStager {
  static constexpr int kPitch = 80;            // 32 data bytes + padding: conflict-free 4-byte writes, 16-byte aligned rows
  static constexpr int kBytes = 16 * kPitch;
  uint32_t* wA;
  uint32_t* wB;
  const uint4* rd;
  int row, ch;
  __device__ __forceinline__ PairStager(unsigned char* buf, int lane) {
    const int g = lane >> 2, t = lane & 3;
    wA = reinterpret_cast<uint32_t*>(buf + g * kPitch) + t;
    wB = reinterpret_cast<uint32_t*>(buf + (g + 8) * kPitch) + t;
    row = lane >> 1;
    ch = lane & 1;
    rd = reinterpret_cast<const uint4*>(buf + row * kPitch + ch * 16);
  }
  __device__ __forceinline__ void put(int h, uint32_t a, uint32_t b) { wA[4 * h] = a; wB[4 * h] = b; }
  // gdst: element (tile row 0, first column of the pair); ld in elements
  __device__ __forceinline__ void flush(__nv_bfloat16* gdst, long long ld, bool row_valid) {
    __syncwarp();
    const uint4 v = *rd;
    if (row_valid) *reinterpret_cast<uint4*>(gdst + row * ld + ch * 8) = v;
    __syncwarp();
  }
};

// ------------------------------------------------------------------------------------------------ back pass (forward and backward)
// k-steps of the up projection that can be non-zero for the n-tile pair starting at column c0 (block-diagonal weights)
struct GroupGeom {
  int cg, crg, kmax;
  uint32_t mg;
  __device__ GroupGeom(int C, int Cr, int G) : cg(C / G), crg(Cr / G), kmax((Cr + 15) >> 4), mg(magic_of(C / G)) {}
};
__device__ __forceinline__ void up_ksteps(int c0, const GroupGeom& gg, int& lo, int& hi) {
  const int g0 = (int)magic_div(c0, gg.mg), g1 = (int)magic_div(c0 + 15, gg.mg);
  lo = (g0 * gg.crg) >> 4;
  hi = min(gg.kmax, ((g1 + 1) * gg.crg + 15) >> 4);
}

// Per-CTA table of the per-channel parameters of the back pass, 12 floats per channel pair (c, c+1) of expert e:
//   [0,1] ybi   [2,3] bn2 weight (1 without BatchNorm)   [4,5] bn2 bias (0)   [6,7] ln_post weight (1)   [8,9] ln_post bias (0)
// read as three 16-byte shared-memory loads per n-tile instead of five global loads.
__device__ __forceinline__ void back_param_table(float* tab, const avmoe_rf_back_args& p, int E, int C) {
  for (int i = threadIdx.x; i < E * (C >> 1); i += blockDim.x) {
    const int e = i / (C >> 1), c = (i - e * (C >> 1)) * 2;
    float* q = tab + (size_t)i * 12;
    const float2 yb = __ldg(reinterpret_cast<const float2*>(p.ybi + (size_t)e * C + c));
    q[0] = yb.x; q[1] = yb.y;
    float2 w = make_float2(1.f, 1.f), b = make_float2(0.f, 0.f);
    if (p.bn2_w[e]) { w = __ldg(reinterpret_cast<const float2*>(p.bn2_w[e] + c)); b = __ldg(reinterpret_cast<const float2*>(p.bn2_b[e] + c)); }
    q[2] = w.x; q[3] = w.y; q[4] = b.x; q[5] = b.y;
    w = make_float2(1.f, 1.f); b = make_float2(0.f, 0.f);
    if (p.ex[e].ln2_w) { w = __ldg(reinterpret_cast<const float2*>(p.ex[e].ln2_w + c)); b = __ldg(reinterpret_cast<const float2*>(p.ex[e].ln2_b + c)); }
    q[6] = w.x; q[7] = w.y; q[8] = b.x; q[9] = b.y;
    q[10] = q[11] = 0.f;
  }
}
struct ColParams { float2 ybi, bw, bb, lw, lb; };
__device__ __forceinline__ ColParams load_col_params(const float* tab, int e, int C, int c) {
  const float4* q = reinterpret_cast<const float4*>(tab + ((size_t)e * (C >> 1) + (c >> 1)) * 12);
  const float4 a = q[0], b = q[1];
  const float2 d = *reinterpret_cast<const float2*>(q + 2);
  ColParams r;
  r.ybi = make_float2(a.x, a.y); r.bw = make_float2(a.z, a.w); r.bb = make_float2(b.x, b.y); r.lw = make_float2(b.z, b.w); r.lb = d;
  return r;
}

// NPR > 0: C == 16 * NPR and the normalised up projection of the whole tile stays in registers (one tensor-core sweep per expert);
// NPR == 0: the tile is recomputed in every sweep (K = bottleneck channels, a handful of MMAs per n-tile pair).
template <int E, int CRK, int NPR, bool BWD>
__global__ void __launch_bounds__(256) rf_back_kernel(const avmoe_rf_back_args p, long long rows, int tiles_per_cta) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int C = p.C, Cr = p.Cr, N = p.N;
  const int pitch = C * 2 + 16;
  const int NP = C >> 4;
  const float invC = 1.f / C;
  const GroupGeom gg(C, Cr, p.G);
  // shared memory: [parameter table E x C/2 x 12 fp32] [BWD: per-warp column accumulators E x 4 x C fp32, E gate sums]
  //                [per warp: one tile -- the out staging tile (fwd) or the dout tile (bwd)]
  float* tab = reinterpret_cast<float*>(smem);
  const size_t tab_floats = (size_t)E * (C >> 1) * 12;
  float* colacc_all = tab + tab_floats;
  const size_t acc_floats = BWD ? (size_t)nwarps * E * 4 * C + 32 : 0;
  float* gateacc = colacc_all + (BWD ? (size_t)nwarps * E * 4 * C : 0);
  float* wacc = colacc_all + (size_t)warp * E * 4 * C;      // this warp's private accumulators: plain read-modify-write
  const size_t warp_bytes = (size_t)16 * pitch + (BWD ? 2 * PairStager::kBytes : 0);
  unsigned char* wbase = smem + (tab_floats + acc_floats) * sizeof(float) + (size_t)warp * warp_bytes;
  unsigned char* t_stage = wbase;                         // fwd: out tile
  unsigned char* t_dout = wbase;                          // bwd: dout tile
  PairStager st_yh(wbase + 16 * pitch, lane), st_dy(wbase + 16 * pitch + PairStager::kBytes, lane);
  back_param_table(tab, p, E, C);
  if (BWD) for (int i = threadIdx.x; i < (int)acc_floats; i += blockDim.x) colacc_all[i] = 0.f;
  __syncthreads();
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
    // normalised up projection yhat of an n-tile pair: acc[h][0..3], h = n-tile within the pair
    auto up_pair = [&](int e, int pr, float (&acc)[2][4]) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[h][i] = 0.f;
      int lo = 0, hi = CRK;
      if (CRK > 1) up_ksteps(pr << 4, gg, lo, hi);
      const uint4* w = wup + e * wu_stride;
#pragma unroll
      for (int s = 0; s < CRK; ++s) {
        if (s >= lo && s < hi) {
          const uint4 b = ldg_b(w, pr, CRK, s, lane);
          mma16816(acc[0], af[e][s], b.x, b.y);
          mma16816(acc[1], af[e][s], b.z, b.w);
        }
      }
    };
    if (BWD) { cp_async_wait<0>(); __syncwarp(); }
    constexpr bool OREG = NPR > 0 && NPR <= 6;      // forward: the combined output tile stays in registers too
    float o[OREG ? NPR : 1][2][4];
    if (!BWD && OREG) {
#pragma unroll
      for (int pr = 0; pr < (OREG ? NPR : 1); ++pr)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int i = 0; i < 4; ++i) o[pr][h][i] = 0.f;
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const bool has_ln = p.ex[e].ln2_w != nullptr;
      const bool has_bn = p.bn2_w[e] != nullptr;
      float yreg[NPR > 0 ? NPR : 1][2][4];
      // ---- sweep 1: yhat (kept in registers when NPR > 0), ln_post row statistics of y2 = bn2 affine(yhat); in the backward also
      //      the row sums of dout against y2 from which <dout, o_e> and the ln_post backward row means follow without another sweep:
      //      S1 = sum d*lw*y2, S2 = sum d*lw, S3 = sum d*lb   ->   <d, o> = rstd*(S1 - mean*S2) + S3,
      //      mean_c(d o_hat) = sc*S2/C,  mean_c(d o_hat * xhat) = sc*rstd*(S1 - mean*S2)/C
      float mean0 = 0.f, mean1 = 0.f, rstd0 = 1.f, rstd1 = 1.f;
      float S1A = 0.f, S2A = 0.f, S3A = 0.f, S1B = 0.f, S2B = 0.f, S3B = 0.f;
      const uint32_t* dpA = reinterpret_cast<const uint32_t*>(t_dout + g * pitch) + t;          // dout at accumulator positions
      const uint32_t* dpB = reinterpret_cast<const uint32_t*>(t_dout + (g + 8) * pitch) + t;
      const float* tabe = tab + ((size_t)e * (C >> 1) + t) * 12;                               // + (8*pr + 4*h) * 12
      {
        float s1A = 0.f, s2A = 0.f, s1B = 0.f, s2B = 0.f;
        auto body = [&](int pr, float (&acc)[2][4]) {
          up_pair(e, pr, acc);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4* q = reinterpret_cast<const float4*>(tabe + ((pr << 3) + (h << 2)) * 12);
            const float4 qa = q[0], qb = q[1];       // ybi.xy bw.xy | bb.xy lw.xy
            acc[h][0] += qa.x; acc[h][1] += qa.y; acc[h][2] += qa.x; acc[h][3] += qa.y;
            const float y0 = fmaf(acc[h][0], qa.z, qb.x), y1 = fmaf(acc[h][1], qa.w, qb.y);
            const float y2 = fmaf(acc[h][2], qa.z, qb.x), y3 = fmaf(acc[h][3], qa.w, qb.y);
            s1A += y0 + y1; s2A = fmaf(y0, y0, fmaf(y1, y1, s2A)); s1B += y2 + y3; s2B = fmaf(y2, y2, fmaf(y3, y3, s2B));
            if (BWD) {
              const float2 lb = *reinterpret_cast<const float2*>(q + 2);
              const float2 dA = unpack_bf16x2(dpA[(pr << 3) + (h << 2)]), dB = unpack_bf16x2(dpB[(pr << 3) + (h << 2)]);
              const float w0 = dA.x * qb.z, w1 = dA.y * qb.w, w2 = dB.x * qb.z, w3 = dB.y * qb.w;
              S1A = fmaf(w0, y0, fmaf(w1, y1, S1A)); S2A += w0 + w1; S3A = fmaf(dA.x, lb.x, fmaf(dA.y, lb.y, S3A));
              S1B = fmaf(w2, y2, fmaf(w3, y3, S1B)); S2B += w2 + w3; S3B = fmaf(dB.x, lb.x, fmaf(dB.y, lb.y, S3B));
            }
          }
        };
        if (NPR > 0) {
#pragma unroll
          for (int pr = 0; pr < (NPR > 0 ? NPR : 1); ++pr) body(pr, yreg[pr]);
        } else if (has_ln || BWD) {
          for (int pr = 0; pr < NP; ++pr) { float acc[2][4]; body(pr, acc); }
        }
        if (has_ln) {
          s1A = quad_sum(s1A); s2A = quad_sum(s2A); s1B = quad_sum(s1B); s2B = quad_sum(s2B);
          mean0 = s1A * invC; mean1 = s1B * invC;
          rstd0 = rsqrtf(fmaxf(s2A * invC - mean0 * mean0, 0.f) + kLnEps);
          rstd1 = rsqrtf(fmaxf(s2B * invC - mean1 * mean1, 0.f) + kLnEps);
        }
      }
      // yhat of a pair for the later sweeps
      auto yhat_pair = [&](int pr, float (&acc)[2][4]) {
        up_pair(e, pr, acc);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float2 yb = *reinterpret_cast<const float2*>(tab + ((size_t)e * (C >> 1) + (pr << 3) + h * 4 + t) * 12);
          acc[h][0] += yb.x; acc[h][1] += yb.y; acc[h][2] += yb.x; acc[h][3] += yb.y;
        }
      };
      if (!BWD) {
        // ---- sweep 2 (forward): out += probs*gate * ln_post(bn2(yhat_e))
        auto body = [&](int pr, const float (&yh)[2][4], float (&oo)[2][4]) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const ColParams cp = load_col_params(tab, e, C, (pr << 4) + h * 8 + 2 * t);
            const float y0 = yh[h][0] * cp.bw.x + cp.bb.x, y1 = yh[h][1] * cp.bw.y + cp.bb.y;
            const float y2 = yh[h][2] * cp.bw.x + cp.bb.x, y3 = yh[h][3] * cp.bw.y + cp.bb.y;
            oo[h][0] += scA[e] * ((y0 - mean0) * rstd0 * cp.lw.x + cp.lb.x);
            oo[h][1] += scA[e] * ((y1 - mean0) * rstd0 * cp.lw.y + cp.lb.y);
            oo[h][2] += scB[e] * ((y2 - mean1) * rstd1 * cp.lw.x + cp.lb.x);
            oo[h][3] += scB[e] * ((y3 - mean1) * rstd1 * cp.lw.y + cp.lb.y);
          }
        };
        if (OREG) {
#pragma unroll
          for (int pr = 0; pr < (OREG ? NPR : 1); ++pr) body(pr, yreg[pr], o[pr]);
        } else {
#pragma unroll
          for (int pr = 0; pr < (NPR > 0 ? NPR : NP); ++pr) {
            if (NPR == 0 && pr >= NP) break;
            float acc[2][4], oo[2][4];
            if (NPR > 0) {
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[h][i] = yreg[NPR > 0 ? pr : 0][h][i];
            } else {
              yhat_pair(pr, acc);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int c = (pr << 4) + h * 8 + 2 * t;
              if (e == 0) { oo[h][0] = oo[h][1] = oo[h][2] = oo[h][3] = 0.f; }
              else {
                const float2 pA = unpack_bf16x2(tile_at(t_stage, pitch, g, c)), pB = unpack_bf16x2(tile_at(t_stage, pitch, g + 8, c));
                oo[h][0] = pA.x; oo[h][1] = pA.y; oo[h][2] = pB.x; oo[h][3] = pB.y;
              }
            }
            body(pr, acc, oo);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int c = (pr << 4) + h * 8 + 2 * t;
              tile_at(t_stage, pitch, g, c) = pack_bf16x2(oo[h][0], oo[h][1]);
              tile_at(t_stage, pitch, g + 8, c) = pack_bf16x2(oo[h][2], oo[h][3]);
            }
          }
        }
      } else {
        float* acc_e = wacc + (size_t)e * 4 * C;       // [4][C]: d ln2_w, d ln2_b, sum dy2, sum dy2*yhat
        S1A = quad_sum(S1A); S2A = quad_sum(S2A); S3A = quad_sum(S3A); S1B = quad_sum(S1B); S2B = quad_sum(S2B); S3B = quad_sum(S3B);
        const float aA = S1A - mean0 * S2A, aB = S1B - mean1 * S2B;
        float dotA = rstd0 * aA + S3A, dotB = rstd1 * aB + S3B;                                 // <dout, o_e> per row
        const float maA = has_ln ? scA[e] * S2A * invC : 0.f, maB = has_ln ? scB[e] * S2B * invC : 0.f;
        const float mbA = has_ln ? scA[e] * rstd0 * aA * invC : 0.f, mbB = has_ln ? scB[e] * rstd1 * aB * invC : 0.f;
        const float nmr0 = -mean0 * rstd0, nmr1 = -mean1 * rstd1;
        // d probs[b,e] += gate * <dout, o_e>; d gate += probs * <dout, o_e>
        if (!vA) dotA = 0.f;
        if (!vB) dotB = 0.f;
        if (t == 0 && p.ex[e].gate) gate_sum[e] += __ldg(p.probs + bA * E + e) * dotA + __ldg(p.probs + bB * E + e) * dotB;
        const long long b_first = r0 / N, b_last = min(r0 + 15, rows - 1) / N;
        if (b_first == b_last) {          // the usual case: the tile lies inside one sample -> one atomic per tile
          float sdot = t == 0 ? dotA + dotB : 0.f;
          sdot += __shfl_xor_sync(0xffffffffu, sdot, 4); sdot += __shfl_xor_sync(0xffffffffu, sdot, 8); sdot += __shfl_xor_sync(0xffffffffu, sdot, 16);
          if (lane == 0) atomicAdd(p.dprobs + b_first * E + e, gate[e] * sdot);
        } else if (t == 0) {
          if (vA) atomicAdd(p.dprobs + bA * E + e, gate[e] * dotA);
          if (vB) atomicAdd(p.dprobs + bB * E + e, gate[e] * dotB);
        }
        // ---- sweep 2: dy2 = gradient at the bn2 output, yhat out, column sums (d ln_post weight / bias, bn2 batch reductions)
        const int ri = rs_index(lane);
        float* accp = acc_e + ((ri >> 1) & 1) * C + (ri >> 2) * 8 + 2 * t + (ri & 1);            // + 16*pr (+ 2*C for the bn2 sums)
        // yhat | dy2 -> yd [rows, 2C] through the per-warp pair stagers (full 32-byte sectors per row)
        __nv_bfloat16* yde = static_cast<__nv_bfloat16*>(p.yd) + (size_t)e * rows * 2 * C + r0 * 2 * C;
        const bool st_valid = r0 + st_yh.row < rows;
        auto body2 = [&](int pr, const float (&yh)[2][4]) {
          float cl[8], cb[8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int o4 = (pr << 3) + (h << 2);
            const float4* q = reinterpret_cast<const float4*>(tabe + o4 * 12);
            const float4 qa = q[0], qb = q[1];
            const float2 dA = unpack_bf16x2(dpA[o4]), dB = unpack_bf16x2(dpB[o4]);
            st_yh.put(h, pack_bf16x2(yh[h][0], yh[h][1]), pack_bf16x2(yh[h][2], yh[h][3]));
            const float t0 = scA[e] * dA.x, t1 = scA[e] * dA.y, t2 = scB[e] * dB.x, t3 = scB[e] * dB.y;      // d o_e
            float dy0 = t0, dy1 = t1, dy2 = t2, dy3 = t3;
            float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
            if (has_ln) {
              x0 = fmaf(fmaf(yh[h][0], qa.z, qb.x), rstd0, nmr0); x1 = fmaf(fmaf(yh[h][1], qa.w, qb.y), rstd0, nmr0);
              x2 = fmaf(fmaf(yh[h][2], qa.z, qb.x), rstd1, nmr1); x3 = fmaf(fmaf(yh[h][3], qa.w, qb.y), rstd1, nmr1);
              dy0 = rstd0 * fmaf(-mbA, x0, fmaf(t0, qb.z, -maA)); dy1 = rstd0 * fmaf(-mbA, x1, fmaf(t1, qb.w, -maA));
              dy2 = rstd1 * fmaf(-mbB, x2, fmaf(t2, qb.z, -maB)); dy3 = rstd1 * fmaf(-mbB, x3, fmaf(t3, qb.w, -maB));
            }
            st_dy.put(h, pack_bf16x2(dy0, dy1), pack_bf16x2(dy2, dy3));
            cl[4 * h + 0] = fmaf(t0, x0, t2 * x2); cl[4 * h + 1] = fmaf(t1, x1, t3 * x3);             // d ln2_w columns c, c+1
            cl[4 * h + 2] = t0 + t2; cl[4 * h + 3] = t1 + t3;                                        // d ln2_b
            cb[4 * h + 0] = dy0 + dy2; cb[4 * h + 1] = dy1 + dy3;                                    // sum dy2
            cb[4 * h + 2] = fmaf(dy0, yh[h][0], dy2 * yh[h][2]); cb[4 * h + 3] = fmaf(dy1, yh[h][1], dy3 * yh[h][3]);   // sum dy2*yhat
          }
          st_yh.flush(yde + (pr << 4), 2 * C, st_valid);
          st_dy.flush(yde + C + (pr << 4), 2 * C, st_valid);
          if (has_ln) accp[pr << 4] += reduce_scatter8(cl, lane);      // index i -> (h = i>>2, quantity (i>>1)&1, column offset i&1)
          if (has_bn) accp[2 * C + (pr << 4)] += reduce_scatter8(cb, lane);
        };
        if (NPR > 0) {
#pragma unroll
          for (int pr = 0; pr < (NPR > 0 ? NPR : 1); ++pr) body2(pr, yreg[pr]);
        } else {
          for (int pr = 0; pr < NP; ++pr) { float acc[2][4]; yhat_pair(pr, acc); body2(pr, acc); }
        }
      }
    }
    if (!BWD) {
      if (OREG) {
#pragma unroll
        for (int pr = 0; pr < (OREG ? NPR : 1); ++pr)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = (pr << 4) + h * 8 + 2 * t;
            tile_at(t_stage, pitch, g, c) = pack_bf16x2(o[pr][h][0], o[pr][h][1]);
            tile_at(t_stage, pitch, g + 8, c) = pack_bf16x2(o[pr][h][2], o[pr][h][3]);
          }
      }
      __syncwarp();
      if (p.res) tile_store<true>(t_stage, pitch, static_cast<__nv_bfloat16*>(p.out), C, static_cast<const __nv_bfloat16*>(p.res), r0, rows, C, lane);
      else tile_store<false>(t_stage, pitch, static_cast<__nv_bfloat16*>(p.out), C, nullptr, r0, rows, C, lane);
      __syncwarp();
    }
  }
  if (BWD) {
    // ---- flush: gate sums, then the per-warp column accumulators summed over the CTA's warps
#pragma unroll
    for (int e = 0; e < E; ++e) {
      float s = gate_sum[e];
      s += __shfl_xor_sync(0xffffffffu, s, 4); s += __shfl_xor_sync(0xffffffffu, s, 8); s += __shfl_xor_sync(0xffffffffu, s, 16);
      if (lane == 0 && p.ex[e].gate) atomicAdd(gateacc + e, s);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < E * 4 * C; i += blockDim.x) {
      float s = 0.f;
      for (int w = 0; w < nwarps; ++w) s += colacc_all[(size_t)w * E * 4 * C + i];
      const int e = i / (4 * C), q = (i - e * 4 * C) / C, c = i - e * 4 * C - q * C;
      if (q < 2) { if (p.ex[e].ln2_w) atomicAdd((q == 0 ? p.ex[e].g_ln2_w : p.ex[e].g_ln2_b) + c, s); }
      else if (p.bn2_w[e]) atomicAdd(p.red2 + ((size_t)e * C + c) * 2 + (q - 2), (double)s);
    }
    if (threadIdx.x < E && p.ex[threadIdx.x].gate) atomicAdd(p.ex[threadIdx.x].g_gate, gateacc[threadIdx.x]);
  }
}

static int back_check(const avmoe_rf_back_args* a, bool bwd, const char* who) {
  AVMOE_CHECK_ARG(a != nullptr, "%s: null args", who);
  AVMOE_CHECK_ARG(rf_capable(a->C, a->Cr, a->E, 0), "%s: geometry not supported by the row-fused passes (C=%d Cr=%d E=%d)", who, a->C,
                  a->Cr, a->E);
  AVMOE_CHECK_ARG(a->z && a->wu_pack && a->ybi && a->probs, "%s: null tensor", who);
  if (bwd) AVMOE_CHECK_ARG(a->dout && a->yd && a->act && a->dprobs, "%s: null backward tensor", who);
  else AVMOE_CHECK_ARG(a->out != nullptr, "%s: null out", who);
  return 0;
}

template <int E, int CRK, int NPR>
static int launch_back(const avmoe_rf_back_args* a, bool bwd, cudaStream_t st) {
  const long long rows = (long long)a->BT * a->N;
  const long long ntiles = (rows + 15) / 16;
  const int pitch = a->C * 2 + 16;
  const size_t tab_bytes = (size_t)E * (a->C / 2) * 12 * sizeof(float);
  const size_t per_warp = (size_t)16 * pitch + (bwd ? (size_t)E * 4 * a->C * sizeof(float) + 2 * PairStager::kBytes : 0);
  const size_t fixed = tab_bytes + (bwd ? 32 * sizeof(float) : 0);
  int warps = 8;
  while (warps > 1 && fixed + warps * per_warp > 100 * 1024) warps >>= 1;    // two CTAs per SM where possible
  if (warps == 1 && fixed + 2 * per_warp <= 227 * 1024) warps = 2;
  const size_t smem = fixed + warps * per_warp;
  AVMOE_CHECK_ARG(smem <= 227 * 1024, "rf_back: tile does not fit shared memory");
  // persistent-style grid: a few CTAs per SM, each owning a contiguous range of tiles (one flush of the column sums per CTA)
  long long ctas = (long long)num_sms() * (bwd ? 3 : 6);
  long long tpc = (ntiles + ctas - 1) / ctas;
  tpc = (tpc + warps - 1) / warps * warps;
  if (tpc < warps) tpc = warps;
  const unsigned grid = (unsigned)((ntiles + tpc - 1) / tpc);
  if (bwd) {
    AVMOE_CUDA(cudaFuncSetAttribute(rf_back_kernel<E, CRK, NPR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    rf_back_kernel<E, CRK, NPR, true><<<grid, warps * 32, smem, st>>>(*a, rows, (int)tpc);
  } else {
    AVMOE_CUDA(cudaFuncSetAttribute(rf_back_kernel<E, CRK, NPR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    rf_back_kernel<E, CRK, NPR, false><<<grid, warps * 32, smem, st>>>(*a, rows, (int)tpc);
  }
  AVMOE_LAUNCH_CHECK("rf_back_kernel");
  return 0;
}

template <int E>
static int dispatch_back_crk(const avmoe_rf_back_args* a, bool bwd, cudaStream_t st) {
  const int crk = (a->Cr + 15) / 16;
  // C = 96 (stage-1 audio): the whole up-projection tile stays in registers (backward 372 vs 538 us, profiles/r02_rowfused.md)
  if (a->C == 96 && crk == 1) return launch_back<E, 1, 6>(a, bwd, st);
  switch (crk) {
    case 1: return launch_back<E, 1, 0>(a, bwd, st);
    case 2: return launch_back<E, 2, 0>(a, bwd, st);
    case 3: return launch_back<E, 3, 0>(a, bwd, st);
    case 6: return launch_back<E, 6, 0>(a, bwd, st);
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

// ------------------------------------------------------------------------------------------------ column sums of a packed operand
__global__ void rf_colsum_kernel(const uint4* __restrict__ packed, int nb, int NP, int KS, float* __restrict__ csum) {
  const int total = nb * NP * 16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / (NP * 16), n = i - b * NP * 16;
    const int p = n >> 4, hi = (n >> 3) & 1, g = n & 7;
    float s = 0.f;
    for (int ks = 0; ks < KS; ++ks)
      for (int t = 0; t < 4; ++t) {
        const uint4 w = packed[(((size_t)b * NP + p) * KS + ks) * 32 + g * 4 + t];
        const float2 a = unpack_bf16x2(hi ? w.z : w.x), c = unpack_bf16x2(hi ? w.w : w.y);
        s += (a.x + a.y) + (c.x + c.y);
      }
    csum[i] = s;
  }
}

// ------------------------------------------------------------------------------------------------ front pass
// Shared by forward and backward: the expert's ln_before input u on this warp's 16-row tile.
//   NONE        u = x                                   (t_u is not written; the x tile is used directly)
//   AUX_ADD     u = x + gate_pre * aux                  AUX_REPLACE  u = aux            (aux tile staged in t_u / t_d first)
//   LATENT      u = x + gate_pre * softmax_T(x tok'^T) tok'      (net_trans_v3.py:386-390)
// Row statistics (mean, rstd over C) of the fp32 u are returned for rows g and g+8.
template <int NT>
struct LatentRegs {
  float S[NT][4];             // logits -> probabilities A2 (fp32), accumulator layout over the T columns
  uint32_t P[NT / 2][4];      // A2 rounded to bf16 as A fragments (k = latent token)
};

template <int NT>
__device__ __forceinline__ void latent_probs(LatentRegs<NT>& L, const unsigned char* t_x, int pitch, int C, int T, const uint4* ts2, int lane) {
  const int t = lane & 3;
  const int KSc = C >> 4;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) L.S[nt][i] = 0.f;
  for (int s = 0; s < KSc; ++s) {
    uint32_t a[4];
    ldmatrix_a(a, t_x, pitch, s << 4, lane);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const uint4 b = ldg_b(ts2, nt, KSc, s, lane);     // (hi, lo) halves of tok' for latent tokens 8nt..8nt+7
      mma16816(L.S[nt], a, b.x, b.y);
      mma16816(L.S[nt], a, b.z, b.w);
    }
  }
  float mA = -3.0e38f, mB = -3.0e38f;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const bool ok = 8 * nt + 2 * t + i < T;
      if (!ok) { L.S[nt][i] = -3.0e38f; L.S[nt][2 + i] = -3.0e38f; }
      mA = fmaxf(mA, L.S[nt][i]); mB = fmaxf(mB, L.S[nt][2 + i]);
    }
  mA = quad_max(mA); mB = quad_max(mB);
  float sA = 0.f, sB = 0.f;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const bool ok = 8 * nt + 2 * t + i < T;
      L.S[nt][i] = ok ? __expf(L.S[nt][i] - mA) : 0.f;
      L.S[nt][2 + i] = ok ? __expf(L.S[nt][2 + i] - mB) : 0.f;
      sA += L.S[nt][i]; sB += L.S[nt][2 + i];
    }
  sA = 1.f / quad_sum(sA); sB = 1.f / quad_sum(sB);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    L.S[nt][0] *= sA; L.S[nt][1] *= sA; L.S[nt][2] *= sB; L.S[nt][3] *= sB;
  }
#pragma unroll
  for (int ks = 0; ks < NT / 2; ++ks) {
    L.P[ks][0] = pack_bf16x2(L.S[2 * ks][0], L.S[2 * ks][1]);
    L.P[ks][1] = pack_bf16x2(L.S[2 * ks][2], L.S[2 * ks][3]);
    L.P[ks][2] = pack_bf16x2(L.S[2 * ks + 1][0], L.S[2 * ks + 1][1]);
    L.P[ks][3] = pack_bf16x2(L.S[2 * ks + 1][2], L.S[2 * ks + 1][3]);
  }
}

// x_res for one n-tile pair of channels: acc[h][0..3] = (A2 tok')[rows g / g+8][16pr + 8h + 2t ..]
template <int NT>
__device__ __forceinline__ void latent_xr_pair(const LatentRegs<NT>& L, const uint4* txr, int pr, int lane, float (&acc)[2][4]) {
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[h][i] = 0.f;
#pragma unroll
  for (int ks = 0; ks < NT / 2; ++ks) {
    const uint4 b = ldg_b(txr, pr, NT / 2, ks, lane);
    mma16816(acc[0], L.P[ks], b.x, b.y);
    mma16816(acc[1], L.P[ks], b.z, b.w);
  }
}

struct RowStats { float muA, muB, rsA, rsB; };

__device__ __forceinline__ RowStats finish_stats(float s1A, float s2A, float s1B, float s2B, float invC, bool has_ln) {
  RowStats r;
  if (!has_ln) { r.muA = r.muB = 0.f; r.rsA = r.rsB = 1.f; return r; }
  s1A = quad_sum(s1A); s2A = quad_sum(s2A); s1B = quad_sum(s1B); s2B = quad_sum(s2B);
  r.muA = s1A * invC; r.muB = s1B * invC;
  r.rsA = rsqrtf(fmaxf(s2A * invC - r.muA * r.muA, 0.f) + kLnEps);
  r.rsB = rsqrtf(fmaxf(s2B * invC - r.muB * r.muB, 0.f) + kLnEps);
  return r;
}

// u (the ln_before input of one expert) for the four elements a lane owns in one n-tile: rows g / g+8, columns c, c+1.
//   NONE u = x    AUX_ADD u = x + gate*aux    AUX_REPLACE u = aux    LATENT u = x + gate*x_res  (xr = the lane's x_res values)
__device__ __forceinline__ void make_u(int kind, float gate, const TileRows& Rx, const TileRows& Raux, int c, const float (&xr)[4], float (&u)[4]) {
  const float2 xA = unpack_bf16x2(Rx.A[c >> 1]), xB = unpack_bf16x2(Rx.B[c >> 1]);
  u[0] = xA.x; u[1] = xA.y; u[2] = xB.x; u[3] = xB.y;
  if (kind == AVMOE_PRE_LATENT) {
    u[0] = fmaf(gate, xr[0], u[0]); u[1] = fmaf(gate, xr[1], u[1]); u[2] = fmaf(gate, xr[2], u[2]); u[3] = fmaf(gate, xr[3], u[3]);
  } else if (kind != AVMOE_PRE_NONE) {
    const float2 aA = unpack_bf16x2(Raux.A[c >> 1]), aB = unpack_bf16x2(Raux.B[c >> 1]);
    if (kind == AVMOE_PRE_AUX_ADD) { u[0] = fmaf(gate, aA.x, u[0]); u[1] = fmaf(gate, aA.y, u[1]); u[2] = fmaf(gate, aB.x, u[2]); u[3] = fmaf(gate, aB.y, u[3]); }
    else { u[0] = aA.x; u[1] = aA.y; u[2] = aB.x; u[3] = aB.y; }
  }
}

// Forward front pass.  Per expert ONE sweep over the n-tile pairs of the channel axis: x_res of the pair (latent experts: A2 tok' on
// the tensor cores), u, its LayerNorm sums, and -- u of a pair being exactly the A fragment of one k-step (accumulator layout ->
// operand layout, the usual attention trick) -- the down projection accumulates in the same sweep.  LayerNorm is folded into the
// operand (z = rstd*(u W'^T - mean*colsum(W')) + W b), so the statistics are only needed in the epilogue; u never touches
// shared memory.  The per-row (mean, rstd) are saved for the backward.
template <int CRK, int NT>
__global__ void __launch_bounds__(256) rf_front_fwd_kernel(const avmoe_rf_front_args p, long long rows, int tiles_per_cta, int ntile_bufs) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int C = p.C, Cr = p.Cr, E = p.E;
  const int pitch = C * 2 + 16;
  const int KSc = C >> 4, NP = C >> 4;
  const float invC = 1.f / C;
  const bool has_ln = p.has_ln != 0;
  unsigned char* t_x = smem + (size_t)warp * ntile_bufs * 16 * pitch;
  unsigned char* t_aux = t_x + (ntile_bufs > 1 ? 16 * pitch : 0);
  const TileRows Rx(t_x, pitch, g), Raux(t_aux, pitch, g);
  const long long ntiles = rows >> 4;
  const long long tile0 = (long long)blockIdx.x * tiles_per_cta;
  const long long tile1 = min(ntiles, tile0 + tiles_per_cta);
  const uint4* wd = static_cast<const uint4*>(p.wd_pack);
  const int wd_stride = CRK * KSc * 32;
  const size_t s2_stride = (size_t)NT * KSc * 32, xr_stride = (size_t)KSc * (NT / 2) * 32;
  for (long long tile = tile0 + warp; tile < tile1; tile += nwarps) {
    const long long r0 = tile << 4;
    const long long b = r0 / p.N;
    const long long rowA = r0 + g, rowB = r0 + g + 8;
    tile_load_async(t_x, pitch, static_cast<const __nv_bfloat16*>(p.x), r0, rows, C, lane);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    for (int e = 0; e < E; ++e) {
      const avmoe_expert& ex = p.ex[e];
      const int kind = ex.pre_kind;
      if (kind == AVMOE_PRE_AUX_ADD || kind == AVMOE_PRE_AUX_REPLACE) {
        __syncwarp();
        tile_load_async(t_aux, pitch, static_cast<const __nv_bfloat16*>(ex.aux), r0, rows, C, lane);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
      }
      const float gate = (kind == AVMOE_PRE_LATENT || kind == AVMOE_PRE_AUX_ADD) ? __ldg(ex.gate_pre) : 0.f;
      LatentRegs<NT> L;
      const uint4* ts2 = nullptr;
      const uint4* txr = nullptr;
      if (kind == AVMOE_PRE_LATENT) {
        ts2 = static_cast<const uint4*>(p.tok_s2[e]) + b * s2_stride;
        txr = static_cast<const uint4*>(p.tok_xr[e]) + b * xr_stride;
        latent_probs<NT>(L, t_x, pitch, C, p.T, ts2, lane);
      }
      float zacc[2 * CRK][4];
#pragma unroll
      for (int j = 0; j < 2 * CRK; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) zacc[j][i] = 0.f;
      const uint4* w = wd + e * wd_stride;
      float s1A = 0.f, s2A = 0.f, s1B = 0.f, s2B = 0.f;
      for (int pr = 0; pr < NP; ++pr) {
        float xr[2][4];
        if (kind == AVMOE_PRE_LATENT) latent_xr_pair<NT>(L, txr, pr, lane, xr);
        uint32_t a[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float u[4];
          make_u(kind, gate, Rx, Raux, (pr << 4) + h * 8 + 2 * t, xr[h], u);
          s1A += u[0] + u[1]; s2A = fmaf(u[0], u[0], fmaf(u[1], u[1], s2A));
          s1B += u[2] + u[3]; s2B = fmaf(u[2], u[2], fmaf(u[3], u[3], s2B));
          a[2 * h] = pack_bf16x2(u[0], u[1]);          // (row g,   k = 8h + 2t ..)
          a[2 * h + 1] = pack_bf16x2(u[2], u[3]);      // (row g+8, k = 8h + 2t ..)
        }
#pragma unroll
        for (int prz = 0; prz < CRK; ++prz) {
          const uint4 bb = ldg_b(w, prz, KSc, pr, lane);
          mma16816(zacc[2 * prz], a, bb.x, bb.y);
          mma16816(zacc[2 * prz + 1], a, bb.z, bb.w);
        }
      }
      const RowStats st = finish_stats(s1A, s2A, s1B, s2B, invC, has_ln);
      if (has_ln && p.ln1_stats && t == 0) {
        *reinterpret_cast<float2*>(p.ln1_stats + ((size_t)e * rows + rowA) * 2) = make_float2(st.muA, st.rsA);
        *reinterpret_cast<float2*>(p.ln1_stats + ((size_t)e * rows + rowB) * 2) = make_float2(st.muB, st.rsB);
      }
#pragma unroll
      for (int j = 0; j < 2 * CRK; ++j) {
        const int ch = 8 * j + 2 * t;
        if (ch < Cr) {
          const float2 cs = __ldg(reinterpret_cast<const float2*>(p.wd_csum + (size_t)e * 16 * CRK + ch));
          const float2 c2 = __ldg(reinterpret_cast<const float2*>(p.c2 + (size_t)e * Cr + ch));
          float2 zA, zB;
          zA.x = st.rsA * (zacc[j][0] - st.muA * cs.x) + c2.x; zA.y = st.rsA * (zacc[j][1] - st.muA * cs.y) + c2.y;
          zB.x = st.rsB * (zacc[j][2] - st.muB * cs.x) + c2.x; zB.y = st.rsB * (zacc[j][3] - st.muB * cs.y) + c2.y;
          *reinterpret_cast<float2*>(p.z + (rowA * E + e) * Cr + ch) = zA;
          *reinterpret_cast<float2*>(p.z + (rowB * E + e) * Cr + ch) = zB;
        }
      }
    }
    __syncwarp();
  }
}

// Backward of the front pass.  Per expert two sweeps over the n-tile pairs, u recomputed in registers each time (x tile + x_res
// on the tensor cores), d xl = dz Wd as K = bottleneck MMAs:
//   sweep A  ln_before backward row sums; d ln_before weight / bias column sums (8-lane reduce-scatter -> per-warp accumulators)
//   sweep B  d u; xl re-emitted for the weight-gradient GEMM (straight from registers); pre-transform backward: d gate_pre,
//            d aux / dXR (registers -> global), and for latent experts dXR of a pair IS the A fragment of one k-step of
//            dA2 = dXR tok'^T, which accumulates in the same sweep
//   tail     latent experts: dS2 = A2 (dA2 - <A2, dA2>), A2 | dS2 out, dx += dS2 tok'
// dx of all experts, the attention path and the router path accumulate in the warp's dx tile and are stored once.
template <int CRK, int NT>
__global__ void __launch_bounds__(256) rf_front_bwd_kernel(const avmoe_rf_front_args p, long long rows, int tiles_per_cta, int ntile_bufs) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int C = p.C, Cr = p.Cr, E = p.E, T = p.T;
  const int pitch = C * 2 + 16;
  const int KSc = C >> 4, NP = C >> 4;
  const float invC = 1.f / C;
  const bool has_ln = p.has_ln != 0;
  const GroupGeom gg(C, Cr, p.G);
  // shared memory: per-warp column accumulators [E][2][C] (d ln1_w, d ln1_b), E gate_pre sums; per warp 2 or 3 tiles (x, dx, aux)
  float* colacc_all = reinterpret_cast<float*>(smem);
  const size_t acc_floats = (size_t)nwarps * E * 2 * C + 32;
  float* gateacc = colacc_all + (size_t)nwarps * E * 2 * C;
  float* wacc = colacc_all + (size_t)warp * E * 2 * C;
  for (int i = threadIdx.x; i < (int)acc_floats; i += blockDim.x) colacc_all[i] = 0.f;
  __syncthreads();
  const size_t warp_bytes = (size_t)ntile_bufs * 16 * pitch + 2 * PairStager::kBytes;
  unsigned char* t_x = smem + acc_floats * sizeof(float) + (size_t)warp * warp_bytes;
  unsigned char* t_dx = t_x + 16 * pitch;
  unsigned char* t_aux = t_x + (ntile_bufs > 2 ? 32 * pitch : 0);
  const TileRows Rx(t_x, pitch, g), Rdx(t_dx, pitch, g), Raux(t_aux, pitch, g);
  PairStager st_xl(t_x + (size_t)ntile_bufs * 16 * pitch, lane), st_ga(t_x + (size_t)ntile_bufs * 16 * pitch + PairStager::kBytes, lane);
  const long long ntiles = rows >> 4;
  const long long tile0 = (long long)blockIdx.x * tiles_per_cta;
  const long long tile1 = min(ntiles, tile0 + tiles_per_cta);
  const uint4* wdt = static_cast<const uint4*>(p.wdt_pack);
  const int wdt_stride = NP * CRK * 32;
  const size_t s2_stride = (size_t)NT * KSc * 32, xr_stride = (size_t)KSc * (NT / 2) * 32;
  const int Tp8 = (T + 7) & ~7;
  const int ri = rs_index(lane);
  for (long long tile = tile0 + warp; tile < tile1; tile += nwarps) {
    const long long r0 = tile << 4;
    const long long b = r0 / p.N;
    const long long rowA = r0 + g, rowB = r0 + g + 8;
    tile_load_async(t_x, pitch, static_cast<const __nv_bfloat16*>(p.x), r0, rows, C, lane);
    if (p.dres) tile_load_async(t_dx, pitch, static_cast<const __nv_bfloat16*>(p.dres), r0, rows, C, lane);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    // dx tile starts from the router-path gradient (d mean_N(x)), identical for every row of the sample, plus -- when the folded
    // residual stream is x itself -- the gradient that passes straight through the residual add
    for (int pr = 0; pr < NP; ++pr)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = (pr << 4) + h * 8 + 2 * t;
        float2 d = make_float2(0.f, 0.f);
        if (p.dpool_x) d = __ldg(reinterpret_cast<const float2*>(p.dpool_x + b * C + c));
        if (p.dres) {
          const float2 rA = unpack_bf16x2(Rdx.A[c >> 1]), rB = unpack_bf16x2(Rdx.B[c >> 1]);
          Rdx.A[c >> 1] = pack_bf16x2(rA.x + d.x, rA.y + d.y);
          Rdx.B[c >> 1] = pack_bf16x2(rB.x + d.x, rB.y + d.y);
        } else {
          const uint32_t v = pack_bf16x2(d.x, d.y);
          Rdx.A[c >> 1] = v;
          Rdx.B[c >> 1] = v;
        }
      }
    __syncwarp();
    for (int e = 0; e < E; ++e) {
      const avmoe_expert& ex = p.ex[e];
      const int kind = ex.pre_kind;
      const bool aux_kind = kind == AVMOE_PRE_AUX_ADD || kind == AVMOE_PRE_AUX_REPLACE;
      if (aux_kind) {
        __syncwarp();
        tile_load_async(t_aux, pitch, static_cast<const __nv_bfloat16*>(ex.aux), r0, rows, C, lane);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
      }
      const float gate = (kind == AVMOE_PRE_LATENT || kind == AVMOE_PRE_AUX_ADD) ? __ldg(ex.gate_pre) : 0.f;
      LatentRegs<NT> L;
      const uint4* ts2 = nullptr;
      const uint4* txr = nullptr;
      if (kind == AVMOE_PRE_LATENT) {
        ts2 = static_cast<const uint4*>(p.tok_s2[e]) + b * s2_stride;
        txr = static_cast<const uint4*>(p.tok_xr[e]) + b * xr_stride;
        latent_probs<NT>(L, t_x, pitch, C, T, ts2, lane);
      }
      RowStats st;
      st.muA = st.muB = 0.f; st.rsA = st.rsB = 1.f;
      if (has_ln) {     // saved by the forward
        const float2 sA = __ldg(reinterpret_cast<const float2*>(p.ln1_stats + ((size_t)e * rows + rowA) * 2));
        const float2 sB = __ldg(reinterpret_cast<const float2*>(p.ln1_stats + ((size_t)e * rows + rowB) * 2));
        st.muA = sA.x; st.rsA = sA.y; st.muB = sB.x; st.rsB = sB.y;
      }
      const float nmA = -st.muA * st.rsA, nmB = -st.muB * st.rsB;
      // dz as A fragments (already bf16 in memory)
      uint32_t dzf[CRK][4];
      {
        const __nv_bfloat16* dz = static_cast<const __nv_bfloat16*>(p.dz) + (size_t)e * rows * p.Crp;
#pragma unroll
        for (int s = 0; s < CRK; ++s)
#pragma unroll
          for (int kh = 0; kh < 2; ++kh) {
            const int k0 = 16 * s + 8 * kh + 2 * t;
            const bool ok = k0 < p.Crp;
            dzf[s][2 * kh] = ok ? __ldg(reinterpret_cast<const uint32_t*>(dz + rowA * p.Crp + k0)) : 0u;
            dzf[s][2 * kh + 1] = ok ? __ldg(reinterpret_cast<const uint32_t*>(dz + rowB * p.Crp + k0)) : 0u;
          }
      }
      const uint4* w = wdt + e * wdt_stride;
      auto dxl_pair = [&](int pr, float (&acc)[2][4]) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[h][i] = 0.f;
        int lo = 0, hi = CRK;
        if (CRK > 1) up_ksteps(pr << 4, gg, lo, hi);
#pragma unroll
        for (int s = 0; s < CRK; ++s)
          if (s >= lo && s < hi) {
            const uint4 bb = ldg_b(w, pr, CRK, s, lane);
            mma16816(acc[0], dzf[s], bb.x, bb.y);
            mma16816(acc[1], dzf[s], bb.z, bb.w);
          }
      };
      float* accp = wacc + (size_t)e * 2 * C + ((ri >> 1) & 1) * C + (ri >> 2) * 8 + 2 * t + (ri & 1);     // + 16*pr
      // ---- sweep A: ln_before backward row sums, d ln_before weight/bias column sums
      float maA = 0.f, maB = 0.f, mbA = 0.f, mbB = 0.f;
      if (has_ln) {
        for (int pr = 0; pr < NP; ++pr) {
          float acc[2][4], xr[2][4];
          dxl_pair(pr, acc);
          if (kind == AVMOE_PRE_LATENT) latent_xr_pair<NT>(L, txr, pr, lane, xr);
          float cs[8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = (pr << 4) + h * 8 + 2 * t;
            float u[4];
            make_u(kind, gate, Rx, Raux, c, xr[h], u);
            const float2 lw = __ldg(reinterpret_cast<const float2*>(ex.ln1_w + c));
            const float xh0 = fmaf(u[0], st.rsA, nmA), xh1 = fmaf(u[1], st.rsA, nmA), xh2 = fmaf(u[2], st.rsB, nmB), xh3 = fmaf(u[3], st.rsB, nmB);
            const float d0 = acc[h][0] * lw.x, d1 = acc[h][1] * lw.y, d2 = acc[h][2] * lw.x, d3 = acc[h][3] * lw.y;
            maA += d0 + d1; maB += d2 + d3;
            mbA = fmaf(d0, xh0, fmaf(d1, xh1, mbA)); mbB = fmaf(d2, xh2, fmaf(d3, xh3, mbB));
            cs[4 * h + 0] = fmaf(acc[h][0], xh0, acc[h][2] * xh2); cs[4 * h + 1] = fmaf(acc[h][1], xh1, acc[h][3] * xh3);
            cs[4 * h + 2] = acc[h][0] + acc[h][2]; cs[4 * h + 3] = acc[h][1] + acc[h][3];
          }
          accp[pr << 4] += reduce_scatter8(cs, lane);
        }
        maA = quad_sum(maA) * invC; maB = quad_sum(maB) * invC; mbA = quad_sum(mbA) * invC; mbB = quad_sum(mbB) * invC;
      }
      // ---- sweep B: d u, xl out, pre-transform backward (latent: dA2 accumulates on the tensor cores)
      float gsum = 0.f;
      float dA[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) dA[nt][i] = 0.f;
      __nv_bfloat16* xl = static_cast<__nv_bfloat16*>(p.xl) + (size_t)e * rows * C;
      __nv_bfloat16* gaux = static_cast<__nv_bfloat16*>(ex.g_aux);
      for (int pr = 0; pr < NP; ++pr) {
        float acc[2][4], xr[2][4];
        dxl_pair(pr, acc);
        if (kind == AVMOE_PRE_LATENT) latent_xr_pair<NT>(L, txr, pr, lane, xr);
        uint32_t ga[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = (pr << 4) + h * 8 + 2 * t;
          float u[4];
          make_u(kind, gate, Rx, Raux, c, xr[h], u);
          float du[4] = {acc[h][0], acc[h][1], acc[h][2], acc[h][3]};
          float x0 = u[0], x1 = u[1], x2 = u[2], x3 = u[3];          // xl without ln_before
          if (has_ln) {
            const float2 lw = __ldg(reinterpret_cast<const float2*>(ex.ln1_w + c)), lb = __ldg(reinterpret_cast<const float2*>(ex.ln1_b + c));
            const float xh0 = fmaf(u[0], st.rsA, nmA), xh1 = fmaf(u[1], st.rsA, nmA), xh2 = fmaf(u[2], st.rsB, nmB), xh3 = fmaf(u[3], st.rsB, nmB);
            du[0] = st.rsA * fmaf(-mbA, xh0, fmaf(du[0], lw.x, -maA)); du[1] = st.rsA * fmaf(-mbA, xh1, fmaf(du[1], lw.y, -maA));
            du[2] = st.rsB * fmaf(-mbB, xh2, fmaf(du[2], lw.x, -maB)); du[3] = st.rsB * fmaf(-mbB, xh3, fmaf(du[3], lw.y, -maB));
            x0 = fmaf(xh0, lw.x, lb.x); x1 = fmaf(xh1, lw.y, lb.y); x2 = fmaf(xh2, lw.x, lb.x); x3 = fmaf(xh3, lw.y, lb.y);
          }
          st_xl.put(h, pack_bf16x2(x0, x1), pack_bf16x2(x2, x3));                        // xl = ln_before(u), operand of dWd
          if (kind == AVMOE_PRE_LATENT) {
            gsum = fmaf(du[0], xr[h][0], fmaf(du[1], xr[h][1], fmaf(du[2], xr[h][2], fmaf(du[3], xr[h][3], gsum))));
            ga[2 * h] = pack_bf16x2(gate * du[0], gate * du[1]);                       // dXR, also the A fragment of dA2's k-step
            ga[2 * h + 1] = pack_bf16x2(gate * du[2], gate * du[3]);
          } else if (aux_kind) {
            const float2 aA = unpack_bf16x2(Raux.A[c >> 1]), aB = unpack_bf16x2(Raux.B[c >> 1]);
            const float sc = kind == AVMOE_PRE_AUX_ADD ? gate : 1.f;
            gsum = fmaf(du[0], aA.x, fmaf(du[1], aA.y, fmaf(du[2], aB.x, fmaf(du[3], aB.y, gsum))));
            ga[2 * h] = pack_bf16x2(sc * du[0], sc * du[1]);                           // d aux
            ga[2 * h + 1] = pack_bf16x2(sc * du[2], sc * du[3]);
          }
          if (kind != AVMOE_PRE_NONE) st_ga.put(h, ga[2 * h], ga[2 * h + 1]);
          if (kind != AVMOE_PRE_AUX_REPLACE) {
            const float2 pA = unpack_bf16x2(Rdx.A[c >> 1]), pB = unpack_bf16x2(Rdx.B[c >> 1]);
            Rdx.A[c >> 1] = pack_bf16x2(pA.x + du[0], pA.y + du[1]);
            Rdx.B[c >> 1] = pack_bf16x2(pB.x + du[2], pB.y + du[3]);
          }
        }
        st_xl.flush(xl + r0 * C + (pr << 4), C, true);
        if (kind != AVMOE_PRE_NONE) st_ga.flush(gaux + r0 * C + (pr << 4), C, true);
        if (kind == AVMOE_PRE_LATENT) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const uint4 bb = ldg_b(ts2, nt, KSc, pr, lane);
            mma16816(dA[nt], ga, bb.x, bb.y);
            mma16816(dA[nt], ga, bb.z, bb.w);
          }
        }
      }
      if (kind != AVMOE_PRE_NONE && kind != AVMOE_PRE_AUX_REPLACE && ex.g_gate_pre) {
        gsum = quad_sum(gsum);
        gsum += __shfl_xor_sync(0xffffffffu, gsum, 4); gsum += __shfl_xor_sync(0xffffffffu, gsum, 8); gsum += __shfl_xor_sync(0xffffffffu, gsum, 16);
        if (lane == 0) atomicAdd(gateacc + e, gsum);
      }
      if (kind == AVMOE_PRE_LATENT) {
        // ---- attention backward on the tile: dS2 = A2 (dA2 - <A2, dA2>), dx += dS2 tok'
        float dotA = 0.f, dotB = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          dotA = fmaf(L.S[nt][0], dA[nt][0], fmaf(L.S[nt][1], dA[nt][1], dotA));
          dotB = fmaf(L.S[nt][2], dA[nt][2], fmaf(L.S[nt][3], dA[nt][3], dotB));
        }
        dotA = quad_sum(dotA); dotB = quad_sum(dotB);
        __nv_bfloat16* a2 = static_cast<__nv_bfloat16*>(p.a2[e]);
        __nv_bfloat16* ds2 = static_cast<__nv_bfloat16*>(p.ds2[e]);
        uint32_t dSf[NT / 2][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const float s0 = L.S[nt][0] * (dA[nt][0] - dotA), s1 = L.S[nt][1] * (dA[nt][1] - dotA);
          const float s2 = L.S[nt][2] * (dA[nt][2] - dotB), s3 = L.S[nt][3] * (dA[nt][3] - dotB);
          const uint32_t pA = pack_bf16x2(s0, s1), pB = pack_bf16x2(s2, s3);
          dSf[nt >> 1][(nt & 1) * 2] = pA;
          dSf[nt >> 1][(nt & 1) * 2 + 1] = pB;
          const int col = 8 * nt + 2 * t;
          if (col < Tp8) {
            *reinterpret_cast<uint32_t*>(a2 + rowA * Tp8 + col) = pack_bf16x2(L.S[nt][0], L.S[nt][1]);
            *reinterpret_cast<uint32_t*>(a2 + rowB * Tp8 + col) = pack_bf16x2(L.S[nt][2], L.S[nt][3]);
            *reinterpret_cast<uint32_t*>(ds2 + rowA * Tp8 + col) = pA;
            *reinterpret_cast<uint32_t*>(ds2 + rowB * Tp8 + col) = pB;
          }
        }
        for (int pr = 0; pr < NP; ++pr) {
          float acc[2][4];
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[h][i] = 0.f;
#pragma unroll
          for (int ks = 0; ks < NT / 2; ++ks) {
            const uint4 bb = ldg_b(txr, pr, NT / 2, ks, lane);
            mma16816(acc[0], dSf[ks], bb.x, bb.y);
            mma16816(acc[1], dSf[ks], bb.z, bb.w);
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = (pr << 4) + h * 8 + 2 * t;
            const float2 pA = unpack_bf16x2(Rdx.A[c >> 1]), pB = unpack_bf16x2(Rdx.B[c >> 1]);
            Rdx.A[c >> 1] = pack_bf16x2(pA.x + acc[h][0], pA.y + acc[h][1]);
            Rdx.B[c >> 1] = pack_bf16x2(pB.x + acc[h][2], pB.y + acc[h][3]);
          }
        }
      }
    }
    __syncwarp();
    tile_store<false>(t_dx, pitch, static_cast<__nv_bfloat16*>(p.dx), C, nullptr, r0, rows, C, lane);
    __syncwarp();
  }
  __syncthreads();
  if (has_ln) {
    for (int i = threadIdx.x; i < E * 2 * C; i += blockDim.x) {
      float sum = 0.f;
      for (int wq = 0; wq < nwarps; ++wq) sum += colacc_all[(size_t)wq * E * 2 * C + i];
      const int e = i / (2 * C), q = (i - e * 2 * C) / C, c = i - e * 2 * C - q * C;
      atomicAdd((q == 0 ? p.ex[e].g_ln1_w : p.ex[e].g_ln1_b) + c, sum);
    }
  }
  if (threadIdx.x < E && p.ex[threadIdx.x].g_gate_pre) atomicAdd(p.ex[threadIdx.x].g_gate_pre, gateacc[threadIdx.x]);
}

static int front_check(const avmoe_rf_front_args* a, bool bwd, const char* who) {
  AVMOE_CHECK_ARG(a != nullptr, "%s: null args", who);
  AVMOE_CHECK_ARG(rf_capable(a->C, a->Cr, a->E, a->T) && a->N % 16 == 0, "%s: geometry not supported (C=%d Cr=%d E=%d T=%d N=%d)", who,
                  a->C, a->Cr, a->E, a->T, a->N);
  AVMOE_CHECK_ARG(a->x && a->c2, "%s: null tensor", who);
  for (int e = 0; e < a->E; ++e) {
    const avmoe_expert& ex = a->ex[e];
    if (ex.pre_kind == AVMOE_PRE_LATENT) AVMOE_CHECK_ARG(a->tok_s2[e] && a->tok_xr[e] && ex.gate_pre, "%s: latent expert %d lacks operands", who, e);
    if (ex.pre_kind == AVMOE_PRE_AUX_ADD) AVMOE_CHECK_ARG(ex.aux && ex.gate_pre, "%s: aux expert %d lacks operands", who, e);
    if (ex.pre_kind == AVMOE_PRE_AUX_REPLACE) AVMOE_CHECK_ARG(ex.aux != nullptr, "%s: aux expert %d lacks operands", who, e);
    if (a->has_ln) AVMOE_CHECK_ARG(ex.ln1_w && ex.ln1_b, "%s: ln_before parameters missing", who);
    if (bwd) {
      if (ex.pre_kind != AVMOE_PRE_NONE) AVMOE_CHECK_ARG(ex.g_aux != nullptr, "%s: expert %d needs g_aux", who, e);
      if (ex.pre_kind == AVMOE_PRE_LATENT) AVMOE_CHECK_ARG(a->a2[e] && a->ds2[e], "%s: latent expert %d needs a2/ds2", who, e);
      if (a->has_ln) AVMOE_CHECK_ARG(ex.g_ln1_w && ex.g_ln1_b, "%s: ln_before gradient accumulators missing", who);
    }
  }
  if (bwd) AVMOE_CHECK_ARG(a->dz && a->dx && a->xl && a->wdt_pack, "%s: null backward tensor", who);
  else AVMOE_CHECK_ARG(a->z && a->wd_pack && a->wd_csum, "%s: null forward tensor", who);
  if (a->has_ln && bwd) AVMOE_CHECK_ARG(a->ln1_stats != nullptr, "%s: ln1_stats (saved by the forward) missing", who);
  return 0;
}

template <int CRK, int NT>
static int launch_front(const avmoe_rf_front_args* a, bool bwd, cudaStream_t st) {
  const long long rows = (long long)a->BT * a->N;
  const long long ntiles = rows / 16;
  const int pitch = a->C * 2 + 16;
  bool has_aux = false;
  for (int e = 0; e < a->E; ++e) has_aux |= a->ex[e].pre_kind == AVMOE_PRE_AUX_ADD || a->ex[e].pre_kind == AVMOE_PRE_AUX_REPLACE;
  const int bufs = (bwd ? 2 : 1) + (has_aux ? 1 : 0);           // x (+ dx in the backward) (+ aux)
  const size_t per_warp = (size_t)bufs * 16 * pitch + (bwd ? (size_t)a->E * 2 * a->C * sizeof(float) + 2 * PairStager::kBytes : 0);
  const size_t fixed = bwd ? 32 * sizeof(float) : 0;
  int warps = 8;
  while (warps > 1 && fixed + warps * per_warp > 100 * 1024) warps >>= 1;
  if (fixed + warps * per_warp > 227 * 1024) { set_error("rf_front: tile does not fit shared memory (C=%d)", a->C); return AVMOE_E_BADARG; }
  const size_t smem = fixed + warps * per_warp;
  // a few CTAs per SM, each owning a contiguous range of tiles; the backward flushes its column sums once per CTA
  long long ctas = (long long)num_sms() * (bwd ? 3 : 6);
  long long tpc = (ntiles + ctas - 1) / ctas;
  tpc = (tpc + warps - 1) / warps * warps;
  if (tpc < warps) tpc = warps;
  const unsigned grid = (unsigned)((ntiles + tpc - 1) / tpc);
  if (bwd) {
    AVMOE_CUDA(cudaFuncSetAttribute(rf_front_bwd_kernel<CRK, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    rf_front_bwd_kernel<CRK, NT><<<grid, warps * 32, smem, st>>>(*a, rows, (int)tpc, bufs);
  } else {
    AVMOE_CUDA(cudaFuncSetAttribute(rf_front_fwd_kernel<CRK, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    rf_front_fwd_kernel<CRK, NT><<<grid, warps * 32, smem, st>>>(*a, rows, (int)tpc, bufs);
  }
  AVMOE_LAUNCH_CHECK("rf_front_kernel");
  return 0;
}

template <int CRK>
static int dispatch_front_nt(const avmoe_rf_front_args* a, bool bwd, cudaStream_t st) {
  const int nt = ((a->T + 15) / 16) * 2;
  if (nt <= 2) return launch_front<CRK, 2>(a, bwd, st);
  if (nt <= 4) return launch_front<CRK, 4>(a, bwd, st);
  if (nt <= 12) return launch_front<CRK, 12>(a, bwd, st);
  set_error("rf_front: unsupported latent token count %d", a->T);
  return AVMOE_E_BADARG;
}

static int dispatch_front(const avmoe_rf_front_args* a, bool bwd, cudaStream_t st) {
  const int rc = front_check(a, bwd, bwd ? "rf_front_bwd" : "rf_front_fwd");
  if (rc) return rc;
  switch ((a->Cr + 15) / 16) {
    case 1: return dispatch_front_nt<1>(a, bwd, st);
    case 2: return dispatch_front_nt<2>(a, bwd, st);
    case 3: return dispatch_front_nt<3>(a, bwd, st);
    case 6: return dispatch_front_nt<6>(a, bwd, st);
    default: break;
  }
  set_error("rf_front: unsupported bottleneck width %d", a->Cr);
  return AVMOE_E_BADARG;
}

}  // namespace rf
}  // namespace avmoe

static bool rf_capable(int C, int Cr, int E, int T) {     // what the kernels can run (unit-tested up to C = 768)
  const int crk = (Cr + 15) / 16;
  return C % 16 == 0 && C <= avmoe::rf::kMaxC && Cr % 2 == 0 && (crk == 1 || crk == 2 || crk == 3 || crk == 6) && E >= 2 && E <= 4 && T <= 96;
}

extern "C" int avmoe_rf_supported(int C, int Cr, int E, int T) {
  const int crk = (Cr + 15) / 16;
  const bool crk_ok = crk == 1 || crk == 2 || crk == 3 || crk == 6;
  // C <= kFastC: a warp's 16-row tiles must be small enough that >= 8 warps (fwd: 16) stay resident per SM; the wide late stages
  // (C >= 384, few rows) are latency-bound with one or two warps per SM (profiles/r02_rowfused.md) and keep the GEMM-based path
  return (C % 16 == 0 && C <= avmoe::rf::kFastC && Cr % 2 == 0 && crk_ok && E >= 2 && E <= 4 && T <= 96) ? 1 : 0;
}

extern "C" int64_t avmoe_rf_pack_b_bytes(int nb, int Nn, int K, int hi_lo) {
  const int64_t NP = hi_lo ? (Nn + 7) / 8 : (Nn + 15) / 16, KS = (K + 15) / 16;
  return (int64_t)nb * NP * KS * 32 * 16;
}

extern "C" int avmoe_rf_pack_b(const float* src, int64_t bstride, int64_t stride_n, int64_t stride_k, int nb, int Nn, int K, int groups,
                               int group_local_k, const float* scale_n, int scale_n_stride, const float* scale_k, int scale_k_stride,
                               int hi_lo, int Nn_pad, void* dst, avmoe_stream_t stream) {
  using namespace avmoe;
  AVMOE_CHECK_ARG(src && dst && nb > 0 && Nn > 0 && K > 0 && groups >= 1, "rf_pack_b: bad arguments");
  AVMOE_CHECK_ARG(Nn % groups == 0 && K % groups == 0, "rf_pack_b: groups must divide both extents");
  const int Nl = Nn_pad > Nn ? Nn_pad : Nn;
  const int NP = hi_lo ? (Nl + 7) / 8 : (Nl + 15) / 16, KS = (K + 15) / 16;
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

extern "C" int avmoe_rf_colsum(const void* packed, int nb, int Nn, int K, float* csum, avmoe_stream_t stream) {
  using namespace avmoe;
  AVMOE_CHECK_ARG(packed && csum && nb > 0 && Nn > 0 && K > 0, "rf_colsum: bad arguments");
  const int NP = (Nn + 15) / 16, KS = (K + 15) / 16;
  const int total = nb * NP * 16;
  rf::rf_colsum_kernel<<<(total + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(packed), nb, NP, KS, csum);
  AVMOE_LAUNCH_CHECK("rf_colsum_kernel");
  return 0;
}
extern "C" int avmoe_rf_front_fwd(const avmoe_rf_front_args* a, avmoe_stream_t stream) {
  return avmoe::rf::dispatch_front(a, false, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int avmoe_rf_front_bwd(const avmoe_rf_front_args* a, avmoe_stream_t stream) {
  return avmoe::rf::dispatch_front(a, true, reinterpret_cast<cudaStream_t>(stream));
}
