// Row-fused passes of the latent-token update  tok' = tok + softmax_Nk(tok · kvᵀ) · kv   (AVMOE/AVE/nets/net_trans_v3.py:379-384).
//
// The GEMM form (functional.py:gemm_latent_fwd/bwd) writes the fp32 logits S1ᵀ [BT,Nk,2Tp] with one skinny GEMM and reads them back
// twice in softmax_cols (and the same again for dA1 in the backward): 3 x 168 MB of logits per stage-1 call around 126 MB of kv.
// Here a warp owns 16 kv rows, recomputes their logits against the packed latent tokens on the tensor cores (mma.sync bf16, hi + lo
// halves of the fp32 tokens: ~16 mantissa bits, what the un-scaled softmax needs) and never stores them:
//   rf_tok_stats   per (sample, row part): online max / sum of exp over the part's rows for every latent token -> partials
//   rf_tok_probs   merges the partials into log-sum-exp, recomputes the logits, writes A1ᵀ = exp(s - lse) as bf16 hi | lo
//                  into A12 [BT,Nk,2Tp] (the operand of the A1 · kv GEMM and of the backward)
//   rf_tok_ds      backward: dA1ᵀ = kv · dRᵀ recomputed per tile against the per-sample packed dR, dS1ᵀ = A1ᵀ ∘ (dA1ᵀ - D[b,t])
//                  written over the lo half of A12 (the K-concatenated dkv GEMM and the d tok GEMM read it there)
// The contractions over the rows (A1 · kv, dkv, d tok) stay on the tcgen05 GEMM.
#include <atomic>

#include "mma_tile.cuh"

namespace avmoe {
namespace rf {

static constexpr int kTokWarps = 4;

__device__ __forceinline__ void tok_tile_load(unsigned char* tile, int pitch, const __nv_bfloat16* src, int C, int lane) {
  const int chunks = C >> 3;
  const uint32_t mg = magic_of(chunks);
  for (int i = lane; i < 16 * chunks; i += 32) {
    const int r = (int)magic_div(i, mg), ch = i - r * chunks;
    cp_async16(tile + r * pitch + ch * 16, src + (long long)r * C + ch * 8);
  }
}

// S[nt][0..3] = (tile rows g / g+8) · (packed operand columns 8nt+2t, +1), hi + lo halves accumulated into the same tile
template <int NT>
__device__ __forceinline__ void tok_logits(float (&S)[NT][4], const unsigned char* tile, int pitch, int C, const uint4* bp, int lane) {
  const int KSc = C >> 4;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) S[nt][i] = 0.f;
  for (int s = 0; s < KSc; ++s) {
    uint32_t a[4];
    ldmatrix_a(a, tile, pitch, s << 4, lane);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const uint4 b = ldg_b(bp, nt, KSc, s, lane);
      mma16816(S[nt], a, b.x, b.y);
      mma16816(S[nt], a, b.z, b.w);
    }
  }
}

__device__ __forceinline__ void merge_ml(float& m, float& l, float m2, float l2) {
  const float M = fmaxf(m, m2);
  l = l * __expf(m - M) + l2 * __expf(m2 - M);
  m = M;
}

struct TokGeom {
  long long tile0, tile_end;
  const __nv_bfloat16* kv;
  const uint4* bp;
};
__device__ __forceinline__ TokGeom tok_geom(const avmoe_rf_tok_args& p, int tiles_per_part, int NT) {
  TokGeom g;
  const int b = blockIdx.y;
  const long long tiles = p.Nk >> 4;
  g.tile0 = (long long)blockIdx.x * tiles_per_part;
  g.tile_end = min(tiles, g.tile0 + tiles_per_part);
  g.kv = static_cast<const __nv_bfloat16*>(p.kv) + (long long)b * p.Nk * p.C;
  const long long per = (long long)NT * (p.C >> 4) * 32;             // uint4 words of one packed hi/lo operand
  g.bp = static_cast<const uint4*>(p.pack) + (p.per_sample ? b * per : 0);
  return g;
}

// ---- pass 1: partial (max, sum exp) per latent token over a part of one sample's rows.  grid (nparts, BT), block 128
template <int NT>
__global__ void __launch_bounds__(32 * kTokWarps) rf_tok_stats_kernel(const avmoe_rf_tok_args p, int tiles_per_part) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ float red[kTokWarps][NT * 8][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = lane & 3;
  const int pitch = 2 * p.C + 16;
  unsigned char* tiles = smem + (size_t)warp * 2 * 16 * pitch;
  const TokGeom G = tok_geom(p, tiles_per_part, NT);
  float m[NT][2], l[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { m[nt][0] = m[nt][1] = -3.0e38f; l[nt][0] = l[nt][1] = 0.f; }
  long long tl = G.tile0 + warp;
  if (tl < G.tile_end) tok_tile_load(tiles, pitch, G.kv + tl * 16 * p.C, p.C, lane);
  cp_async_commit();
  int buf = 0;
  for (; tl < G.tile_end; tl += kTokWarps) {
    const long long nx = tl + kTokWarps;
    if (nx < G.tile_end) tok_tile_load(tiles + (buf ^ 1) * 16 * pitch, pitch, G.kv + nx * 16 * p.C, p.C, lane);
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    float S[NT][4];
    tok_logits<NT>(S, tiles + buf * 16 * pitch, pitch, p.C, G.bp, lane);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float sA = S[nt][i], sB = S[nt][2 + i];
        const float M = fmaxf(m[nt][i], fmaxf(sA, sB));
        l[nt][i] = l[nt][i] * __expf(m[nt][i] - M) + __expf(sA - M) + __expf(sB - M);
        m[nt][i] = M;
      }
    __syncwarp();
    buf ^= 1;
  }
  cp_async_wait<0>();
  // the 8 lanes that share t hold the same columns for different rows
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int off = 4; off < 32; off <<= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m[nt][i], off), l2 = __shfl_xor_sync(0xffffffffu, l[nt][i], off);
        merge_ml(m[nt][i], l[nt][i], m2, l2);
      }
      if (lane < 4) { red[warp][8 * nt + 2 * t + i][0] = m[nt][i]; red[warp][8 * nt + 2 * t + i][1] = l[nt][i]; }
    }
  __syncthreads();
  if (threadIdx.x < NT * 8) {
    float M = red[0][threadIdx.x][0], L = red[0][threadIdx.x][1];
    for (int w = 1; w < kTokWarps; ++w) merge_ml(M, L, red[w][threadIdx.x][0], red[w][threadIdx.x][1]);
    float* out = p.part + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 2 * (NT * 8);
    out[threadIdx.x] = M;
    out[NT * 8 + threadIdx.x] = L;
  }
}

// per-warp staging of one 16-row slab of A12 (row = 2*Tp bf16): accumulator-layout 4-byte words in, 16-byte row chunks out
struct A12Stage {
  unsigned char* buf;
  int spitch, cpr;       // bytes per staged row, 16-byte chunks per row
  __device__ __forceinline__ A12Stage(unsigned char* b, int Tp) : buf(b), spitch(4 * Tp + 16), cpr(Tp >> 2) {}
  __device__ __forceinline__ uint32_t& word(int row, int col) { return *reinterpret_cast<uint32_t*>(buf + row * spitch + col * 2); }
};

// ---- pass 2: A12[b][row][0..T) = bf16(p), A12[b][row][Tp..Tp+T) = bf16(p - hi),  p = exp(s - lse[b][t]).  grid (nparts, BT)
template <int NT>
__global__ void __launch_bounds__(32 * kTokWarps) rf_tok_probs_kernel(const avmoe_rf_tok_args p, int tiles_per_part, int nstat_parts) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ float lse_s[NT * 8];
  constexpr int Tp = NT * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int pitch = 2 * p.C + 16;
  unsigned char* tiles = smem + (size_t)warp * 2 * 16 * pitch;
  A12Stage st(smem + (size_t)kTokWarps * 2 * 16 * pitch + (size_t)warp * 16 * (4 * Tp + 16), Tp);
  const TokGeom G = tok_geom(p, tiles_per_part, NT);
  long long tl = G.tile0 + warp;
  if (tl < G.tile_end) tok_tile_load(tiles, pitch, G.kv + tl * 16 * p.C, p.C, lane);
  cp_async_commit();
  if (threadIdx.x < Tp) {
    const float* part = p.part + (long long)blockIdx.y * nstat_parts * 2 * Tp;
    float M = part[threadIdx.x], L = part[Tp + threadIdx.x];
    for (int q = 1; q < nstat_parts; ++q) merge_ml(M, L, part[q * 2 * Tp + threadIdx.x], part[q * 2 * Tp + Tp + threadIdx.x]);
    const float v = M + __logf(L);
    lse_s[threadIdx.x] = v;
    if (p.lse && blockIdx.x == 0 && threadIdx.x < p.T) p.lse[(long long)blockIdx.y * p.T + threadIdx.x] = v;
  }
  __syncthreads();
  float lse[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { lse[nt][0] = lse_s[8 * nt + 2 * t]; lse[nt][1] = lse_s[8 * nt + 2 * t + 1]; }
  __nv_bfloat16* A12 = static_cast<__nv_bfloat16*>(p.A12) + (long long)blockIdx.y * p.Nk * 2 * Tp;
  int buf = 0;
  for (; tl < G.tile_end; tl += kTokWarps) {
    const long long nx = tl + kTokWarps;
    if (nx < G.tile_end) tok_tile_load(tiles + (buf ^ 1) * 16 * pitch, pitch, G.kv + nx * 16 * p.C, p.C, lane);
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    float S[NT][4];
    tok_logits<NT>(S, tiles + buf * 16 * pitch, pitch, p.C, G.bp, lane);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float pv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ok = 8 * nt + 2 * t + (i & 1) < p.T;
        pv[i] = ok ? __expf(S[nt][i] - lse[nt][i & 1]) : 0.f;
      }
      const uint32_t hA = pack_bf16x2(pv[0], pv[1]), hB = pack_bf16x2(pv[2], pv[3]);
      const float2 fA = unpack_bf16x2(hA), fB = unpack_bf16x2(hB);
      st.word(g, 8 * nt + 2 * t) = hA;
      st.word(g + 8, 8 * nt + 2 * t) = hB;
      st.word(g, Tp + 8 * nt + 2 * t) = pack_bf16x2(pv[0] - fA.x, pv[1] - fA.y);      // rounding residual: (hi + lo) · kv keeps ~16 bits of A1
      st.word(g + 8, Tp + 8 * nt + 2 * t) = pack_bf16x2(pv[2] - fB.x, pv[3] - fB.y);
    }
    __syncwarp();
    {
      __nv_bfloat16* dst = A12 + tl * 16 * 2 * Tp;
      for (int i = lane; i < 16 * st.cpr; i += 32) {
        const int r = i / st.cpr, ch = i - r * st.cpr;
        *reinterpret_cast<uint4*>(dst + (long long)r * 2 * Tp + ch * 8) = *reinterpret_cast<const uint4*>(st.buf + r * st.spitch + ch * 16);
      }
    }
    __syncwarp();
    buf ^= 1;
  }
  cp_async_wait<0>();
}

// ---- backward: A12[b][row][Tp + t] = A1ᵀ[row][t] * (dA1ᵀ[row][t] - Dt[b][t]),  dA1ᵀ = kv · dRᵀ (per-sample packed dR).  grid (nparts, BT)
template <int NT>
__global__ void __launch_bounds__(32 * kTokWarps) rf_tok_ds_kernel(const avmoe_rf_tok_args p, int tiles_per_part) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int Tp = NT * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int pitch = 2 * p.C + 16;
  const int apitch = 4 * Tp + 16;                       // staged A12 rows (hi | second half)
  unsigned char* tiles = smem + (size_t)warp * 2 * 16 * (pitch + apitch);
  const TokGeom G = tok_geom(p, tiles_per_part, NT);
  __nv_bfloat16* A12 = static_cast<__nv_bfloat16*>(p.A12) + (long long)blockIdx.y * p.Nk * 2 * Tp;
  const int cpr = Tp >> 2;                               // 16-byte chunks per A12 row
  auto load = [&](unsigned char* dstbuf, long long tile) {
    tok_tile_load(dstbuf, pitch, G.kv + tile * 16 * p.C, p.C, lane);
    unsigned char* ab = dstbuf + 16 * pitch;
    const __nv_bfloat16* src = A12 + tile * 16 * 2 * Tp;
    for (int i = lane; i < 16 * (cpr >> 1); i += 32) {   // the hi half of each row only
      const int r = i / (cpr >> 1), ch = i - r * (cpr >> 1);
      cp_async16(ab + r * apitch + ch * 16, src + (long long)r * 2 * Tp + ch * 8);
    }
  };
  float dt[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = 8 * nt + 2 * t + i;
      dt[nt][i] = c < p.T ? __ldg(p.Dt + (long long)blockIdx.y * p.T + c) : 0.f;
    }
  long long tl = G.tile0 + warp;
  if (tl < G.tile_end) load(tiles, tl);
  cp_async_commit();
  int buf = 0;
  for (; tl < G.tile_end; tl += kTokWarps) {
    const long long nx = tl + kTokWarps;
    if (nx < G.tile_end) load(tiles + (buf ^ 1) * 16 * (pitch + apitch), nx);
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    unsigned char* kvt = tiles + buf * 16 * (pitch + apitch);
    unsigned char* ab = kvt + 16 * pitch;
    float S[NT][4];
    tok_logits<NT>(S, kvt, pitch, p.C, G.bp, lane);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int c = 8 * nt + 2 * t;
      const float2 aA = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(ab + g * apitch + c * 2));
      const float2 aB = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(ab + (g + 8) * apitch + c * 2));
      const bool ok0 = c < p.T, ok1 = c + 1 < p.T;
      const uint32_t wA = pack_bf16x2(ok0 ? aA.x * (S[nt][0] - dt[nt][0]) : 0.f, ok1 ? aA.y * (S[nt][1] - dt[nt][1]) : 0.f);
      const uint32_t wB = pack_bf16x2(ok0 ? aB.x * (S[nt][2] - dt[nt][0]) : 0.f, ok1 ? aB.y * (S[nt][3] - dt[nt][1]) : 0.f);
      *reinterpret_cast<uint32_t*>(ab + g * apitch + (Tp + c) * 2) = wA;
      *reinterpret_cast<uint32_t*>(ab + (g + 8) * apitch + (Tp + c) * 2) = wB;
    }
    __syncwarp();
    {
      __nv_bfloat16* dst = A12 + tl * 16 * 2 * Tp + Tp;   // second half of each row
      for (int i = lane; i < 16 * (cpr >> 1); i += 32) {
        const int r = i / (cpr >> 1), ch = i - r * (cpr >> 1);
        *reinterpret_cast<uint4*>(dst + (long long)r * 2 * Tp + ch * 8) = *reinterpret_cast<const uint4*>(ab + r * apitch + (Tp * 2) + ch * 16);
      }
    }
    __syncwarp();
    buf ^= 1;
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------ host side
static int tok_check(const avmoe_rf_tok_args* a, const char* who) {
  AVMOE_CHECK_ARG(a && a->kv && a->pack && a->A12, "%s: null tensor", who);
  AVMOE_CHECK_ARG(a->BT > 0 && a->Nk > 0 && a->Nk % 16 == 0 && a->C >= 16 && a->C % 16 == 0 && a->C <= 384, "%s: needs Nk %% 16 == 0, C %% 16 == 0, C <= 384", who);
  AVMOE_CHECK_ARG(a->T >= 1 && a->T <= 96 && a->Tp == (a->T + 7) / 8 * 8, "%s: T <= 96 and Tp = T rounded up to 8 (got T=%d Tp=%d)", who, a->T, a->Tp);
  AVMOE_CHECK_ARG(a->BT <= 65535, "%s: BT too large for the grid", who);
  return 0;
}

struct TokPlan { int nparts, tiles_per_part; };
static TokPlan tok_plan(const avmoe_rf_tok_args* a) {
  const int tiles = a->Nk / 16;
  int want = (8 * num_sms() + a->BT - 1) / a->BT;              // ~8 CTAs of 4 warps per SM
  int maxp = tiles / (2 * kTokWarps);                           // at least two tiles per warp
  if (maxp < 1) maxp = 1;
  if (want > maxp) want = maxp;
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  TokPlan pl;
  pl.tiles_per_part = (tiles + want - 1) / want;
  pl.nparts = (tiles + pl.tiles_per_part - 1) / pl.tiles_per_part;
  return pl;
}

template <typename Kern>
static int tok_smem_attr(Kern k, size_t smem, const char* who) {
  AVMOE_CHECK_ARG(smem <= 200 * 1024, "%s: tile does not fit shared memory (%zu bytes)", who, smem);
  if (smem > 48 * 1024) AVMOE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return 0;
}

#define TOK_DISPATCH(NTV, ...)                                           \
  switch (NTV) {                                                         \
    case 1: { constexpr int NT = 1; __VA_ARGS__ } break;                 \
    case 2: { constexpr int NT = 2; __VA_ARGS__ } break;                 \
    case 4: { constexpr int NT = 4; __VA_ARGS__ } break;                 \
    case 6: { constexpr int NT = 6; __VA_ARGS__ } break;                 \
    case 8: { constexpr int NT = 8; __VA_ARGS__ } break;                 \
    case 11: { constexpr int NT = 11; __VA_ARGS__ } break;               \
    case 12: { constexpr int NT = 12; __VA_ARGS__ } break;               \
    default: set_error("rf_tok: unsupported latent-token count"); return AVMOE_E_UNSUPPORTED; \
  }

static bool tok_nt_ok(int nt) { return nt == 1 || nt == 2 || nt == 4 || nt == 6 || nt == 8 || nt == 11 || nt == 12; }

}  // namespace rf
}  // namespace avmoe

using namespace avmoe;
using namespace avmoe::rf;

extern "C" int avmoe_rf_tok_supported(int Nk, int C, int T) {
  return (Nk >= 16 && Nk % 16 == 0 && C >= 16 && C % 16 == 0 && C <= 384 && T >= 1 && T <= 96 && tok_nt_ok((T + 7) / 8)) ? 1 : 0;   // wider rows: the double-buffered tiles of 4 warps exceed shared memory
}

extern "C" int avmoe_rf_tok_parts(int BT, int Nk) {
  avmoe_rf_tok_args a = {};
  a.BT = BT; a.Nk = Nk;
  if (BT <= 0 || Nk < 16) return 1;
  return tok_plan(&a).nparts;
}

extern "C" int avmoe_rf_tok_probs(const avmoe_rf_tok_args* a, avmoe_stream_t stream) {
  if (int rc = tok_check(a, "rf_tok_probs")) return rc;
  AVMOE_CHECK_ARG(a->part && !a->per_sample, "rf_tok_probs: needs the partials workspace and a shared packed operand");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const TokPlan pl = tok_plan(a);
  const int NTV = a->Tp / 8;
  const int pitch = 2 * a->C + 16;
  const size_t smem1 = (size_t)kTokWarps * 2 * 16 * pitch;
  const size_t smem2 = smem1 + (size_t)kTokWarps * 16 * (4 * a->Tp + 16);
  dim3 grid(pl.nparts, a->BT);
  TOK_DISPATCH(NTV,
    if (int rc = tok_smem_attr(rf_tok_stats_kernel<NT>, smem1, "rf_tok_stats")) return rc;
    if (int rc = tok_smem_attr(rf_tok_probs_kernel<NT>, smem2, "rf_tok_probs")) return rc;
    rf_tok_stats_kernel<NT><<<grid, 32 * kTokWarps, smem1, st>>>(*a, pl.tiles_per_part);
    AVMOE_LAUNCH_CHECK("rf_tok_stats_kernel");
    rf_tok_probs_kernel<NT><<<grid, 32 * kTokWarps, smem2, st>>>(*a, pl.tiles_per_part, pl.nparts);
    AVMOE_LAUNCH_CHECK("rf_tok_probs_kernel");
  )
  return 0;
}

extern "C" int avmoe_rf_tok_ds(const avmoe_rf_tok_args* a, avmoe_stream_t stream) {
  if (int rc = tok_check(a, "rf_tok_ds")) return rc;
  AVMOE_CHECK_ARG(a->Dt && a->per_sample, "rf_tok_ds: needs D[b,t] and a per-sample packed dR operand");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const TokPlan pl = tok_plan(a);
  const int NTV = a->Tp / 8;
  const int pitch = 2 * a->C + 16, apitch = 4 * a->Tp + 16;
  const size_t smem = (size_t)kTokWarps * 2 * 16 * (pitch + apitch);
  dim3 grid(pl.nparts, a->BT);
  TOK_DISPATCH(NTV,
    if (int rc = tok_smem_attr(rf_tok_ds_kernel<NT>, smem, "rf_tok_ds")) return rc;
    rf_tok_ds_kernel<NT><<<grid, 32 * kTokWarps, smem, st>>>(*a, pl.tiles_per_part);
    AVMOE_LAUNCH_CHECK("rf_tok_ds_kernel");
  )
  return 0;
}
