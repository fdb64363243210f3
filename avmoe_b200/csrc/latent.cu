// Latent-token attention, first half: the T learned tokens of an expert attend to a key/value token
// stream kv (the resampled other modality for multimodal experts, x itself for the AVS-v2 latent
// self-attention) and are updated residually:
//     tok' = tok + softmax_N(tok · kvᵀ) · kv            (no 1/sqrt(d) scaling)
// Reference: AVMOE/AVE/nets/net_trans_v3.py:379-383 (rep_token / att_v2tk / rep_token_res) and
// AVMOE/AVS/avs_scripts/avs_s4/model/PVT_AVSModel_v2.py:217-222.
//
// Flash-style: the [T,N] attention matrix is never written to HBM.  Forward splits N over CTAs with an
// online softmax and merges the partial (max, sum, acc) triples; backward recomputes the probabilities
// from the saved log-sum-exp.  fp32 math; kv is read once per pass in its storage dtype.
#include <type_traits>

#include "block_mma.cuh"
#include "common.cuh"

namespace avmoe {

struct LatentParams {
  const void* kv;                          // [BT,Nk,C]
  const float* tok[AVMOE_MAX_EXPERTS];     // [T,C]
  int slot[AVMOE_MAX_EXPERTS];
  int nexp, nl;
  float* tokp;                             // [BT,nl,T,C]
  float* lse;                              // [BT,nl,T]
  float* part;                             // [BT][nexp][nsplit][T][C]
  float* part_ml;                          // [BT][nexp][nsplit][T][2]
  int BT, Nk, C, T, nsplit, tokens_per_split, TN, CC, csplit;
  // backward only
  const float* dtokp;                      // [BT,nl,T,C]
  const float* dpool;                      // [BT,C] or null
  void* dkv;                               // [BT,Nk,C]
  int accumulate;
  float* dtok[AVMOE_MAX_EXPERTS];          // [T,C] accumulators
};

template <typename T>
__device__ __forceinline__ void load_kv_tile(float* kvs, const T* kvb, int n0, int n1, int TN, int C, int ld) {
  const int C4 = C >> 2;
  for (int i = threadIdx.x; i < TN * C4; i += blockDim.x) {
    const int n = i / C4, c4 = i - n * C4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (n0 + n < n1) load4<T>(kvb + (long long)(n0 + n) * C + c4 * 4, v);
    float* d = kvs + n * ld + c4 * 4;
    d[0] = v[0]; d[1] = v[1]; d[2] = v[2]; d[3] = v[3];
  }
}

// grid (nsplit*csplit, nexp, BT), block 256
template <typename T>
__global__ void __launch_bounds__(256, 2) latent_fwd_kernel(const LatentParams p) {
  extern __shared__ float sm[];
  constexpr bool kExact = std::is_same<T, float>::value;
  const int ld = p.C + 4, lds = p.TN + 4;
  float* kvs = sm;                            // [TN][C+4]
  float* S = kvs + p.TN * ld;                 // [T][TN+4]
  float* acc = S + p.T * lds;                 // [T][CC]
  float* mrun = acc + p.T * p.CC;             // [T]
  float* lrun = mrun + p.T;
  float* alpha = lrun + p.T;
  const int split = blockIdx.x / p.csplit, cs = blockIdx.x % p.csplit;
  const int ei = blockIdx.y, b = blockIdx.z;
  const int c_lo = cs * p.CC, c_hi = min(p.C, c_lo + p.CC);
  const int cw = c_hi - c_lo;
  const float* tok = p.tok[ei];
  const T* kvb = reinterpret_cast<const T*>(p.kv) + (long long)b * p.Nk * p.C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < p.T * p.CC; i += blockDim.x) acc[i] = 0.f;
  for (int t = threadIdx.x; t < p.T; t += blockDim.x) { mrun[t] = -INFINITY; lrun[t] = 0.f; }
  const int nb = split * p.tokens_per_split, ne = min(p.Nk, nb + p.tokens_per_split);
  __syncthreads();
  for (int n0 = nb; n0 < ne; n0 += p.TN) {
    load_kv_tile<T>(kvs, kvb, n0, ne, p.TN, p.C, ld);
    __syncthreads();
    // logits S[t][n] = tok_t . kv_n
    contract_to<kExact, Out::Store>(p.T, p.TN, p.C, tok, p.C, 1, kvs, 1, ld, S, lds, 1);
    __syncthreads();
    const int nvalid = min(p.TN, ne - n0);   // columns >= nvalid are padding: treated as -inf below
    // online softmax per latent token (one warp per t)
    for (int t = warp; t < p.T; t += nwarps) {
      const float mold = mrun[t];                 // read by every lane before lane 0 updates it below
      float mx = -INFINITY;
      for (int n = lane; n < nvalid; n += 32) mx = fmaxf(mx, S[t * lds + n]);
      mx = warp_max(mx);
      const float mnew = fmaxf(mold, mx);
      float sum = 0.f;
      for (int n = lane; n < p.TN; n += 32) {
        const float e = n < nvalid ? expf(S[t * lds + n] - mnew) : 0.f;
        S[t * lds + n] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      const float a = expf(mold - mnew);           // first tile: exp(-inf) = 0
      if (lane == 0) {
        lrun[t] = lrun[t] * a + sum;
        mrun[t] = mnew;
      }
      // rescale this row of the accumulator for the new running maximum
      for (int c = lane; c < cw; c += 32) acc[t * p.CC + c] *= a;
    }
    __syncthreads();
    // acc[t][c] = alpha[t]*acc[t][c] + sum_n P[t][n]*kv[n][c]   (this CTA's channel range only)
    contract_to<kExact, Out::Add>(p.T, cw, p.TN, S, lds, 1, kvs + c_lo, ld, 1, acc, p.CC, 1);
    __syncthreads();
  }
  float* pp = p.part + ((((long long)b * p.nexp + ei) * p.nsplit + split) * p.T) * p.C;
  for (int i = threadIdx.x; i < p.T * cw; i += blockDim.x) {
    const int t = i / cw, c = i - t * cw;
    pp[(long long)t * p.C + c_lo + c] = acc[t * p.CC + c];
  }
  if (cs == 0) {
    float* pm = p.part_ml + ((((long long)b * p.nexp + ei) * p.nsplit + split) * p.T) * 2;
    for (int t = threadIdx.x; t < p.T; t += blockDim.x) { pm[t * 2] = mrun[t]; pm[t * 2 + 1] = lrun[t]; }
  }
}

// grid (T, nexp, BT), block 128: merge the split partials
__global__ void latent_merge_kernel(const LatentParams p) {
  const int t = blockIdx.x, ei = blockIdx.y, b = blockIdx.z;
  const long long base = ((long long)b * p.nexp + ei) * p.nsplit;
  float M = -INFINITY;
  for (int s = 0; s < p.nsplit; ++s) M = fmaxf(M, p.part_ml[((base + s) * p.T + t) * 2]);
  float L = 0.f;
  for (int s = 0; s < p.nsplit; ++s) {
    const float* ml = p.part_ml + ((base + s) * p.T + t) * 2;
    L += ml[1] * expf(ml[0] - M);
  }
  const float invL = 1.f / L;
  const int slot = p.slot[ei];
  float* outp = p.tokp + (((long long)b * p.nl + slot) * p.T + t) * p.C;
  const float* tok = p.tok[ei] + (long long)t * p.C;
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    float a = 0.f;
    for (int s = 0; s < p.nsplit; ++s) {
      const float w = expf(p.part_ml[((base + s) * p.T + t) * 2] - M);
      a = fmaf(p.part[((base + s) * p.T + t) * p.C + c], w, a);
    }
    outp[c] = tok[c] + a * invL;
  }
  if (threadIdx.x == 0) p.lse[((long long)b * p.nl + slot) * p.T + t] = M + logf(L);
}

// grid (nsplit*csplit, BT), block 256; loops over the call's experts so dkv is written exactly once
template <typename T>
__global__ void __launch_bounds__(256, 2) latent_bwd_kernel(const LatentParams p) {
  extern __shared__ float sm[];
  constexpr bool kExact = std::is_same<T, float>::value;
  const int ld = p.C + 4, lds = p.TN + 4;
  float* kvs = sm;                            // [TN][C+4]
  float* A1 = kvs + p.TN * ld;                // [T][TN+4]  probabilities
  float* dS = A1 + p.T * lds;                 // [T][TN+4]
  float* dkvt = dS + p.T * lds;               // [TN][CC]
  float* dtk = dkvt + p.TN * p.CC;            // [nexp][T][CC]
  float* Dt = dtk + p.nexp * p.T * p.CC;      // [nexp][T]
  const int split = blockIdx.x / p.csplit, cs = blockIdx.x % p.csplit;
  const int b = blockIdx.y;
  const int c_lo = cs * p.CC, c_hi = min(p.C, c_lo + p.CC);
  const int cw = c_hi - c_lo;
  const T* kvb = reinterpret_cast<const T*>(p.kv) + (long long)b * p.Nk * p.C;
  T* dkvb = reinterpret_cast<T*>(p.dkv) + (long long)b * p.Nk * p.C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < p.nexp * p.T * p.CC; i += blockDim.x) dtk[i] = 0.f;
  // D[t] = dR_t . R_t  with R = tok' - tok
  for (int i = warp; i < p.nexp * p.T; i += nwarps) {
    const int ei = i / p.T, t = i - ei * p.T;
    const long long off = (((long long)b * p.nl + p.slot[ei]) * p.T + t) * p.C;
    const float* tk = p.tok[ei] + (long long)t * p.C;
    float s = 0.f;
    for (int c = lane; c < p.C; c += 32) s = fmaf(p.dtokp[off + c], p.tokp[off + c] - tk[c], s);
    s = warp_sum(s);
    if (lane == 0) Dt[i] = s;
  }
  const int nb = split * p.tokens_per_split, ne = min(p.Nk, nb + p.tokens_per_split);
  __syncthreads();
  for (int n0 = nb; n0 < ne; n0 += p.TN) {
    load_kv_tile<T>(kvs, kvb, n0, ne, p.TN, p.C, ld);
    for (int i = threadIdx.x; i < p.TN * p.CC; i += blockDim.x) dkvt[i] = 0.f;
    __syncthreads();
    for (int ei = 0; ei < p.nexp; ++ei) {
      const float* tok = p.tok[ei];
      const long long sbase = ((long long)b * p.nl + p.slot[ei]) * p.T;
      const float* dR = p.dtokp + sbase * p.C;
      // probabilities A1 = exp(tok.kv - lse) and dS = A1 * (dR.kv - D)
      contract_to<kExact, Out::Store>(p.T, p.TN, p.C, tok, p.C, 1, kvs, 1, ld, A1, lds, 1);
      contract_to<kExact, Out::Store>(p.T, p.TN, p.C, dR, p.C, 1, kvs, 1, ld, dS, lds, 1);
      __syncthreads();
      {
        const int nvalid = min(p.TN, ne - n0);
        for (int t = warp; t < p.T; t += nwarps) {
          const float l = p.lse[sbase + t], dt = Dt[ei * p.T + t];
          for (int n = lane; n < p.TN; n += 32) {
            const float a1 = n < nvalid ? expf(A1[t * lds + n] - l) : 0.f;
            A1[t * lds + n] = a1;
            dS[t * lds + n] = a1 * (dS[t * lds + n] - dt);
          }
        }
      }
      __syncthreads();
      // dkv[n][c] += sum_t A1[t][n]*dR[t][c] + dS[t][n]*tok[t][c]
      contract_to<kExact, Out::Add>(p.TN, cw, p.T, A1, 1, lds, dR + c_lo, p.C, 1, dkvt, p.CC, 1);
      contract_to<kExact, Out::Add>(p.TN, cw, p.T, dS, 1, lds, tok + c_lo, p.C, 1, dkvt, p.CC, 1);
      // dtok[t][c] += sum_n dS[t][n]*kv[n][c]
      contract_to<kExact, Out::Add>(p.T, cw, p.TN, dS, lds, 1, kvs + c_lo, ld, 1, dtk + ei * p.T * p.CC, p.CC, 1);
      __syncthreads();
    }
    // write the dkv tile
    for (int i = threadIdx.x; i < p.TN * cw; i += blockDim.x) {
      const int n = i / cw, c = i - n * cw;
      if (n0 + n < ne) {
        const long long off = (long long)(n0 + n) * p.C + c_lo + c;
        float v = dkvt[n * p.CC + c];
        if (p.dpool) v += p.dpool[(long long)b * p.C + c_lo + c];
        if (p.accumulate) v += to_f32(dkvb[off]);
        dkvb[off] = from_f32<T>(v);
      }
    }
    __syncthreads();
  }
  for (int ei = 0; ei < p.nexp; ++ei) {
    const long long sbase = ((long long)b * p.nl + p.slot[ei]) * p.T;
    for (int i = threadIdx.x; i < p.T * cw; i += blockDim.x) {
      const int t = i / cw, c = i - t * cw;
      float v = dtk[(ei * p.T + t) * p.CC + c];
      if (split == 0) v += p.dtokp[(sbase + t) * p.C + c_lo + c];   // d tok' / d tok = I (residual)
      atomicAdd(p.dtok[ei] + (long long)t * p.C + c_lo + c, v);
    }
  }
}

static int plan_latent(LatentParams& p, bool bwd, size_t* smem) {
  // kv tile: keep TN*(C+1) floats under 48 KiB
  int TN = 12288 / (p.C + 4);
  if (TN > 64) TN = 64;
  if (TN >= 32) TN = TN / 32 * 32; else if (TN >= 8) TN = TN / 8 * 8; else TN = TN < 1 ? 1 : TN;
  p.TN = TN;
  // channel range per CTA so that the [T][CC] (fwd) / [nexp][T][CC]+[TN][CC] (bwd) accumulators stay under 64 KiB
  const int rows = bwd ? p.nexp * p.T + TN : p.T;
  int CC = 16384 / rows;
  if (CC >= p.C) CC = p.C;
  else CC = CC / 4 * 4;
  if (CC < 4) CC = 4;
  p.CC = CC;
  p.csplit = (p.C + CC - 1) / CC;
  // token splits: aim at ~3 CTAs per SM in total
  const long long ctas_per_split = (long long)p.BT * p.csplit * (bwd ? 1 : p.nexp);
  int nsplit = (int)((3LL * num_sms() + ctas_per_split - 1) / ctas_per_split);
  const int max_split = (p.Nk + TN - 1) / TN;
  if (nsplit > max_split) nsplit = max_split;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > 32) nsplit = 32;
  p.tokens_per_split = ((p.Nk + nsplit - 1) / nsplit + TN - 1) / TN * TN;
  p.nsplit = (p.Nk + p.tokens_per_split - 1) / p.tokens_per_split;
  const size_t lds = TN + 4;
  if (!bwd) *smem = sizeof(float) * ((size_t)TN * (p.C + 4) + (size_t)p.T * lds + (size_t)p.T * CC + 3 * (size_t)p.T);
  else *smem = sizeof(float) * ((size_t)TN * (p.C + 4) + 2 * (size_t)p.T * lds + (size_t)TN * CC + (size_t)p.nexp * p.T * CC + (size_t)p.nexp * p.T);
  return 0;
}

}  // namespace avmoe

using namespace avmoe;

// workspace query: floats needed for `part` and `part_ml` of avmoe_latent_fwd
extern "C" int64_t avmoe_latent_fwd_workspace(int BT, int Nk, int C, int T, int nexp) {
  LatentParams p{};
  p.BT = BT; p.Nk = Nk; p.C = C; p.T = T; p.nexp = nexp;
  size_t smem;
  plan_latent(p, false, &smem);
  return (int64_t)BT * nexp * p.nsplit * T * (C + 2);
}

extern "C" int avmoe_latent_fwd(const void* kv, const float* const* tokens, const int* slots, int nexp, int nl, float* tokp, float* lse,
                                float* workspace, int BT, int Nk, int C, int T, int dtype, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(kv && tokens && slots && tokp && lse && workspace, "latent_fwd: null pointer");
  AVMOE_CHECK_ARG(nexp >= 1 && nexp <= AVMOE_MAX_EXPERTS && nl >= nexp && BT > 0 && Nk > 0 && T > 0 && C > 0 && (C % 4) == 0,
                  "latent_fwd: bad sizes (C must be a multiple of 4)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LatentParams p{};
  p.kv = kv; p.nexp = nexp; p.nl = nl; p.tokp = tokp; p.lse = lse;
  p.BT = BT; p.Nk = Nk; p.C = C; p.T = T;
  for (int i = 0; i < nexp; ++i) { p.tok[i] = tokens[i]; p.slot[i] = slots[i]; }
  size_t smem;
  plan_latent(p, false, &smem);
  p.part = workspace;
  p.part_ml = workspace + (size_t)BT * nexp * p.nsplit * T * C;
  AVMOE_CHECK_ARG(smem <= 227 * 1024, "latent_fwd: tile does not fit shared memory (T=%d C=%d)", T, C);
  dim3 grid(p.nsplit * p.csplit, nexp, BT);
  if (dtype == AVMOE_F32) {
    AVMOE_CUDA(cudaFuncSetAttribute(latent_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    latent_fwd_kernel<float><<<grid, 256, smem, st>>>(p);
  } else {
    AVMOE_CUDA(cudaFuncSetAttribute(latent_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    latent_fwd_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(p);
  }
  AVMOE_LAUNCH_CHECK("latent_fwd_kernel");
  latent_merge_kernel<<<dim3(T, nexp, BT), 128, 0, st>>>(p);
  AVMOE_LAUNCH_CHECK("latent_merge_kernel");
  return 0;
}

extern "C" int avmoe_latent_bwd(const void* kv, const float* const* tokens, const int* slots, int nexp, int nl, const float* tokp,
                                const float* lse, const float* dtokp, const float* dpool, void* dkv, int accumulate,
                                float* const* dtokens, int BT, int Nk, int C, int T, int dtype, avmoe_stream_t stream) {
  AVMOE_CHECK_ARG(kv && tokens && slots && tokp && lse && dtokp && dkv && dtokens, "latent_bwd: null pointer");
  AVMOE_CHECK_ARG(nexp >= 1 && nexp <= AVMOE_MAX_EXPERTS && nl >= nexp && BT > 0 && Nk > 0 && T > 0 && C > 0 && (C % 4) == 0,
                  "latent_bwd: bad sizes (C must be a multiple of 4)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LatentParams p{};
  p.kv = kv; p.nexp = nexp; p.nl = nl; p.tokp = const_cast<float*>(tokp); p.lse = const_cast<float*>(lse);
  p.dtokp = dtokp; p.dpool = dpool; p.dkv = dkv; p.accumulate = accumulate;
  p.BT = BT; p.Nk = Nk; p.C = C; p.T = T;
  for (int i = 0; i < nexp; ++i) { p.tok[i] = tokens[i]; p.slot[i] = slots[i]; p.dtok[i] = dtokens[i]; }
  size_t smem;
  plan_latent(p, true, &smem);
  AVMOE_CHECK_ARG(smem <= 227 * 1024, "latent_bwd: tile does not fit shared memory (T=%d C=%d)", T, C);
  dim3 grid(p.nsplit * p.csplit, BT);
  if (dtype == AVMOE_F32) {
    AVMOE_CUDA(cudaFuncSetAttribute(latent_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    latent_bwd_kernel<float><<<grid, 256, smem, st>>>(p);
  } else {
    AVMOE_CUDA(cudaFuncSetAttribute(latent_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    latent_bwd_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(p);
  }
  AVMOE_LAUNCH_CHECK("latent_bwd_kernel");
  return 0;
}
