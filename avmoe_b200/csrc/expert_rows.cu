// Row kernels of the GEMM-based expert path (bf16 streams; AVMOE/AVE/nets/net_trans_v3.py:386-434, 481-486).
//
// The four channel contractions of an expert (grouped 1x1 down/up projection, forward and backward, plus their weight
// gradients) run on the tcgen05 GEMM as dense GEMMs against block-diagonal weights; everything between them is a streaming
// pass over token rows:
//
//   xg_ln_pre_fwd     x (+ gate*aux | aux) -> ln_before -> xl[e]                       (GEMM: z[e]  = xl[e] · Wdᵀ)
//   xg_z_stats        column sums of z, z^2 (bn1 batch statistics)
//   xg_bn_act         a[e] = act(bn1(z[e]))                                            (GEMM: yhat[e] = bn2-normalised a[e] · Wuᵀ)
//   xg_post_fwd       out = sum_e probs[b,e] * gate_e * ln_post(bn2 affine(yhat[e]))
//   xg_post_bwd       dout -> dy2[e] (grad at the bn2 output), d ln_post, d gate, d probs, bn2 batch reductions
//   xg_prep_*         block-diagonal / bn-folded weight operands                       (GEMMs: da = [yhat|dy2] · W2ᵀ,  P = [yhat|dy2]ᵀ · a)
//   xg_bn1_bwd_*      relu mask + bn1 backward on the bottleneck                       (GEMMs: dxl = dz · Wd,  dWd = dzᵀ · xl)
//   xg_ln_pre_bwd     dxl[e] -> ln_before backward -> dx, d aux, d gate_pre, d ln_before
//   xg_finish_wgrads  dense weight-gradient products -> grouped conv weight gradients
//
// Rows are handled by "teams" of 1, 2 or 4 warps (C <= 384 / 768 / 1536) so that a lane never owns more than 3 float4 column
// groups; per-column gradient sums live in registers across a CTA's row chunk and are flushed once per CTA.
#include "row_pass.cuh"

namespace avmoe {

static constexpr int kWarps = 8;

template <int kTeam>
struct Team {
  int lane, warp, team, tlane;
  __device__ Team() {
    lane = threadIdx.x & 31;
    warp = threadIdx.x >> 5;
    team = warp / kTeam;
    tlane = (warp % kTeam) * 32 + lane;
  }
};

// plain (coherent) 4-element load: for data this thread stored earlier in the same kernel
template <typename T>
__device__ __forceinline__ float4 ld_act4(const T* p) {
  float v[4];
  load4<T>(p, v);
  return make_float4(v[0], v[1], v[2], v[3]);
}

// Packed 4-element loads kept in their storage form (2 registers for bf16): the next row's operands are requested before the
// current row is processed, so every warp keeps two rows of loads in flight (the kernels are latency-, not bandwidth-limited
// at one row per warp).
template <typename T> struct Raw4;
template <> struct Raw4<float> { float4 v; };
template <> struct Raw4<__nv_bfloat16> { uint2 v; };
__device__ __forceinline__ Raw4<float> ldg_raw(const float* p, bool pred) {
  Raw4<float> r;
  r.v = pred ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
  return r;
}
__device__ __forceinline__ Raw4<__nv_bfloat16> ldg_raw(const __nv_bfloat16* p, bool pred) {
  Raw4<__nv_bfloat16> r;
  r.v = pred ? __ldg(reinterpret_cast<const uint2*>(p)) : make_uint2(0u, 0u);
  return r;
}
__device__ __forceinline__ Raw4<float> ld_raw(const float* p, bool pred) {   // coherent variant
  Raw4<float> r;
  r.v = pred ? *reinterpret_cast<const float4*>(p) : make_float4(0.f, 0.f, 0.f, 0.f);
  return r;
}
__device__ __forceinline__ Raw4<__nv_bfloat16> ld_raw(const __nv_bfloat16* p, bool pred) {
  Raw4<__nv_bfloat16> r;
  r.v = pred ? *reinterpret_cast<const uint2*>(p) : make_uint2(0u, 0u);
  return r;
}
__device__ __forceinline__ float4 cvt(const Raw4<float>& r) { return r.v; }
__device__ __forceinline__ float4 cvt(const Raw4<__nv_bfloat16>& r) {
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&r.v.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&r.v.y);
  return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}

#define TEAM_VECS(k, c) _Pragma("unroll") for (int k = 0, c = tm.tlane * 4; k < kMaxV; ++k, c += 128 * kTeam) if (c < C)

// sum of up to four values over the lanes of a team.  kTeam > 1 goes through shared memory with ONE __syncthreads per call
// (two buffers alternate), so every thread of the CTA must make the same sequence of calls.
template <int kTeam>
__device__ __forceinline__ float4 team_sum(float4 v, float4 (*red)[kWarps], int& par, const Team<kTeam>& tm) {
  v.x = warp_sum(v.x); v.y = warp_sum(v.y); v.z = warp_sum(v.z); v.w = warp_sum(v.w);
  if (kTeam == 1) return v;
  if (tm.lane == 0) red[par][tm.warp] = v;
  __syncthreads();
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < kTeam; ++j) s = s + red[par][tm.team * kTeam + j];
  par ^= 1;
  return s;
}

// flush one per-lane column accumulator of every team: out[c * stride] += sum over teams (OutT = float or double atomics)
template <int kMaxV, int kTeam, typename OutT>
__device__ __forceinline__ void flush_cols(const float4 (&acc)[kMaxV], float* buf, OutT* out, int stride, int C, const Team<kTeam>& tm) {
  constexpr int kTeams = kWarps / kTeam;
  __syncthreads();
  TEAM_VECS(k, c) st4(buf + tm.team * C + c, acc[k]);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < kTeams; ++t) s += buf[t * C + c];
    atomicAdd(out + (long long)c * stride, (OutT)s);
  }
}

// ------------------------------------------------------------------------------------------------ forward: pre-transform + ln_before
template <typename T, int kMaxV, int kTeam>
__global__ void __launch_bounds__(256) xg_ln_pre_fwd_kernel(const avmoe_xg_args p, long long rows, int rpb) {
  constexpr int kTeams = kWarps / kTeam;
  const Team<kTeam> tm;
  __shared__ float4 red[2][kWarps];
  int par = 0;
  const int C = p.C;
  const float invC = 1.f / C;
  const long long r0 = blockIdx.x * (long long)rpb, r1 = min(rows, r0 + rpb);
  // next-row prefetch of x and of the first expert's aux stream (the common case: one aux-carrying expert)
  const T* xin = static_cast<const T*>(p.x);
  const T* aux0 = p.ex[0].pre_kind == AVMOE_PRE_NONE ? nullptr : static_cast<const T*>(p.ex[0].aux);
  Raw4<T> nx[kMaxV], na[kMaxV];
  {
    const long long row = r0 + tm.team;
    const bool act = row < r1;
    TEAM_VECS(k, c) { nx[k] = ldg_raw(xin + row * C + c, act); na[k] = ldg_raw(aux0 + row * C + c, act && aux0); }
  }
  for (long long rb = r0; rb < r1; rb += kTeams) {
    const long long row = rb + tm.team;
    const bool act = row < r1;
    float4 xv[kMaxV], a0[kMaxV];
    TEAM_VECS(k, c) { xv[k] = cvt(nx[k]); a0[k] = cvt(na[k]); }
    {
      const long long nrow = row + kTeams;
      const bool nact = nrow < r1;
      TEAM_VECS(k, c) { nx[k] = ldg_raw(xin + nrow * C + c, nact); na[k] = ldg_raw(aux0 + nrow * C + c, nact && aux0); }
    }
    for (int e = 0; e < p.E; ++e) {
      const avmoe_expert& ex = p.ex[e];
      float4 v[kMaxV];
      if (ex.pre_kind == AVMOE_PRE_NONE) {
        TEAM_VECS(k, c) v[k] = xv[k];
      } else {
        const T* aux = static_cast<const T*>(ex.aux) + row * C;
        const float g = ex.pre_kind == AVMOE_PRE_AUX_ADD ? __ldg(ex.gate_pre) : 1.f;
        TEAM_VECS(k, c) {
          const float4 a = e == 0 ? a0[k] : (act ? ldg_act4<T>(aux + c) : f4(0.f));
          v[k] = ex.pre_kind == AVMOE_PRE_AUX_ADD ? xv[k] + g * a : a;
        }
      }
      if (ex.ln1_w) {
        float s = 0.f;
        TEAM_VECS(k, c) s += hsum(v[k]);
        const float mean = team_sum<kTeam>(make_float4(s, 0.f, 0.f, 0.f), red, par, tm).x * invC;
        float ss = 0.f;
        TEAM_VECS(k, c) { const float4 d = v[k] - f4(mean); ss += hsum(d * d); }
        const float rstd = rsqrtf(team_sum<kTeam>(make_float4(ss, 0.f, 0.f, 0.f), red, par, tm).x * invC + kLnEps);
        TEAM_VECS(k, c) v[k] = (rstd * (v[k] - f4(mean))) * ldg4(ex.ln1_w + c) + ldg4(ex.ln1_b + c);
        if (act && tm.tlane == 0) {
          float* st = p.ln1_stats + ((long long)e * rows + row) * 2;
          st[0] = mean; st[1] = rstd;
        }
      }
      if (act) {
        T* dst = static_cast<T*>(p.xl) + ((long long)e * rows + row) * C;
        TEAM_VECS(k, c) st_act4<T>(dst + c, v[k]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward: bn2 affine + ln_post + combine
template <typename T, int kMaxV, int kTeam>
__global__ void __launch_bounds__(256, 3) xg_post_fwd_kernel(const avmoe_xg_args p, long long rows, int rpb) {
  constexpr int kTeams = kWarps / kTeam;
  const Team<kTeam> tm;
  __shared__ float4 red[2][kWarps];
  int par = 0;
  const int C = p.C;
  const float invC = 1.f / C;
  const long long r0 = blockIdx.x * (long long)rpb, r1 = min(rows, r0 + rpb);
  for (long long rb = r0; rb < r1; rb += kTeams) {
    const long long row = rb + tm.team;
    const bool act = row < r1;
    const long long b = act ? row / p.N : 0;
    float4 acc[kMaxV];
    TEAM_VECS(k, c) acc[k] = f4(0.f);
    // all experts' rows are requested up front (independent loads in flight), then reduced one expert at a time
    Raw4<T> ry[2][kMaxV];
#pragma unroll
    for (int e = 0; e < 2; ++e)
      TEAM_VECS(k, c) ry[e][k] = ldg_raw(static_cast<const T*>(p.yd) + ((long long)e * rows + row) * 2 * C + c, act && e < p.E);
    for (int e = 0; e < p.E; ++e) {
      const avmoe_expert& ex = p.ex[e];
      const T* yh = static_cast<const T*>(p.yd) + ((long long)e * rows + row) * 2 * C;
      float4 t[kMaxV];
      TEAM_VECS(k, c) {
        if (e == 0) t[k] = cvt(ry[0][k]);
        else if (e == 1) t[k] = cvt(ry[1][k]);
        else t[k] = act ? ldg_act4<T>(yh + c) : f4(0.f);
        if (p.bn2_w[e]) t[k] = t[k] * ldg4(p.bn2_w[e] + c) + ldg4(p.bn2_b[e] + c);
      }
      if (ex.ln2_w) {
        float s = 0.f;
        TEAM_VECS(k, c) s += hsum(t[k]);
        const float mean = team_sum<kTeam>(make_float4(s, 0.f, 0.f, 0.f), red, par, tm).x * invC;
        float ss = 0.f;
        TEAM_VECS(k, c) { const float4 d = t[k] - f4(mean); ss += hsum(d * d); }
        const float rstd = rsqrtf(team_sum<kTeam>(make_float4(ss, 0.f, 0.f, 0.f), red, par, tm).x * invC + kLnEps);
        TEAM_VECS(k, c) t[k] = (rstd * (t[k] - f4(mean))) * ldg4(ex.ln2_w + c) + ldg4(ex.ln2_b + c);
        if (act && tm.tlane == 0) {
          float* st = p.ln2_stats + ((long long)e * rows + row) * 2;
          st[0] = mean; st[1] = rstd;
        }
      }
      const float sc = __ldg(p.probs + b * p.E + e) * (ex.gate ? __ldg(ex.gate) : 1.f);
      TEAM_VECS(k, c) acc[k] = acc[k] + sc * t[k];
    }
    if (act) {
      T* dst = static_cast<T*>(p.out) + row * C;
      if (p.res) {   // residual stream folded into the combine epilogue (row a9)
        const T* rs = static_cast<const T*>(p.res) + row * C;
        TEAM_VECS(k, c) acc[k] = acc[k] + ldg_act4<T>(rs + c);
      }
      TEAM_VECS(k, c) st_act4<T>(dst + c, acc[k]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward: combine + ln_post + bn2 sums
template <typename T, int kMaxV, int kTeam>
__global__ void __launch_bounds__(256) xg_post_bwd_kernel(const avmoe_xg_args p, long long rows, int rpb) {
  constexpr int kTeams = kWarps / kTeam;
  const Team<kTeam> tm;
  extern __shared__ float colbuf[];          // [kTeams][C]
  __shared__ float4 red[2][kWarps];
  __shared__ float sred[kWarps];
  int par = 0;
  const int C = p.C;
  const float invC = 1.f / C;
  const long long r0 = blockIdx.x * (long long)rpb, r1 = min(rows, r0 + rpb);
  for (int e = 0; e < p.E; ++e) {
    const avmoe_expert& ex = p.ex[e];
    const bool has_bn = p.bn2_w[e] != nullptr;
    const float gate = ex.gate ? __ldg(ex.gate) : 1.f;
    float4 a_lw[kMaxV], a_lb[kMaxV], a_r0[kMaxV], a_r1[kMaxV];
    TEAM_VECS(k, c) a_lw[k] = a_lb[k] = a_r0[k] = a_r1[k] = f4(0.f);
    float a_gate = 0.f, dp_acc = 0.f;
    long long dp_b = -1;
    const T* dout = static_cast<const T*>(p.dout);
    T* yde = static_cast<T*>(p.yd) + (long long)e * rows * 2 * C;
    Raw4<T> nd[kMaxV], ny[kMaxV];
    float2 nst = make_float2(0.f, 0.f);
    {
      const long long row = r0 + tm.team;
      const bool act = row < r1;
      TEAM_VECS(k, c) { nd[k] = ldg_raw(dout + row * C + c, act); ny[k] = ldg_raw(yde + row * 2 * C + c, act); }
      if (ex.ln2_w && act) nst = __ldg(reinterpret_cast<const float2*>(p.ln2_stats) + (long long)e * rows + row);
    }
    for (long long rb = r0; rb < r1; rb += kTeams) {
      const long long row = rb + tm.team;
      const bool act = row < r1;
      const long long b = act ? row / p.N : 0;
      T* ydr = yde + row * 2 * C;
      float4 d[kMaxV], yh[kMaxV], xh[kMaxV];
      TEAM_VECS(k, c) { d[k] = cvt(nd[k]); yh[k] = cvt(ny[k]); }
      const float mean = nst.x, rstd = nst.y;
      {   // request the next row of this team
        const long long nrow = row + kTeams;
        const bool nact = nrow < r1;
        TEAM_VECS(k, c) { nd[k] = ldg_raw(dout + nrow * C + c, nact); ny[k] = ldg_raw(yde + nrow * 2 * C + c, nact); }
        nst = (ex.ln2_w && nact) ? __ldg(reinterpret_cast<const float2*>(p.ln2_stats) + (long long)e * rows + nrow) : make_float2(0.f, 0.f);
      }
      float dot = 0.f, s1 = 0.f, s2 = 0.f;
      TEAM_VECS(k, c) {
        float4 t = yh[k];
        if (has_bn) t = t * ldg4(p.bn2_w[e] + c) + ldg4(p.bn2_b[e] + c);
        if (ex.ln2_w) {
          xh[k] = rstd * (t - f4(mean));
          const float4 lw = ldg4(ex.ln2_w + c);
          dot += hsum(d[k] * (xh[k] * lw + ldg4(ex.ln2_b + c)));
          const float4 dxh = d[k] * lw;
          s1 += hsum(dxh);
          s2 += hsum(dxh * xh[k]);
        } else {
          xh[k] = t;
          dot += hsum(d[k] * t);
        }
      }
      const float4 sums = team_sum<kTeam>(make_float4(dot, s1, s2, 0.f), red, par, tm);
      const float pg = act ? __ldg(p.probs + b * p.E + e) : 0.f;
      const float sc = pg * gate;
      // d probs[b,e] = gate * <dout, t_e>   (accumulated while the team stays inside one sample)
      if (act && b != dp_b) {
        if (dp_b >= 0 && tm.tlane == 0) atomicAdd(p.dprobs + dp_b * p.E + e, dp_acc);
        dp_b = b; dp_acc = 0.f;
      }
      dp_acc += gate * sums.x;
      a_gate += pg * sums.x;
      const float m1 = sc * sums.y * invC, m2 = sc * sums.z * invC;
      TEAM_VECS(k, c) {
        const float4 dt = sc * d[k];
        float4 dy2;
        if (ex.ln2_w) {
          a_lw[k] = a_lw[k] + dt * xh[k];
          a_lb[k] = a_lb[k] + dt;
          dy2 = rstd * (dt * ldg4(ex.ln2_w + c) - f4(m1) - m2 * xh[k]);
        } else {
          dy2 = dt;
        }
        if (has_bn) { a_r0[k] = a_r0[k] + dy2; a_r1[k] = a_r1[k] + dy2 * yh[k]; }
        if (act) st_act4<T>(ydr + C + c, dy2);
      }
    }
    if (dp_b >= 0 && tm.tlane == 0) atomicAdd(p.dprobs + dp_b * p.E + e, dp_acc);
    if (ex.ln2_w) {
      flush_cols<kMaxV, kTeam, float>(a_lw, colbuf, ex.g_ln2_w, 1, C, tm);
      flush_cols<kMaxV, kTeam, float>(a_lb, colbuf, ex.g_ln2_b, 1, C, tm);
    }
    if (has_bn) {
      flush_cols<kMaxV, kTeam, double>(a_r0, colbuf, p.red2 + (long long)e * C * 2, 2, C, tm);
      flush_cols<kMaxV, kTeam, double>(a_r1, colbuf, p.red2 + (long long)e * C * 2 + 1, 2, C, tm);
    }
    if (ex.gate) {   // every lane of a team holds the same a_gate: one contribution per team
      __syncthreads();
      if (tm.tlane == 0) sred[tm.team] = a_gate;
      __syncthreads();
      if (threadIdx.x == 0) {
        float s = 0.f;
        for (int t = 0; t < kTeams; ++t) s += sred[t];
        atomicAdd(ex.g_gate, s);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward: ln_before + pre-transform
template <typename T, int kMaxV, int kTeam>
__global__ void __launch_bounds__(256) xg_ln_pre_bwd_kernel(const avmoe_xg_args p, long long rows, int rpb) {
  constexpr int kTeams = kWarps / kTeam;
  const Team<kTeam> tm;
  extern __shared__ float colbuf[];
  __shared__ float4 red[2][kWarps];
  __shared__ float sred[kWarps];
  int par = 0;
  const int C = p.C;
  const float invC = 1.f / C;
  const long long r0 = blockIdx.x * (long long)rpb, r1 = min(rows, r0 + rpb);
  for (int e = 0; e < p.E; ++e) {
    const avmoe_expert& ex = p.ex[e];
    const int kind = ex.pre_kind;
    const float gp = kind == AVMOE_PRE_AUX_ADD ? __ldg(ex.gate_pre) : 1.f;
    float4 a_w[kMaxV], a_b[kMaxV];
    TEAM_VECS(k, c) a_w[k] = a_b[k] = f4(0.f);
    float a_gp = 0.f;
    const T* dxl = static_cast<const T*>(p.dxl) + (long long)e * rows * C;
    const T* xin = static_cast<const T*>(p.x);
    const T* auxp = kind == AVMOE_PRE_NONE ? nullptr : static_cast<const T*>(ex.aux);
    T* dxo = static_cast<T*>(p.dx);
    const bool need_x = ex.ln1_w && kind != AVMOE_PRE_AUX_REPLACE;
    const bool need_prev = e > 0 && kind != AVMOE_PRE_AUX_REPLACE;     // dx already holds the earlier experts' sum
    Raw4<T> ng[kMaxV], nx[kMaxV], na[kMaxV], np[kMaxV];
    float2 nst = make_float2(0.f, 0.f);
    {
      const long long row = r0 + tm.team;
      const bool act = row < r1;
      TEAM_VECS(k, c) {
        ng[k] = ldg_raw(dxl + row * C + c, act);
        nx[k] = ldg_raw(xin + row * C + c, act && need_x);
        na[k] = ldg_raw(auxp + row * C + c, act && auxp);
        np[k] = ld_raw(dxo + row * C + c, act && need_prev);
      }
      if (ex.ln1_w && act) nst = __ldg(reinterpret_cast<const float2*>(p.ln1_stats) + (long long)e * rows + row);
    }
    for (long long rb = r0; rb < r1; rb += kTeams) {
      const long long row = rb + tm.team;
      const bool act = row < r1;
      const long long b = act ? row / p.N : 0;
      float4 g[kMaxV], xh[kMaxV], av[kMaxV], prev[kMaxV];
      TEAM_VECS(k, c) { g[k] = cvt(ng[k]); xh[k] = cvt(nx[k]); av[k] = cvt(na[k]); prev[k] = cvt(np[k]); }
      const float mean = nst.x, rstd = nst.y;
      {   // request the next row of this team (its dx was stored by this same thread while processing the previous expert)
        const long long nrow = row + kTeams;
        const bool nact = nrow < r1;
        TEAM_VECS(k, c) {
          ng[k] = ldg_raw(dxl + nrow * C + c, nact);
          nx[k] = ldg_raw(xin + nrow * C + c, nact && need_x);
          na[k] = ldg_raw(auxp + nrow * C + c, nact && auxp);
          np[k] = ld_raw(dxo + nrow * C + c, nact && need_prev);
        }
        nst = (ex.ln1_w && nact) ? __ldg(reinterpret_cast<const float2*>(p.ln1_stats) + (long long)e * rows + nrow) : make_float2(0.f, 0.f);
      }
      if (ex.ln1_w) {
        float s1 = 0.f, s2 = 0.f;
        TEAM_VECS(k, c) {
          float4 xa = xh[k];                                   // x (or unused for aux-replace)
          if (kind == AVMOE_PRE_AUX_REPLACE) xa = av[k];
          else if (kind == AVMOE_PRE_AUX_ADD) xa = xa + gp * av[k];
          xh[k] = rstd * (xa - f4(mean));
          a_w[k] = a_w[k] + g[k] * xh[k];
          a_b[k] = a_b[k] + g[k];
          g[k] = g[k] * ldg4(ex.ln1_w + c);      // d xhat
          s1 += hsum(g[k]);
          s2 += hsum(g[k] * xh[k]);
        }
        const float4 sums = team_sum<kTeam>(make_float4(s1, s2, 0.f, 0.f), red, par, tm);
        const float m1 = sums.x * invC, m2 = sums.y * invC;
        TEAM_VECS(k, c) g[k] = rstd * (g[k] - f4(m1) - m2 * xh[k]);   // d xa
      }
      TEAM_VECS(k, c) {
        if (kind == AVMOE_PRE_AUX_ADD) a_gp += hsum(g[k] * av[k]);
        if (!act) continue;
        if (kind != AVMOE_PRE_NONE) st_act4<T>(static_cast<T*>(ex.g_aux) + row * C + c, kind == AVMOE_PRE_AUX_ADD ? gp * g[k] : g[k]);
        float4 base = e == 0 ? (p.dpool_x ? ldg4(p.dpool_x + b * C + c) : f4(0.f)) : prev[k];
        if (kind != AVMOE_PRE_AUX_REPLACE) base = base + g[k];
        if (e == 0 || kind != AVMOE_PRE_AUX_REPLACE) st_act4<T>(dxo + row * C + c, base);
      }
    }
    if (ex.ln1_w) {
      flush_cols<kMaxV, kTeam, float>(a_w, colbuf, ex.g_ln1_w, 1, C, tm);
      flush_cols<kMaxV, kTeam, float>(a_b, colbuf, ex.g_ln1_b, 1, C, tm);
    }
    if (kind == AVMOE_PRE_AUX_ADD) {
      a_gp = warp_sum(a_gp);
      __syncthreads();
      if (tm.lane == 0) sred[tm.warp] = a_gp;
      __syncthreads();
      if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < kWarps; ++w) s += sred[w];
        atomicAdd(ex.g_gate_pre, s);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ bottleneck-sized passes
// column sums of z [rows][W] (W = E*Cr): thread = (column, row lane); fp32 inside a CTA's chunk, fp64 atomics across CTAs
__global__ void xg_z_stats_kernel(const float* __restrict__ z, double* __restrict__ stats, long long rows, int W, int R, int rpb) {
  extern __shared__ float zbuf[];   // [2][R][W]
  const int col = threadIdx.x % W, rl = threadIdx.x / W;
  const long long r0 = blockIdx.x * (long long)rpb, r1 = min(rows, r0 + rpb);
  float s = 0.f, ss = 0.f;
  for (long long r = r0 + rl; r < r1; r += R) {
    const float v = z[r * W + col];
    s += v;
    ss = fmaf(v, v, ss);
  }
  zbuf[rl * W + col] = s;
  zbuf[(R + rl) * W + col] = ss;
  __syncthreads();
  if (rl == 0) {
    float a = 0.f, b = 0.f;
    for (int j = 0; j < R; ++j) { a += zbuf[j * W + col]; b += zbuf[(R + j) * W + col]; }
    atomicAdd(stats + 2 * col, (double)a);
    atomicAdd(stats + 2 * col + 1, (double)b);
  }
}

// a[e][row][cp] = act(bn1(z[row][e][cp]))  (cp < Cr), zero in the padding columns
template <typename T>
__global__ void xg_bn_act_kernel(const avmoe_xg_args p, long long rows) {
  const int e = blockIdx.y;
  const unsigned total = (unsigned)(rows * p.Crp);
  const bool relu = p.ex[e].relu != 0;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned row = i / p.Crp, cp = i - row * p.Crp;
    float v = 0.f;
    if ((int)cp < p.Cr) {
      v = p.z[((long long)row * p.E + e) * p.Cr + cp];
      if (p.aff1) { const float2 sb = reinterpret_cast<const float2*>(p.aff1)[e * p.Cr + cp]; v = fmaf(v, sb.x, sb.y); }
      if (relu) v = fmaxf(v, 0.f);
    }
    static_cast<T*>(p.a)[(long long)e * total + i] = from_f32<T>(v);
  }
}

// red1[e][c] = (sum dh, sum dh*zhat),  dh = da * relu mask
__global__ void xg_bn1_bwd_reduce_kernel(const avmoe_xg_args p, long long rows, int R, int rpb) {
  extern __shared__ float zbuf[];
  const int e = blockIdx.y, W = p.Cr;
  const int col = threadIdx.x % W, rl = threadIdx.x / W;
  const long long r0 = blockIdx.x * (long long)rpb, r1 = min(rows, r0 + rpb);
  const float2 sb = reinterpret_cast<const float2*>(p.aff1)[e * W + col];
  const float2 mi = reinterpret_cast<const float2*>(p.save1)[e * W + col];
  const bool relu = p.ex[e].relu != 0;
  float s0 = 0.f, s1 = 0.f;
  for (long long r = r0 + rl; r < r1; r += R) {
    const float zz = p.z[(r * p.E + e) * W + col];
    float dh = p.da[((long long)e * rows + r) * p.Crp + col];
    if (relu && fmaf(zz, sb.x, sb.y) <= 0.f) dh = 0.f;
    s0 += dh;
    s1 = fmaf(dh, (zz - mi.x) * mi.y, s1);
  }
  zbuf[rl * W + col] = s0;
  zbuf[(R + rl) * W + col] = s1;
  __syncthreads();
  if (rl == 0) {
    float a = 0.f, b = 0.f;
    for (int j = 0; j < R; ++j) { a += zbuf[j * W + col]; b += zbuf[(R + j) * W + col]; }
    atomicAdd(p.red1 + 2 * (e * W + col), (double)a);
    atomicAdd(p.red1 + 2 * (e * W + col) + 1, (double)b);
  }
}

// dz[e][row][cp] = bn1 backward of dh  (train: batch-statistics form; eval: scale only; off: identity)
template <typename T>
__global__ void xg_bn1_bwd_apply_kernel(const avmoe_xg_args p, long long rows) {
  const int e = blockIdx.y;
  const unsigned total = (unsigned)(rows * p.Crp);
  const bool relu = p.ex[e].relu != 0;
  const float inv_n = 1.f / (float)rows;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned row = i / p.Crp, cp = i - row * p.Crp;
    float v = 0.f;
    if ((int)cp < p.Cr) {
      const int ch = e * p.Cr + cp;
      const float zz = p.z[((long long)row * p.E + e) * p.Cr + cp];
      v = p.da[(long long)e * total + i];
      if (p.bn_mode == AVMOE_BN_OFF) {
        if (relu && zz <= 0.f) v = 0.f;
      } else {
        const float2 sb = reinterpret_cast<const float2*>(p.aff1)[ch];
        if (relu && fmaf(zz, sb.x, sb.y) <= 0.f) v = 0.f;
        if (p.bn_mode == AVMOE_BN_TRAIN) {
          const float2 mi = reinterpret_cast<const float2*>(p.save1)[ch];
          const float m1 = (float)(p.red1[2 * ch] * (double)inv_n), m2 = (float)(p.red1[2 * ch + 1] * (double)inv_n);
          v = sb.x * (v - m1 - (zz - mi.x) * mi.y * m2);
        } else {
          v *= sb.x;
        }
      }
    }
    static_cast<T*>(p.dz)[(long long)e * total + i] = from_f32<T>(v);
  }
}

// 4 bottleneck channels per thread (Cr % 4 == 0, Crp % 4 == 0): 16-byte loads of da and z, one 8/16-byte store of dz
template <typename T>
__global__ void xg_bn1_bwd_apply_vec_kernel(const avmoe_xg_args p, long long rows) {
  const int e = blockIdx.y;
  const unsigned q4 = (unsigned)p.Crp >> 2;
  const unsigned total4 = (unsigned)(rows * q4);
  const bool relu = p.ex[e].relu != 0;
  const float inv_n = 1.f / (float)rows;
  const float* da = p.da + (long long)e * rows * p.Crp;
  T* dz = static_cast<T*>(p.dz) + (long long)e * rows * p.Crp;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
    const unsigned row = i / q4, cp = (i - row * q4) << 2;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if ((int)cp < p.Cr) {
      const float4 zz4 = *reinterpret_cast<const float4*>(p.z + ((long long)row * p.E + e) * p.Cr + cp);
      const float4 d4 = *reinterpret_cast<const float4*>(da + (long long)row * p.Crp + cp);
      const float zz[4] = {zz4.x, zz4.y, zz4.z, zz4.w};
      const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ch = e * p.Cr + cp + j;
        float t = dd[j];
        if (p.bn_mode == AVMOE_BN_OFF) {
          if (relu && zz[j] <= 0.f) t = 0.f;
        } else {
          const float2 sb = reinterpret_cast<const float2*>(p.aff1)[ch];
          if (relu && fmaf(zz[j], sb.x, sb.y) <= 0.f) t = 0.f;
          if (p.bn_mode == AVMOE_BN_TRAIN) {
            const float2 mi = reinterpret_cast<const float2*>(p.save1)[ch];
            const float m1 = (float)(p.red1[2 * ch] * (double)inv_n), m2 = (float)(p.red1[2 * ch + 1] * (double)inv_n);
            t = sb.x * (t - m1 - (zz[j] - mi.x) * mi.y * m2);
          } else {
            t *= sb.x;
          }
        }
        v[j] = t;
      }
    }
    st_act4<T>(dz + (long long)row * p.Crp + cp, make_float4(v[0], v[1], v[2], v[3]));
  }
}

// ------------------------------------------------------------------------------------------------ weight operands / weight gradients
__device__ __forceinline__ float wd_dense(const float* wd, int r, int c, int Cr, int C, int G) {   // wd [Cr][C/G]
  const int Cg = C / G, Crg = Cr / G, g = r / Crg;
  return (c / Cg == g) ? wd[(long long)r * Cg + (c - g * Cg)] : 0.f;
}
__device__ __forceinline__ float wu_dense(const float* wu, int c, int r, int Cr, int C, int G) {   // wu [C][Cr/G]
  const int Cg = C / G, Crg = Cr / G, g = c / Cg;
  return (r / Crg == g) ? wu[(long long)c * Crg + (r - g * Crg)] : 0.f;
}

// wdd [E][Cr][C], wdt [E][C][Crp] in the activation dtype (zero outside the group blocks / padding)
template <typename T>
__global__ void xg_prep_weights_kernel(const avmoe_xg_args p, T* __restrict__ wdd, T* __restrict__ wdt) {
  const int e = blockIdx.y, C = p.C, Cr = p.Cr, Crp = p.Crp;
  const float* wd = p.ex[e].wd;
  const int n1 = Cr * C, n2 = C * Crp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += gridDim.x * blockDim.x) {
    if (i < n1) {
      const int r = i / C, c = i - r * C;
      wdd[(long long)e * n1 + i] = from_f32<T>(wd_dense(wd, r, c, Cr, C, p.G));
    } else {
      const int j = i - n1, c = j / Crp, r = j - c * Crp;
      wdt[(long long)e * n2 + j] = from_f32<T>(r < Cr ? wd_dense(wd, r, c, Cr, C, p.G) : 0.f);
    }
  }
}

// up-projection operand with bn2's normalisation folded in:  yhat = a · wudᵀ + ybi,
//   wud [E][C][Crp] = invstd_c * Wu (dense, padded),  ybi [E][C] = -mean_c * invstd_c     (no BatchNorm: invstd = 1, mean = 0)
template <typename T>
__global__ void xg_prep_up_kernel(const avmoe_xg_args p, T* __restrict__ wud, float* __restrict__ ybi) {
  const int e = blockIdx.y, C = p.C, Cr = p.Cr, Crp = p.Crp;
  const float* wu = p.ex[e].wu;
  const int n2 = C * Crp;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n2; j += gridDim.x * blockDim.x) {
    const int c = j / Crp, r = j - c * Crp;
    float mean = 0.f, invstd = 1.f;
    if (p.bn_mode != AVMOE_BN_OFF) { const float2 mi = reinterpret_cast<const float2*>(p.save2)[e * C + c]; mean = mi.x; invstd = mi.y; }
    wud[(long long)e * n2 + j] = from_f32<T>(r < Cr ? invstd * wu_dense(wu, c, r, Cr, C, p.G) : 0.f);
    if (r == 0) ybi[e * C + c] = -mean * invstd;
  }
}

// bn2 backward folded into the operand of  da = [yhat | dy2] · W2ᵀ + bias :
//   dy = alpha*dy2 + beta*yhat + gamma,  alpha = invstd*w2, beta = -alpha*mean(dy2*yhat), gamma = -alpha*mean(dy2)   (train)
__device__ __forceinline__ void bn2_bwd_coeffs(const avmoe_xg_args& p, int e, int c, float inv_n, float& alpha, float& beta, float& gamma) {
  alpha = 1.f; beta = 0.f; gamma = 0.f;
  if (p.bn_mode == AVMOE_BN_OFF) return;
  const float invstd = p.save2[2 * (e * p.C + c) + 1];
  alpha = invstd * __ldg(p.bn2_w[e] + c);
  if (p.bn_mode == AVMOE_BN_TRAIN) {
    const double* r = p.red2 + 2 * ((long long)e * p.C + c);
    gamma = -alpha * (float)(r[0] * (double)inv_n);
    beta = -alpha * (float)(r[1] * (double)inv_n);
  }
}

// w2 [E][Crp][2C] (activation dtype), bias [E][Crp] fp32; one CTA per (r, e)
template <typename T>
__global__ void xg_prep_bwd_kernel(const avmoe_xg_args p, long long rows, T* __restrict__ w2, float* __restrict__ bias) {
  const int e = blockIdx.y, r = blockIdx.x, C = p.C;
  const float inv_n = 1.f / (float)rows;
  __shared__ float sred[32];
  float bsum = 0.f;
  T* dst = w2 + ((long long)e * p.Crp + r) * 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float w = 0.f, alpha, beta, gamma;
    bn2_bwd_coeffs(p, e, c, inv_n, alpha, beta, gamma);
    if (r < p.Cr) w = wu_dense(p.ex[e].wu, c, r, p.Cr, C, p.G);
    dst[c] = from_f32<T>(beta * w);
    dst[C + c] = from_f32<T>(alpha * w);
    bsum = fmaf(gamma, w, bsum);
  }
  bsum = warp_sum(bsum);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = bsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sred[w];
    bias[e * p.Crp + r] = s;
  }
}

// pwd [E][Crp][C] = dzᵀ·xl,  pwu [E][2C][Crp] = [yhat|dy2]ᵀ·a,  suma [E][Cr] = column sums of a (fp64, bn train only)
//   g_wd[r][ci] = pwd[r][g*Cg + ci] ;  g_wu[c][ri] = alpha_c*pwu[C+c][r] + beta_c*pwu[c][r] + gamma_c*suma[r],  r = g*Crg + ri
__global__ void xg_finish_wgrads_kernel(const avmoe_xg_args p, long long rows, const float* __restrict__ pwd, const float* __restrict__ pwu,
                                        const double* __restrict__ suma) {
  const int e = blockIdx.y, C = p.C, Cr = p.Cr, Crp = p.Crp, G = p.G;
  const int Cg = C / G, Crg = Cr / G;
  const float inv_n = 1.f / (float)rows;
  const int n1 = Cr * Cg, n2 = C * Crg;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += gridDim.x * blockDim.x) {
    if (i < n1) {
      const int r = i / Cg, ci = i - r * Cg, g = r / Crg;
      p.ex[e].g_wd[i] = pwd[((long long)e * Crp + r) * C + g * Cg + ci];
    } else {
      const int j = i - n1, c = j / Crg, ri = j - c * Crg, r = (c / Cg) * Crg + ri;
      float alpha, beta, gamma;
      bn2_bwd_coeffs(p, e, c, inv_n, alpha, beta, gamma);
      const float* pe = pwu + (long long)e * 2 * C * Crp;
      float v = alpha * pe[((long long)C + c) * Crp + r];
      if (p.bn_mode == AVMOE_BN_TRAIN) v += beta * pe[(long long)c * Crp + r] + gamma * (float)suma[e * Cr + r];
      p.ex[e].g_wu[j] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
static int xg_check(const avmoe_xg_args* a, const char* who) {
  AVMOE_CHECK_ARG(a, "%s: null args", who);
  AVMOE_CHECK_ARG(a->BT > 0 && a->N > 0 && a->C > 0 && a->C % 4 == 0 && a->C <= 1536, "%s: C must be a multiple of 4, <= 1536", who);
  AVMOE_CHECK_ARG(a->E >= 1 && a->E <= AVMOE_MAX_EXPERTS && a->G >= 1 && a->Cr > 0 && a->Cr % a->G == 0 && a->C % a->G == 0 && a->Crp >= a->Cr &&
                      a->Crp % 8 == 0,
                  "%s: bad expert geometry", who);
  AVMOE_CHECK_ARG(a->dtype == AVMOE_F32 || a->dtype == AVMOE_BF16, "%s: bad dtype", who);
  AVMOE_CHECK_ARG((long long)a->BT * a->N * a->Crp < (1ll << 31), "%s: bottleneck tensor too large for 32-bit indexing", who);
  for (int e = 0; e < a->E; ++e)
    AVMOE_CHECK_ARG(a->ex[e].pre_kind != AVMOE_PRE_LATENT, "%s: the GEMM expert path takes latent attention as an aux-add (pre_kind 2)", who);
  return 0;
}

// rows per CTA so that the grid is ~4 CTAs per SM (column sums are flushed once per CTA), a multiple of the teams per CTA
static int rows_per_cta(long long rows, int teams, bool has_col_sums) {
  long long rpb = has_col_sums ? (rows + 4 * num_sms() - 1) / (4 * num_sms()) : 4 * teams;
  rpb = (rpb + teams - 1) / teams * teams;
  return (int)(rpb < teams ? teams : rpb);
}

// team width / vectors per lane for a channel count: C <= 384 -> 1 warp, <= 768 -> 2 warps, else 4 warps per row
#define XG_DISPATCH(C_, T_, KERNEL, grid, smem, st, ...)                                                         \
  do {                                                                                                           \
    const int _c = (C_);                                                                                         \
    if (_c <= 128) KERNEL<T_, 1, 1><<<grid, 256, smem, st>>>(__VA_ARGS__);                                       \
    else if (_c <= 256) KERNEL<T_, 2, 1><<<grid, 256, smem, st>>>(__VA_ARGS__);                                  \
    else if (_c <= 384) KERNEL<T_, 3, 1><<<grid, 256, smem, st>>>(__VA_ARGS__);                                  \
    else if (_c <= 512) KERNEL<T_, 2, 2><<<grid, 256, smem, st>>>(__VA_ARGS__);                                  \
    else if (_c <= 768) KERNEL<T_, 3, 2><<<grid, 256, smem, st>>>(__VA_ARGS__);                                  \
    else if (_c <= 1024) KERNEL<T_, 2, 4><<<grid, 256, smem, st>>>(__VA_ARGS__);                                 \
    else KERNEL<T_, 3, 4><<<grid, 256, smem, st>>>(__VA_ARGS__);                                                 \
  } while (0)

static int teams_for(int C) { return C <= 384 ? 8 : C <= 768 ? 4 : 2; }

}  // namespace avmoe

using namespace avmoe;

#define XG_ROW_ENTRY(NAME, KERNEL, COLSUMS, CHECKS)                                                              \
  extern "C" int NAME(const avmoe_xg_args* a, avmoe_stream_t stream) {                                           \
    if (int rc = xg_check(a, #NAME)) return rc;                                                                  \
    CHECKS                                                                                                       \
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);                                                    \
    const long long rows = (long long)a->BT * a->N;                                                              \
    const int teams = teams_for(a->C);                                                                           \
    const int rpb = rows_per_cta(rows, teams, COLSUMS);                                                          \
    const unsigned grid = (unsigned)((rows + rpb - 1) / rpb);                                                    \
    const size_t smem = COLSUMS ? sizeof(float) * (size_t)teams * a->C : 0;                                      \
    if (a->dtype == AVMOE_F32) XG_DISPATCH(a->C, float, KERNEL, grid, smem, st, *a, rows, rpb);                  \
    else XG_DISPATCH(a->C, __nv_bfloat16, KERNEL, grid, smem, st, *a, rows, rpb);                                \
    AVMOE_LAUNCH_CHECK(#KERNEL);                                                                                 \
    return 0;                                                                                                    \
  }

XG_ROW_ENTRY(avmoe_xg_ln_pre_fwd, xg_ln_pre_fwd_kernel, false, AVMOE_CHECK_ARG(a->x && a->xl, "xg_ln_pre_fwd: null tensor");)
XG_ROW_ENTRY(avmoe_xg_post_fwd, xg_post_fwd_kernel, false, AVMOE_CHECK_ARG(a->yd && a->probs && a->out, "xg_post_fwd: null tensor");)
XG_ROW_ENTRY(avmoe_xg_post_bwd, xg_post_bwd_kernel, true,
             AVMOE_CHECK_ARG(a->yd && a->probs && a->dout && a->dprobs && (a->bn_mode == AVMOE_BN_OFF || a->red2), "xg_post_bwd: null tensor");)
XG_ROW_ENTRY(avmoe_xg_ln_pre_bwd, xg_ln_pre_bwd_kernel, true, AVMOE_CHECK_ARG(a->x && a->dxl && a->dx, "xg_ln_pre_bwd: null tensor");)

static int col_geometry(int W, long long rows, int* R, int* rpb, unsigned* grid) {
  if (W > 1024) return 1;
  *R = W >= 256 ? 1 : 256 / W;
  long long per = (rows + 4 * num_sms() - 1) / (4 * num_sms());
  if (per < *R) per = *R;
  *rpb = (int)per;
  *grid = (unsigned)((rows + per - 1) / per);
  return 0;
}

extern "C" int avmoe_xg_z_stats(const avmoe_xg_args* a, avmoe_stream_t stream) {
  if (int rc = xg_check(a, "xg_z_stats")) return rc;
  AVMOE_CHECK_ARG(a->z && a->stats1, "xg_z_stats: null tensor");
  const long long rows = (long long)a->BT * a->N;
  const int W = a->E * a->Cr;
  int R, rpb; unsigned grid;
  AVMOE_CHECK_ARG(col_geometry(W, rows, &R, &rpb, &grid) == 0, "xg_z_stats: E*Cr must be <= 1024");
  xg_z_stats_kernel<<<grid, W * R, sizeof(float) * 2 * R * W, reinterpret_cast<cudaStream_t>(stream)>>>(a->z, a->stats1, rows, W, R, rpb);
  AVMOE_LAUNCH_CHECK("xg_z_stats_kernel");
  return 0;
}

extern "C" int avmoe_xg_bn_act(const avmoe_xg_args* a, avmoe_stream_t stream) {
  if (int rc = xg_check(a, "xg_bn_act")) return rc;
  AVMOE_CHECK_ARG(a->z && a->a && (a->bn_mode == AVMOE_BN_OFF || a->aff1), "xg_bn_act: null tensor");
  const long long rows = (long long)a->BT * a->N;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)min((long long)(8 * num_sms()), (rows * a->Crp + 255) / 256), a->E);
  if (a->dtype == AVMOE_F32) xg_bn_act_kernel<float><<<grid, 256, 0, st>>>(*a, rows);
  else xg_bn_act_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(*a, rows);
  AVMOE_LAUNCH_CHECK("xg_bn_act_kernel");
  return 0;
}

extern "C" int avmoe_xg_bn1_bwd(const avmoe_xg_args* a, avmoe_stream_t stream) {
  if (int rc = xg_check(a, "xg_bn1_bwd")) return rc;
  AVMOE_CHECK_ARG(a->z && a->da && a->dz, "xg_bn1_bwd: null tensor");
  const long long rows = (long long)a->BT * a->N;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->bn_mode != AVMOE_BN_OFF) {
    AVMOE_CHECK_ARG(a->aff1 && a->save1 && a->red1, "xg_bn1_bwd: null bn tensors");
    int R, rpb; unsigned gx;
    AVMOE_CHECK_ARG(col_geometry(a->Cr, rows, &R, &rpb, &gx) == 0, "xg_bn1_bwd: Cr must be <= 1024");
    xg_bn1_bwd_reduce_kernel<<<dim3(gx, a->E), a->Cr * R, sizeof(float) * 2 * R * a->Cr, st>>>(*a, rows, R, rpb);
    AVMOE_LAUNCH_CHECK("xg_bn1_bwd_reduce_kernel");
  }
  dim3 grid((unsigned)min((long long)(8 * num_sms()), (rows * a->Crp + 255) / 256), a->E);
  if (a->Cr % 4 == 0 && a->Crp % 4 == 0) {
    if (a->dtype == AVMOE_F32) xg_bn1_bwd_apply_vec_kernel<float><<<grid, 256, 0, st>>>(*a, rows);
    else xg_bn1_bwd_apply_vec_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(*a, rows);
  } else if (a->dtype == AVMOE_F32) xg_bn1_bwd_apply_kernel<float><<<grid, 256, 0, st>>>(*a, rows);
  else xg_bn1_bwd_apply_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(*a, rows);
  AVMOE_LAUNCH_CHECK("xg_bn1_bwd_apply_kernel");
  return 0;
}

extern "C" int avmoe_xg_prep_weights(const avmoe_xg_args* a, void* wdd, void* wdt, avmoe_stream_t stream) {
  if (int rc = xg_check(a, "xg_prep_weights")) return rc;
  AVMOE_CHECK_ARG(wdd && wdt, "xg_prep_weights: null tensor");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int n = a->Cr * a->C + a->C * a->Crp;
  dim3 grid((n + 255) / 256, a->E);
  if (a->dtype == AVMOE_F32) xg_prep_weights_kernel<float><<<grid, 256, 0, st>>>(*a, (float*)wdd, (float*)wdt);
  else xg_prep_weights_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(*a, (__nv_bfloat16*)wdd, (__nv_bfloat16*)wdt);
  AVMOE_LAUNCH_CHECK("xg_prep_weights_kernel");
  return 0;
}

extern "C" int avmoe_xg_prep_up(const avmoe_xg_args* a, void* wud, float* ybi, avmoe_stream_t stream) {
  if (int rc = xg_check(a, "xg_prep_up")) return rc;
  AVMOE_CHECK_ARG(wud && ybi && (a->bn_mode == AVMOE_BN_OFF || a->save2), "xg_prep_up: null tensor");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((a->C * a->Crp + 255) / 256, a->E);
  if (a->dtype == AVMOE_F32) xg_prep_up_kernel<float><<<grid, 256, 0, st>>>(*a, (float*)wud, ybi);
  else xg_prep_up_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(*a, (__nv_bfloat16*)wud, ybi);
  AVMOE_LAUNCH_CHECK("xg_prep_up_kernel");
  return 0;
}

extern "C" int avmoe_xg_prep_bwd(const avmoe_xg_args* a, void* w2, float* bias, avmoe_stream_t stream) {
  if (int rc = xg_check(a, "xg_prep_bwd")) return rc;
  AVMOE_CHECK_ARG(w2 && bias && (a->bn_mode == AVMOE_BN_OFF || a->save2) && (a->bn_mode != AVMOE_BN_TRAIN || a->red2), "xg_prep_bwd: null tensor");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long rows = (long long)a->BT * a->N;
  dim3 grid(a->Crp, a->E);
  if (a->dtype == AVMOE_F32) xg_prep_bwd_kernel<float><<<grid, 256, 0, st>>>(*a, rows, (float*)w2, bias);
  else xg_prep_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(*a, rows, (__nv_bfloat16*)w2, bias);
  AVMOE_LAUNCH_CHECK("xg_prep_bwd_kernel");
  return 0;
}

extern "C" int avmoe_xg_finish_wgrads(const avmoe_xg_args* a, const float* pwd, const float* pwu, const double* suma, avmoe_stream_t stream) {
  if (int rc = xg_check(a, "xg_finish_wgrads")) return rc;
  AVMOE_CHECK_ARG(pwd && pwu && (a->bn_mode != AVMOE_BN_TRAIN || suma), "xg_finish_wgrads: null tensor");
  const long long rows = (long long)a->BT * a->N;
  const int n = a->Cr * (a->C / a->G) + a->C * (a->Cr / a->G);
  xg_finish_wgrads_kernel<<<dim3((n + 255) / 256, a->E), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*a, rows, pwd, pwu, suma);
  AVMOE_LAUNCH_CHECK("xg_finish_wgrads_kernel");
  return 0;
}
