// Op-level entry points of the token resampler  vis' = fc(conv_adapter(other))  and its backward
// (AVMOE/AVE/nets/net_trans_v3.py:445-446,469-471): the whole launch sequence -- weight casts, layout transposes, tcgen05 GEMMs,
// bias reductions -- is issued from C++ against a caller-provided workspace, one C-ABI call per direction.
//
// The sequence is the one avmoe_b200/functional.py:resample_fwd / resample_bwd spells out op by op (that Python form stays: it is
// what the CPU emulation executes and what the profiling mode of bench.py times GEMM by GEMM; tests check both give the same bits):
//
//   Cx > C  ("fc first", fewer FLOPs):   P  = o · W_fcᵀ                         [BT,Nx,C]     (kept for the backward)
//                                         v  = W_tok · P + b_tok ⊗ s + 1 ⊗ b_fc   s = rowsum(W_fc)
//   else     ("tokens first"):            U  = W_tok · o + b_tok ⊗ 1              [BT,N,Cx]     (kept)
//                                         v  = U · W_fcᵀ + 1 ⊗ b_fc
//
// The token GEMM always runs K-major x K-major: its activation operand is re-laid out channel-major by a transpose whose row
// pitch is the token count rounded up to 16 bytes (zero columns), so that the 36-token stage-4 stream is a valid TMA tensor map.
#include "common.cuh"

namespace avmoe {

// out[r][0..cols) = in[r][0..cols) converted, out[r][cols..ld) = 0
template <typename T>
__global__ void cast_rows_kernel(const float* __restrict__ in, T* __restrict__ out, long long rows, int cols, int ld) {
  const long long total = rows * ld;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ld;
    const int c = (int)(i - r * ld);
    out[i] = from_f32<T>(c < cols ? in[r * cols + c] : 0.f);
  }
}

namespace {

struct Arena {   // bump allocator over the caller's workspace (256-byte granules); base == nullptr only measures
  char* base;
  long long used = 0;
  explicit Arena(void* b) : base(static_cast<char*>(b)) {}
  void* take(long long bytes) {
    void* p = base ? base + used : nullptr;
    used += (bytes + 255) & ~255ll;
    return p;
  }
};

inline int es_of(int dtype) { return dtype == AVMOE_BF16 ? 2 : 4; }
inline long long pad_k(long long k, int dtype) {
  const long long m = 16 / es_of(dtype);
  return (k + m - 1) / m * m;
}
// split-K factor for the skinny weight-gradient GEMMs ([m,n] output, k_rows reduction): enough CTAs for two per SM
inline int split_k_hint(long long k_rows, long long m, long long n) {
  const long long tiles = ((m + 127) / 128) * ((n + 255) / 256);
  long long want = (2 * 148) / tiles;
  if (want < 1) want = 1;
  long long cap = k_rows / 256;
  if (cap < 1) cap = 1;
  return (int)(want < cap ? want : cap);
}

struct Seq {   // launch sequence state: stream, error code, launch count
  cudaStream_t st;
  int impl;
  int dtype;
  int launches = 0;
  int rc = 0;

  void gemm(avmoe_gemm_args& g) {
    if (rc) return;
    g.dtype = dtype;
    g.impl = impl;
    if (g.alpha == 0.f) g.alpha = 1.f;
    if (g.a_batch < 1) g.a_batch = 1;
    if (g.b_batch < 1) g.b_batch = 1;
    if (g.split_k > 1) ++launches;   // the zero-fill of the split-K output
    rc = avmoe_gemm(&g, st);
    ++launches;
  }
  // fp32 master [rows][cols] -> activation dtype with row pitch ld (zero pad); returns the operand to use
  const void* weight(const float* w, long long rows, int cols, int ld, void* buf) {
    if (rc) return buf;
    if (dtype == AVMOE_F32 && ld == cols) return w;
    if (ld == cols) {   // dense: the vectorised flat cast
      rc = avmoe_cast(w, buf, rows * cols, AVMOE_F32, dtype, st);
      ++launches;
      return buf;
    }
    const long long total = rows * ld;
    const int grid = (int)((total + 255) / 256 < 4LL * num_sms() ? (total + 255) / 256 : 4LL * num_sms());
    if (dtype == AVMOE_F32) cast_rows_kernel<float><<<grid, 256, 0, st>>>(w, static_cast<float*>(buf), rows, cols, ld);
    else cast_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(w, static_cast<__nv_bfloat16*>(buf), rows, cols, ld);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("launch of cast_rows_kernel failed: %s", cudaGetErrorString(e)); rc = (int)e; }
    ++launches;
    return buf;
  }
  // out[b][c][0..R) = in[b][r][c], row pitch ld >= R with zero pad columns
  void transpose(const void* in, void* out, int batch, int R, int C, long long ld, int in_dtype) {
    if (rc) return;
    if (ld != R) {
      cudaError_t e = cudaMemsetAsync(out, 0, (size_t)batch * C * ld * es_of(dtype), st);
      if (e != cudaSuccess) { set_error("cudaMemsetAsync failed: %s", cudaGetErrorString(e)); rc = (int)e; return; }
      ++launches;
    }
    rc = avmoe_transpose(in, out, batch, R, C, C, (int64_t)R * C, ld, (int64_t)C * ld, in_dtype, dtype, st);
    ++launches;
  }
  void reduce_bias(const void* d, const float* colw, const float* roww, float* rows_w, float* cols, float* cols_w, int BT, int N, int C) {
    if (rc) return;
    rc = avmoe_reduce_bias(d, colw, roww, rows_w, cols, cols_w, BT, N, C, dtype, st);
    launches += 1 + (rows_w != nullptr) + (cols != nullptr) + (cols_w != nullptr);
  }
};

int check(const avmoe_resample_args* a, const char* who) {
  AVMOE_CHECK_ARG(a, "%s: null args", who);
  AVMOE_CHECK_ARG(a->dtype == AVMOE_F32 || a->dtype == AVMOE_BF16, "%s: bad dtype", who);
  AVMOE_CHECK_ARG(a->BT > 0 && a->N > 0 && a->Nx > 0 && a->C > 0 && a->Cx > 0, "%s: bad sizes", who);
  AVMOE_CHECK_ARG(a->o && a->w_tok && a->b_tok && a->w_fc && a->mid, "%s: null tensor", who);
  AVMOE_CHECK_ARG(a->Cx <= a->C || a->s, "%s: the fc-first association (Cx > C) needs s", who);
  return 0;
}

// one pass over the sequence: with a null arena base it only sizes the workspace
long long plan_or_run_fwd(avmoe_resample_args* a, Seq* q, void* ws) {
  const int dt = a->dtype, es = es_of(dt);
  const long long BT = a->BT, N = a->N, Nx = a->Nx, C = a->C, Cx = a->Cx;
  const long long Kp = pad_k(Nx, dt);
  Arena ar(ws);
  void* wt_buf = ar.take(N * Kp * es);
  void* wf_buf = ar.take(C * Cx * es);
  const bool fc_first = Cx > C;
  void* tbuf = ar.take(BT * (fc_first ? C : Cx) * Kp * es);       // channel-major activation operand of the token GEMM
  if (!q) return ar.used;
  const void* wt = q->weight(a->w_tok, N, (int)Nx, (int)Kp, wt_buf);
  const void* wf = q->weight(a->w_fc, C, (int)Cx, (int)Cx, wf_buf);
  if (fc_first) {
    avmoe_gemm_args g = {};
    g.A = a->o; g.B = wf; g.D = a->mid; g.d_dtype = dt;                      // P = o · W_fcᵀ
    g.M = BT * Nx; g.N = C; g.K = Cx; g.a_kmajor = g.b_kmajor = 1; g.lda = Cx; g.ldb = Cx; g.ldd = C;
    q->gemm(g);
    q->transpose(a->mid, tbuf, (int)BT, (int)Nx, (int)C, Kp, dt);            // Pt [BT,C,Kp]
    if (!q->rc) { q->rc = avmoe_rowsum(a->w_fc, a->s, (int)C, (int)Cx, q->st); ++q->launches; }
    avmoe_gemm_args h = {};
    h.A = wt; h.B = tbuf; h.D = a->v; h.d_dtype = dt;                        // v = W_tok · P + b_tok ⊗ s + b_fc
    h.M = N; h.N = BT * C; h.K = Kp; h.a_kmajor = h.b_kmajor = 1; h.lda = Kp; h.ldb = Kp; h.Nb = C; h.ldd = C; h.d_bstride = N * C;
    h.row_bias = a->b_tok; h.col_scale = a->s; h.col_bias = a->b_fc;
    q->gemm(h);
  } else {
    q->transpose(a->o, tbuf, (int)BT, (int)Nx, (int)Cx, Kp, dt);             // Ot [BT,Cx,Kp]
    avmoe_gemm_args g = {};
    g.A = wt; g.B = tbuf; g.D = a->mid; g.d_dtype = dt;                      // U = W_tok · o + b_tok
    g.M = N; g.N = BT * Cx; g.K = Kp; g.a_kmajor = g.b_kmajor = 1; g.lda = Kp; g.ldb = Kp; g.Nb = Cx; g.ldd = Cx; g.d_bstride = N * Cx;
    g.row_bias = a->b_tok;
    q->gemm(g);
    avmoe_gemm_args h = {};
    h.A = a->mid; h.B = wf; h.D = a->v; h.d_dtype = dt;                      // v = U · W_fcᵀ + b_fc
    h.M = BT * N; h.N = C; h.K = Cx; h.a_kmajor = h.b_kmajor = 1; h.lda = Cx; h.ldb = Cx; h.ldd = C;
    h.col_bias = a->b_fc;
    q->gemm(h);
  }
  return ar.used;
}

long long plan_or_run_bwd(avmoe_resample_args* a, Seq* q, void* ws) {
  const int dt = a->dtype, es = es_of(dt);
  const long long BT = a->BT, N = a->N, Nx = a->Nx, C = a->C, Cx = a->Cx;
  const long long Np = pad_k(N, dt);
  const bool fc_first = Cx > C;
  Arena ar(ws);
  void* wtT = ar.take(Nx * Np * es);                                         // W_tokᵀ [Nx,Np]
  void* wfT = ar.take(Cx * C * es);                                          // W_fcᵀ  [Cx,C]
  float* ds = static_cast<float*>(ar.take(C * 4));
  void* tbuf = ar.take(BT * (fc_first ? C : Cx) * Np * es);                  // dVt [BT,C,Np] / dUt [BT,Cx,Np]
  void* dmid = ar.take(fc_first ? BT * Nx * C * es : BT * N * Cx * es);      // dP [BT,Nx,C] / dU [BT,N,Cx]
  if (!q) return ar.used;
  q->transpose(a->w_tok, wtT, 1, (int)N, (int)Nx, Np, AVMOE_F32);
  q->transpose(a->w_fc, wfT, 1, (int)C, (int)Cx, C, AVMOE_F32);
  if (fc_first) {
    // db_fc[c] = sum dv ; db_tok[n] = sum_c dv * s[c] ; ds[c] = sum_n dv * b_tok[n]   (gradient of rowsum(W_fc))
    q->reduce_bias(a->dv, a->s, a->b_tok, a->db_tok, a->db_fc, ds, (int)BT, (int)N, (int)C);
    avmoe_gemm_args g = {};
    g.A = a->dv; g.B = a->mid; g.D = a->dw_tok; g.d_dtype = AVMOE_F32;       // dW_tok = sum_b dv[b] · P[b]ᵀ
    g.M = N; g.N = Nx; g.K = C; g.a_kmajor = g.b_kmajor = 1; g.lda = C; g.ldb = C; g.ldd = Nx;
    g.a_batch = g.b_batch = (int)BT; g.a_bstride = N * C; g.b_bstride = Nx * C; g.Nb = Nx;
    g.split_k = split_k_hint(BT * C, N, Nx);
    q->gemm(g);
    q->transpose(a->dv, tbuf, (int)BT, (int)N, (int)C, Np, dt);              // dVt
    avmoe_gemm_args h = {};
    h.A = wtT; h.B = tbuf; h.D = dmid; h.d_dtype = dt;                       // dP = W_tokᵀ · dv[b]
    h.M = Nx; h.N = BT * C; h.K = Np; h.a_kmajor = h.b_kmajor = 1; h.lda = Np; h.ldb = Np; h.Nb = C; h.ldd = C; h.d_bstride = Nx * C;
    q->gemm(h);
    avmoe_gemm_args w = {};
    w.A = dmid; w.B = a->o; w.D = a->dw_fc; w.d_dtype = AVMOE_F32;           // dW_fc = dPᵀ · o + ds ⊗ 1
    w.M = C; w.N = Cx; w.K = BT * Nx; w.a_kmajor = w.b_kmajor = 0; w.lda = C; w.ldb = Cx; w.ldd = Cx; w.Nb = Cx;
    w.split_k = split_k_hint(BT * Nx, C, Cx); w.row_bias = ds;
    q->gemm(w);
    if (a->d_o) {
      avmoe_gemm_args d = {};
      d.A = dmid; d.B = wfT; d.D = a->d_o; d.d_dtype = dt;                   // d other = dP · W_fc
      d.M = BT * Nx; d.N = Cx; d.K = C; d.a_kmajor = d.b_kmajor = 1; d.lda = C; d.ldb = C; d.ldd = Cx;
      q->gemm(d);
    }
  } else {
    q->reduce_bias(a->dv, nullptr, nullptr, nullptr, a->db_fc, nullptr, (int)BT, (int)N, (int)C);
    avmoe_gemm_args w = {};
    w.A = a->dv; w.B = a->mid; w.D = a->dw_fc; w.d_dtype = AVMOE_F32;        // dW_fc = dvᵀ · U
    w.M = C; w.N = Cx; w.K = BT * N; w.a_kmajor = w.b_kmajor = 0; w.lda = C; w.ldb = Cx; w.ldd = Cx; w.Nb = Cx;
    w.split_k = split_k_hint(BT * N, C, Cx);
    q->gemm(w);
    avmoe_gemm_args u = {};
    u.A = a->dv; u.B = wfT; u.D = dmid; u.d_dtype = dt;                      // dU = dv · W_fc
    u.M = BT * N; u.N = Cx; u.K = C; u.a_kmajor = u.b_kmajor = 1; u.lda = C; u.ldb = C; u.ldd = Cx;
    q->gemm(u);
    q->reduce_bias(dmid, nullptr, nullptr, a->db_tok, nullptr, nullptr, (int)BT, (int)N, (int)Cx);
    avmoe_gemm_args g = {};
    g.A = dmid; g.B = a->o; g.D = a->dw_tok; g.d_dtype = AVMOE_F32;          // dW_tok = sum_b dU[b] · o[b]ᵀ
    g.M = N; g.N = Nx; g.K = Cx; g.a_kmajor = g.b_kmajor = 1; g.lda = Cx; g.ldb = Cx; g.ldd = Nx;
    g.a_batch = g.b_batch = (int)BT; g.a_bstride = N * Cx; g.b_bstride = Nx * Cx; g.Nb = Nx;
    g.split_k = split_k_hint(BT * Cx, N, Nx);
    q->gemm(g);
    if (a->d_o) {
      q->transpose(dmid, tbuf, (int)BT, (int)N, (int)Cx, Np, dt);            // dUt
      avmoe_gemm_args d = {};
      d.A = wtT; d.B = tbuf; d.D = a->d_o; d.d_dtype = dt;                   // d other = W_tokᵀ · dU[b]
      d.M = Nx; d.N = BT * Cx; d.K = Np; d.a_kmajor = d.b_kmajor = 1; d.lda = Np; d.ldb = Np; d.Nb = Cx; d.ldd = Cx; d.d_bstride = Nx * Cx;
      q->gemm(d);
    }
  }
  return ar.used;
}

}  // namespace
}  // namespace avmoe

using namespace avmoe;

extern "C" int64_t avmoe_resample_workspace_bytes(const avmoe_resample_args* a, int backward) {
  if (!a || a->BT <= 0 || a->N <= 0 || a->Nx <= 0 || a->C <= 0 || a->Cx <= 0) return -1;
  avmoe_resample_args tmp = *a;
  return backward ? plan_or_run_bwd(&tmp, nullptr, nullptr) : plan_or_run_fwd(&tmp, nullptr, nullptr);
}

extern "C" int avmoe_resample_fwd(avmoe_resample_args* a, avmoe_stream_t stream) {
  if (int rc = check(a, "resample_fwd")) return rc;
  AVMOE_CHECK_ARG(a->v && a->b_fc, "resample_fwd: null tensor");
  AVMOE_CHECK_ARG(a->workspace && a->workspace_bytes >= avmoe_resample_workspace_bytes(a, 0) && ((uintptr_t)a->workspace % 256) == 0,
                  "resample_fwd: workspace too small or not 256-byte aligned (need %lld bytes)", (long long)avmoe_resample_workspace_bytes(a, 0));
  Seq q{reinterpret_cast<cudaStream_t>(stream), a->gemm_impl, a->dtype};
  plan_or_run_fwd(a, &q, a->workspace);
  a->launches = q.launches;
  return q.rc;
}

extern "C" int avmoe_resample_bwd(avmoe_resample_args* a, avmoe_stream_t stream) {
  if (int rc = check(a, "resample_bwd")) return rc;
  AVMOE_CHECK_ARG(a->dv && a->dw_tok && a->db_tok && a->dw_fc && a->db_fc, "resample_bwd: null tensor");
  AVMOE_CHECK_ARG(a->workspace && a->workspace_bytes >= avmoe_resample_workspace_bytes(a, 1) && ((uintptr_t)a->workspace % 256) == 0,
                  "resample_bwd: workspace too small or not 256-byte aligned (need %lld bytes)", (long long)avmoe_resample_workspace_bytes(a, 1));
  Seq q{reinterpret_cast<cudaStream_t>(stream), a->gemm_impl, a->dtype};
  plan_or_run_bwd(a, &q, a->workspace);
  a->launches = q.launches;
  return q.rc;
}
