/*
 * avmoe_b200 — C ABI of the B200-native AVMoE adapter hot path.
 *
 * Every entry point borrows raw DEVICE pointers owned by the caller (torch's caching
 * allocator in the Python host layer) for the duration of one call, enqueues its kernels
 * on the given cudaStream_t of the CURRENT device and returns immediately:
 *   0  = enqueued,  >0 = cudaError_t,  <0 = AVMOE_E_* (bad argument / unsupported shape).
 * Nothing is allocated that the caller must free; workspaces are passed in.  The library
 * keeps no mutable global state besides lazily-resolved driver entry points and per-function
 * attribute flags, so it is re-entrant across host threads (reference DataParallel runs one
 * thread per GPU: AVMOE/AVVP/main.py:420-421).  There is NO CPU fallback.
 *
 * Layout convention: all activations are TOKEN-MAJOR  [BT, N, C]  (C contiguous).  The
 * reference passes x as the permuted view f.permute(0,2,1).unsqueeze(-1) of exactly such a
 * tensor (AVMOE/AVE/nets/net_trans_v3.py:695-697), so no transposes happen anywhere.
 *
 * dtype codes: 0 = float32 (tensor-core math in TF32), 1 = bfloat16.  Parameters, statistics,
 * router and all accumulation are always float32.
 *
 * Reference interface each group replaces (AVE copy; the other four variants cite the same
 * lines of their own files):
 *   avmoe_gemm / avmoe_resample_*  <- MoEAdapter.conv_adapter + fc      net_trans_v3.py:445-446,469-471
 *   avmoe_pool / avmoe_router_*    <- mean-pool, router MLP, softmax    net_trans_v3.py:460-466,472-479
 *   avmoe_latent_* / avmoe_rf_tok_* <- ExpertAdapter latent attention   net_trans_v3.py:379-384
 *   avmoe_expert_front_*           <- token<-latent attention, gate_av,
 *                                     ln_before, grouped down_sampler   net_trans_v3.py:386-395
 *   avmoe_bn_*                     <- bn1 / bn2 (train + eval)          net_trans_v3.py:397-403
 *   avmoe_expert_back_*            <- activation, up_sampler, bn2,
 *                                     ln_post, gate, prob-weighted sum  net_trans_v3.py:400-403,430-434,481-486
 *   avmoe_selfattn_*               <- AVVP NxN token self-attention     AVMOE/AVVP/nets/mgn.py:132-139
 *   avmoe_mha_*                    <- AVS v1 MHA over the BT axis       avs_s4/model/PVT_AVSModel_v2.py:212-215
 *   avmoe_lb_loss_*                <- compute_load_balancing_loss       AVMOE/AVVP/nets/mgn.py:220-224
 */
#ifndef AVMOE_B200_H
#define AVMOE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* avmoe_stream_t; /* cudaStream_t */

#define AVMOE_F32 0
#define AVMOE_BF16 1

#define AVMOE_E_BADARG (-1)
#define AVMOE_E_UNSUPPORTED (-2)
#define AVMOE_E_NODRIVER (-3)

#define AVMOE_MAX_EXPERTS 8

/* library / device probes (no GPU work) */
int avmoe_abi_version(void);
const char* avmoe_build_info(void);
const char* avmoe_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Generic batched GEMM on the 5th-gen tensor cores (tcgen05.mma, TMA-fed, TMEM accumulators).
 *
 *   D[m, n] = alpha * sum_k A[m,k] * B[n,k]  (+ row_bias[m] * col_scale[n%ldn] + col_bias[n%ldn])
 *
 * Operand storage (element type = dtype; TF32 math for f32, BF16 for bf16, fp32 accumulate):
 *   a_kmajor=1: A is [a_batch][M][Ka]  (Ka contiguous, row pitch lda, batch pitch a_bstride);
 *               a_batch>1 means K is the concatenation (b,k): the product is summed over b.
 *   a_kmajor=0: A is stored transposed [K][M] (M contiguous, pitch lda); a_batch must be 1.
 *   b_kmajor=1: B is [b_batch][N][Kb]  (same rules; b_batch must equal a_batch).
 *   b_kmajor=0: B is [b_batch][K][Nb]  (Nb contiguous, pitch ldb, batch pitch b_bstride); the
 *               GEMM N axis is the concatenation (b, nb), N = b_batch*Nb, and D is written
 *               as [b_batch][M][Nb] (pitch ldd, batch pitch d_bstride).
 *   Nb: 0 (or N) = D is [M][N] with pitch ldd.  0 < Nb < N = the N axis is a concatenation of N/Nb batches
 *               and D is written as [N/Nb][M][Nb] (pitch ldd, batch pitch d_bstride) — also valid with a
 *               K-major B whose rows are the flattened (batch, n) pairs.  d_dtype: AVMOE_F32 or AVMOE_BF16.
 *   split_k>1: K range is split over split_k CTAs that atomically add into a zero-initialised
 *               fp32 D (the call zeroes it); biases are added once.
 *   impl: 0 = tcgen05 whenever TMA/UMMA can express the operands (16-byte pitches, per-batch MN extents
 *             that are multiples of the swizzle chunk), else the generic SIMT kernel (tiny/odd shapes only);
 *         1 = force the SIMT kernel (tests cross-check the tensor-core path against it);
 *         2 = tcgen05 or fail with AVMOE_E_UNSUPPORTED (tests prove the hot shapes stay on tensor cores).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* A; const void* B; void* D;
  int32_t dtype, d_dtype;
  int64_t M, N, K;            /* K = per-batch K when a_batch>1 */
  int32_t a_kmajor, b_kmajor;
  int64_t lda, ldb, ldd;
  int32_t a_batch, b_batch;   /* see above */
  int64_t a_bstride, b_bstride, d_bstride;
  int64_t Nb;                 /* per-batch N extent of D (see above); 0 = unbatched */
  float alpha;
  const float* row_bias;      /* [M] or NULL */
  const float* col_scale;     /* [Nb] or NULL (=1) ; multiplies row_bias */
  const float* col_bias;      /* [Nb] or NULL */
  int32_t split_k;
  int32_t accumulate;         /* 1: D += result (fp32 D: atomics; bf16 D: read-modify-write in the epilogue, no split-K) */
  int32_t impl;
  int32_t full_batch;         /* >1: that many independent problems; A, B, D advance by a_bstride/b_bstride/d_bstride
                                 (a_batch = b_batch = 1, no N-batching, no split-K) */
  int64_t col_bias_bstride;   /* full_batch: col_bias advances by this many floats per problem (0 = one bias shared by all) */
} avmoe_gemm_args;

int avmoe_gemm(const avmoe_gemm_args* args, avmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Op-level entry: the token resampler  vis' = fc(conv_adapter(other))  and its backward
 * (MoEAdapter.conv_adapter + fc, AVMOE/AVE/nets/net_trans_v3.py:445-446,469-471; SURVEY.md §8a row a1, §8b "avmoe_resample_fwd/bwd").
 * One call issues the whole launch sequence from C++ (weight casts, layout transposes, tcgen05 GEMMs, bias reductions) against a
 * caller-provided workspace of avmoe_resample_workspace_bytes() bytes (256-byte aligned).  Association by FLOPs:
 *   Cx > C :  mid = P = o·W_fcᵀ [BT][Nx][C],  v = W_tok·P + b_tok⊗s + 1⊗b_fc  with  s = rowsum(W_fc) [C]
 *   else   :  mid = U = W_tok·o + b_tok⊗1 [BT][N][Cx],  v = U·W_fcᵀ + 1⊗b_fc.
 * `mid` (and `s`) are written by the forward and must be handed back unchanged to the backward.  Activations o, v, mid, dv, d_o
 * have element type `dtype`; weights, biases and their gradients are fp32.  d_o may be NULL (no gradient wanted for `other`).
 * `launches` returns how many kernels / memsets the call enqueued.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t dtype;
  int32_t BT, N, Nx, C, Cx;
  int32_t gemm_impl;          /* forwarded to avmoe_gemm_args.impl (0 default, 1 exact SIMT, 2 tensor cores or fail) */
  const void* o;              /* [BT][Nx][Cx] tokens of the other modality */
  const float* w_tok;         /* conv_adapter.weight [N][Nx] */
  const float* b_tok;         /* conv_adapter.bias   [N] */
  const float* w_fc;          /* fc.weight [C][Cx] */
  const float* b_fc;          /* fc.bias   [C]  (forward only) */
  void* v;                    /* forward out [BT][N][C] */
  void* mid;                  /* forward out / backward in (see above) */
  float* s;                   /* forward out / backward in: rowsum(W_fc) [C]; required when Cx > C */
  const void* dv;             /* backward in  [BT][N][C] */
  void* d_o;                  /* backward out [BT][Nx][Cx] or NULL */
  float* dw_tok;              /* backward out [N][Nx] */
  float* db_tok;              /* backward out [N] */
  float* dw_fc;               /* backward out [C][Cx] */
  float* db_fc;               /* backward out [C] */
  void* workspace;
  int64_t workspace_bytes;
  int32_t launches;           /* out */
} avmoe_resample_args;

int64_t avmoe_resample_workspace_bytes(const avmoe_resample_args* args, int backward);
int avmoe_resample_fwd(avmoe_resample_args* args, avmoe_stream_t stream);
int avmoe_resample_bwd(avmoe_resample_args* args, avmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Layout helpers feeding the GEMMs, router pooling, bias-gradient reductions.
 * ------------------------------------------------------------------------------------------ */
/* out[b][c][r] = in[b][r][c]  (dtype pairs: f32->f32, f32->bf16, bf16->bf16) */
int avmoe_transpose(const void* in, void* out, int batch, int R, int C, int64_t in_ld, int64_t in_bstride, int64_t out_ld,
                    int64_t out_bstride, int in_dtype, int out_dtype, avmoe_stream_t stream);
int avmoe_cast(const void* in, void* out, int64_t n, int in_dtype, int out_dtype, avmoe_stream_t stream);
/* partial[b][chunk][c] = sum of x[b][n][c] over the chunk's tokens; chunk count from avmoe_pool_chunks */
int avmoe_pool_chunks(int BT, int N);
int avmoe_pool(const void* x, float* partial, int BT, int N, int C, int chunks, int dtype, avmoe_stream_t stream);
/* rows_w[n] = sum_{b,c} d*col_weight[c] ; cols[c] = sum_{b,n} d ; cols_w[c] = sum_{b,n} d*row_weight[n]  (outputs may be NULL) */
int avmoe_reduce_bias(const void* d, const float* col_weight, const float* row_weight, float* rows_w, float* cols, float* cols_w,
                      int BT, int N, int C, int dtype, avmoe_stream_t stream);
int avmoe_rowsum(const float* w, float* out, int rows, int cols, avmoe_stream_t stream);
int avmoe_broadcast_rows(const float* src, void* out, int BT, int N, int C, int dtype, avmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Router (fp32, fixed reduction order; the wide first layer is tiled over samples) and the load-balancing loss.
 *   rin = [mean_N x | mean_N vis'] ; probs = softmax(W4 relu(W2 relu(W0 rin + b0) + b2) + b4 (+noise))
 *   idx = first argmax of probs (int64, bit-exact contract).
 * ------------------------------------------------------------------------------------------ */
int avmoe_router_fwd(const float* pool_x, const float* pool_v, int chunks_x, int chunks_v, int BT, int C, int E, float inv_n,
                     const float* w0, const float* b0, const float* w2, const float* b2, const float* w4, const float* b4,
                     const float* noise, float* probs, int64_t* idx, float* h1, float* h2, float* rin, avmoe_stream_t stream);
/* ws: BT*(E+32+128) floats of scratch; dpool_* already carry the 1/N of the mean */
int avmoe_router_bwd(const float* dprobs, const float* probs, const float* h1, const float* h2, const float* rin, int BT, int C,
                     int E, float inv_n, const float* w0, const float* w2, const float* w4, float* ws, float* dpool_x,
                     float* dpool_v, float* dw0, float* db0, float* dw2, float* db2, float* dw4, float* db4, avmoe_stream_t stream);
int avmoe_lb_loss_fwd(const float* probs, float* loss, int BT, int E, avmoe_stream_t stream);
int avmoe_lb_loss_bwd(const float* probs, const float* dloss, float* dprobs /* += */, int BT, int E, avmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Latent tokens attend to a key/value stream:  tok' = tok + softmax_N(tok kv^T) kv ; lse = logsumexp_N.
 *   tokens[i] [T,C] fp32 is written to slot slots[i] of tokp [BT,nl,T,C] / lse [BT,nl,T].
 *   Backward: dkv (= or +=) d/dkv, plus dpool[b][c] broadcast to every token (router path);
 *             dtokens[i] [T,C] += d/d tokens[i]  (zero-initialised by the caller).
 * ------------------------------------------------------------------------------------------ */
int64_t avmoe_latent_fwd_workspace(int BT, int Nk, int C, int T, int nexp);   /* floats */
int avmoe_latent_fwd(const void* kv, const float* const* tokens, const int* slots, int nexp, int nl, float* tokp, float* lse,
                     float* workspace, int BT, int Nk, int C, int T, int dtype, avmoe_stream_t stream);
int avmoe_latent_bwd(const void* kv, const float* const* tokens, const int* slots, int nexp, int nl, const float* tokp,
                     const float* lse, const float* dtokp, const float* dpool, void* dkv, int accumulate,
                     float* const* dtokens, int BT, int Nk, int C, int T, int dtype, avmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Expert kernels.  All experts of a MoEAdapter are evaluated per token tile in one launch.
 * ------------------------------------------------------------------------------------------ */
#define AVMOE_PRE_NONE 0        /* single-modal expert: x goes straight to ln_before                        */
#define AVMOE_PRE_LATENT 1      /* x + gate_pre * softmax_T(x tok'^T) tok'   (multimodal / AVS-v2 latent-self) */
#define AVMOE_PRE_AUX_ADD 2     /* x + gate_pre * aux                        (AVVP NxN self-attention result)  */
#define AVMOE_PRE_AUX_REPLACE 3 /* aux replaces x                            (AVS-v1 MHA output)              */
#define AVMOE_BN_OFF 0
#define AVMOE_BN_TRAIN 1
#define AVMOE_BN_EVAL 2

typedef struct {
  int32_t pre_kind, relu, latent_slot, _pad;
  const float* gate_pre;               /* [1] gate_av / gate_self, or NULL */
  const void* aux;                     /* [BT,N,C] activation dtype, or NULL */
  const float* ln1_w; const float* ln1_b;   /* ln_before (NULL = absent) */
  const float* wd;                     /* down_sampler.weight [Cr, C/G] */
  const float* wu;                     /* up_sampler.weight   [C, Cr/G] */
  const float* ln2_w; const float* ln2_b;   /* ln_post (NULL = absent) */
  const float* gate;                   /* [1] or NULL */
  /* fp32 gradient accumulators (+=, backward entry points only) */
  float* g_gate_pre; void* g_aux; float* g_ln1_w; float* g_ln1_b; float* g_wd; float* g_wu;
  float* g_ln2_w; float* g_ln2_b; float* g_gate;
} avmoe_expert;

typedef struct {
  int32_t BT, N, C, Cr, G, T, E, nl;
  int32_t dtype, bn_mode;
  avmoe_expert ex[AVMOE_MAX_EXPERTS];
  const void* x;            /* [BT,N,C] */
  const float* tokp;        /* [BT,nl,T,C] updated latent tokens */
  float* z;                 /* [BT,N,E,Cr] bottleneck activations before bn1 */
  double* stats1;           /* [E,Cr,2] sum z, sum z^2 (bn train) */
  const float* aff1; const float* save1;   /* [E,Cr,2] (scale,shift) / (mean,invstd) of bn1 */
  const float* aff2; const float* save2;   /* [E,C,2] of bn2 */
  const float* probs;       /* [BT,E] */
  void* out;                /* [BT,N,C] */
  const void* dout;         /* [BT,N,C] */
  double* red2;             /* [E,C,2]  sum dyt, sum dyt*yhat over all tokens (bn train: input of the main backward) */
  double* red2b;            /* [E,C,2]  same sums produced in eval mode (bn2 weight/bias grads) */
  double* red1;             /* [E,Cr,2] sum dzt, sum dzt*zhat */
  float* dzt;               /* [BT,N,E,Cr] grad wrt bn1 output (after the relu mask) */
  float* dprobs;            /* [BT,E] += */
  const float* dpool_x;     /* [BT,C] router-path grad per token (already /N) */
  void* dx;                 /* [BT,N,C] */
  float* dtokp;             /* [BT,nl,T,C] += */
} avmoe_expert_args;

int avmoe_expert_front_fwd(const avmoe_expert_args* a, avmoe_stream_t stream);
int avmoe_expert_back_fwd(const avmoe_expert_args* a, avmoe_stream_t stream);
int avmoe_expert_back_bwd_stats(const avmoe_expert_args* a, avmoe_stream_t stream);
int avmoe_expert_back_bwd(const avmoe_expert_args* a, avmoe_stream_t stream);
int avmoe_expert_front_bwd(const avmoe_expert_args* a, avmoe_stream_t stream);
/* sums of a = act(bn1(z)) and of a a^T per conv group, used to derive bn2's batch statistics algebraically */
int avmoe_bn_moments(const avmoe_expert_args* a, double* m1 /* [E,Cr] */, double* m2 /* [E,G,Cr/G,Cr/G] */, avmoe_stream_t stream);

typedef struct {
  int32_t E, C, Cr, G, train, _pad;
  int64_t n;                                   /* BT*N */
  const float* w1[AVMOE_MAX_EXPERTS]; const float* b1[AVMOE_MAX_EXPERTS];   /* bn1 weight / bias */
  const float* w2[AVMOE_MAX_EXPERTS]; const float* b2[AVMOE_MAX_EXPERTS];   /* bn2 weight / bias */
  float* rm1[AVMOE_MAX_EXPERTS]; float* rv1[AVMOE_MAX_EXPERTS];             /* running stats (updated when train) */
  float* rm2[AVMOE_MAX_EXPERTS]; float* rv2[AVMOE_MAX_EXPERTS];
  const float* wu[AVMOE_MAX_EXPERTS];          /* up_sampler weights [C, Cr/G] (bn2 train) */
  const double* stats1;                        /* [E,Cr,2] */
  const double* m1; const double* m2;          /* from avmoe_bn_moments */
  float* aff1; float* save1; float* aff2; float* save2;
} avmoe_bn_args;
int avmoe_bn1_finalize(const avmoe_bn_args* a, avmoe_stream_t stream);
int avmoe_bn2_finalize(const avmoe_bn_args* a, avmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Variant pre-transforms of the single-modal expert.
 *   AVVP  N x N token self-attention (AVMOE/AVVP/nets/mgn.py:132-139): the two N x N contractions are avmoe_gemm calls
 *         with full_batch; these kernels are the row softmax between them and its backward.
 *   AVS-v1 nn.MultiheadAttention over the BT axis (avs_s4/model/PVT_AVSModel_v2.py:212-215): per (token position, head)
 *         attention core on the fp32 in_proj output qkv [BT][N][3C]; probabilities are recomputed in the backward.
 * ------------------------------------------------------------------------------------------ */
int avmoe_nxn_softmax_rows(const float* S, void* P, float* lse, int64_t rows, int n, int p_dtype, avmoe_stream_t stream);
int avmoe_nxn_bwd_rows(const float* S, const float* dAtt, const float* lse, void* att, void* dS, int64_t rows, int n, int p_dtype,
                       avmoe_stream_t stream);
int avmoe_add_into(void* dst, const void* a, const void* b, const void* c, int64_t n, int dtype, avmoe_stream_t stream);
/* keep [NB,heads,L,L] uint8 or NULL: attention-probability dropout mask of nn.MultiheadAttention(dropout=0.2)
 * (avs_s4/model/PVT_AVSModel_v2.py:139-142); kept probabilities are scaled by keep_scale = 1/(1-p).  The host layer draws the mask
 * from torch's generator (or takes an injected one for parity runs) and hands the same mask to the backward. */
int avmoe_mha_attn_fwd(const float* qkv, void* ctx, int L, int NB, int C, int heads, int dtype, const unsigned char* keep, float keep_scale,
                       avmoe_stream_t stream);
int avmoe_mha_attn_bwd(const float* qkv, const float* dctx, void* dqkv, int L, int NB, int C, int heads, int dtype, const unsigned char* keep,
                       float keep_scale, avmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * GEMM-based expert path (bf16 streams).  The grouped 1x1 down/up projections of an expert (net_trans_v3.py:393-401, "down_sampler" /
 * "up_sampler") and their backward run as avmoe_gemm calls against block-diagonal dense weights; these entry points are the streaming
 * row passes between the GEMMs (pre-transform + ln_before, bn1 + activation, bn2 affine + ln_post + gate + probability-weighted
 * combine, and the hand-derived backward of each).  rows = BT*N.  Buffers (caller-owned, activation dtype unless noted):
 *   xl [E,rows,C]  ln_before output            z [rows,E,Cr] fp32  bottleneck before bn1      a [E,rows,Crp]  act(bn1(z)), zero padded
 *   yd [E,rows,2C] columns [0,C) = bn2-normalised up projection (written by the GEMM epilogue), columns [C,2C) = gradient at the bn2
 *                  output (written by avmoe_xg_post_bwd): one K-concatenated operand for  da = yd·W2^T  and  [2C,Crp] = yd^T·a
 *   da [E,rows,Crp] fp32   dz [E,rows,Crp]   dxl [E,rows,C]   ln1_stats / ln2_stats [E,rows,2] fp32 (mean, rstd)
 * avmoe_expert.pre_kind must be NONE / AUX_ADD / AUX_REPLACE (latent attention arrives as an aux-add).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t BT, N, C, Cr, Crp, G, E, dtype, bn_mode, _pad;
  avmoe_expert ex[AVMOE_MAX_EXPERTS];
  const float* bn2_w[AVMOE_MAX_EXPERTS]; const float* bn2_b[AVMOE_MAX_EXPERTS];   /* bn2 weight / bias (NULL: no BatchNorm) */
  const void* x; void* xl; float* ln1_stats;
  float* z; void* a; void* yd; float* ln2_stats;
  const float* aff1; const float* save1; const float* save2;   /* [E,Cr,2] (scale,shift), [E,Cr,2] / [E,C,2] (mean,invstd) */
  const float* probs; void* out; const void* dout;
  double* stats1;          /* [E,Cr,2] sum z, sum z^2 */
  double* red1;            /* [E,Cr,2] sum dh, sum dh*zhat  (dh = relu-masked da) */
  double* red2;            /* [E,C,2]  sum dy2, sum dy2*yhat */
  const float* da; void* dz; const void* dxl;
  float* dprobs;           /* [BT,E] += */
  const float* dpool_x;    /* [BT,C] router-path gradient per token (already /N), or NULL */
  void* dx;                /* [rows,C] */
  const void* res;         /* [rows,C] or NULL: residual stream folded into the combine epilogue, out = res + sum_e ... (the call-site add
                            * f = f + adapter(...), AVMOE/AVE/nets/net_trans_v3.py:703,706,718,720 -- SURVEY.md §8 row a9) */
} avmoe_xg_args;

int avmoe_xg_ln_pre_fwd(const avmoe_xg_args* a, avmoe_stream_t stream);     /* x, ex[].aux -> xl, ln1_stats */
int avmoe_xg_z_stats(const avmoe_xg_args* a, avmoe_stream_t stream);        /* z -> stats1 (+=) */
int avmoe_xg_bn_act(const avmoe_xg_args* a, avmoe_stream_t stream);         /* z, aff1 -> a */
int avmoe_xg_post_fwd(const avmoe_xg_args* a, avmoe_stream_t stream);       /* yd[:, :C], probs -> out, ln2_stats */
/* dout, yd[:, :C], ln2_stats, probs -> yd[:, C:], red2 (+=), dprobs (+=), ex[].g_ln2_w / g_ln2_b / g_gate (+=) */
int avmoe_xg_post_bwd(const avmoe_xg_args* a, avmoe_stream_t stream);
int avmoe_xg_bn1_bwd(const avmoe_xg_args* a, avmoe_stream_t stream);        /* da, z -> red1 (+=), dz */
/* dxl, x, ex[].aux, ln1_stats, dpool_x -> dx, ex[].g_aux, ex[].g_gate_pre / g_ln1_w / g_ln1_b (+=) */
int avmoe_xg_ln_pre_bwd(const avmoe_xg_args* a, avmoe_stream_t stream);
/* dense block-diagonal operands of the down projection from ex[].wd [Cr,C/G]:  wdd [E,Cr,C] (forward), wdt [E,C,Crp] (dgrad) */
int avmoe_xg_prep_weights(const avmoe_xg_args* a, void* wdd, void* wdt, avmoe_stream_t stream);
/* up-projection operand with bn2's normalisation folded in (yhat = a·wud^T + ybi):  wud [E,C,Crp] = invstd_c * dense(ex[].wu),
 * ybi [E,C] = -mean_c*invstd_c from save2; bn off: plain dense weight, zero bias */
int avmoe_xg_prep_up(const avmoe_xg_args* a, void* wud, float* ybi, avmoe_stream_t stream);
/* bn2 backward folded into the operand of  da = yd·W2^T + bias:  w2 [E,Crp,2C] (activation dtype), bias [E,Crp] fp32 */
int avmoe_xg_prep_bwd(const avmoe_xg_args* a, void* w2, float* bias, avmoe_stream_t stream);
/* pwd [E,Crp,C] = dz^T·xl, pwu [E,2C,Crp] = yd^T·a, suma [E,Cr] fp64 column sums of a (bn train)  ->  ex[].g_wd, ex[].g_wu (=) */
int avmoe_xg_finish_wgrads(const avmoe_xg_args* a, const float* pwd, const float* pwu, const double* suma, avmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * GEMM-based latent attention (the product path for multimodal / latent-self experts; AVMOE/AVE/nets/net_trans_v3.py:379-390):
 * the contractions are avmoe_gemm calls (flat, fully-batched); these are the streaming softmax passes between them on the
 * token-major logits S1^T [BT,Nk,T] (softmax over Nk, per column) and S2 [BT,N,T] (softmax over T, per row).
 * ------------------------------------------------------------------------------------------ */
/* off_lo: logits produced against a bf16 operand split into hi + lo halves (avmoe_cast_split) arrive as two column blocks of
 * the fp32 GEMM output, columns [0,T) and [off_lo, off_lo+T); the kernels add them on load.  off_lo = 0: one block. */
/* p_lo > 0 (bf16 P): also write the rounding residual  P - bf16(P)  at column p_lo, so (hi + lo)·kv carries ~16 bits of A1 */
int avmoe_softmax_cols(const float* S, void* P, float* lse, int BT, int Nk, int T, int lds, int ldp, int off_lo, int p_lo, int p_dtype,
                       avmoe_stream_t stream);
int avmoe_softmax_cols_bwd(void* A12 /* [BT,Nk,ld12]: A1^T at column 0, dS1^T written at column off2 */, const float* dA, const float* Dt, int BT,
                           int Nk, int T, int ld12, int off2, int ldd, int off_lo, int p_dtype, avmoe_stream_t stream);
int avmoe_softmax_rows(const float* S, void* A, int64_t rows, int T, int lds, int off_lo, int ld, int p_dtype, avmoe_stream_t stream);
int avmoe_softmax_rows_bwd(const void* A, const float* dA, void* dS, int64_t rows, int T, int ld, int ldd, int off_lo, int p_dtype,
                           avmoe_stream_t stream);
/* fp32 [groups][T][C] -> bf16 [groups][2*Tp][C]: rows [0,T) = hi = bf16(v), rows [Tp,Tp+T) = lo = bf16(v - hi); padding rows untouched */
int avmoe_cast_split(const float* in, void* out_bf16, int64_t groups, int T, int Tp, int C, avmoe_stream_t stream);
/* out[b][t][c] = in[b][t][c] (+ in[b][off_lo+t][c]) + tile[t][c], fp32; `in` has `rows` rows per sample */
int avmoe_tile_add(const float* in, const float* tile, float* out, int BT, int T, int C, int rows, int off_lo, avmoe_stream_t stream);
int avmoe_latent_dots(const float* dR, const float* tokp, const float* tok, float* Dt, int BT, int T, int C, avmoe_stream_t stream);
int avmoe_add_rowvec(void* d, const float* vec, int BT, int N, int C, int dtype, avmoe_stream_t stream);
int avmoe_batch_sum_add(const float* in, float* out /* += */, int BT, int64_t L, avmoe_stream_t stream);


/* ------------------------------------------------------------------------------------------
 * Row-fused expert passes (bf16 streams; csrc/rowfused.cu).  One warp owns 16 token rows and keeps every intermediate of
 * those rows in registers / its shared-memory tile; the skinny contractions (K or N = T latent tokens, C/r bottleneck
 * channels) run on warp-level bf16 tensor-core tiles against "packed B" weight operands (mma fragment order, see
 * csrc/mma_tile.cuh), so only the row's inputs and outputs touch HBM.  They replace chains of GEMM + row-kernel launches:
 *
 *   avmoe_rf_front_fwd   x (+aux), tok' -> [softmax_T(x tok'^T) tok', gate_pre residual, ln_before, grouped down projection] -> z
 *                        (net_trans_v3.py:386-395; replaces S2 GEMM + softmax_rows + A2.tok' GEMM + xg_ln_pre_fwd + down GEMM)
 *   avmoe_rf_back_fwd    z -> [bn1 affine, act, grouped up projection, bn2, ln_post, gate, probability-weighted combine over
 *                        experts, + residual stream] -> out      (net_trans_v3.py:398-403,430-434,481-486 and the call-site add
 *                        :703-720; replaces xg_bn_act + up GEMM + xg_post_fwd)
 *   avmoe_rf_back_bwd    dout, z -> recompute of the above, yd = [yhat | dy2], a, d ln_post / d gate / d probs, bn2 batch reductions
 *                        (replaces xg_post_bwd and the saved yd / a tensors)
 *   avmoe_rf_front_bwd   dz, x, tok' -> recompute of the front pass, dx, xl (for the weight-gradient GEMM), d aux / dXR,
 *                        A2 | dS2 (for the d tok' GEMMs), d ln_before, d gate_pre
 *                        (replaces dxl GEMM + xg_ln_pre_bwd + dA2 GEMM + softmax_rows_bwd + dS2.tok' GEMM)
 * ------------------------------------------------------------------------------------------ */
/* fp32 weight matrix -> packed B operand.  Logical operand B[k][n], k < K, n < Nn, zero padded to KS = ceil(K/16) k-steps and
 * NP = ceil(Nn/16) n-tile pairs; dst holds nb * NP * KS * 32 uint4.  Source element of B[k][n] for batch b:
 *   src[b*bstride + n*stride_n + k*stride_k]                              (groups == 1), or the block-diagonal operand of a grouped
 *   1x1 convolution (groups > 1): non-zero only where n and k fall in the same group (n / (Nn/groups) == k / (K/groups));
 *   group_local = 1: the k index (0: the n index) is taken relative to its group start, as conv weights are stored.
 * scale_n [nb,Nn] / scale_k [nb,K] (optional, fp32, element stride scale_*_stride) multiply the element.  hi_lo = 1: n-tile pairs hold (hi, lo) halves of ONE
 * n-tile each (NP = ceil(Nn/8)). */
int avmoe_rf_pack_b(const float* src, int64_t bstride, int64_t stride_n, int64_t stride_k, int nb, int Nn, int K, int groups,
                    int group_local_k, const float* scale_n, int scale_n_stride, const float* scale_k, int scale_k_stride, int hi_lo,
                    int Nn_pad /* layout extent of n (>= Nn, extra columns zero); 0 = Nn */, void* dst, avmoe_stream_t stream);
int64_t avmoe_rf_pack_b_bytes(int nb, int Nn, int K, int hi_lo);
/* csum[b][n] = sum_k of the bf16 elements of a packed (not hi/lo) operand, n < 16*ceil(Nn/16): the LayerNorm mean fold of the
 * front pass uses the column sums of exactly the rounded weights the tensor cores see */
int avmoe_rf_colsum(const void* packed, int nb, int Nn, int K, float* csum, avmoe_stream_t stream);

typedef struct {
  int32_t BT, N, C, Cr, Crp, G, E, bn_mode;
  avmoe_expert ex[AVMOE_MAX_EXPERTS];                    /* relu, ln2_w/ln2_b, gate (+ g_ln2_w, g_ln2_b, g_gate in the backward) */
  const float* bn2_w[AVMOE_MAX_EXPERTS]; const float* bn2_b[AVMOE_MAX_EXPERTS];   /* NULL: no BatchNorm */
  const float* z;            /* [rows,E,Cr] fp32 bottleneck before bn1 */
  const float* aff1;         /* [E,Cr,2] bn1 (scale, shift) or NULL */
  const void* wu_pack;       /* E packed operands B[k=j][n=c] = invstd2_c * up_sampler (block diagonal), K = Cr, Nn = C */
  const float* ybi;          /* [E,C] -mean2_c * invstd2_c (zeros without BatchNorm) */
  const float* probs;        /* [BT,E] */
  const void* res;           /* [rows,C] residual stream or NULL */
  void* out;                 /* [rows,C] */
  const void* dout;          /* [rows,C]  (backward) */
  void* yd;                  /* [E,rows,2C]: columns [0,C) = yhat, [C,2C) = gradient at the bn2 output (backward) */
  void* act;                 /* [E,rows,Crp] a = act(bn1(z)), zero padded (backward) */
  float* dprobs;             /* [BT,E] += */
  double* red2;              /* [E,C,2] += sum dy2, sum dy2*yhat (BatchNorm only) */
} avmoe_rf_back_args;
int avmoe_rf_back_fwd(const avmoe_rf_back_args* a, avmoe_stream_t stream);
int avmoe_rf_back_bwd(const avmoe_rf_back_args* a, avmoe_stream_t stream);

typedef struct {
  int32_t BT, N, C, Cr, Crp, G, E, T, has_ln, _pad;
  avmoe_expert ex[AVMOE_MAX_EXPERTS];      /* pre_kind, gate_pre, aux, ln1_w, ln1_b (+ g_gate_pre, g_aux, g_ln1_w, g_ln1_b in the backward;
                                              for AVMOE_PRE_LATENT experts g_aux receives dXR = gate_pre * d(x_pre), [rows,C]) */
  const void* tok_s2[AVMOE_MAX_EXPERTS];   /* PRE_LATENT: per-sample packed hi/lo operand B[k=c][n=t] = tok'[b][t][c] (K=C, Nn=T padded to 16) */
  const void* tok_xr[AVMOE_MAX_EXPERTS];   /* PRE_LATENT: per-sample packed operand B[k=t][n=c] = tok'[b][t][c] (K=T, Nn=C) */
  const void* wd_pack;       /* E operands B[k=c][n=j] = ln1_w[c] * dense down_sampler[j][c] (K=C, Nn=Cr) */
  const float* wd_csum;      /* [E,16*ceil(Cr/16)] avmoe_rf_colsum of wd_pack */
  const float* c2;           /* [E,Cr] dense down_sampler · ln1_b (zeros without ln_before) */
  const void* wdt_pack;      /* backward: E operands B[k=j][n=c] = dense down_sampler[j][c] (K=Cr, Nn=C) */
  const void* x;             /* [rows,C] */
  float* z;                  /* [rows,E,Cr] fp32 (forward output) */
  const void* dz;            /* [E,rows,Crp] gradient at the down-projection output (backward) */
  const float* dpool_x;      /* [BT,C] router-path gradient per token (already /N) or NULL */
  void* dx;                  /* [rows,C] */
  void* xl;                  /* [E,rows,C] ln_before output, operand of the weight-gradient GEMM */
  void* a2[AVMOE_MAX_EXPERTS];   /* PRE_LATENT: [rows,Tp8] softmax_T probabilities (operand of the d tok' GEMM) */
  void* ds2[AVMOE_MAX_EXPERTS];  /* PRE_LATENT: [rows,Tp8] gradient at the logits */
  float* ln1_stats;          /* [E,rows,2] ln_before (mean, rstd) per row: written by the forward, read by the backward */
  const void* dres;          /* [rows,C] or NULL (backward): gradient of the folded residual when the residual stream IS x -- added into dx
                              * here instead of by a separate autograd accumulation pass */
} avmoe_rf_front_args;
int avmoe_rf_front_fwd(const avmoe_rf_front_args* a, avmoe_stream_t stream);
int avmoe_rf_front_bwd(const avmoe_rf_front_args* a, avmoe_stream_t stream);
/* 1 when the row-fused passes support this geometry (C % 16 == 0, C/r <= 96, 2 <= E <= 4, T <= 96), else 0 (callers use the GEMM path);
 * the front passes additionally need N % 16 == 0 (a 16-row tile never straddles two samples' latent tokens) */
int avmoe_rf_supported(int C, int Cr, int E, int T);

/* ------------------------------------------------------------------------------------------
 * Row-fused passes of the latent-token update  tok' = tok + softmax_Nk(tok · kvᵀ) · kv  (net_trans_v3.py:379-384): the logits are
 * recomputed per 16-row tile on the tensor cores and never stored (csrc/rf_tokens.cu).
 *   avmoe_rf_tok_probs: A12[b][n][0..T) = bf16(A1ᵀ), A12[b][n][Tp..Tp+T) = its bf16 rounding residual; `pack` = the shared latent
 *                       tokens [T][C] as a packed hi/lo operand (avmoe_rf_pack_b, nb = 1, Nn_pad = Tp); `part` = workspace of
 *                       BT * avmoe_rf_tok_parts(BT, Nk) * 2 * Tp floats; `lse` [BT][T] optional output.
 *   avmoe_rf_tok_ds:    A12[b][n][Tp..Tp+T) = A1ᵀ ∘ (kv · dRᵀ − Dt[b][t]); `pack` = dR [BT][T][C] packed per sample (nb = BT), per_sample = 1.
 * kv [BT][Nk][C] bf16, Nk % 16 == 0, C % 16 == 0, C <= 384, Tp = T rounded up to 8; avmoe_rf_tok_supported says whether a geometry is covered.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t BT, Nk, C, T, Tp;
  int32_t per_sample;        /* 0: one packed operand shared by all samples; 1: one per sample */
  const void* kv;
  const void* pack;
  float* part;
  void* A12;                 /* [BT][Nk][2*Tp] bf16 */
  const float* Dt;           /* [BT][T] (backward) */
  float* lse;                /* [BT][T] or NULL (forward) */
} avmoe_rf_tok_args;
int avmoe_rf_tok_supported(int Nk, int C, int T);
int avmoe_rf_tok_parts(int BT, int Nk);
int avmoe_rf_tok_probs(const avmoe_rf_tok_args* a, avmoe_stream_t stream);
int avmoe_rf_tok_ds(const avmoe_rf_tok_args* a, avmoe_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused Adam on the flat parameter buffer (torch.optim.Adam semantics, AVMOE/AVE/main_trans_v3.py:311-323: default betas/eps,
 * L2 weight decay added to the gradient, bias correction by `step` >= 1).  p, g, m, v: n fp32 each (16-byte aligned).
 * grad_scale multiplies the gradient first (1/world of the data-parallel average, 1/accum of gradient accumulation);
 * zero_grad != 0 clears g in the same pass.
 * ------------------------------------------------------------------------------------------ */
int avmoe_adam_flat(float* p, float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                    float grad_scale, int64_t step, int zero_grad, avmoe_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AVMOE_B200_H */
