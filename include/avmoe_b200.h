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
 *   avmoe_latent_*                 <- ExpertAdapter latent attention    net_trans_v3.py:379-384
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
  int32_t accumulate;         /* 1: D += result (fp32 D only, atomics) */
  int32_t impl;
} avmoe_gemm_args;

int avmoe_gemm(const avmoe_gemm_args* args, avmoe_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AVMOE_B200_H */
