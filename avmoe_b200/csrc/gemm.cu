// Batched GEMM for the AVMoE token resampler (conv_adapter + fc) and its gradients.
//
// Product path: a persistent, warp-specialised tcgen05 kernel for sm_100a
//   warp 0     TMA producer   (cp.async.bulk.tensor 3-D tiles, 128/64/32-byte swizzle, OOB zero fill)
//   warp 1     MMA issuer     (one elected lane: tcgen05.mma.cta_group::1, M=128, N<=256, fp32 accum in TMEM)
//   warps 2-9  epilogue       (tcgen05.ld 32x32b -> smem transpose -> bias/scale -> coalesced 16-byte or atomic global stores)
// with a multi-stage smem ring (full/empty mbarriers) and TWO accumulator stages in TMEM so the
// epilogue of tile i overlaps the main loop of tile i+1.
//
// Both operands may be K-major or MN-major (UMMA descriptors support either for BF16 and TF32), which
// is what lets every gradient GEMM of the resampler consume the token-major activations in place,
// with no transposed copies (see DESIGN.md "resampler").
//
// Replaces: nn.Conv2d(Nx->N,1x1) over the token axis + nn.Linear(Cx->C)
//           (AVMOE/AVE/nets/net_trans_v3.py:445-446,469-471) and their autograd.
#include <atomic>
#include <stdlib.h>
#include <mutex>

#include "common.cuh"

namespace avmoe {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (cudaErrorLaunchFailure), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {  // ~2 s at 2 GHz
      printf("avmoe gemm: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool kTf32>
__device__ __forceinline__ void tcgen05_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kTf32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// explicit shared-space accesses for the epilogue staging slabs (the dynamic-smem carve-up hides the address space from the
// compiler, which otherwise emits generic LD.E/ST.E with their longer latency)
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------
// kernel parameters
// ------------------------------------------------------------------------------------------------
struct OperandTile {
  int kmajor;          // 1: [rows][K] tile (one TMA box); 0: MN-major, tile = nchunks boxes of [BK rows][chunk]
  int nchunks;         // boxes per stage
  int chunk_elems;     // MN elements per box (MN-major only)
  int batch_extent;    // MN-major batched over N: per-batch MN extent (Nb); else 0
  uint32_t chunk_bytes;  // smem bytes per box
  uint32_t k_advance;    // descriptor start-address advance per MMA (bytes)
  uint64_t desc_base;    // descriptor without the start address
};

struct GemmParams {
  CUtensorMap tmA, tmB;
  int slab_path;       // 1: epilogue transposes 32x32 slabs through smem and writes coalesced 16-byte rows (slabs never straddle a batch)
  int vec_ok;          // D pitches / base allow 16-byte vector stores
  OperandTile a, b;
  int BN;              // tile N (multiple of 32, <= 256)
  int BK;              // K elements per stage
  int mma_per_stage;
  int num_stages;
  uint32_t stage_bytes_a, stage_bytes_b;
  uint32_t idesc;
  int tmem_cols;       // power of two >= nacc*BN
  int nacc;            // accumulator stages in TMEM: 4 when BN <= 128 (narrow tiles: epilogues of consecutive tiles overlap), else 2
  int tiles_m, tiles_n, splits;
  int full_batch;      // >1: A, B and D all advance over this many independent problems (third TMA coordinate)
  int k_iters;         // total stages over the whole (batch,K) reduction
  int k_iters_per_batch;  // batched-K: stages per batch element; else k_iters
  int k_iters_per_split;
  // epilogue
  void* D;
  int d_bf16;
  long long M, N, Nb, ldd, d_bstride;
  int n_batched;       // D is [b][M][Nb]
  float alpha;
  const float* row_bias;
  const float* col_scale;
  const float* col_bias;
  long long cb_bstride;   // full_batch: col_bias advances by this many floats per problem (0 = shared)
  int acc_bf16;           // bf16 D, accumulate: D = D + result (read-modify-write in the epilogue, one rounding)
  int atomic;
};

static constexpr int kMaxAcc = 4;
static constexpr int kEpiWarps = 16;        // epilogue warps: kEpiWarps/4 per TMEM lane quarter, taking interleaved 32-column slabs
static constexpr int kGemmThreads = 64 + 32 * kEpiWarps;   // TMA warp, MMA warp, epilogue warps
static constexpr int kBM = 128;
static constexpr int kStagingBytes = ((kEpiWarps * 32 * 33 * 4 + 1023) / 1024) * 1024;   // per-warp [32][33] fp32 slabs

template <bool kTf32>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A tile][B tile] (1024-aligned), then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = p.stage_bytes_a + p.stage_bytes_b;
  uint8_t* staging = smem + (size_t)p.num_stages * stage_bytes;          // 4 warps x 2 x 4 KiB (1024-aligned)
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kStagingBytes);
  uint64_t* full_bar = bars;                       // [num_stages]
  uint64_t* empty_bar = bars + p.num_stages;       // [num_stages]
  uint64_t* tmem_full = bars + 2 * p.num_stages;   // [kMaxAcc]
  uint64_t* tmem_empty = tmem_full + kMaxAcc;      // [kMaxAcc]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + kMaxAcc);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmB) : "memory");

    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < p.nacc; ++s) {
      mbar_init(smem_u32(&tmem_full[s]), 1);
      mbar_init(smem_u32(&tmem_empty[s]), kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int tiles_mn = p.tiles_m * p.tiles_n;
  const int tiles_per_batch = tiles_mn * p.splits;
  const int num_tiles = tiles_per_batch * p.full_batch;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tb = tile / tiles_per_batch;
        const int trem = tile - tb * tiles_per_batch;
        const int split = trem / tiles_mn;
        const int mn = trem - split * tiles_mn;
        const int m0 = (mn % p.tiles_m) * kBM;
        const int n0 = (mn / p.tiles_m) * p.BN;
        const int k_begin = split * p.k_iters_per_split;
        const int k_end = min(p.k_iters, k_begin + p.k_iters_per_split);
        for (int kit = k_begin; kit < k_end; ++kit) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, stage_bytes);
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t sb = sa + p.stage_bytes_a;
          int kb = kit / p.k_iters_per_batch;
          const int kc = (kit - kb * p.k_iters_per_batch) * p.BK;
          if (p.full_batch > 1) kb = tb;
          if (p.a.kmajor) {
            tma_load_3d(sa, &p.tmA, fb, kc, m0, kb);
          } else {
            for (int i = 0; i < p.a.nchunks; ++i)
              tma_load_3d(sa + i * p.a.chunk_bytes, &p.tmA, fb, m0 + i * p.a.chunk_elems, kc, tb);
          }
          if (p.b.kmajor) {
            tma_load_3d(sb, &p.tmB, fb, kc, n0, kb);
          } else {
            for (int i = 0; i < p.b.nchunks; ++i) {
              int n = n0 + i * p.b.chunk_elems;
              int bb = tb;
              if (p.b.batch_extent) { bb = n / p.b.batch_extent; n -= bb * p.b.batch_extent; }
              tma_load_3d(sb + i * p.b.chunk_bytes, &p.tmB, fb, n, kc, bb);
            }
          }
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int split = (tile % tiles_per_batch) / tiles_mn;
      const int k_begin = split * p.k_iters_per_split;
      const int k_end = min(p.k_iters, k_begin + p.k_iters_per_split);
      mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
      for (int kit = k_begin; kit < k_end; ++kit) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tcgen05_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t sb = sa + p.stage_bytes_a;
          for (int j = 0; j < p.mma_per_stage; ++j) {
            const uint64_t adesc = p.a.desc_base | (uint64_t)(((sa + j * p.a.k_advance) & 0x3FFFF) >> 4);
            const uint64_t bdesc = p.b.desc_base | (uint64_t)(((sb + j * p.b.k_advance) & 0x3FFFF) >> 4);
            tcgen05_mma<kTf32>(d_tmem, adesc, bdesc, p.idesc, (kit > k_begin || j > 0) ? 1u : 0u);
          }
          tcgen05_commit(smem_u32(&empty_bar[stage]));
          if (kit == k_end - 1) tcgen05_commit(smem_u32(&tmem_full[acc]));
        }
        __syncwarp();
        if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
      }
      if (k_end <= k_begin && lane == 0) tcgen05_commit(smem_u32(&tmem_full[acc]));  // empty split: still signal
      __syncwarp();
      if (++acc == p.nacc) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // the warps of a quarter take interleaved 32-column slabs
    int tseq = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t stg = smem_u32(staging) + (uint32_t)((warp - 2) * (32 * 33) * 4);   // per-warp [32][33] fp32 transpose slab (shared address)
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int tb = tile / tiles_per_batch;
      const int trem = tile - tb * tiles_per_batch;
      const int split = trem / tiles_mn;
      const int mn = trem - split * tiles_mn;
      const int m0 = (mn % p.tiles_m) * kBM;
      const int n0 = (mn / p.tiles_m) * p.BN;
      const int k_begin = split * p.k_iters_per_split;
      const bool has_k = k_begin < p.k_iters;
      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
      tcgen05_fence_after();
      const int mrow0 = m0 + quarter * 32;
      const long long m = mrow0 + lane;
      const bool add_bias = (split == 0);
      const float* cbias = p.col_bias ? p.col_bias + (long long)tb * p.cb_bstride : nullptr;
      const float rb = (add_bias && p.row_bias && m < p.M) ? p.row_bias[m] : 0.f;
      // slab g of this CTA's tseq-th tile goes to warp group (g + tseq) % groups: consecutive narrow tiles are drained by
      // different warps, so their epilogues overlap (two accumulator stages in flight) instead of queueing on one group
      constexpr int kGroups = kEpiWarps / 4;
      const int gstart = (half + kGroups - (tseq % kGroups)) % kGroups;
      ++tseq;
      for (int g = gstart; g < p.BN / 32; g += kGroups) {
        const long long ng = (long long)n0 + g * 32;
        if (ng >= p.N || mrow0 >= p.M) break;   // warp-uniform
        long long bb = tb, c0 = ng;
        if (p.n_batched) { bb = ng / p.Nb; c0 = ng - bb * p.Nb; }
        // interior bf16 slab: the 8 bias values this lane adds are requested BEFORE the TMEM load so their latency hides under it
        const bool interior = p.slab_path && p.d_bf16 && p.vec_ok && !p.row_bias && mrow0 + 32 <= p.M && ng + 32 <= p.N;
        const bool packed = interior && !p.acc_bf16;        // convert to bf16 BEFORE the transpose: half the smem traffic, no re-pack
        float cbv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) cbv[j] = (interior && !packed && add_bias && cbias) ? __ldg(cbias + c0 + (lane & 3) * 8 + j) : 0.f;
        const bool pbias = packed && add_bias && cbias;
        const float cbl = pbias ? __ldg(cbias + c0 + lane) : 0.f;   // lane e holds the bias of column e (broadcast by shuffles below)
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * p.BN + g * 32);
        tmem_ld_32x32b_x32(taddr, r);
        if (packed) {
          // ---- interior bf16 slab, packed: lane = row holds 32 columns -> 16 bf16x2 words -> per-warp [32][17]-word slab ->
          //      each lane reads 4 consecutive words (8 columns) of 4 rows -> 16-byte global stores ----
          const float al = has_k ? p.alpha : 0.f;
          uint32_t pk[16];
          if (pbias) {
#pragma unroll
            for (int e2 = 0; e2 < 16; ++e2) {
              const float b0 = __shfl_sync(0xffffffffu, cbl, 2 * e2), b1 = __shfl_sync(0xffffffffu, cbl, 2 * e2 + 1);
              const float v0 = has_k ? fmaf(__uint_as_float(r[2 * e2]), al, b0) : b0;
              const float v1 = has_k ? fmaf(__uint_as_float(r[2 * e2 + 1]), al, b1) : b1;
              pk[e2] = pack_bf16x2(v0, v1);
            }
          } else if (has_k) {
#pragma unroll
            for (int e2 = 0; e2 < 16; ++e2) pk[e2] = pack_bf16x2(__uint_as_float(r[2 * e2]) * al, __uint_as_float(r[2 * e2 + 1]) * al);
          } else {
#pragma unroll
            for (int e2 = 0; e2 < 16; ++e2) pk[e2] = 0u;
          }
          __syncwarp();                                   // the previous slab has been read out
#pragma unroll
          for (int e2 = 0; e2 < 16; ++e2) sts_u32(stg + (uint32_t)((lane * 17 + e2) * 4), pk[e2]);
          __syncwarp();
          const int chunk = lane & 3, rsub = lane >> 2;
          __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.D) + bb * p.d_bstride + (long long)(mrow0 + rsub) * p.ldd + c0 + chunk * 8;
          const uint32_t sp = stg + (uint32_t)((rsub * 17 + chunk * 4) * 4);
          const long long step = 8 * p.ldd;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint4 q;
            q.x = lds_u32(sp + (uint32_t)((k * (8 * 17) + 0) * 4));
            q.y = lds_u32(sp + (uint32_t)((k * (8 * 17) + 1) * 4));
            q.z = lds_u32(sp + (uint32_t)((k * (8 * 17) + 2) * 4));
            q.w = lds_u32(sp + (uint32_t)((k * (8 * 17) + 3) * 4));
            *reinterpret_cast<uint4*>(dp + k * step) = q;
          }
        } else if (p.slab_path) {
          // ---- fast path: registers -> per-warp smem slab [32][33] -> coalesced 16-byte global stores (or fp32 atomics) ----
          __syncwarp();                                   // the previous slab has been read out
          if (has_k) {
#pragma unroll
            for (int e = 0; e < 32; ++e) sts_f32(stg + (uint32_t)((lane * 33 + e) * 4), __uint_as_float(r[e]) * p.alpha);
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) sts_f32(stg + (uint32_t)((lane * 33 + e) * 4), 0.f);
          }
          __syncwarp();
          if (interior) {
            // ---- interior bf16 slab: no predicates, one 64-bit address per lane, 4 x (8 LDS + pack + 16-byte store) ----
            const int chunk = lane & 3, rsub = lane >> 2;
            const float (&cb)[8] = cbv;
            __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.D) + bb * p.d_bstride + (long long)(mrow0 + rsub) * p.ldd + c0 + chunk * 8;
            const uint32_t sp = stg + (uint32_t)((rsub * 33 + chunk * 8) * 4);
            const long long step = 8 * p.ldd;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = lds_f32(sp + (uint32_t)((k * (8 * 33) + j) * 4)) + cb[j];
              if (p.acc_bf16) {
                const uint4 old = *reinterpret_cast<const uint4*>(dp + k * step);
                const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&old);
#pragma unroll
                for (int j = 0; j < 4; ++j) { o[2 * j] += __low2float(oh[j]); o[2 * j + 1] += __high2float(oh[j]); }
              }
              uint4 q;
              __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(o[4], o[5]), h3 = __floats2bfloat162_rn(o[6], o[7]);
              q.x = *reinterpret_cast<uint32_t*>(&h0); q.y = *reinterpret_cast<uint32_t*>(&h1);
              q.z = *reinterpret_cast<uint32_t*>(&h2); q.w = *reinterpret_cast<uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(dp + k * step) = q;
            }
          } else if (p.d_bf16) {
            const int chunk = lane & 3, rsub = lane >> 2;          // 8 columns per lane, 8 rows per instruction
            const long long cc = c0 + chunk * 8;
            const bool cvalid = ng + chunk * 8 < p.N;
            float cs[8], cb[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const long long ce = cc + j < p.Nb ? cc + j : 0;
              cs[j] = (add_bias && p.row_bias) ? (p.col_scale ? __ldg(p.col_scale + ce) : 1.f) : 0.f;
              cb[j] = (add_bias && cbias) ? __ldg(cbias + ce) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int row = rsub + 8 * k;
              const long long mm = mrow0 + row;
              if (mm < p.M && cvalid) {
                const float rbv = (add_bias && p.row_bias) ? __ldg(p.row_bias + mm) : 0.f;
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = lds_f32(stg + (uint32_t)((row * 33 + chunk * 8 + j) * 4)) + rbv * cs[j] + cb[j];
                __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.D) + bb * p.d_bstride + mm * p.ldd + cc;
                const int valid = (int)min((long long)8, p.N - (ng + chunk * 8));
                if (p.acc_bf16) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) if (j < valid) o[j] += __bfloat162float(dp[j]);
                }
                if (valid == 8 && p.vec_ok) {
                  uint4 q;
                  __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]);
                  __nv_bfloat162 h2 = __floats2bfloat162_rn(o[4], o[5]), h3 = __floats2bfloat162_rn(o[6], o[7]);
                  q.x = *reinterpret_cast<uint32_t*>(&h0); q.y = *reinterpret_cast<uint32_t*>(&h1);
                  q.z = *reinterpret_cast<uint32_t*>(&h2); q.w = *reinterpret_cast<uint32_t*>(&h3);
                  *reinterpret_cast<uint4*>(dp) = q;
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) if (j < valid) dp[j] = __float2bfloat16_rn(o[j]);
                }
              }
            }
          } else {
            const int chunk = lane & 7, rsub = lane >> 3;          // 4 columns per lane, 4 rows per instruction
            const long long cc = c0 + chunk * 4;
            const bool cvalid = ng + chunk * 4 < p.N;
            float cs[4], cb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const long long ce = cc + j < p.Nb ? cc + j : 0;
              cs[j] = (add_bias && p.row_bias) ? (p.col_scale ? __ldg(p.col_scale + ce) : 1.f) : 0.f;
              cb[j] = (add_bias && cbias) ? __ldg(cbias + ce) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int row = rsub + 4 * k;
              const long long mm = mrow0 + row;
              if (mm < p.M && cvalid) {
                const float rbv = (add_bias && p.row_bias) ? __ldg(p.row_bias + mm) : 0.f;
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = lds_f32(stg + (uint32_t)((row * 33 + chunk * 4 + j) * 4)) + rbv * cs[j] + cb[j];
                float* dp = reinterpret_cast<float*>(p.D) + bb * p.d_bstride + mm * p.ldd + cc;
                const int valid = (int)min((long long)4, p.N - (ng + chunk * 4));
                if (p.atomic) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) if (j < valid) atomicAdd(dp + j, o[j]);
                } else if (valid == 4 && p.vec_ok) {
                  *reinterpret_cast<float4*>(dp) = make_float4(o[0], o[1], o[2], o[3]);
                } else {
#pragma unroll
                  for (int j = 0; j < 4; ++j) if (j < valid) dp[j] = o[j];
                }
              }
            }
          }
        } else if (m < p.M) {
          // ---- generic path (per-batch N not a multiple of 32): direct 4-wide stores ----
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const long long n = ng + v * 4;
            if (n >= p.N) continue;        // (no break: keeps r[] statically indexed, i.e. in registers)
            long long b2 = tb, c = n;
            if (p.n_batched) { b2 = n / p.Nb; c = n - b2 * p.Nb; }
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float val = has_k ? __uint_as_float(r[v * 4 + e]) * p.alpha : 0.f;
              if (add_bias) {
                const long long ce = c + e;
                if (p.row_bias) val += rb * (p.col_scale ? p.col_scale[ce < p.Nb ? ce : 0] : 1.f);
                if (cbias) val += cbias[ce < p.Nb ? ce : 0];
              }
              o[e] = val;
            }
            const long long off = b2 * p.d_bstride + m * p.ldd + c;
            const int valid = (int)min((long long)4, p.N - n);
            if (p.atomic) {
              float* dp = reinterpret_cast<float*>(p.D) + off;
#pragma unroll
              for (int e = 0; e < 4; ++e) if (e < valid) atomicAdd(dp + e, o[e]);
            } else if (p.d_bf16) {
              __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.D) + off;
              if (p.acc_bf16) {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (e < valid) o[e] += __bfloat162float(dp[e]);
              }
              if (valid == 4) store4<__nv_bfloat16>(dp, o);
              else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (e < valid) dp[e] = __float2bfloat16_rn(o[e]);
              }
            } else {
              float* dp = reinterpret_cast<float*>(p.D) + off;
              if (valid == 4) store4<float>(dp, o);
              else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (e < valid) dp[e] = o[e];
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));
      if (++acc == p.nacc) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// SIMT cross-check kernel (same semantics, arbitrary strides).  impl=1 only; not a product path.
// ------------------------------------------------------------------------------------------------
struct SimtParams {
  const void* A; const void* B; void* D;
  int bf16, d_bf16;
  long long M, N, K, Nb;
  int a_kmajor, b_kmajor, a_batch, b_batch, n_batched, b_nbatched, full_batch;
  long long lda, ldb, ldd, a_bstride, b_bstride, d_bstride;
  float alpha;
  const float* row_bias; const float* col_scale; const float* col_bias; long long cb_bstride;
  int accumulate;
};

template <typename T>
__global__ void gemm_simt_kernel(const SimtParams p) {
  const long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long m = blockIdx.y * (long long)blockDim.y + threadIdx.y;
  if (m >= p.M || n >= p.N) return;
  const T* A = reinterpret_cast<const T*>(p.A);
  const T* B = reinterpret_cast<const T*>(p.B);
  const long long fbz = p.full_batch > 1 ? blockIdx.z : 0;
  if (p.full_batch > 1) { A += fbz * p.a_bstride; B += fbz * p.b_bstride; }
  long long bb = fbz, c = n;
  if (p.n_batched) { bb = n / p.Nb; c = n - bb * p.Nb; }
  const int kbatches = (p.a_kmajor && p.b_kmajor && p.full_batch <= 1) ? p.a_batch : 1;
  float acc = 0.f;
  for (int kb = 0; kb < kbatches; ++kb) {
    for (long long k = 0; k < p.K; ++k) {
      float a = p.a_kmajor ? to_f32(A[kb * p.a_bstride + m * p.lda + k]) : to_f32(A[k * p.lda + m]);
      float b = p.b_kmajor ? to_f32(B[kb * p.b_bstride + n * p.ldb + k])
                           : (p.b_nbatched ? to_f32(B[bb * p.b_bstride + k * p.ldb + c]) : to_f32(B[k * p.ldb + n]));
      acc = fmaf(a, b, acc);
    }
  }
  float val = acc * p.alpha;
  if (p.row_bias) val += p.row_bias[m] * (p.col_scale ? p.col_scale[c] : 1.f);
  if (p.col_bias) val += p.col_bias[fbz * p.cb_bstride + c];
  const long long off = bb * p.d_bstride + m * p.ldd + c;
  if (p.d_bf16) {
    __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.D) + off;
    *dp = __float2bfloat16_rn(p.accumulate ? val + __bfloat162float(*dp) : val);
  }
  else if (p.accumulate) reinterpret_cast<float*>(p.D)[off] += val;
  else reinterpret_cast<float*>(p.D)[off] = val;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  });
  return fn;
}

static int swizzle_for(long long contiguous_bytes_per_batch, bool may_overrun) {
  // may_overrun: the extent is bounded by the tensor map (OOB zero fill) so any swizzle is correct.
  if (may_overrun || contiguous_bytes_per_batch % 128 == 0) return 128;
  if (contiguous_bytes_per_batch % 64 == 0) return 64;
  if (contiguous_bytes_per_batch % 32 == 0) return 32;
  return 0;
}

// swizzle code: 128/64/32 = classic 16-byte-atom swizzles; 129 = 128-byte span with 32-byte atoms
// (UMMA SWIZZLE_128B_BASE32B / TMA SWIZZLE_128B_ATOM_32B), the only layout tcgen05 accepts for MN-major TF32.
static CUtensorMapSwizzle cu_swizzle(int s) {
  return s == 129 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
       : s == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : s == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}
static uint64_t layout_bits(int s) { return (uint64_t)(s == 129 ? 1 : s == 128 ? 2 : s == 64 ? 4 : 6) << 61; }
static int swz_bytes(int s) { return s == 129 ? 128 : s; }
static int swz_atom_rows(int s) { return s == 129 ? 4 : 8; }

static int encode_map(CUtensorMap* map, const void* base, bool bf16, const cuuint64_t dims[3], const cuuint64_t strides_bytes[2],
                      const cuuint32_t box[3], int swz_code) {
  const int swz = swz_code;
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available from the driver"); return AVMOE_E_NODRIVER; }
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, const_cast<void*>(base), dims,
                  strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, cu_swizzle(swz), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(%u,%u,%u) swz=%d", (int)r,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
              (unsigned long long)strides_bytes[0], (unsigned long long)strides_bytes[1], box[0], box[1], box[2], swz);
    return AVMOE_E_BADARG;
  }
  return 0;
}

static int launch_simt(const avmoe_gemm_args* a, cudaStream_t st) {
  SimtParams p;
  p.A = a->A; p.B = a->B; p.D = a->D;
  p.bf16 = a->dtype == AVMOE_BF16; p.d_bf16 = a->d_dtype == AVMOE_BF16;
  p.M = a->M; p.N = a->N; p.K = a->K; p.Nb = a->Nb > 0 ? a->Nb : a->N;
  p.a_kmajor = a->a_kmajor; p.b_kmajor = a->b_kmajor; p.a_batch = a->a_batch; p.b_batch = a->b_batch;
  p.n_batched = (!a->b_kmajor && a->b_batch > 1) || (a->Nb > 0 && a->Nb < a->N);
  p.b_nbatched = (!a->b_kmajor && a->b_batch > 1);
  p.lda = a->lda; p.ldb = a->ldb; p.ldd = a->ldd;
  p.a_bstride = a->a_bstride; p.b_bstride = a->b_bstride; p.d_bstride = a->d_bstride;
  p.alpha = a->alpha; p.row_bias = a->row_bias; p.col_scale = a->col_scale; p.col_bias = a->col_bias;
  p.cb_bstride = a->full_batch > 1 ? a->col_bias_bstride : 0;
  p.accumulate = a->accumulate;
  p.full_batch = a->full_batch > 1 ? a->full_batch : 1;
  if (p.full_batch > 1) { p.a_batch = p.b_batch = 1; }
  dim3 block(32, 8);
  dim3 grid((unsigned)((a->N + 31) / 32), (unsigned)((a->M + 7) / 8), (unsigned)p.full_batch);
  if (p.bf16) gemm_simt_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(p);
  else gemm_simt_kernel<float><<<grid, block, 0, st>>>(p);
  AVMOE_LAUNCH_CHECK("gemm_simt_kernel");
  return 0;
}

// Tile plan of one operand.  `code` is a swizzle code (see cu_swizzle); returns false when tcgen05/TMA cannot
// express the operand (the caller then uses the generic SIMT kernel).
struct OperandPlan {
  int code = 0;       // swizzle code
  int chunk = 0;      // MN elements per TMA box (MN-major)
};

static bool plan_operand(bool kmajor, bool bf16, long long contiguous_extent, bool bounded_by_map, OperandPlan* out) {
  const int es = bf16 ? 2 : 4;
  if (kmajor) { out->code = 128; return true; }  // narrowed later if the per-batch K needs it
  if (!bf16) {  // TF32 MN-major: only the 128B-span / 32B-atom layout exists
    if (!bounded_by_map && (contiguous_extent * es) % 128 != 0) return false;
    out->code = 129; out->chunk = 32;
    return true;
  }
  int sw = swizzle_for(contiguous_extent * es, bounded_by_map);
  if (!sw) return false;
  out->code = sw; out->chunk = sw / es;
  return true;
}

static uint64_t make_desc_base(int code, bool kmajor, int BK) {
  const int sb = swz_bytes(code);
  const uint64_t sbo = (uint64_t)((swz_atom_rows(code) * sb) >> 4);
  const uint64_t lbo = kmajor ? 1ull : (uint64_t)((BK * sb) >> 4);
  return layout_bits(code) | (1ull << 46) | (sbo << 32) | (lbo << 16);
}

// cuTensorMapEncodeTiled is a DRIVER call and needs a current context on the calling thread.  A thread that has only ever been
// handed a device by torch (nn.DataParallel's per-replica worker threads, AVMOE/AVVP/main.py:420-421) may not have one bound yet
// (CUDA_ERROR_INVALID_CONTEXT); a runtime call that touches the device binds the primary context.  Once per thread and device.
static int ensure_context() {
  static thread_local int bound_dev = -1;
  int dev = 0;
  AVMOE_CUDA(cudaGetDevice(&dev));
  if (dev != bound_dev) {
    // never during a stream capture (cudaFree is a prohibited call there, and a capturing thread necessarily has a context)
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    const bool capturing = cudaStreamIsCapturing(cudaStreamPerThread, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone;
    (void)cudaGetLastError();
    if (!capturing && getenv("AVMOE_NO_CTX_BIND") == nullptr) AVMOE_CUDA(cudaFree(nullptr));
    bound_dev = dev;
  }
  return 0;
}

int gemm_dispatch(const avmoe_gemm_args* a, cudaStream_t st) {
  if (int rc = ensure_context()) return rc;
  AVMOE_CHECK_ARG(a && a->A && a->B && a->D, "gemm: null pointer");
  AVMOE_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "gemm: empty problem M=%lld N=%lld K=%lld", (long long)a->M, (long long)a->N, (long long)a->K);
  AVMOE_CHECK_ARG(a->dtype == AVMOE_F32 || a->dtype == AVMOE_BF16, "gemm: bad dtype");
  AVMOE_CHECK_ARG(a->a_batch >= 1 && a->b_batch >= 1, "gemm: batch counts must be >= 1");
  const int fbatch = a->full_batch > 1 ? a->full_batch : 1;
  if (fbatch > 1) AVMOE_CHECK_ARG(a->a_batch == 1 && a->b_batch == 1 && (a->Nb == 0 || a->Nb == a->N) && a->split_k <= 1,
                                  "gemm: full_batch excludes K-/N-batching and split-K");
  const bool kbatched = a->a_kmajor && a->b_kmajor && a->a_batch > 1;
  if (a->a_kmajor && a->b_kmajor) AVMOE_CHECK_ARG(a->a_batch == a->b_batch, "gemm: K-batched operands need equal batch counts");
  if (!a->a_kmajor) AVMOE_CHECK_ARG(a->a_batch == 1, "gemm: MN-major A cannot be batched");
  if (!a->b_kmajor && a->a_kmajor) AVMOE_CHECK_ARG(a->a_batch == 1, "gemm: N-batched B needs an unbatched A");
  const bool b_nbatched = (!a->b_kmajor && a->b_batch > 1);      // B is [b][K][Nb]
  const bool n_batched = b_nbatched || (a->Nb > 0 && a->Nb < a->N);  // D is [N/Nb][M][Nb]
  const long long Nb = n_batched ? a->Nb : a->N;
  if (n_batched) AVMOE_CHECK_ARG(Nb > 0 && a->N % Nb == 0, "gemm: N must be a multiple of Nb");
  if (b_nbatched) AVMOE_CHECK_ARG(Nb * a->b_batch == a->N, "gemm: N must equal b_batch*Nb");
  const bool acc_bf16 = a->accumulate && a->d_dtype == AVMOE_BF16;     // plain read-modify-write in the epilogue
  if (acc_bf16) AVMOE_CHECK_ARG(a->split_k <= 1, "gemm: accumulate into a bf16 D excludes split-K");
  const bool atomic = a->split_k > 1 || (a->accumulate && !acc_bf16);
  if (atomic) AVMOE_CHECK_ARG(a->d_dtype == AVMOE_F32, "gemm: split-K needs an fp32 D");
  if (a->impl == 1) return launch_simt(a, st);

  const bool bf16 = a->dtype == AVMOE_BF16;
  const int es = bf16 ? 2 : 4;
  const int umma_k = 32 / es;  // 16 (bf16) / 8 (tf32)
  // ---- can the tensor-core path express this problem? ------------------------------------------
  OperandPlan pa, pb;
  const char* why = nullptr;
  if (!plan_operand(a->a_kmajor, bf16, a->a_kmajor ? a->K : a->M, true, &pa)) why = "operand A extent";
  if (!why && !plan_operand(a->b_kmajor, bf16, a->b_kmajor ? a->K : Nb, !b_nbatched, &pb)) why = "operand B per-batch extent";
  if (!why && ((a->lda * es) % 16 || (a->ldb * es) % 16 || ((uintptr_t)a->A % 16) || ((uintptr_t)a->B % 16))) why = "operand alignment";
  if (!why && ((a->a_batch > 1 || fbatch > 1) && (a->a_bstride * es) % 16)) why = "A batch pitch";
  if (!why && ((a->b_batch > 1 || fbatch > 1) && (a->b_bstride * es) % 16)) why = "B batch pitch";
  if (!why && (a->ldd % 4 || a->d_bstride % 4 || ((uintptr_t)a->D % 16))) why = "D alignment";
  if (!why && n_batched && Nb % 4) why = "per-batch N not a multiple of 4";
  if (why) {
    if (a->impl == 2) { set_error("gemm: tcgen05 path unsupported (%s)", why); return AVMOE_E_UNSUPPORTED; }
    return launch_simt(a, st);  // generic small-shape kernel (still CUDA; there is no CPU path)
  }
  // K-batched with a per-batch K that is not a multiple of the 128-byte stage: use a narrower swizzle to avoid
  // multiplying zero padding (K=96 bf16 -> 64B rows).
  if (kbatched && (a->K * es) % 128 != 0) {
    int s = swizzle_for(a->K * es, false);
    if (s) { pa.code = s; pb.code = s; }
  }
  int BK = 128 / es;
  if (a->a_kmajor && pa.code / es < BK) BK = pa.code / es;
  if (a->b_kmajor && pb.code / es < BK) BK = pb.code / es;
  if (a->a_kmajor) pa.code = BK * es;
  if (a->b_kmajor) pb.code = BK * es;
  // ---- tile N ---------------------------------------------------------------------------------
  int BN = 256;
  if (a->N < 256) {
    BN = (int)((a->N + 31) / 32) * 32;
  } else {
    // equal-width column tiles: N = 384 runs as 2 x 192 instead of 256 + a half-empty 128 (same tile count, a third less MMA and
    // epilogue work on the critical CTA)
    static const bool balance = []() { const char* e = getenv("AVMOE_GEMM_BN_BALANCE"); return !(e && e[0] == '0'); }();
    if (balance) {
      const long long nt = (a->N + 255) / 256;
      BN = (int)((((a->N + nt - 1) / nt) + 31) / 32) * 32;
    }
  }
  if (!a->b_kmajor) {
    BN = (BN / pb.chunk) * pb.chunk;
    if (BN == 0) BN = pb.chunk;
    if (BN % 32) BN = ((BN + 31) / 32) * 32;   // chunk is 16 only for 32B swizzle; keep the epilogue granularity
  }

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.BN = BN; p.BK = BK; p.mma_per_stage = BK / umma_k;
  // ---- operand A ------------------------------------------------------------------------------
  {
    cuuint64_t dims[3]; cuuint64_t strides[2]; cuuint32_t box[3];
    const int sb = swz_bytes(pa.code);
    p.a.desc_base = make_desc_base(pa.code, a->a_kmajor, BK);
    if (a->a_kmajor) {
      dims[0] = a->K; dims[1] = a->M; dims[2] = fbatch > 1 ? fbatch : a->a_batch;
      strides[0] = a->lda * es; strides[1] = ((a->a_batch > 1 || fbatch > 1) ? a->a_bstride : a->lda * a->M) * es;
      box[0] = BK; box[1] = kBM; box[2] = 1;
      p.a.kmajor = 1; p.a.nchunks = 1; p.a.chunk_bytes = kBM * sb; p.a.k_advance = 32;
      p.stage_bytes_a = kBM * sb;
    } else {
      dims[0] = a->M; dims[1] = a->K; dims[2] = fbatch;
      strides[0] = a->lda * es; strides[1] = (fbatch > 1 ? a->a_bstride : a->lda * a->K) * es;
      box[0] = pa.chunk; box[1] = BK; box[2] = 1;
      p.a.kmajor = 0; p.a.nchunks = kBM / pa.chunk; p.a.chunk_elems = pa.chunk; p.a.chunk_bytes = BK * sb;
      p.a.k_advance = umma_k * sb;
      p.stage_bytes_a = p.a.nchunks * p.a.chunk_bytes;
    }
    int rc = encode_map(&p.tmA, a->A, bf16, dims, strides, box, pa.code);
    if (rc) return rc;
  }
  // ---- operand B ------------------------------------------------------------------------------
  {
    cuuint64_t dims[3]; cuuint64_t strides[2]; cuuint32_t box[3];
    const int sb = swz_bytes(pb.code);
    p.b.desc_base = make_desc_base(pb.code, a->b_kmajor, BK);
    if (a->b_kmajor) {
      dims[0] = a->K; dims[1] = a->N; dims[2] = fbatch > 1 ? fbatch : a->b_batch;
      strides[0] = a->ldb * es; strides[1] = ((a->b_batch > 1 || fbatch > 1) ? a->b_bstride : a->ldb * a->N) * es;
      box[0] = BK; box[1] = BN; box[2] = 1;
      p.b.kmajor = 1; p.b.nchunks = 1; p.b.chunk_bytes = BN * sb; p.b.k_advance = 32;
      p.stage_bytes_b = BN * sb;
    } else {
      dims[0] = Nb; dims[1] = a->K; dims[2] = fbatch > 1 ? fbatch : a->b_batch;
      strides[0] = a->ldb * es; strides[1] = ((a->b_batch > 1 || fbatch > 1) ? a->b_bstride : a->ldb * a->K) * es;
      box[0] = pb.chunk; box[1] = BK; box[2] = 1;
      p.b.kmajor = 0; p.b.nchunks = BN / pb.chunk; p.b.chunk_elems = pb.chunk; p.b.chunk_bytes = BK * sb;
      p.b.batch_extent = b_nbatched ? (int)Nb : 0;
      p.b.k_advance = umma_k * sb;
      p.stage_bytes_b = p.b.nchunks * p.b.chunk_bytes;
    }
    int rc = encode_map(&p.tmB, a->B, bf16, dims, strides, box, pb.code);
    if (rc) return rc;
  }
  // stage sizes must keep every tile base 1024-byte aligned (swizzle atom <= 8 rows x 128 bytes)
  p.stage_bytes_a = (p.stage_bytes_a + 1023) & ~1023u;
  p.stage_bytes_b = (p.stage_bytes_b + 1023) & ~1023u;
  // ---- instruction descriptor -----------------------------------------------------------------
  //  [4,6) D fmt = 1 (f32) | [7,10) A fmt | [10,13) B fmt (bf16=1, tf32=2) | bit15 A MN-major | bit16 B MN-major
  //  [17,23) N>>3 | [24,29) M>>4
  const uint32_t fmt = bf16 ? 1u : 2u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((a->a_kmajor ? 0u : 1u) << 15) | ((a->b_kmajor ? 0u : 1u) << 16) |
            ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
  // ---- K schedule -----------------------------------------------------------------------------
  const int kpb = (int)((a->K + BK - 1) / BK);
  p.k_iters_per_batch = kpb;
  p.k_iters = kbatched ? kpb * a->a_batch : kpb;
  p.splits = a->split_k > 1 ? a->split_k : 1;
  if (p.splits > p.k_iters) p.splits = p.k_iters;
  p.k_iters_per_split = (p.k_iters + p.splits - 1) / p.splits;
  p.splits = (p.k_iters + p.k_iters_per_split - 1) / p.k_iters_per_split;
  p.tiles_m = (int)((a->M + kBM - 1) / kBM);
  p.tiles_n = (int)((a->N + BN - 1) / BN);
  p.full_batch = fbatch;
  p.nacc = BN <= 128 ? kMaxAcc : 2;
  int tc = 32;
  while (tc < p.nacc * BN) tc <<= 1;
  p.tmem_cols = tc;
  // ---- epilogue -------------------------------------------------------------------------------
  p.D = a->D; p.d_bf16 = a->d_dtype == AVMOE_BF16;
  p.M = a->M; p.N = a->N; p.Nb = Nb; p.ldd = a->ldd; p.d_bstride = a->d_bstride; p.n_batched = n_batched;
  p.alpha = a->alpha; p.row_bias = a->row_bias; p.col_scale = a->col_scale; p.col_bias = a->col_bias;
  p.cb_bstride = a->full_batch > 1 ? a->col_bias_bstride : 0;
  p.atomic = atomic;
  p.acc_bf16 = acc_bf16 ? 1 : 0;
  if (p.splits > 1 && !a->accumulate) {
    if (n_batched) {
      AVMOE_CHECK_ARG(a->d_bstride == a->M * a->ldd && a->ldd == Nb, "gemm: split-K needs a dense D");
      AVMOE_CUDA(cudaMemsetAsync(a->D, 0, sizeof(float) * a->M * a->N, st));
    } else {
      AVMOE_CUDA(cudaMemset2DAsync(a->D, a->ldd * sizeof(float), 0, a->N * sizeof(float), a->M, st));
    }
  }
  // ---- epilogue mode: 32x32 slabs must never straddle a batch of an N-batched D ----
  p.slab_path = (!n_batched || Nb % 32 == 0) ? 1 : 0;
  {
    const int vec = p.d_bf16 ? 8 : 4;
    p.vec_ok = (a->ldd % vec == 0) && (a->d_bstride % vec == 0) && (((uintptr_t)a->D % 16) == 0) && (!n_batched || Nb % vec == 0);
  }
  // ---- smem / launch --------------------------------------------------------------------------
  const uint32_t stage_bytes = p.stage_bytes_a + p.stage_bytes_b;
  const int max_smem = 227 * 1024;
  int stages = (max_smem - 1024 /*align slack*/ - 512 /*barriers*/ - kStagingBytes) / (int)stage_bytes;
  if (stages > 8) stages = 8;
  AVMOE_CHECK_ARG(stages >= 2, "gemm: tile does not fit shared memory");
  p.num_stages = stages;
  const size_t smem_bytes = 1024 + (size_t)stages * stage_bytes + kStagingBytes + 512;
  const int tiles = p.tiles_m * p.tiles_n * p.splits * p.full_batch;
  int grid = num_sms();
  if (grid > tiles) grid = tiles;
  // The dynamic-shared-memory opt-in is a per-device (per-context) function attribute: a single-process multi-GPU host
  // (nn.DataParallel, AVMOE/AVVP/main.py:420-421) needs it set once on EVERY device it launches on.
  {
    static std::atomic<unsigned char> attr_done[64];
    int dev = 0;
    AVMOE_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev].load(std::memory_order_acquire)) {
      AVMOE_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      AVMOE_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      if (dev >= 0 && dev < 64) attr_done[dev].store(1, std::memory_order_release);
    }
  }
  if (bf16) gemm_tcgen05_kernel<false><<<grid, kGemmThreads, smem_bytes, st>>>(p);
  else gemm_tcgen05_kernel<true><<<grid, kGemmThreads, smem_bytes, st>>>(p);
  AVMOE_LAUNCH_CHECK("gemm_tcgen05_kernel");
  return 0;
}

}  // namespace avmoe

extern "C" int avmoe_gemm(const avmoe_gemm_args* args, avmoe_stream_t stream) {
  return avmoe::gemm_dispatch(args, reinterpret_cast<cudaStream_t>(stream));
}
