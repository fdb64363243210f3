// Warp-level bf16 tensor-core tiles for the row-fused expert kernels (csrc/rowfused.cu).
//
// The expert / latent-attention passes between the big GEMMs are HBM-bound row passes whose contractions are skinny
// (K or N = T <= 96 latent tokens, C/r <= 96 bottleneck channels): each warp owns 16 token rows and keeps every intermediate of
// those rows in registers or in its private shared-memory tile, so nothing but the row's input and output touches HBM.
// For such tiles `mma.sync.m16n8k16` (one warp, operands in registers) is the right instrument: a tcgen05 tile is 128 rows
// issued by one thread with operands in shared memory and the accumulator in TMEM, which would force every per-row
// statistic (softmax over T, LayerNorm over C) through a TMEM round trip per step.  The dense contractions (token resampler,
// weight gradients over all rows) stay on the tcgen05 GEMM (csrc/gemm.cu).
//
// Fragment layouts (PTX ISA, mma.m16n8k16 .bf16; g = lane >> 2, t = lane & 3):
//   A 16x16 row-major : a0 (row g,   k 2t..2t+1)  a1 (row g+8, k 2t..2t+1)  a2 (row g,   k 2t+8..2t+9)  a3 (row g+8, k 2t+8..2t+9)
//   B 16x8            : b0 (k 2t..2t+1, n g)      b1 (k 2t+8..2t+9, n g)
//   C/D 16x8 fp32     : c0 (row g, n 2t)  c1 (row g, n 2t+1)  c2 (row g+8, n 2t)  c3 (row g+8, n 2t+1)
//
// B operands are small weight matrices shared by every row tile.  They are stored in global memory already in fragment
// order ("packed B", written by rf_pack_b_kernel): for n-tile pair p and k-step s, lane l owns one 16-byte word
//   { b0(tile 2p), b1(tile 2p), b0(tile 2p+1), b1(tile 2p+1) }   at   ((p * KS + s) * 32 + l) * 16 bytes,
// so a warp fetches the operands of two MMAs with ONE coalesced 512-byte request that hits L1/L2 (the packed weights of a
// call are a few tens of KB).  In "hi/lo" mode a pair is (hi(tile p), lo(tile p)): hi = bf16(w), lo = bf16(w - hi), and the two
// MMAs accumulate into the same tile, giving ~16 mantissa bits of the fp32 operand (un-scaled softmax logits need them).
#pragma once
#include "common.cuh"

namespace avmoe {
namespace rf {

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);   // .x (low half) = lo, .y = hi
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v);
  return make_float2(__low2float(b), __high2float(b));
}

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragment of a 16-row bf16 tile in shared memory (row pitch `pitch` bytes, 16-byte aligned rows), columns [k0, k0+16)
__device__ __forceinline__ void ldmatrix_a(uint32_t (&a)[4], const unsigned char* tile, int pitch, int k0, int lane) {
  const int row = (lane & 7) + ((lane >> 3) & 1) * 8;
  const int col = k0 + (lane >> 4) * 8;
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(tile + row * pitch + col * 2);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
               : "r"(addr)
               : "memory");
}

__device__ __forceinline__ uint4 ldg_b(const uint4* packed, int pair, int ks, int s, int lane) {
  return __ldg(packed + (((pair * ks + s) << 5) + lane));      // 32-bit index: a packed operand is far below 2^31 words
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// floor(i / d) for 0 <= i < 2^15 and 1 <= d <= 2^9 by a precomputed reciprocal (tile loops: i / chunks-per-row)
__device__ __forceinline__ uint32_t magic_of(uint32_t d) { return ((1u << 24) + d - 1) / d; }
__device__ __forceinline__ uint32_t magic_div(uint32_t i, uint32_t magic) { return (uint32_t)(((uint64_t)i * magic) >> 24); }

// sum over the 4 lanes of a quad (the lanes that share accumulator rows g / g+8)
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}

// Column sums over the 16 rows of a tile for 8 per-thread values v[0..7] (each already the sum of the thread's two rows):
// a reduce-scatter over the 8 lanes that share t (lane bits 2..4) in 4 + 2 + 1 shuffles.  Returns the sum of value index
// `rs_index(lane)` over the 8 row-lanes; every lane ends up owning exactly one of the 8 values.
__device__ __forceinline__ int rs_index(int lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }
__device__ __forceinline__ float reduce_scatter8(const float (&v)[8], int lane) {
  const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
  float a[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {   // keep half (bit 4 selects 4..7), send the other half
    const float keep = h4 ? v[4 + i] : v[i];
    const float send = h4 ? v[i] : v[4 + i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float b[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = h3 ? a[2 + i] : a[i];
    const float send = h3 ? a[i] : a[2 + i];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const float keep = h2 ? b[1] : b[0];
  const float send = h2 ? b[0] : b[1];
  return keep + __shfl_xor_sync(0xffffffffu, send, 4);
}

}  // namespace rf
}  // namespace avmoe
