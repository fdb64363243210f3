// Row-pass vocabulary shared by the expert kernels: one warp per token row, float4 column groups per lane.
#pragma once
#include "common.cuh"

namespace avmoe {

static constexpr float kLnEps = 1e-5f;
static constexpr double kBnEps = 1e-5;
static constexpr double kBnMomentum = 0.1;
static constexpr int kThreads = 256;

static constexpr int kMaxVLimit = 12;   // float4 vectors per lane per row: C <= 1536 (kernels are instantiated for 1,2,3,6,12)

// Row passes: one warp per token row, lane l owns the float4 column groups c = 4*l + 128*k  (k < kMaxV, c < C).
// No integer divisions, 4 elements per instruction, row reductions by warp shuffles.
#define ROW_VECS(k, c) _Pragma("unroll") for (int k = 0, c = lane * 4; k < kMaxV; ++k, c += 128) if (c < C)

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float hsum(float4 v) { return (v.x + v.y) + (v.z + v.w); }
__device__ __forceinline__ float4 f4(float a) { return make_float4(a, a, a, a); }
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator-(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 operator*(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 operator*(float a, float4 b) { return make_float4(a * b.x, a * b.y, a * b.z, a * b.w); }

// (scale, shift) or (mean, invstd) pairs are stored interleaved per channel: 4 channels = 2 float4
__device__ __forceinline__ void ldg_pairs4(const float* p, float4& first, float4& second) {
  const float4 a = ldg4(p), b = ldg4(p + 4);
  first = make_float4(a.x, a.z, b.x, b.z);
  second = make_float4(a.y, a.w, b.y, b.w);
}

template <typename T>
__device__ __forceinline__ float4 ldg_act4(const T* p);
template <>
__device__ __forceinline__ float4 ldg_act4<float>(const float* p) { return ldg4(p); }
template <>
__device__ __forceinline__ float4 ldg_act4<__nv_bfloat16>(const __nv_bfloat16* p) {
  float v[4];
  load4<__nv_bfloat16>(p, v);
  return make_float4(v[0], v[1], v[2], v[3]);
}
template <typename T>
__device__ __forceinline__ void st_act4(T* p, float4 v) {
  const float a[4] = {v.x, v.y, v.z, v.w};
  store4<T>(p, a);
}

// instantiate for the smallest per-lane vector count that covers C (register arrays are sized by it)
#define AVMOE_DISPATCH_V(C_, ...)                                 \
  do {                                                           \
    const int _v = ((C_) + 127) / 128;                           \
    if (_v <= 1) { constexpr int kV = 1; __VA_ARGS__ }           \
    else if (_v <= 2) { constexpr int kV = 2; __VA_ARGS__ }      \
    else if (_v <= 3) { constexpr int kV = 3; __VA_ARGS__ }      \
    else if (_v <= 6) { constexpr int kV = 6; __VA_ARGS__ }      \
    else { constexpr int kV = 12; __VA_ARGS__ }                  \
  } while (0)

}  // namespace avmoe
