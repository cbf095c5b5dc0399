// Consumer side of the deferred split-K reduction (include/b200_tgis.h "B200SplitK"): device helpers that read an element
// of a GEMM output as the sum of its contributors' fp32 partials, in contributor order, exactly like the GEMMs' own fix-up.
#pragma once
#include "common.cuh"
#include "../../include/b200_tgis.h"

namespace b200 {

struct SplitKRef {
  const float* base;  // first contributor's value
  int stride;         // floats between contributors
  int n_contrib;
};

// element (t, n); n is a column of the GEMM output [T, N] in its natural numbering
__device__ __forceinline__ SplitKRef splitk_ref(const B200SplitK& d, int t, int n) {
  const int tile = n >> 7;
  int unit, r;
  if (d.half_tiles > 0) {  // gate|up layout: unit u = tiles (u, u + half_tiles)
    r = tile >= d.half_tiles ? 1 : 0;
    unit = tile - r * d.half_tiles;
  } else {
    unit = tile / d.tiles_per_unit;
    r = tile - unit * d.tiles_per_unit;
  }
  const int c_first = (unit * d.nkb) / d.units_per_cta;
  const int c_last = ((unit + 1) * d.nkb - 1) / d.units_per_cta;
  SplitKRef ref;
  ref.stride = d.tiles_per_unit * d.tn * 128;
  ref.base = d.partial + ((size_t)unit * d.max_contrib * d.tiles_per_unit + r) * (size_t)(d.tn * 128) + (size_t)t * 128 + (n & 127);
  ref.n_contrib = c_last - c_first + 1;
  return ref;
}

// 4 consecutive columns n .. n + 3 (n % 4 == 0), bias included: the fp32 value the GEMM would round to fp16
__device__ __forceinline__ float4 splitk_sum4(const B200SplitK& d, int t, int n) {
  const SplitKRef ref = splitk_ref(d, t, n);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c0 = 0; c0 < ref.n_contrib; c0 += 8) {  // 8 loads in flight (L2 latency bound), summed in contributor order
    float4 ld[8];
#pragma unroll
    for (int cc = 0; cc < 8; ++cc)
      ld[cc] = c0 + cc < ref.n_contrib ? __ldcg(reinterpret_cast<const float4*>(ref.base + (size_t)(c0 + cc) * ref.stride))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      if (c0 + cc < ref.n_contrib) { acc.x += ld[cc].x; acc.y += ld[cc].y; acc.z += ld[cc].z; acc.w += ld[cc].w; }
    }
  }
  if (d.bias) {
    const __half* b = reinterpret_cast<const __half*>(d.bias) + n;
    acc.x += __half2float(b[0]); acc.y += __half2float(b[1]); acc.z += __half2float(b[2]); acc.w += __half2float(b[3]);
  }
  return acc;
}

__device__ __forceinline__ float splitk_sum1(const B200SplitK& d, int t, int n) {
  const SplitKRef ref = splitk_ref(d, t, n);
  float acc = 0.f;
  for (int c0 = 0; c0 < ref.n_contrib; c0 += 8) {
    float ld[8];
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) ld[cc] = c0 + cc < ref.n_contrib ? __ldcg(ref.base + (size_t)(c0 + cc) * ref.stride) : 0.f;
#pragma unroll
    for (int cc = 0; cc < 8; ++cc)
      if (c0 + cc < ref.n_contrib) acc += ld[cc];
  }
  if (d.bias) acc += __half2float(reinterpret_cast<const __half*>(d.bias)[n]);
  return acc;
}

}  // namespace b200
