// Micro-benchmark: issue rate of the int4 -> fp16 dequant sequence (registers only) at 1..16 warps per SM,
// plus variants, to size the dequant warps of gemm_w4a16.cu.   nvcc -arch=sm_100a -O3 -o dequant_rate dequant_rate.cu
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ void dequant_word(uint32_t w, __half2 z1024, __half2 z64, __half2 scale, uint32_t* out) {
  const uint32_t kMagic = 0x64006400u;
  const __half2 k16th = __floats2half2_rn(0.0625f, 0.0625f);
  uint32_t q0 = lop3_and_or(w, 0x000f000fu, kMagic);
  uint32_t q1 = lop3_and_or(w, 0x00f000f0u, kMagic);
  const uint32_t w8 = w >> 8;
  uint32_t q2 = lop3_and_or(w8, 0x000f000fu, kMagic);
  uint32_t q3 = lop3_and_or(w8, 0x00f000f0u, kMagic);
  __half2 h0 = __hsub2(*reinterpret_cast<__half2*>(&q0), z1024);
  __half2 h1 = __hfma2(*reinterpret_cast<__half2*>(&q1), k16th, z64);
  __half2 h2 = __hsub2(*reinterpret_cast<__half2*>(&q2), z1024);
  __half2 h3 = __hfma2(*reinterpret_cast<__half2*>(&q3), k16th, z64);
  h0 = __hmul2(h0, scale); h1 = __hmul2(h1, scale); h2 = __hmul2(h2, scale); h3 = __hmul2(h3, scale);
  out[0] = *reinterpret_cast<uint32_t*>(&h0); out[1] = *reinterpret_cast<uint32_t*>(&h1);
  out[2] = *reinterpret_cast<uint32_t*>(&h2); out[3] = *reinterpret_cast<uint32_t*>(&h3);
}

template <int MODE>
__global__ void k(uint32_t* out, long long* cyc, int iters, uint32_t seed) {
  uint32_t w[8];
  for (int r = 0; r < 8; ++r) w[r] = seed * (threadIdx.x + 1) + r * 0x9e3779b9u;
  __half2 z1024 = __floats2half2_rn(1029.f, 1029.f), z64 = __floats2half2_rn(-69.f, -69.f), sc = __floats2half2_rn(0.01f, 0.01f);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t v[32];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (MODE == 0) dequant_word(w[r], z1024, z64, sc, &v[r * 4]);
      if (MODE == 1) {  // LOP3/SHF only
        v[r*4] = lop3_and_or(w[r], 0x000f000fu, 0x64006400u); v[r*4+1] = lop3_and_or(w[r], 0x00f000f0u, 0x64006400u);
        uint32_t w8 = w[r] >> 8; v[r*4+2] = lop3_and_or(w8, 0x000f000fu, 0x64006400u); v[r*4+3] = lop3_and_or(w8, 0x00f000f0u, 0x64006400u);
      }
      if (MODE == 2) {  // half2 math only (8 ops)
        __half2 a = *reinterpret_cast<__half2*>(&w[r]);
        __half2 h0 = __hsub2(a, z1024), h1 = __hfma2(a, sc, z64), h2 = __hsub2(a, z64), h3 = __hfma2(a, z64, z1024);
        h0 = __hmul2(h0, sc); h1 = __hmul2(h1, sc); h2 = __hmul2(h2, sc); h3 = __hmul2(h3, sc);
        v[r*4] = *reinterpret_cast<uint32_t*>(&h0); v[r*4+1] = *reinterpret_cast<uint32_t*>(&h1);
        v[r*4+2] = *reinterpret_cast<uint32_t*>(&h2); v[r*4+3] = *reinterpret_cast<uint32_t*>(&h3);
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) acc ^= v[j];
#pragma unroll
    for (int r = 0; r < 8; ++r) w[r] += acc;   // dependency to the next iteration, keeps the work alive
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode) {
    for (int warps : {1, 4, 8, 16, 32}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, iters, 12345u);
        if (mode == 1) k<1><<<148, warps * 32>>>(out, cyc, iters, 12345u);
        if (mode == 2) k<2><<<148, warps * 32>>>(out, cyc, iters, 12345u);
        cudaDeviceSynchronize();
      }
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double c = (double)h[0] / iters;
      printf("mode %d warps/SM %2d: %.1f cycles per 8-word iteration per warp -> %.1f cycles per SM per 128x128 unit\n", mode, warps, c,
             c * 8.0 / (warps < 4 ? 1 : 1) * (8.0 / warps) );
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
