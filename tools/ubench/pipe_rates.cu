// Micro-benchmark: issue rates of the instruction mixes the int4 dequant path can be built from, at 16 warps per SM
// (4 per SM sub-partition), to size gemm_w4a16.cu against the HBM budget (one 128x128 unit per 0.19 us per SM).
//   nvcc -arch=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t hmul2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t shr8(uint32_t a) {
  uint32_t r;
  asm volatile("shr.b32 %0, %1, 8;" : "=r"(r) : "r"(a));
  return r;
}
// ptxas removes dead results, so every pair of values is folded into an accumulator with one 3-input LOP3 (xor3);
// these count as ALU-pipe ops in the "ops" column below.
#define keep2(a, b) acc = xor3(acc, (a), (b))
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

// MODE 0: exact dequant of one word (13 ops)   1: LOP3/SHF part only (5)   2: FP part only (8)
//      3: 1-FMA-per-pair dequant (9 ops)        4: 8 HFMA2   5: 8 HMUL2   6: 8 HADD2   7: 8 LOP3   8: 4 LOP3 + 4 HFMA2 interleaved
template <int MODE>
__global__ void k(uint32_t* out, long long* cyc, int iters, uint32_t seed) {
  uint32_t w[8];
  for (int r = 0; r < 8; ++r) w[r] = seed * (threadIdx.x + 1) + r * 0x9e3779b9u;
  const uint32_t z1024 = 0xE405E405u, z64 = 0xD450D450u, sc = 0x21002100u, k16 = 0x2C002C00u, M = 0x64006400u;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      uint32_t x = w[r];
      if (MODE == 0) {
        uint32_t q0 = lop3_and_or(x, 0x000f000fu, M), q1 = lop3_and_or(x, 0x00f000f0u, M);
        uint32_t x8 = shr8(x);
        uint32_t q2 = lop3_and_or(x8, 0x000f000fu, M), q3 = lop3_and_or(x8, 0x00f000f0u, M);
        keep2(hmul2(hadd2(q0, z1024), sc), hmul2(hfma2(q1, k16, z64), sc));
        keep2(hmul2(hadd2(q2, z1024), sc), hmul2(hfma2(q3, k16, z64), sc));
      } else if (MODE == 1) {
        uint32_t q0 = lop3_and_or(x, 0x000f000fu, M), q1 = lop3_and_or(x, 0x00f000f0u, M);
        uint32_t x8 = shr8(x);
        uint32_t q2 = lop3_and_or(x8, 0x000f000fu, M), q3 = lop3_and_or(x8, 0x00f000f0u, M);
        keep2(q0, q1); keep2(q2, q3);
      } else if (MODE == 2) {
        keep2(hmul2(hadd2(x, z1024), sc), hmul2(hfma2(x, k16, z64), sc));
        keep2(hmul2(hadd2(x, z64), sc), hmul2(hfma2(x, sc, z64), sc));
      } else if (MODE == 3) {
        uint32_t q0 = lop3_and_or(x, 0x000f000fu, M), q1 = lop3_and_or(x, 0x00f000f0u, M);
        uint32_t x8 = shr8(x);
        uint32_t q2 = lop3_and_or(x8, 0x000f000fu, M), q3 = lop3_and_or(x8, 0x00f000f0u, M);
        keep2(hfma2(q0, sc, z1024), hfma2(q1, k16, z64)); keep2(hfma2(q2, sc, z1024), hfma2(q3, k16, z64));
      } else if (MODE == 4) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) keep2(hfma2(x, sc + j, z64), hfma2(x, sc + j + 1, z64));
      } else if (MODE == 5) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) keep2(hmul2(x, sc + j), hmul2(x, sc + j + 1));
      } else if (MODE == 6) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) keep2(hadd2(x, sc + j), hadd2(x, sc + j + 1));
      } else if (MODE == 7) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) keep2(lop3_and_or(x, 0x000f000fu + j, M), lop3_and_or(x, 0x000f000fu + j + 1, M));
      } else if (MODE == 8) {
#pragma unroll
        for (int j = 0; j < 4; ++j) keep2(lop3_and_or(x, 0x000f000fu + j, M), hfma2(x, sc + j, z64));
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) w[r] += acc;  // loop-carried: keeps every iteration alive
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = w[0] ^ acc;
}

template <int MODE>
void run(const char* name, int ops, uint32_t* out, long long* cyc) {
  const int iters = 4000;
  for (int warps : {4, 8, 16, 32}) {
    for (int rep = 0; rep < 2; ++rep) {
      k<MODE><<<148, warps * 32>>>(out, cyc, iters, 12345u);
      cudaDeviceSynchronize();
    }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double c = (double)h[0] / iters / 8.0;  // cycles per word-iteration per warp
    const double per_smsp = c / (warps / 4.0);    // cycles per warp-word per SM sub-partition
    printf("%-28s warps/SM %2d: %6.2f cyc per word per SMSP  (%4.2f instr/clk/SMSP)  -> %6.1f cyc per 128x128 unit per SM\n", name, warps,
           per_smsp, ops / per_smsp, per_smsp * 16.0);
  }
}

int main() {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  run<0>("exact dequant (13+2 ops)", 15, out, cyc);
  run<1>("LOP3x4+SHF (5+2)", 7, out, cyc);
  run<2>("FP part (8+2)", 10, out, cyc);
  run<3>("1-FMA dequant (9+2)", 11, out, cyc);
  run<4>("8 HFMA2 (+4)", 12, out, cyc);
  run<5>("8 HMUL2 (+4)", 12, out, cyc);
  run<6>("8 HADD2 (+4)", 12, out, cyc);
  run<7>("8 LOP3 (+4)", 12, out, cyc);
  run<8>("4 LOP3 + 4 HFMA2 (+4)", 12, out, cyc);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
