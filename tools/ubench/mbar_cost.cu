// Micro-benchmark: cost of a SATISFIED mbarrier wait (and of alternatives) for one warp, nothing else running.
#include "common.cuh"
#include <cstdio>
using namespace b200;

__device__ __forceinline__ uint32_t lds_volatile(const void* p) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_acquire(const void* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}

// mode 0: try_wait loop, all lanes   1: try_wait, lane 0 only   2: test_wait   3: ld.volatile.shared flag poll
// mode 4: ld.acquire.cta flag poll   5: try_wait on a barrier that has really completed phases (arrive each iteration)
template <int mode>
__global__ void k(long long* cyc, int iters) {
  __shared__ uint64_t bar[2];
  __shared__ uint32_t flag;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); flag = 1; mbar_fence_init(); }
  __syncthreads();
  uint32_t acc = 0;
  long long t0 = clock64();
  uint32_t ph = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
   for (int rep8 = 0; rep8 < 8; ++rep8) {
    if (mode == 0) mbar_wait(&bar[0], 1);
    else if (mode == 1) { if (threadIdx.x == 0) mbar_wait(&bar[0], 1); __syncwarp(); }
    else if (mode == 2) {
      uint32_t o;
      do {
        asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(o) : "r"(smem_u32(&bar[0])), "r"(1u) : "memory");
      } while (!o);
    } else if (mode == 3) { while (lds_volatile(&flag) == 0) {} acc += 1; }
    else if (mode == 4) { while (lds_acquire(&flag) == 0) {} acc += 1; }
    else if (mode == 5) {
      if (threadIdx.x == 0) mbar_arrive(&bar[1]);
      mbar_wait(&bar[1], ph);
      ph ^= 1;
    }
   }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0 + (acc == 12345u);
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 8);
  const char* names[] = {"try_wait loop (32 lanes, satisfied)", "try_wait loop (lane 0) + syncwarp", "test_wait (32 lanes)",
                         "ld.volatile.shared poll", "ld.acquire.cta.shared poll", "arrive + try_wait (real phases)"};
  for (int mode = 0; mode < 6; ++mode) {
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) {
      switch (mode) {
        case 0: k<0><<<148, 32>>>(cyc, iters); break;
        case 1: k<1><<<148, 32>>>(cyc, iters); break;
        case 2: k<2><<<148, 32>>>(cyc, iters); break;
        case 3: k<3><<<148, 32>>>(cyc, iters); break;
        case 4: k<4><<<148, 32>>>(cyc, iters); break;
        default: k<5><<<148, 32>>>(cyc, iters); break;
      }
      cudaDeviceSynchronize();
    }
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-40s %7.1f cycles per wait\n", names[mode], (double)h / iters / 8);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
