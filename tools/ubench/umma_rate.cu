// Micro-benchmark: tcgen05.mma issue/execute rate for M = 128, K = 16 steps, N = 16..256, A operand in TMEM (.ts) vs in
// shared memory (.ss), one CTA per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../text-generation-inference_b200/csrc
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
using namespace b200;

template <int N, int TS>
__global__ void __launch_bounds__(128, 1) k(long long* cyc, int iters, int per_commit) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = umma_idesc_f16_f32acc(128, N);
    const uint64_t bdesc = umma_desc_kmajor_sw128(smem_u32(smem));
    const uint64_t adesc = umma_desc_kmajor_sw128(smem_u32(smem + 32 * 1024));
    long long t0 = clock64();
    int ph = 0;
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
        for (int k = 0; k < per_commit; ++k) {
          if (TS) umma_f16_ts(tm, tm + 256 + (k & 7) * 8, bdesc + (uint64_t)((k & 3) * 2), idesc, 1u);
          else umma_f16_ss(tm, adesc + (uint64_t)((k & 3) * 2), bdesc + (uint64_t)((k & 3) * 2), idesc, 1u);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, ph);
      ph ^= 1;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

// the GEMM's MMA-warp loop shape: per unit WAITS satisfied waits, fence, elect, 8 MMAs, one commit; completion is only
// awaited at the end.  MODE 0: mbarrier try_wait loop   1: ld.acquire.cta.shared flag poll   2: as 0, waits placed between the
// MMAs and the commit
__device__ __forceinline__ uint32_t lds_acquire(const void* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
template <int N, int WAITS, int MODE>
__global__ void __launch_bounds__(128, 1) kloop(long long* cyc, int units) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[8];
  __shared__ uint32_t slot, flag[4];
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); for (int i = 0; i < 4; ++i) flag[i] = 1; mbar_fence_init(); }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = umma_idesc_f16_f32acc(128, N);
    const uint64_t bdesc = umma_desc_kmajor_sw128(smem_u32(smem));
    long long t0 = clock64();
    for (int u = 0; u < units; ++u) {
      if (MODE != 2) {
#pragma unroll
        for (int w = 0; w < WAITS; ++w) {
          if (MODE == 0) mbar_wait(&bar[w], 1);  // fresh barrier: parity 1 is already complete
          else while (lds_acquire(&flag[w]) == 0) {}
        }
      }
      tcgen05_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_f16_ts(tm, tm + 256 + (u & 3) * 64 + k * 8, bdesc + (uint64_t)((k & 3) * 2), idesc, 1u);
      }
      if (MODE == 2) {
#pragma unroll
        for (int w = 0; w < WAITS; ++w) mbar_wait(&bar[w], 1);
      }
      if (elect_one()) umma_commit(&bar[4 + (u & 1)]);  // nobody waits on these until the end
      __syncwarp();
    }
    long long t1 = clock64();
    if (elect_one()) umma_commit(&bar[7]);
    __syncwarp();
    mbar_wait(&bar[7], 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) { cyc[blockIdx.x * 2] = t1 - t0; cyc[blockIdx.x * 2 + 1] = t2 - t0; }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

template <int N, int WAITS, int MODE>
void runloop1(long long* cyc) {
  cudaFuncSetAttribute(kloop<N, WAITS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int units = 400;
  for (int rep = 0; rep < 2; ++rep) { kloop<N, WAITS, MODE><<<148, 128, 64 * 1024>>>(cyc, units); cudaDeviceSynchronize(); }
  long long h[296];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("loop mode %d N=%3d waits/unit=%d: issue loop %6.1f cycles per unit of 8 MMAs, until complete %6.1f\n", MODE, N, WAITS,
         (double)h[0] / units, (double)h[1] / units);
}
template <int N>
void runloop(long long* cyc) {
  runloop1<N, 0, 0>(cyc); runloop1<N, 1, 0>(cyc); runloop1<N, 2, 0>(cyc); runloop1<N, 3, 0>(cyc);
  runloop1<N, 1, 1>(cyc); runloop1<N, 2, 1>(cyc); runloop1<N, 3, 1>(cyc);
  runloop1<N, 1, 2>(cyc); runloop1<N, 2, 2>(cyc);
}

template <int N, int TS>
void run(long long* cyc) {
  cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int per : {1, 8, 64}) {
    const int iters = 2000 / per + 10;
    for (int rep = 0; rep < 2; ++rep) { k<N, TS><<<148, 128, 64 * 1024>>>(cyc, iters, per); cudaDeviceSynchronize(); }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("N=%3d %s  %2d MMAs per commit+wait: %7.1f cycles per MMA (%6.1f per batch)\n", N, TS ? "A in TMEM" : "A in smem", per,
           (double)h[0] / iters / per, (double)h[0] / iters);
  }
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 16);
  runloop<64>(cyc);
  if (getenv("UMMA_ALL")) { run<16, 1>(cyc); run<32, 1>(cyc); run<64, 1>(cyc); run<128, 1>(cyc); run<256, 1>(cyc); run<64, 0>(cyc); run<128, 0>(cyc); run<256, 0>(cyc); }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
