// One-shot all-reduce over NVLink peer memory for the tensor-parallel layer boundary.
//
// Replaces (when enabled, B200_P2P_ALLREDUCE=1): torch.distributed.all_reduce after the row-parallel o_proj / down_proj and the
// vocab-parallel embedding (/root/reference/server/text_generation_server/utils/layers.py:303-306 TensorParallelRowLinear,
// :343-345 TensorParallelEmbedding; flash_llama_modeling.py:296, :335) for the decode-sized messages of the step
// (bs 64 x hidden 4096 fp16 = 512 KB): NCCL costs 10-15 us per call there, twice per layer.
//
// EXPERIMENTAL: written without multi-GPU time left in the round, off by default, NCCL stays the product path until
// tests/test_gpu_experimental.py has passed on a 2-GPU box.
//
// Design.  Every rank owns a *window* in its own HBM, allocated here with cudaMalloc and exported over CUDA IPC:
//     [kBlocks][kMaxWorld] u32 arrival flags | [kBlocks] u32 epochs | pad to 4 KB | 2 slots x max_bytes of fp16 data
// The message is cut into fixed 8 KB chunks; chunk j always belongs to block j % kBlocks and always lives at offset j * 8 KB
// of a slot, so a block only ever races with itself.  One launch, per block:
//   1. copy the block's chunks of the local partial sums into the local window (slot = epoch & 1), system-scope fence,
//   2. store `epoch` into the block's flag in every peer's window (st.release.sys over NVLink),
//   3. spin (bounded) until every peer's flag in the local window reached `epoch` (ld.acquire.sys),
//   4. read the chunks of all ranks' windows and add them in rank order 0 .. world-1 with fp32 accumulation, round once:
//      every rank computes bit-identical sums, which greedy decoding across lock-step shards relies on.
// Two slots make an end-of-call barrier unnecessary: a rank can only reach epoch e + 1 (and overwrite slot (e + 1) & 1, last
// used at e - 1) after every peer signalled epoch e, i.e. after every peer finished reading epoch e - 1.
// The epoch lives in device memory and is advanced by the kernel itself, so the launch is CUDA-graph replayable.
#include "common.cuh"
#include "../../include/b200_tgis.h"

#include <cstring>

namespace b200 {

constexpr int kP2PBlocks = 64;
constexpr int kP2PThreads = 512;
constexpr int kP2PMaxWorld = 8;
constexpr int kP2PChunkBytes = kP2PThreads * 16;  // one 16-byte vector per thread
constexpr int kP2PHeaderBytes = 4096;
constexpr unsigned long long kP2PSpinNs = 4000000000ull;  // a peer that is 4 s late is gone: trap instead of hanging the GPU

struct P2PContext {
  int world = 0, rank = 0;
  int64_t max_bytes = 0;
  unsigned char* window = nullptr;                   // local, cudaMalloc
  unsigned char* peer[kP2PMaxWorld] = {};            // every rank's window mapped here (peer[rank] == window)
  bool opened[kP2PMaxWorld] = {};
  unsigned char** peer_table = nullptr;              // device copy of `peer`
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_peer_v4(const void* p) {  // never served from a stale L1 line: the slots are re-used
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void add_h8(float (&acc)[8], const uint4& v) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    acc[2 * i] += f.x;
    acc[2 * i + 1] += f.y;
  }
}

__global__ void __launch_bounds__(kP2PThreads, 1)
p2p_allreduce_f16_kernel(unsigned char* const* __restrict__ peers, __half* __restrict__ data, int64_t n, int world, int rank,
                         int64_t slot_bytes) {
  __shared__ uint32_t s_epoch;
  pdl_launch_dependents();
  const int blk = blockIdx.x;
  unsigned char* mine = peers[rank];
  uint32_t* my_flags = reinterpret_cast<uint32_t*>(mine) + blk * kP2PMaxWorld;
  uint32_t* my_epoch = reinterpret_cast<uint32_t*>(mine) + kP2PBlocks * kP2PMaxWorld + blk;
  pdl_wait();  // `data` is the previous kernel's output; the epoch was written by the previous all-reduce of the stream
  if (threadIdx.x == 0) s_epoch = *my_epoch + 1;
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const int64_t slot_off = kP2PHeaderBytes + (int64_t)(epoch & 1) * slot_bytes;
  const int64_t n_bytes = n * 2;
  const int64_t n_chunks = (n_bytes + kP2PChunkBytes - 1) / kP2PChunkBytes;

  // 1. local partial sums -> local window
  for (int64_t c = blk; c < n_chunks; c += kP2PBlocks) {
    const int64_t off = c * kP2PChunkBytes + (int64_t)threadIdx.x * 16;
    if (off < n_bytes) {
      const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(data) + off);
      *reinterpret_cast<uint4*>(mine + slot_off + off) = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  // 2. tell every peer, 3. wait for every peer
  if (threadIdx.x < world && threadIdx.x != rank) {
    uint32_t* theirs = reinterpret_cast<uint32_t*>(peers[threadIdx.x]) + blk * kP2PMaxWorld + rank;
    st_release_sys(theirs, epoch);
    const unsigned long long t0 = global_timer_ns();
    while ((int32_t)(ld_acquire_sys(my_flags + threadIdx.x) - epoch) < 0) {
      if (global_timer_ns() - t0 > kP2PSpinNs) __trap();
    }
  }
  __syncthreads();
  // 4. sum in rank order
  for (int64_t c = blk; c < n_chunks; c += kP2PBlocks) {
    const int64_t off = c * kP2PChunkBytes + (int64_t)threadIdx.x * 16;
    if (off < n_bytes) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int r = 0; r < world; ++r) add_h8(acc, ld_peer_v4(peers[r] + slot_off + off));
      uint4 out;
      __half2* h = reinterpret_cast<__half2*>(&out);
#pragma unroll
      for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(acc[2 * i], acc[2 * i + 1]);
      *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(data) + off) = out;
    }
  }
  if (threadIdx.x == 0) *my_epoch = epoch;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_p2p_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

// Allocates this rank's window for messages of up to max_bytes and writes its IPC handle (b200_p2p_handle_bytes bytes).
extern "C" int b200_p2p_create(int64_t max_bytes, int world, int rank, void** ctx_out, void* handle_out) {
  if (!ctx_out || !handle_out || world < 2 || world > kP2PMaxWorld || rank < 0 || rank >= world || max_bytes <= 0) {
    b200_set_last_error("p2p_create: need 2 <= world <= 8, 0 <= rank < world, max_bytes > 0");
    return B200_ERR_ARG;
  }
  P2PContext* ctx = new P2PContext();
  ctx->world = world;
  ctx->rank = rank;
  ctx->max_bytes = (max_bytes + kP2PChunkBytes - 1) / kP2PChunkBytes * kP2PChunkBytes;
  const size_t bytes = kP2PHeaderBytes + 2 * (size_t)ctx->max_bytes;
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaMalloc(&ctx->window, bytes);
  if (e == cudaSuccess) e = cudaMemset(ctx->window, 0, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&ctx->peer_table, sizeof(unsigned char*) * kP2PMaxWorld);
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, ctx->window);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    b200_set_last_error(cudaGetErrorString(e));
    if (ctx->window) cudaFree(ctx->window);
    if (ctx->peer_table) cudaFree(ctx->peer_table);
    delete ctx;
    return B200_ERR_CUDA;
  }
  memcpy(handle_out, &h, sizeof(h));
  ctx->peer[rank] = ctx->window;
  *ctx_out = ctx;
  return B200_OK;
}

// handles: [world][b200_p2p_handle_bytes] gathered from every rank (any transport: torch.distributed all_gather).
extern "C" int b200_p2p_connect(void* ctx_, const void* handles) {
  P2PContext* ctx = (P2PContext*)ctx_;
  if (!ctx || !handles) { b200_set_last_error("p2p_connect: null argument"); return B200_ERR_ARG; }
  for (int r = 0; r < ctx->world; ++r) {
    if (r == ctx->rank || ctx->opened[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const unsigned char*)handles + (size_t)r * sizeof(h), sizeof(h));
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { b200_set_last_error(cudaGetErrorString(e)); return B200_ERR_CUDA; }
    ctx->peer[r] = (unsigned char*)p;
    ctx->opened[r] = true;
  }
  const cudaError_t e = cudaMemcpy(ctx->peer_table, ctx->peer, sizeof(unsigned char*) * kP2PMaxWorld, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { b200_set_last_error(cudaGetErrorString(e)); return B200_ERR_CUDA; }
  return B200_OK;
}

extern "C" int64_t b200_p2p_max_bytes(void* ctx_) { return ctx_ ? ((P2PContext*)ctx_)->max_bytes : 0; }

// In-place sum over ranks of `data` (fp16 [n], 16-byte aligned, n % 8 == 0, n * 2 <= max_bytes).  Every rank must make
// the same sequence of calls with the same n.
extern "C" int b200_p2p_allreduce_f16(void* ctx_, void* data, int64_t n, void* stream) {
  P2PContext* ctx = (P2PContext*)ctx_;
  if (!ctx || !data || n < 0) { b200_set_last_error("p2p_allreduce: null argument"); return B200_ERR_ARG; }
  if (n == 0) return B200_OK;
  if (n % 8 != 0 || ((uintptr_t)data & 15) != 0 || n * 2 > ctx->max_bytes) {
    b200_set_last_error("p2p_allreduce: need n % 8 == 0, 16-byte aligned data and n * 2 <= max_bytes");
    return B200_ERR_ARG;
  }
  for (int r = 0; r < ctx->world; ++r)
    if (!ctx->peer[r]) { b200_set_last_error("p2p_allreduce: b200_p2p_connect has not run"); return B200_ERR_ARG; }
  const int64_t n_chunks = (n * 2 + kP2PChunkBytes - 1) / kP2PChunkBytes;
  const int blocks = (int)(n_chunks < kP2PBlocks ? n_chunks : kP2PBlocks);
  B200_LAUNCH(p2p_allreduce_f16_kernel, dim3(blocks), dim3(kP2PThreads), 0, (cudaStream_t)stream,
              (unsigned char* const*)ctx->peer_table, (__half*)data, n, ctx->world, ctx->rank, ctx->max_bytes);
  b200_count_launches(1);
  return B200_OK;
}

extern "C" void b200_p2p_destroy(void* ctx_) {
  P2PContext* ctx = (P2PContext*)ctx_;
  if (!ctx) return;
  for (int r = 0; r < ctx->world; ++r)
    if (ctx->opened[r]) cudaIpcCloseMemHandle(ctx->peer[r]);
  if (ctx->peer_table) cudaFree(ctx->peer_table);
  if (ctx->window) cudaFree(ctx->window);
  delete ctx;
}
