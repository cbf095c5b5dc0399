// One-shot all-reduce over NVLink peer memory for the tensor-parallel layer boundary.
//
// Replaces (default for decode-sized messages; B200_P2P_ALLREDUCE=0 keeps NCCL): torch.distributed.all_reduce after the row-parallel o_proj / down_proj and the
// vocab-parallel embedding (/root/reference/server/text_generation_server/utils/layers.py:303-306 TensorParallelRowLinear,
// :343-345 TensorParallelEmbedding; flash_llama_modeling.py:296, :335) for the decode-sized messages of the step
// (bs 64 x hidden 4096 fp16 = 512 KB): NCCL costs 10-15 us per call there, twice per layer.
//
// Validated on a 2-GPU B200 box in round 2 (tests/test_gpu_p2p.py, tests/test_gpu_tp.py; profiles/r2_tp2_and_splitk_validation.txt).
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
#include "splitk.cuh"
#include "../../include/b200_tgis.h"

#include <cstring>

namespace b200 {

constexpr int kP2PBlocks = 64;
constexpr int kP2PThreads = 512;
constexpr int kP2PMaxWorld = 8;
constexpr int kP2PChunkBytes = kP2PThreads * 16;  // one 16-byte vector per thread
constexpr int kP2PMaxRows = 256;  // fused all-reduce + norm: one block per token row
// header: [kP2PMaxRows][kP2PMaxWorld] u32 arrival flags (the chunked kernel uses the first kP2PBlocks rows) | [kP2PMaxRows] u32 epochs
constexpr int kP2PHeaderBytes = 16384;
static_assert(kP2PMaxRows * kP2PMaxWorld * 4 + kP2PMaxRows * 4 <= kP2PHeaderBytes, "p2p header");
constexpr unsigned long long kP2PSpinNs = 4000000000ull;  // a peer that is 4 s late is gone: trap instead of hanging the GPU

struct P2PContext {
  int world = 0, rank = 0;
  int family = 0;  // 0 unused yet, 1 chunked all-reduce, 2 fused row kernel: a window serves one kernel family (its slot / epoch
                   // bookkeeping is per block index and the two families cut the message differently)
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
  uint32_t* my_epoch = reinterpret_cast<uint32_t*>(mine) + kP2PMaxRows * kP2PMaxWorld + blk;
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

// ------------------------------------------------------------------------------------------------------------------
// Fused tensor-parallel layer boundary: all-reduce + residual add + RMSNorm in ONE kernel (decode-sized T).
//
// Replaces, per half-layer of the sharded Llama graph, three things of the reference: the row-parallel linear's
// torch.distributed.all_reduce (utils/layers.py:318-322), and the fused residual + RMSNorm that consumes its result
// (flash_llama_modeling.py:132-148) - and, when the row-parallel GEMM ran with deferred split-K reduction, the GEMM's own
// cross-CTA fix-up as well: the rank-partial row is summed from the fp32 partials straight into an NVLink window.
//
// Nothing downstream ever reads the all-reduced hidden state itself, only norm(hidden + residual) and the new residual, and
// the residual stream is only ever read by the next boundary.  So the exchange is a reduce-scatter by TOKEN ROW followed by a
// broadcast of the NORMED row: row t belongs to rank t % world, which alone keeps the residual stream of that row.
// Per row t, on every rank:
//   1. this rank's partial row (fp16 input, or sum of split-K partials rounded to fp16) -> stored straight into the OWNER's
//      window (inbox 1: [slot][t / world][source rank][H]), 16-byte stores over NVLink, no flag and no fence
//   2. owner block: poll inbox 1 until the row of every source rank has landed, add in rank order with fp32 accumulation, ONE
//      rounding to fp16 = the all-reduced hidden state; x = that + residual (fp32) -> residual_out[t]; RMSNorm(x) -> normed[t]
//      locally and stored into every peer's window (inbox 2: [slot][t][H])
//      other blocks: poll inbox 2 until the normed row landed, copy it to normed[t]
// "Landed" is read off the data itself (the Lamport trick): an inbox cell holds fp16 -0.0 in every element until written,
// senders write +0.0 for -0.0, and a receiver re-reads a 16-byte vector until none of its 8 elements is -0.0 - one NVLink
// one-way latency per hop instead of store / fence / flag / load round trips, and 1.75x the message per rank on the wire
// instead of 7x (world = 8).  Three slots per row, slot = calls % 3: a block clears the slot of the PREVIOUS call of its row
// (read only by itself, in an earlier kernel of the stream) for the call after the next; a peer can only write that slot again
// after it completed the next call, which needs this rank's push of the next call, which is stream-ordered after this
// kernel.  The per-row call counter lives in the window header and is advanced by the kernel: CUDA-graph replayable.
// Up to 256 rows every row has its own block; longer steps (prefill chunks up to 2048 rows) give each of <= 256 blocks a run of
// R = k * world consecutive rows, so every block owns R / world of its rows and the owner work is spread over all blocks.  A
// block pushes all its partial rows first, then does its owner work, then waits for the rows it does not own: owner work only
// ever waits for pushes, and all blocks of a launch are resident at once, so the wait graph has no cycle.
// Every rank computes nothing twice and every rank ends with bit-identical normed rows (they are copies).
// residual / residual_out are only read / written for the rows this rank owns.
// ------------------------------------------------------------------------------------------------------------------
constexpr uint32_t kP2PUnwritten = 0x80008000u;  // two fp16 -0.0

__device__ __forceinline__ void st_peer_v4(void* p, const uint4& v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t no_neg_zero(uint32_t w) { return w ^ (__vcmpeq2(w, kP2PUnwritten) & kP2PUnwritten); }
__device__ __forceinline__ uint4 no_neg_zero(uint4 v) {
  v.x = no_neg_zero(v.x);
  v.y = no_neg_zero(v.y);
  v.z = no_neg_zero(v.z);
  v.w = no_neg_zero(v.w);
  return v;
}
__device__ __forceinline__ bool landed(const uint4& v) {
  return (__vcmpeq2(v.x, kP2PUnwritten) | __vcmpeq2(v.y, kP2PUnwritten) | __vcmpeq2(v.z, kP2PUnwritten) | __vcmpeq2(v.w, kP2PUnwritten)) == 0;
}
__device__ __forceinline__ uint4 poll_v4(const void* p, unsigned long long t0) {
  uint4 v = ld_peer_v4(p);
  while (!landed(v)) {
    if (global_timer_ns() - t0 > kP2PSpinNs) __trap();
    v = ld_peer_v4(p);
  }
  return v;
}

// The fused boundary's own header layout: [kP2PBoundaryRows] u32 per-row call counters.  Inbox 1 holds ceil(rows / world) owned
// rows x world sources; both inboxes are laid out for h_cap elements per row.
constexpr int kP2PBoundaryRows = 2048;
constexpr int kP2PBoundaryGrid = 256;   // blocks of one launch: all resident at once (2 per SM), so no block waits for an unscheduled one
constexpr int kP2PBoundaryMaxR = 16;    // rows per block
static_assert(kP2PBoundaryRows * 4 <= kP2PHeaderBytes, "p2p header");
__host__ __device__ inline int64_t p2p_inbox1_rows(int world) { return (int64_t)((kP2PBoundaryRows + world - 1) / world) * world; }
// rows per block: one while every row can have its own block, else a multiple of world (every block then owns R / world rows
// and all blocks share the owner work)
__host__ __device__ inline int p2p_rows_per_block(int64_t T, int world) {
  if (T <= kP2PBoundaryGrid) return 1;
  return world * (int)((T + (int64_t)kP2PBoundaryGrid * world - 1) / ((int64_t)kP2PBoundaryGrid * world));
}

template <bool kSplitK>
__global__ void __launch_bounds__(kP2PThreads, 2)
p2p_allreduce_rmsnorm_kernel(unsigned char* const* __restrict__ peers, const __half* __restrict__ h, const B200SplitK parts,
                             const __half* __restrict__ residual, const __half* __restrict__ gamma, __half* __restrict__ normed,
                             __half* __restrict__ res_out, int T, int R, int H, int h_cap, float eps, int world, int rank) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* xs = reinterpret_cast<float*>(smem_raw);  // [H] fp32 copy of a row after the residual add (owner work)
  __shared__ float red[32];
  __shared__ uint32_t s_calls[kP2PBoundaryMaxR];
  pdl_launch_dependents();
  const int row0 = blockIdx.x * R;
  const int n_rows = min(R, T - row0);
  unsigned char* mine = peers[rank];
  uint32_t* my_calls = reinterpret_cast<uint32_t*>(mine);
  pdl_wait();  // the input is the previous kernel's output; the counters were written by the previous fused launch of the stream
  if (threadIdx.x < n_rows) s_calls[threadIdx.x] = my_calls[row0 + threadIdx.x];
  __syncthreads();
  const int64_t row_bytes = (int64_t)h_cap * 2;
  const int64_t in1_rows = p2p_inbox1_rows(world);
  const int64_t in2_base = kP2PHeaderBytes + 3 * in1_rows * row_bytes;
  auto inbox1 = [&](int sl, int row, int src) {
    return kP2PHeaderBytes + ((int64_t)sl * in1_rows + (int64_t)(row / world) * world + src) * row_bytes;
  };
  auto inbox2 = [&](int sl, int row) { return in2_base + ((int64_t)sl * kP2PBoundaryRows + row) * row_bytes; };
  const uint4 blank = make_uint4(kP2PUnwritten, kP2PUnwritten, kP2PUnwritten, kP2PUnwritten);

  // 1. this rank's partial rows -> their owners' inbox 1 (nothing here waits)
  for (int j = 0; j < n_rows; ++j) {
    const int row = row0 + j;
    unsigned char* to_owner = peers[row % world] + inbox1((int)(s_calls[j] % 3u), row, rank);
    for (int i = threadIdx.x * 8; i < H; i += kP2PThreads * 8) {
      uint4 v;
      if constexpr (kSplitK) {
        const float4 a = splitk_sum4(parts, row, i), b = splitk_sum4(parts, row, i + 4);
        v.x = pack_half2(a.x, a.y);
        v.y = pack_half2(a.z, a.w);
        v.z = pack_half2(b.x, b.y);
        v.w = pack_half2(b.z, b.w);
      } else {
        v = *reinterpret_cast<const uint4*>(h + (size_t)row * H + i);
      }
      st_peer_v4(to_owner + (int64_t)i * 2, no_neg_zero(v));
    }
  }
  const unsigned long long t0 = global_timer_ns();

  // 2. the rows this rank owns: sum in rank order, residual add, RMSNorm, normed row to everyone (waits for step 1 of the
  //    same block index on the other ranks only)
  for (int j = 0; j < n_rows; ++j) {
    const int row = row0 + j;
    if (row % world != rank) continue;
    const int slot = (int)(s_calls[j] % 3u), prev = (int)((s_calls[j] + 2u) % 3u);
    float ss = 0.f;
    for (int i = threadIdx.x * 8; i < H; i += kP2PThreads * 8) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      uint4 ld[kP2PMaxWorld];
#pragma unroll
      for (int r = 0; r < kP2PMaxWorld; ++r)
        if (r < world) ld[r] = ld_peer_v4(mine + inbox1(slot, row, r) + (int64_t)i * 2);  // all sources in flight together
#pragma unroll
      for (int r = 0; r < kP2PMaxWorld; ++r)
        if (r < world) {
          if (!landed(ld[r])) ld[r] = poll_v4(mine + inbox1(slot, row, r) + (int64_t)i * 2, t0);
          add_h8(acc, ld[r]);
        }
      float x[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = __half2float(__float2half_rn(acc[k]));  // the all-reduce result is an fp16 tensor
      uint4 ov;
      __half2* o2 = reinterpret_cast<__half2*>(&ov);
      if (residual) {
        const uint4 rv = *reinterpret_cast<const uint4*>(residual + (size_t)row * H + i);
        const __half2* r2 = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(r2[k]);
          x[2 * k] += f.x;
          x[2 * k + 1] += f.y;
        }
      }
      // residual == NULL (first layer, flash_llama_modeling.py:149-150): residual_out = the reduced hidden state itself
#pragma unroll
      for (int k = 0; k < 4; ++k) o2[k] = __floats2half2_rn(x[2 * k], x[2 * k + 1]);
      *reinterpret_cast<uint4*>(res_out + (size_t)row * H + i) = ov;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        xs[i + k] = x[k];
        ss += x[k] * x[k];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    __syncthreads();  // red[] of the previous owned row has been read by everyone
    if (lane_id() == 0) red[warp_id()] = ss;
    __syncthreads();
    if (warp_id() == 0) {
      float v = lane_id() < kP2PThreads / 32 ? red[lane_id()] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane_id() == 0) red[0] = v;
    }
    __syncthreads();
    const float rstd = rsqrtf(red[0] / (float)H + eps);
    const int64_t out_off = inbox2(slot, row);
    for (int i = threadIdx.x * 8; i < H; i += kP2PThreads * 8) {
      const uint4 gv = *reinterpret_cast<const uint4*>(gamma + i);
      const __half2* g2 = reinterpret_cast<const __half2*>(&gv);
      uint4 ov;
      __half2* o2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 g = __half22float2(g2[k]);
        o2[k] = __floats2half2_rn(xs[i + 2 * k] * rstd * g.x, xs[i + 2 * k + 1] * rstd * g.y);
      }
      ov = no_neg_zero(ov);
      for (int r = 1; r < world; ++r) {  // peers first, starting with the next rank: the ranks' pushes spread over the links
        const int dst = rank + r < world ? rank + r : rank + r - world;
        st_peer_v4(peers[dst] + out_off + (int64_t)i * 2, ov);
      }
      *reinterpret_cast<uint4*>(normed + (size_t)row * H + i) = ov;
    }
    // clear inbox 1 of this row's previous call (all sources); each thread re-reads only its own xs[] entries, no barrier needed
    for (int r = 0; r < world; ++r) {
      unsigned char* old = mine + inbox1(prev, row, r);
      for (int i = threadIdx.x * 8; i < h_cap; i += kP2PThreads * 8) *reinterpret_cast<uint4*>(old + (int64_t)i * 2) = blank;
    }
  }

  // 3. the other rows: clear the cell of the row's previous call, then wait for the owner's normed row
  for (int j = 0; j < n_rows; ++j) {
    const int row = row0 + j;
    if (row % world == rank) continue;
    unsigned char* old = mine + inbox2((int)((s_calls[j] + 2u) % 3u), row);
    for (int i = threadIdx.x * 8; i < h_cap; i += kP2PThreads * 8) *reinterpret_cast<uint4*>(old + (int64_t)i * 2) = blank;
  }
  for (int j = 0; j < n_rows; ++j) {
    const int row = row0 + j;
    if (row % world == rank) continue;
    const unsigned char* in = mine + inbox2((int)(s_calls[j] % 3u), row);
    for (int i = threadIdx.x * 8; i < H; i += kP2PThreads * 8)
      *reinterpret_cast<uint4*>(normed + (size_t)row * H + i) = poll_v4(in + (int64_t)i * 2, t0);
  }
  if (threadIdx.x < n_rows) my_calls[row0 + threadIdx.x] = s_calls[threadIdx.x] + 1;
}

__global__ void p2p_fill_kernel(uint4* p, int64_t n, uint32_t word) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = make_uint4(word, word, word, word);
}

// ------------------------------------------------------------------------------------------------------------------
// Tensor-parallel greedy head without the logits all-gather (replaces TensorParallelHead's
// all_gather_into_tensor of [T, V/tp] fp16 per rank, utils/layers.py:249-269, followed by Greedy's argmax, utils/tokens.py:44-46,
// for all-greedy batches): every rank takes the arg-max of its own vocabulary shard, the (value, index) pairs cross NVLink
// (8 bytes per row and rank instead of V/tp logits), and every rank picks the same winner: the largest value, ties to the
// lowest global token id - exactly torch.argmax over the concatenated row.  One block per row, window family 3:
// slot layout [2][kP2PMaxRows][2] u32 (value bits, global index).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kP2PArgmaxThreads = 1024;
__global__ void __launch_bounds__(kP2PArgmaxThreads, 1)
p2p_argmax_kernel(unsigned char* const* __restrict__ peers, const __half* __restrict__ logits, int64_t V_local, int64_t ld,
                  const int64_t* __restrict__ banned, int64_t* __restrict__ out, int world, int rank) {
  __shared__ float sv[kP2PArgmaxThreads / 32];
  __shared__ int si[kP2PArgmaxThreads / 32];
  __shared__ uint32_t s_epoch;
  pdl_launch_dependents();
  const int row = blockIdx.x;
  unsigned char* mine = peers[rank];
  uint32_t* my_flags = reinterpret_cast<uint32_t*>(mine) + row * kP2PMaxWorld;
  uint32_t* my_epoch = reinterpret_cast<uint32_t*>(mine) + kP2PMaxRows * kP2PMaxWorld + row;
  pdl_wait();
  if (threadIdx.x == 0) s_epoch = *my_epoch + 1;
  // local arg-max over this rank's shard (first index wins ties); the banned token id is global
  const __half* lrow = logits + (size_t)row * ld;
  const int64_t v0 = (int64_t)rank * V_local;
  const int ban = banned ? (int)(banned[row] - v0) : -1;  // negative or >= V_local: not in this shard
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  const bool vec = (ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  const int V8 = vec ? (int)(V_local & ~7LL) : 0;
  for (int i = threadIdx.x * 8; i < V8; i += kP2PArgmaxThreads * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(lrow + i);
    const __half* hv = reinterpret_cast<const __half*>(&v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float f = (i + j == ban) ? -INFINITY : __half2float(hv[j]);
      if (f > best) { best = f; best_i = i + j; }
    }
  }
  for (int i = V8 + threadIdx.x; i < (int)V_local; i += kP2PArgmaxThreads) {
    const float f = (i == ban) ? -INFINITY : __half2float(lrow[i]);
    if (f > best || (f == best && i < best_i)) { best = f; best_i = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
  }
  if (lane_id() == 0) { sv[warp_id()] = best; si[warp_id()] = best_i; }
  __syncthreads();
  if (warp_id() == 0) {
    best = sv[lane_id()];
    best_i = si[lane_id()];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    const uint32_t epoch = s_epoch;
    const int64_t slot_off = kP2PHeaderBytes + (int64_t)(epoch & 1) * (kP2PMaxRows * 8) + (int64_t)row * 8;
    if (lane_id() == 0) {
      // an all -inf / all-banned shard still takes part with (-inf, its first index)
      const uint32_t gi = (uint32_t)(v0 + (best_i == 0x7fffffff ? 0 : best_i));
      *reinterpret_cast<uint2*>(mine + slot_off) = make_uint2(__float_as_uint(best), gi);
      __threadfence_system();
    }
    __syncwarp();
    if (lane_id() < world && lane_id() != rank) {
      uint32_t* theirs = reinterpret_cast<uint32_t*>(peers[lane_id()]) + row * kP2PMaxWorld + rank;
      st_release_sys(theirs, epoch);
      const unsigned long long t0 = global_timer_ns();
      while ((int32_t)(ld_acquire_sys(my_flags + lane_id()) - epoch) < 0) {
        if (global_timer_ns() - t0 > kP2PSpinNs) __trap();
      }
    }
    __syncwarp();
    float wv = -INFINITY;
    uint32_t wi = 0xffffffffu;
    if (lane_id() < world) {
      uint2 pr;
      asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(pr.x), "=r"(pr.y) : "l"(peers[lane_id()] + slot_off) : "memory");
      wv = __uint_as_float(pr.x);
      wi = pr.y;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {  // world <= 8 lanes hold candidates
      const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
      const uint32_t oi = __shfl_xor_sync(0xffffffffu, wi, o);
      if (ov > wv || (ov == wv && oi < wi)) { wv = ov; wi = oi; }
    }
    if (lane_id() == 0) {
      out[row] = (int64_t)wi;
      *my_epoch = epoch;
    }
  }
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
  // the chunked all-reduce and the arg-max use two slots of max_bytes; the fused boundary uses 3 slots x (inbox 1 + inbox 2),
  // each 2048 (+ up to 7) rows of max_bytes / 2048, and wants every data cell to read "unwritten" (fp16 -0.0) before its first call
  const size_t bytes = kP2PHeaderBytes + 7 * (size_t)ctx->max_bytes;
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaMalloc(&ctx->window, bytes);
  if (e == cudaSuccess) e = cudaMemset(ctx->window, 0, kP2PHeaderBytes);
  if (e == cudaSuccess) {
    p2p_fill_kernel<<<256, 256>>>(reinterpret_cast<uint4*>(ctx->window + kP2PHeaderBytes), (int64_t)(bytes - kP2PHeaderBytes) / 16, kP2PUnwritten);
    e = cudaGetLastError();
  }
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
  if (ctx->family == 2) { b200_set_last_error("p2p_allreduce: this window already serves b200_p2p_allreduce_rmsnorm"); return B200_ERR_ARG; }
  ctx->family = 1;
  const int64_t n_chunks = (n * 2 + kP2PChunkBytes - 1) / kP2PChunkBytes;
  const int blocks = (int)(n_chunks < kP2PBlocks ? n_chunks : kP2PBlocks);
  B200_LAUNCH(p2p_allreduce_f16_kernel, dim3(blocks), dim3(kP2PThreads), 0, (cudaStream_t)stream,
              (unsigned char* const*)ctx->peer_table, (__half*)data, n, ctx->world, ctx->rank, ctx->max_bytes);
  b200_count_launches(1);
  return B200_OK;
}

// Fused layer boundary (see the kernel).  Exactly one of h (fp16 [T, H], this rank's partial sums) and h_parts (a deferred
// row-parallel GEMM, H = h_parts->N, T = h_parts->T) is given.  residual may be NULL (first layer): residual_out then receives
// the reduced hidden state.  residual / residual_out are read / written ONLY for the rows this rank owns (t % world == rank):
// the residual stream of a row lives on its owner, every rank gets every normed row.  All boundaries of a step must go through
// this call for that to hold.  T <= 2048, 2048 * H * 2 <= max_bytes, H % 8 == 0, H <= 16384.
extern "C" int b200_p2p_allreduce_rmsnorm(void* ctx_, const void* h, const B200SplitK* h_parts, const void* residual, const void* gamma,
                                          void* normed_out, void* residual_out, int64_t T, int64_t H, float eps, void* stream) {
  P2PContext* ctx = (P2PContext*)ctx_;
  if (!ctx || (!h) == (!h_parts) || !gamma || !normed_out || !residual_out) {
    b200_set_last_error("p2p_allreduce_rmsnorm: need a context, exactly one of h / h_parts, gamma and both outputs");
    return B200_ERR_ARG;
  }
  if (h_parts) {
    if (!h_parts->partial || h_parts->half_tiles != 0 || h_parts->T <= 0 || h_parts->T > h_parts->tn) {
      b200_set_last_error("p2p_allreduce_rmsnorm: h_parts is not a plain-layout B200SplitK");
      return B200_ERR_ARG;
    }
    T = h_parts->T;
    H = h_parts->N;
  }
  if (T == 0) return B200_OK;
  // the window is laid out for 2048 rows of h_cap elements: 3 x (2048 + 7) + 3 x 2048 rows fit in the 7 * max_bytes behind the header
  const int64_t h_cap = ctx->max_bytes / (2 * kP2PBoundaryRows) / 8 * 8;
  if (T > kP2PBoundaryRows || H % 8 != 0 || H > 16384 || H > h_cap || (h && ((uintptr_t)h & 15) != 0)) {
    b200_set_last_error("p2p_allreduce_rmsnorm: need T <= 2048, H % 8 == 0, H <= 16384, 2048 * H * 2 <= max_bytes, 16-byte aligned h");
    return B200_ERR_ARG;
  }
  const int R = p2p_rows_per_block(T, ctx->world);
  const unsigned grid = (unsigned)((T + R - 1) / R);
  for (int r = 0; r < ctx->world; ++r)
    if (!ctx->peer[r]) { b200_set_last_error("p2p_allreduce_rmsnorm: b200_p2p_connect has not run"); return B200_ERR_ARG; }
  if (ctx->family == 1) { b200_set_last_error("p2p_allreduce_rmsnorm: this window already serves b200_p2p_allreduce_f16"); return B200_ERR_ARG; }
  ctx->family = 2;
  const size_t smem = (size_t)H * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (h_parts) {
    constexpr auto kernel = p2p_allreduce_rmsnorm_kernel<true>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    B200_LAUNCH_AS("p2p_allreduce_rmsnorm_kernel<splitk>", kernel, dim3(grid), dim3(kP2PThreads), smem, st,
                   (unsigned char* const*)ctx->peer_table, (const __half*)nullptr, *h_parts, (const __half*)residual, (const __half*)gamma,
                   (__half*)normed_out, (__half*)residual_out, (int)T, R, (int)H, (int)h_cap, eps, ctx->world, ctx->rank);
  } else {
    constexpr auto kernel = p2p_allreduce_rmsnorm_kernel<false>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    B200_LAUNCH_AS("p2p_allreduce_rmsnorm_kernel", kernel, dim3(grid), dim3(kP2PThreads), smem, st,
                   (unsigned char* const*)ctx->peer_table, (const __half*)h, B200SplitK{}, (const __half*)residual, (const __half*)gamma,
                   (__half*)normed_out, (__half*)residual_out, (int)T, R, (int)H, (int)h_cap, eps, ctx->world, ctx->rank);
  }
  b200_count_launches(1);
  return B200_OK;
}

// Greedy ids of a vocabulary-sharded head (see the kernel): logits [B, V_local] fp16 with row stride ld (halves), this rank's
// shard = global token ids [rank * V_local, (rank + 1) * V_local); banned_ids (optional, [B]) as in b200_argmax, global ids.
// out_ids [B] int64 is identical on every rank.  B <= 256.  The window (b200_p2p_create with any max_bytes >= 4096) must not
// be shared with the other p2p kernels.
extern "C" int b200_p2p_argmax(void* ctx_, const void* logits, int64_t* out_ids, int64_t B, int64_t V_local, int64_t ld,
                               const int64_t* banned_ids, void* stream) {
  P2PContext* ctx = (P2PContext*)ctx_;
  if (!ctx || !logits || !out_ids || B < 0 || V_local <= 0 || ld < V_local) {
    b200_set_last_error("p2p_argmax: null argument or bad shape");
    return B200_ERR_ARG;
  }
  if (B == 0) return B200_OK;
  if (B > kP2PMaxRows || 2 * ctx->max_bytes < 2 * kP2PMaxRows * 8 || V_local * ctx->world > 0x7fffffffLL) {
    b200_set_last_error("p2p_argmax: need B <= 256 and a window of at least 4096 bytes");
    return B200_ERR_ARG;
  }
  for (int r = 0; r < ctx->world; ++r)
    if (!ctx->peer[r]) { b200_set_last_error("p2p_argmax: b200_p2p_connect has not run"); return B200_ERR_ARG; }
  if (ctx->family != 0 && ctx->family != 3) { b200_set_last_error("p2p_argmax: this window already serves another p2p kernel"); return B200_ERR_ARG; }
  ctx->family = 3;
  B200_LAUNCH(p2p_argmax_kernel, dim3((unsigned)B), dim3(kP2PArgmaxThreads), 0, (cudaStream_t)stream,
              (unsigned char* const*)ctx->peer_table, (const __half*)logits, V_local, ld, banned_ids, out_ids, ctx->world, ctx->rank);
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
