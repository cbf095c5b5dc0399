// Paged decode attention (one query token per sequence) — the dominant HBM consumer of the decode step.
//
// Replaces: utils/flash_attn.py:43-127 `attention(...)` in its decode form (flash_llama_modeling.py:285-295,
// flash_attn_2_cuda.varlen_fwd with max_s_q = 1, causal = False) and fms-extras `paged_attention_v1/v2`
// (paged_llama_modeling.py:267) of /root/reference/server/text_generation_server.
//
// Design:
//   * work item = (sequence b, kv head, 512-token chunk): split-KV so B*h_kv*chunks >> 148 SMs, 2 CTAs/SM
//   * KV pages [16 tokens][d] fp16 are stored in HBM already XOR-swizzled; a producer warp streams them with
//     cp.async.bulk (TMA, one 2-4 KB copy per page-head tile for K and V) into a 3/4-stage ring, mbarrier-tracked
//   * the GQA group's G query heads form the (padded) 16-row A operand of m16n8k16 HMMA so every K/V byte is read
//     once per group; fp32 online softmax with quad shuffles; P rounded to fp16 before P.V like flash-attn
//   * per-chunk (m, l, O) partials -> fp32 workspace, merged by attn_decode_combine_kernel
// The kernel is bandwidth-bound: algorithmic bytes = 2 * L * d * 2 B per (sequence, kv head).
#include "common.cuh"

#include <cstdlib>

namespace b200 {

constexpr int kMaxChunkTokens = 512;  // <= 32 pages: the producer warp holds one block id per lane
constexpr int kMinChunkTokens = 128;
constexpr int kPagesPerStage = 4;
constexpr int kConsumerWarps = 4;
constexpr int kDecodeThreads = (kConsumerWarps + 1) * 32;
constexpr float kNegBig = -1.0e30f;

template <int D>
struct DecodeSmem {
  static constexpr int kPageBytes = kPageTokens * D * 2;
  static constexpr int kStages = D == 128 ? 3 : 4;
  static constexpr int kStageBytes = 2 * kPagesPerStage * kPageBytes;
  static constexpr int kMergeStride = D + 8;  // floats
  static constexpr int kBytes = kStages * kStageBytes + 2 * kStages * 8 + 16;
};

template <int D>
__global__ void __launch_bounds__(kDecodeThreads, 2)
attn_decode_paged_kernel(const __half* __restrict__ q, int64_t q_token_stride, const __half* __restrict__ k_pool,
                         const __half* __restrict__ v_pool, const int32_t* __restrict__ block_table, int64_t bt_stride,
                         const int32_t* __restrict__ context_lens, float* __restrict__ part_o, float* __restrict__ part_ml,
                         int n_heads, int n_kv, int n_chunks_max, int chunk_tokens, float scale_log2) {
  using S = DecodeSmem<D>;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* stages = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kStages * S::kStageBytes);
  uint64_t* empty_bar = full_bar + S::kStages;

  pdl_launch_dependents();
  pdl_wait();
  const int chunk = blockIdx.x, hk = blockIdx.y, b = blockIdx.z;
  const int L = context_lens[b];
  const int tok0 = chunk * chunk_tokens;
  if (tok0 >= L) return;
  const int n_tok = min(chunk_tokens, L - tok0);
  const int n_pages = (n_tok + kPageTokens - 1) / kPageTokens;
  const int n_iters = (n_pages + kPagesPerStage - 1) / kPagesPerStage;
  const int G = n_heads / n_kv;
  const int warp = warp_id(), lane = lane_id();

  if (threadIdx.x == 0) {
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == kConsumerWarps) {
    // ------------------------------------------------------------------ producer warp
    const int32_t* bt = block_table + (int64_t)b * bt_stride + tok0 / kPageTokens;
    int my_block = lane < n_pages ? bt[lane] : 0;  // chunk has <= 32 pages
    const size_t tile_halves = (size_t)kPageTokens * D;
    for (int it = 0; it < n_iters; ++it) {
      const int s = it % S::kStages;
      const int np = min(kPagesPerStage, n_pages - it * kPagesPerStage);
      if (lane == 0) {
        mbar_wait(&empty_bar[s], ((it / S::kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(np * 2 * S::kPageBytes));
      }
      for (int p = 0; p < np; ++p) {
        const int blk = __shfl_sync(0xffffffffu, my_block, it * kPagesPerStage + p);
        if (lane == 0) {
          const size_t off = ((size_t)blk * n_kv + hk) * tile_halves;
          unsigned char* dstK = stages + s * S::kStageBytes + p * S::kPageBytes;
          unsigned char* dstV = dstK + kPagesPerStage * S::kPageBytes;
          tma_bulk_g2s(dstK, k_pool + off, S::kPageBytes, &full_bar[s]);
          tma_bulk_g2s(dstV, v_pool + off, S::kPageBytes, &full_bar[s]);
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumer warps
  const int g = lane >> 2, tig = lane & 3;
  // Q fragments (A operand, rows = heads of the GQA group padded to 16)
  uint32_t qf[D / 16][4];
  {
    const __half* qb = q + (int64_t)b * q_token_stride + (int64_t)hk * G * D;
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
      const int col = ks * 16 + tig * 2;
      qf[ks][0] = g < G ? *reinterpret_cast<const uint32_t*>(qb + g * D + col) : 0u;
      qf[ks][1] = g + 8 < G ? *reinterpret_cast<const uint32_t*>(qb + (g + 8) * D + col) : 0u;
      qf[ks][2] = g < G ? *reinterpret_cast<const uint32_t*>(qb + g * D + col + 8) : 0u;
      qf[ks][3] = g + 8 < G ? *reinterpret_cast<const uint32_t*>(qb + (g + 8) * D + col + 8) : 0u;
    }
  }
  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = kNegBig, m1 = kNegBig, l0 = 0.f, l1 = 0.f;

  const int mi = lane >> 3, r8 = lane & 7;
  for (int it = 0; it < n_iters; ++it) {
    const int s = it % S::kStages;
    const int page = it * kPagesPerStage + warp;
    mbar_wait(&full_bar[s], (it / S::kStages) & 1);
    if (page < n_pages) {
      const int n_valid = min(kPageTokens, n_tok - page * kPageTokens);
      const uint32_t kb = smem_u32(stages + s * S::kStageBytes + warp * S::kPageBytes);
      const uint32_t vb = kb + kPagesPerStage * S::kPageBytes;
      float sc[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
        const int row = nt * 8 + r8;
#pragma unroll
        for (int kc = 0; kc < D / 32; ++kc) {
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4(b0, b1, b2, b3, kb + kv_swizzled_chunk_offset<D>(row, kc * 4 + mi));
          mma_m16n8k16_f16f32(sc[nt], qf[kc * 2], b0, b1);
          mma_m16n8k16_f16f32(sc[nt], qf[kc * 2 + 1], b2, b3);
        }
      }
      // mask + online softmax (log2 domain)
      float mx0 = kNegBig, mx1 = kNegBig;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const bool ok = nt * 8 + tig * 2 + e < n_valid;
          sc[nt][e] = ok ? sc[nt][e] * scale_log2 : -INFINITY;
          sc[nt][2 + e] = ok ? sc[nt][2 + e] * scale_log2 : -INFINITY;
          mx0 = fmaxf(mx0, sc[nt][e]);
          mx1 = fmaxf(mx1, sc[nt][2 + e]);
        }
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
      const float a0 = fast_exp2(m0 - mn0), a1 = fast_exp2(m1 - mn1);
      m0 = mn0;
      m1 = mn1;
      float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          sc[nt][e] = fast_exp2(sc[nt][e] - mn0);
          sc[nt][2 + e] = fast_exp2(sc[nt][2 + e] - mn1);
          ps0 += sc[nt][e];
          ps1 += sc[nt][2 + e];
        }
      }
      l0 = l0 * a0 + ps0;
      l1 = l1 * a1 + ps1;
      uint32_t pa[4];
      pa[0] = pack_half2(sc[0][0], sc[0][1]);
      pa[1] = pack_half2(sc[0][2], sc[0][3]);
      pa[2] = pack_half2(sc[1][0], sc[1][1]);
      pa[3] = pack_half2(sc[1][2], sc[1][3]);
#pragma unroll
      for (int dc = 0; dc < D / 16; ++dc) {
        o[2 * dc][0] *= a0; o[2 * dc][1] *= a0; o[2 * dc][2] *= a1; o[2 * dc][3] *= a1;
        o[2 * dc + 1][0] *= a0; o[2 * dc + 1][1] *= a0; o[2 * dc + 1][2] *= a1; o[2 * dc + 1][3] *= a1;
        uint32_t v0, v1, v2, v3;
        const int row = (mi & 1) * 8 + r8;
        ldmatrix_x4_trans(v0, v1, v2, v3, vb + kv_swizzled_chunk_offset<D>(row, 2 * dc + (mi >> 1)));
        mma_m16n8k16_f16f32(o[2 * dc], pa, v0, v1);
        mma_m16n8k16_f16f32(o[2 * dc + 1], pa, v2, v3);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }

  // ---------------------------------------------------------------------- merge the 4 warps' partials
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32));  // every stage consumed by every warp
  float* mo = reinterpret_cast<float*>(stages);                               // [4][16][D+8]
  float* mml = mo + kConsumerWarps * 16 * S::kMergeStride;                     // [4][16][2]
  if (g < G) {
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt)
      *reinterpret_cast<float2*>(&mo[(warp * 16 + g) * S::kMergeStride + nt * 8 + tig * 2]) = make_float2(o[nt][0], o[nt][1]);
    if (tig == 0) { mml[(warp * 16 + g) * 2] = m0; mml[(warp * 16 + g) * 2 + 1] = l0; }
  }
  if (g + 8 < G) {
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt)
      *reinterpret_cast<float2*>(&mo[(warp * 16 + g + 8) * S::kMergeStride + nt * 8 + tig * 2]) = make_float2(o[nt][2], o[nt][3]);
    if (tig == 0) { mml[(warp * 16 + g + 8) * 2] = m1; mml[(warp * 16 + g + 8) * 2 + 1] = l1; }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32));
  for (int idx = threadIdx.x; idx < G * D; idx += kConsumerWarps * 32) {
    const int r = idx / D, d = idx % D;
    float M = kNegBig;
#pragma unroll
    for (int w = 0; w < kConsumerWarps; ++w) M = fmaxf(M, mml[(w * 16 + r) * 2]);
    float acc = 0.f, lsum = 0.f;
#pragma unroll
    for (int w = 0; w < kConsumerWarps; ++w) {
      const float f = fast_exp2(mml[(w * 16 + r) * 2] - M);
      acc += f * mo[(w * 16 + r) * S::kMergeStride + d];
      lsum += f * mml[(w * 16 + r) * 2 + 1];
    }
    const int64_t slot = ((int64_t)b * n_heads + hk * G + r) * n_chunks_max + chunk;
    part_o[slot * D + d] = acc;
    if (d == 0) { part_ml[slot * 2] = M; part_ml[slot * 2 + 1] = lsum; }
  }
}

template <int D>
__global__ void attn_decode_combine_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml,
                                           const int32_t* __restrict__ context_lens, __half* __restrict__ out,
                                           int64_t out_token_stride, int n_heads, int n_chunks_max, int chunk_tokens) {
  pdl_launch_dependents();
  pdl_wait();
  const int head = blockIdx.x, b = blockIdx.y, d = threadIdx.x;
  const int L = context_lens[b];
  const int nc = (L + chunk_tokens - 1) / chunk_tokens;
  const int64_t base = ((int64_t)b * n_heads + head) * n_chunks_max;
  float M = kNegBig;
  for (int c = 0; c < nc; ++c) M = fmaxf(M, part_ml[(base + c) * 2]);
  float acc = 0.f, l = 0.f;
  for (int c = 0; c < nc; ++c) {
    const float f = fast_exp2(part_ml[(base + c) * 2] - M);
    acc += f * part_o[(base + c) * D + d];
    l += f * part_ml[(base + c) * 2 + 1];
  }
  out[(int64_t)b * out_token_stride + head * D + d] = __float2half_rn(nc > 0 ? acc / l : 0.f);
}

}  // namespace b200

using namespace b200;

extern "C" int64_t b200_attn_decode_workspace_bytes(int B, int n_heads, int head_dim, int max_context_len) {
  const int64_t nc = (max_context_len + kMinChunkTokens - 1) / kMinChunkTokens;  // sized for the smallest chunk
  return (int64_t)B * n_heads * (nc > 0 ? nc : 1) * (head_dim + 2) * 4;
}

static int decode_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int D>
static int launch_decode(const void* q, int64_t q_token_stride, const void* k_pool, const void* v_pool, const int32_t* block_table,
                         int64_t bt_stride, const int32_t* context_lens, void* out, int64_t out_token_stride, void* workspace,
                         int B, int n_heads, int n_kv, int n_chunks, int chunk_tokens, float scale, cudaStream_t st) {
  using S = DecodeSmem<D>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_decode_paged_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kBytes);
    if (e != cudaSuccess) { b200_set_last_error(cudaGetErrorString(e)); return B200_ERR_CUDA; }
    configured = true;
  }
  float* part_o = (float*)workspace;
  float* part_ml = part_o + (int64_t)B * n_heads * n_chunks * D;
  dim3 grid(n_chunks, n_kv, B);
  b200_timing_mark(B200_TIME_ATTN_DECODE, 0, st);
  B200_LAUNCH(attn_decode_paged_kernel<D>, grid, dim3(kDecodeThreads), (size_t)S::kBytes, st, (const __half*)q, q_token_stride,
              (const __half*)k_pool, (const __half*)v_pool, block_table, bt_stride, context_lens, part_o, part_ml, n_heads, n_kv,
                n_chunks, chunk_tokens, scale * 1.4426950408889634f);
  b200_timing_mark(B200_TIME_ATTN_DECODE, 1, st);
  b200_count_launches(1);
  B200_LAUNCH(attn_decode_combine_kernel<D>, dim3(n_heads, B), dim3(D), 0, st, (const float*)part_o, (const float*)part_ml, context_lens,
              (__half*)out, out_token_stride, n_heads, n_chunks, chunk_tokens);
  b200_count_launches(1);
  return B200_OK;
}

extern "C" int b200_attn_decode_paged(const void* q, int64_t q_token_stride, const void* k_pool, const void* v_pool,
                                      const int32_t* block_table, int64_t block_table_stride, const int32_t* context_lens,
                                      void* out, int64_t out_token_stride, void* workspace, int64_t workspace_bytes, int B,
                                      int n_heads, int n_kv_heads, int head_dim, int max_context_len, float softmax_scale,
                                      void* stream) {
  if (B == 0) return B200_OK;
  if (n_kv_heads <= 0 || n_heads % n_kv_heads != 0 || n_heads / n_kv_heads > 16) {
    b200_set_last_error("attn_decode_paged: need n_heads % n_kv_heads == 0 and group size <= 16");
    return B200_ERR_ARG;
  }
  if (workspace_bytes < b200_attn_decode_workspace_bytes(B, n_heads, head_dim, max_context_len)) {
    b200_set_last_error("attn_decode_paged: workspace too small");
    return B200_ERR_ARG;
  }
  if ((q_token_stride & 1) || ((uintptr_t)q & 3)) { b200_set_last_error("attn_decode_paged: q must be 4-byte aligned"); return B200_ERR_ARG; }
  // split-KV granularity: the largest chunk size that still gives every SM a CTA (measured, profiles/r2_attn_chunk_sweep.txt:
  // asking for more CTAs than that - this used to be 4 x 296 - cuts tensor-parallel shards, whose B * n_kv is small, into
  // 128-token chunks whose per-CTA set-up dominates: 8B tp4 30.6 -> 23.0 us, 70B tp8 32.9 -> 24.2 us per launch), then
  // EQUAL chunks of that count (page aligned): with a fixed 512 a context of 1563 tokens would be cut 512/512/512/27 and the
  // last quarter of the CTAs would do almost nothing while the others set the kernel's duration
  int target = kMaxChunkTokens;
  {
    static int env_target = -1;
    if (env_target < 0) {
      const char* e = getenv("B200_ATTN_CHUNK");
      env_target = e ? atoi(e) : 0;
    }
    if (env_target >= kMinChunkTokens && env_target <= kMaxChunkTokens) target = env_target;
  }
  static int min_ctas = -1;
  if (min_ctas < 0) {
    const char* e = getenv("B200_ATTN_MIN_CTAS");
    min_ctas = e ? atoi(e) : 148;
  }
  while (target > kMinChunkTokens && (int64_t)B * n_kv_heads * ((max_context_len + target - 1) / target) < min_ctas) target >>= 1;
  int n_chunks = (max_context_len + target - 1) / target;
  if (n_chunks < 1) n_chunks = 1;
  int chunk_tokens = ((max_context_len + n_chunks - 1) / n_chunks + kPageTokens - 1) / kPageTokens * kPageTokens;
  if (chunk_tokens < kMinChunkTokens) chunk_tokens = kMinChunkTokens;
  n_chunks = (max_context_len + chunk_tokens - 1) / chunk_tokens;
  if (n_chunks < 1) n_chunks = 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (head_dim == 128)
    return launch_decode<128>(q, q_token_stride, k_pool, v_pool, block_table, block_table_stride, context_lens, out,
                              out_token_stride, workspace, B, n_heads, n_kv_heads, n_chunks, chunk_tokens, softmax_scale, st);
  if (head_dim == 64)
    return launch_decode<64>(q, q_token_stride, k_pool, v_pool, block_table, block_table_stride, context_lens, out,
                             out_token_stride, workspace, B, n_heads, n_kv_heads, n_chunks, chunk_tokens, softmax_scale, st);
  b200_set_last_error("attn_decode_paged: head_dim must be 64 or 128");
  return B200_ERR_UNSUPPORTED;
}
