// Prefill / chunked-prefill attention over the PAGED KV pool on the 5th-generation tensor cores.
//
// Replaces: utils/flash_attn.py:43-127 `attention(q, k, v, cu_seqlens, max_s, softmax_scale)` in its prefill form
// (flash_llama_modeling.py:271-278; flash_attn_2_cuda.varlen_fwd, causal) of /root/reference/server/text_generation_server, and
// the prefill half of the paged twin (paged_llama_modeling.py:253-264).  Unlike the reference's prefill it reads K and V from the
// block pool (the step's own tokens have just been appended by b200_rope_kv_write_paged), so the queries of a step may follow a
// CACHED context: query i of sequence b sits at absolute position context_lens[b] - n_q(b) + i and attends keys 0 .. that position.
// That is what an add-on prefill over a prompt prefix, a chunked prefill and a mixed prefill + decode step need.
//
// Design (one CTA = 128 queries of one sequence x one query head; 6 warps):
//   warp 0      TMA producer.  Q tile [128 x d] through a 128B-swizzled tensor map; per 128-key tile 8 KV pages x (K, V) x d/64
//               slabs, each one cp.async.bulk.tensor box [16 tokens x 64 halves] of the pool: the pool stores every token row
//               with its 16-byte chunks XOR-swizzled by (token & 7), which IS the tcgen05 SWIZZLE_128B image, so
//               pages land in shared memory ready for the tensor core without any software swizzle.  2-stage ring.
//   warp 1      MMA issuer (one elected thread).  S = Q K^T: tcgen05.mma kind::f16, both operands K-major from shared memory,
//               fp32 accumulator in TMEM (two S buffers: QK^T of tile j+1 overlaps the softmax of tile j).
//               O += P V: A = P from TMEM (fp16, written by the softmax warps over the S buffer), B = V as an MN-major
//               shared-memory operand (keys are the MMA K dimension, d contiguous), accumulator O in TMEM.
//   warps 2..5  softmax: thread = one query row = one TMEM lane (no shuffles).  tcgen05.ld S, causal mask, fp32 online softmax
//               in the log2 domain, P rounded to fp16 (flash-attn semantics) and stored back to TMEM; when the running maximum
//               moved, O is rescaled in TMEM (tcgen05.ld / st) before the next P V accumulates into it.
// Causality prunes whole key tiles (a query tile only visits tiles up to its last row's position).
#include "common.cuh"
#include "tmap.cuh"
#include "../../include/b200_tgis.h"

namespace b200 {

constexpr int kPfM = 128;        // queries per CTA
constexpr int kPfN = 128;        // keys per tile = 8 pages
constexpr int kPfThreads = 192;  // warp 0 TMA, warp 1 MMA, warps 2..5 softmax
constexpr int kPfSlabBytes = 128 * 128;  // 128 rows x 64 halves
constexpr int kPfPageSlabBytes = kPageTokens * 128;
constexpr float kPfNegBig = -1.0e30f;

template <int D>
struct PfCfg {
  static constexpr int kSlabs = D / 64;
  static constexpr int kTileBytes = kSlabs * kPfSlabBytes;  // a Q, K or V tile
  static constexpr int kStages = 2;
  static constexpr int kNumBars = 1 + 3 * kStages + 2 + 2 + 1;  // q_full | k_full, v_full, kv_empty | s_full[2] | p_ready[2] | o_done
  static constexpr int kSmemBytes = (1 + 2 * kStages) * kTileBytes + kNumBars * 8 + 16 + 1024;
  static constexpr int kTmemCols = 512;  // S0 [0,128) | S1 [128,256) | O [256, 256 + D)
};

// MN-major shared-memory operand, 128-byte swizzle: rows along K (here: keys) of 128 bytes = 64 MN elements, 8-row swizzle atoms
// `sbo` bytes apart, blocks of 64 MN elements `lbo` bytes apart.  Bit layout as umma_desc_kmajor_sw128 (common.cuh).
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int D>
__global__ void __launch_bounds__(kPfThreads, 1)
attn_prefill_paged_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                          const __grid_constant__ CUtensorMap tmap_v, const int32_t* __restrict__ block_table, int64_t bt_stride,
                          const int32_t* __restrict__ context_lens, const int32_t* __restrict__ cu_q, __half* __restrict__ out,
                          int64_t out_stride, int n_heads, int n_kv, int tiles_per_seq, float scale_log2) {
  using C = PfCfg<D>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sQ = smem;
  unsigned char* sKV = smem + C::kTileBytes;  // stage s: K at sKV + s * 2 * kTileBytes, V right behind it
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + C::kStages * 2 * C::kTileBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = k_full + C::kStages;
  uint64_t* kv_empty = v_full + C::kStages;
  uint64_t* s_full = kv_empty + C::kStages;
  uint64_t* p_ready = s_full + 2;
  uint64_t* o_done = p_ready + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

  pdl_launch_dependents();
  pdl_wait();  // q and the pool's newest tokens are the previous kernels' output
  const int warp = warp_id(), lane = lane_id();
  const int b = blockIdx.x / tiles_per_seq, tile = blockIdx.x % tiles_per_seq;
  const int head = blockIdx.y;
  const int hk = head / (n_heads / n_kv);
  const int q_begin = cu_q[b];
  const int n_q = cu_q[b + 1] - q_begin;
  const int q0 = tile * kPfM;
  if (q0 >= n_q) return;  // the whole CTA: this sequence has fewer query tiles
  const int L = context_lens[b];
  const int past = max(L - n_q, 0);  // cached tokens in front of this step's first query
  const int rows = min(kPfM, n_q - q0);
  const int n_kt = (past + q0 + rows - 1) / kPfN + 1;  // key tiles the last query row reaches
  const int last_page = (L - 1) / kPageTokens;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 4);
    }
    mbar_init(o_done, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<C::kTmemCols>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 256;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, C::kTileBytes);
      for (int s = 0; s < C::kSlabs; ++s) tma_load_2d(sQ + s * kPfSlabBytes, &tmap_q, head * D + s * 64, q_begin + q0, q_full);
    }
    const int32_t* bt = block_table + (int64_t)b * bt_stride;
    for (int j = 0; j < n_kt; ++j) {
      const int st = j % C::kStages;
      // 8 pages of this key tile; pages past the sequence's last one repeat it (finite data, masked by the softmax warps)
      const int page = min(j * (kPfN / kPageTokens) + (lane & 7), last_page);
      const int blk = bt[page];
      if (lane == 0) mbar_wait(&kv_empty[st], ((j / C::kStages) & 1) ^ 1);
      __syncwarp();
      unsigned char* sK = sKV + st * 2 * C::kTileBytes;
      unsigned char* sV = sK + C::kTileBytes;
      if (lane == 0) {
        mbar_arrive_expect_tx(&k_full[st], C::kTileBytes);
        mbar_arrive_expect_tx(&v_full[st], C::kTileBytes);
      }
      for (int p = 0; p < kPfN / kPageTokens; ++p) {
        const int bp = __shfl_sync(0xffffffffu, blk, p);
        if (lane == 0) {
          const int row = (bp * n_kv + hk) * kPageTokens;  // row of the pool viewed as [blocks * n_kv * 16, d]
          for (int s = 0; s < C::kSlabs; ++s)
            tma_load_2d(sK + s * kPfSlabBytes + p * kPfPageSlabBytes, &tmap_k, s * 64, row, &k_full[st]);
        }
      }
      for (int p = 0; p < kPfN / kPageTokens; ++p) {
        const int bp = __shfl_sync(0xffffffffu, blk, p);
        if (lane == 0) {
          const int row = (bp * n_kv + hk) * kPageTokens;
          for (int s = 0; s < C::kSlabs; ++s)
            tma_load_2d(sV + s * kPfSlabBytes + p * kPfPageSlabBytes, &tmap_v, s * 64, row, &v_full[st]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_qk = umma_idesc_f16_f32acc(kPfM, kPfN);
    constexpr uint32_t idesc_pv = umma_idesc_f16_f32acc(kPfM, D) | (1u << 16);  // B (= V) is MN-major
    const uint64_t qdesc = umma_desc_kmajor_sw128(smem_u32(sQ));
    auto issue_qk = [&](int j) {
      const int st = j % C::kStages, buf = j & 1;
      mbar_wait(&k_full[st], (j / C::kStages) & 1);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint64_t kdesc = umma_desc_kmajor_sw128(smem_u32(sKV + st * 2 * C::kTileBytes));
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint64_t off = (uint64_t)((k >> 2) * (kPfSlabBytes >> 4) + (k & 3) * 2);
          umma_f16_ss(tmem_base + buf * kPfN, qdesc + off, kdesc + off, idesc_qk, k > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[buf]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_qk(0);
    for (int j = 0; j < n_kt; ++j) {
      const int st = j % C::kStages, buf = j & 1;
      if (j + 1 < n_kt) issue_qk(j + 1);  // overlaps the softmax of tile j (the other S buffer)
      mbar_wait(&p_ready[buf], (j >> 1) & 1);
      mbar_wait(&v_full[st], (j / C::kStages) & 1);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t v_addr = smem_u32(sKV + st * 2 * C::kTileBytes + C::kTileBytes);
        const uint64_t vdesc = umma_desc_mnmajor_sw128(v_addr, kPfSlabBytes, 1024);
        const uint32_t tmem_p = tmem_base + buf * kPfN;  // fp16 P over the first 64 columns of the S buffer
#pragma unroll
        for (int kk = 0; kk < kPfN / 16; ++kk)  // 16 keys = 2 swizzle atoms of 8 key rows = 2048 bytes per step
          umma_f16_ts(tmem_o, tmem_p + kk * 8, vdesc + (uint64_t)((kk * 2048) >> 4), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&kv_empty[st]);
        umma_commit(o_done);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax warps: thread = query row = TMEM lane
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int qpos = past + q0 + row;  // absolute position of this row's query (rows >= `rows` are padding: computed, not stored)
    float m = kPfNegBig, l = 0.f;
    for (int j = 0; j < n_kt; ++j) {
      const int buf = j & 1;
      const uint32_t tmem_s = tmem_base + lane_base + buf * kPfN;
      mbar_wait(&s_full[buf], (j >> 1) & 1);
      tcgen05_fence_after();
      float s[kPfN];
#pragma unroll
      for (int c = 0; c < kPfN / 16; ++c) tmem_ld_32x32b_x16(tmem_s + c * 16, reinterpret_cast<uint32_t(&)[16]>(s[c * 16]));
      tmem_ld_wait();
      const int key0 = j * kPfN;
      float mx = kPfNegBig;
      if (key0 + kPfN - 1 > past + q0 + quarter * 32) {  // warp-uniform: some row of this warp has keys to mask in this tile
#pragma unroll
        for (int c = 0; c < kPfN; ++c) {
          s[c] = (key0 + c <= qpos) ? s[c] * scale_log2 : -INFINITY;
          mx = fmaxf(mx, s[c]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < kPfN; ++c) {
          s[c] *= scale_log2;
          mx = fmaxf(mx, s[c]);
        }
      }
      const float m_new = fmaxf(m, mx);
      const float alpha = fast_exp2(m - m_new);
      m = m_new;
      float sum = 0.f;
      uint32_t p[kPfN / 2];
#pragma unroll
      for (int c = 0; c < kPfN; c += 2) {
        const float p0 = fast_exp2(s[c] - m_new), p1 = fast_exp2(s[c + 1] - m_new);
        sum += p0 + p1;
        p[c / 2] = pack_half2(p0, p1);
      }
      l = l * alpha + sum;
      if (j > 0) {
        mbar_wait(o_done, (j - 1) & 1);  // P V of tile j-1 has completed: O is quiescent and the P buffer is free
        tcgen05_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
          for (int c = 0; c < D / 16; ++c) {
            uint32_t o[16];
            tmem_ld_32x32b_x16(tmem_o + lane_base + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x16(tmem_o + lane_base + c * 16, o);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < kPfN / 32; ++c) tmem_st_32x32b_x16(tmem_s + c * 16, reinterpret_cast<const uint32_t(&)[16]>(p[c * 16]));
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[buf]);
    }
    mbar_wait(o_done, (n_kt - 1) & 1);
    tcgen05_fence_after();
    const float inv = l > 0.f ? 1.f / l : 0.f;
    __half* op = out + (int64_t)(q_begin + q0 + row) * out_stride + head * D;
#pragma unroll
    for (int c = 0; c < D / 16; ++c) {
      uint32_t o[16];
      tmem_ld_32x32b_x16(tmem_o + lane_base + c * 16, o);
      tmem_ld_wait();
      if (row < rows) {
        uint4 lo, hi;
        lo.x = pack_half2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
        lo.y = pack_half2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
        lo.z = pack_half2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
        lo.w = pack_half2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
        hi.x = pack_half2(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
        hi.y = pack_half2(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
        hi.z = pack_half2(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
        hi.w = pack_half2(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
        *reinterpret_cast<uint4*>(op + c * 16) = lo;
        *reinterpret_cast<uint4*>(op + c * 16 + 8) = hi;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<C::kTmemCols>(tmem_base);
}

}  // namespace b200

using namespace b200;

template <int D>
static int launch_prefill_paged(const void* q, int64_t q_stride, int64_t T_q, const void* k_pool, const void* v_pool, int64_t num_blocks,
                                const int32_t* block_table, int64_t bt_stride, const int32_t* context_lens, const int32_t* cu_q, void* out,
                                int64_t out_stride, int B, int max_q, int n_heads, int n_kv, float scale, cudaStream_t st) {
  using C = PfCfg<D>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_prefill_paged_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) { b200_set_last_error(cudaGetErrorString(e)); return B200_ERR_CUDA; }
    configured = true;
  }
  const CUtensorMap* mq = get_tmap_2d(q, (uint64_t)T_q, (uint64_t)n_heads * D, (uint64_t)q_stride, kPfM, 64, TmapDtype::kF16, TmapSwizzle::k128B);
  const uint64_t pool_rows = (uint64_t)num_blocks * n_kv * kPageTokens;
  const CUtensorMap* mk = get_tmap_2d(k_pool, pool_rows, D, D, kPageTokens, 64, TmapDtype::kF16, TmapSwizzle::kNone);
  const CUtensorMap* mv = get_tmap_2d(v_pool, pool_rows, D, D, kPageTokens, 64, TmapDtype::kF16, TmapSwizzle::kNone);
  if (!mq || !mk || !mv) return B200_ERR_CUDA;
  const int tiles = (max_q + kPfM - 1) / kPfM;
  dim3 grid((unsigned)(B * tiles), (unsigned)n_heads);
  B200_LAUNCH(attn_prefill_paged_kernel<D>, grid, dim3(kPfThreads), (size_t)C::kSmemBytes, st, *mq, *mk, *mv, block_table, bt_stride,
              context_lens, cu_q, (__half*)out, out_stride, n_heads, n_kv, tiles, scale * 1.4426950408889634f);
  b200_count_launches(1);
  return B200_OK;
}

// q: this step's query tokens, head-major [n_heads][d] at q + t * q_token_stride (halves), T_q tokens in all, sequence b owning
// tokens cu_seqlens_q[b] .. cu_seqlens_q[b+1].  context_lens[b] = tokens of sequence b in the pool INCLUDING this step's
// (already written).  Causal over absolute positions.  max_q = the largest number of query tokens of a sequence.
extern "C" int b200_attn_prefill_paged(const void* q, int64_t q_token_stride, int64_t T_q, const void* k_pool, const void* v_pool,
                                       int64_t num_blocks, const int32_t* block_table, int64_t block_table_stride,
                                       const int32_t* context_lens, const int32_t* cu_seqlens_q, void* out, int64_t out_token_stride,
                                       int B, int max_q, int n_heads, int n_kv_heads, int head_dim, float softmax_scale, void* stream) {
  if (B == 0 || max_q == 0 || T_q == 0) return B200_OK;
  if (n_kv_heads <= 0 || n_heads % n_kv_heads != 0) { b200_set_last_error("attn_prefill_paged: bad head counts"); return B200_ERR_ARG; }
  if (((q_token_stride | out_token_stride) & 7) || ((uintptr_t)q & 15) || ((uintptr_t)out & 15) || num_blocks <= 0) {
    b200_set_last_error("attn_prefill_paged: q / out must be 16-byte aligned with token strides that are multiples of 8 halves");
    return B200_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (head_dim == 128)
    return launch_prefill_paged<128>(q, q_token_stride, T_q, k_pool, v_pool, num_blocks, block_table, block_table_stride, context_lens,
                                     cu_seqlens_q, out, out_token_stride, B, max_q, n_heads, n_kv_heads, softmax_scale, st);
  if (head_dim == 64)
    return launch_prefill_paged<64>(q, q_token_stride, T_q, k_pool, v_pool, num_blocks, block_table, block_table_stride, context_lens,
                                    cu_seqlens_q, out, out_token_stride, B, max_q, n_heads, n_kv_heads, softmax_scale, st);
  b200_set_last_error("attn_prefill_paged: head_dim must be 64 or 128");
  return B200_ERR_UNSUPPORTED;
}
