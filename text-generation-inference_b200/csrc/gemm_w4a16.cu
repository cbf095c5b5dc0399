// GPTQ int4 linear  Y[T,N] = X[T,K] . dequant(Wq)[K,N]  on tcgen05 tensor cores, weights de-quantised in-kernel
// straight into TMEM (no fp16 scratch matrix, for any T).
//
// Replaces `exllamav2_kernels.make_q_matrix` / `gemm_half_q_half`
// (/root/reference/server/text_generation_server/utils/gptq/exllamav2.py:14-62,139-144), including its M > 50
// "dequantise everything to temp_dq, then cuBLAS" branch (:87), and follows the dequant formula of record
// utils/gptq/quant_linear.py:184-192:  W[k,n] = fp16( scales[g,n] * (q[k,n] - (qzeros[g,n] + 1)) ).
//
// Swap-AB like gemm_f16.cu: 128 output features = UMMA M.  Per 64-wide k-block
//   warp 0      TMA: packed int4 tile [8 words x 128 features] (4 KB) + activation tile [TN x 64] fp16 (128B swizzle)
//   warps 2..9  dequant: thread = one feature row; LOP3 nibble-pair extraction with the 0x6400 magic bias, exact
//               zero-point subtraction (HSUB2 / HFMA2 x 1/16), one HMUL2 by the group scale -> 32 packed fp16 pairs
//               -> tcgen05.st into a 4-deep ring of A-operand tiles in TMEM (32 columns each)
//   warp 1      one thread issues tcgen05.mma.kind::f16 with A from TMEM, B from shared memory, D (fp32) in TMEM
//   warps 2..9  epilogue after the K loop: tcgen05.ld -> fp16 -> HBM (or fp32 split-K partials + ordered reduce)
// `b200_gptq_repack` re-orders the 8 nibbles of every qweight word once at load time (k0 k2 k4 k6 | k1 k3 k5 k7)
// so that one LOP3 yields a (k, k+1) half2 pair; scales / qzeros / g_idx stay in the checkpoint layout.
#include "common.cuh"
#include "tmap.cuh"

int b200_pick_splits(int n_tiles, int n_kblocks);

namespace b200 {

constexpr int kW4Threads = 320;  // TMA warp, MMA warp, 8 dequant/epilogue warps
constexpr int kW4TileM = 128;
constexpr int kW4TileK = 64;
constexpr int kW4AStages = 4;        // TMEM A-operand ring
constexpr int kW4AColsPerStage = 32; // 64 fp16 per row = 32 x 32-bit columns
constexpr int64_t kW4CounterBytes = 64 * 1024;

template <int TN>
struct GemmW4Cfg {
  static constexpr int kQBytes = (kW4TileK / 8) * kW4TileM * 4;  // 4096
  static constexpr int kBBytes = TN * kW4TileK * 2;
  static constexpr int kStageBytes = kQBytes + kBBytes;
  static constexpr int kStages = TN <= 64 ? 8 : (TN == 128 ? 6 : 4);
  static constexpr int kTmemColsNeeded = TN + kW4AStages * kW4AColsPerStage;
  static constexpr int kTmemCols = kTmemColsNeeded <= 256 ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + (2 * kStages + 2 * kW4AStages + 1) * 8 + 16 + 1024;
};

__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(b), "r"(c));  // (a & b) | c
  return r;
}

// one repacked word (8 weights along k) -> 4 half2 (k,k+1) pairs, each fp16(scale * (q - zero))
__device__ __forceinline__ void dequant_word(uint32_t w, __half2 z1024, __half2 z64, __half2 scale, uint32_t* out) {
  const uint32_t kMagic = 0x64006400u;               // half2(1024, 1024)
  const __half2 k16th = __floats2half2_rn(0.0625f, 0.0625f);
  uint32_t q0 = lop3_and_or(w, 0x000f000fu, kMagic);  // 1024 + q      (k0, k1)
  uint32_t q1 = lop3_and_or(w, 0x00f000f0u, kMagic);  // 1024 + 16 q   (k2, k3)
  const uint32_t w8 = w >> 8;
  uint32_t q2 = lop3_and_or(w8, 0x000f000fu, kMagic);  // (k4, k5)
  uint32_t q3 = lop3_and_or(w8, 0x00f000f0u, kMagic);  // (k6, k7)
  __half2 h0 = __hsub2(*reinterpret_cast<__half2*>(&q0), z1024);
  __half2 h1 = __hfma2(*reinterpret_cast<__half2*>(&q1), k16th, z64);
  __half2 h2 = __hsub2(*reinterpret_cast<__half2*>(&q2), z1024);
  __half2 h3 = __hfma2(*reinterpret_cast<__half2*>(&q3), k16th, z64);
  h0 = __hmul2(h0, scale);
  h1 = __hmul2(h1, scale);
  h2 = __hmul2(h2, scale);
  h3 = __hmul2(h3, scale);
  out[0] = *reinterpret_cast<uint32_t*>(&h0);
  out[1] = *reinterpret_cast<uint32_t*>(&h1);
  out[2] = *reinterpret_cast<uint32_t*>(&h2);
  out[3] = *reinterpret_cast<uint32_t*>(&h3);
}

template <int TN>
__global__ void __launch_bounds__(kW4Threads, 1)
gemm_w4a16_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                  const int32_t* __restrict__ qzeros, const __half* __restrict__ scales, __half* __restrict__ y,
                  float* __restrict__ partial, int* __restrict__ counters, const __half* __restrict__ bias, int T, int N,
                  int n_kblocks, int kblocks_per_split, int groupsize) {
  using C = GemmW4Cfg<TN>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* a_full = empty_bar + C::kStages;
  uint64_t* a_empty = a_full + kW4AStages;
  uint64_t* tmem_full = a_empty + kW4AStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  __shared__ int s_is_last;

  const int warp = warp_id(), lane = lane_id();
  const int n0 = blockIdx.x * kW4TileM, t0 = blockIdx.y * TN, split = blockIdx.z, n_splits = gridDim.z;
  const int kb0 = split * kblocks_per_split;
  const int kb1 = min(n_kblocks, kb0 + kblocks_per_split);
  const int nkb = kb1 - kb0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1 + 8);  // MMA commit (activation tile) + 8 dequant warps (packed tile)
    }
    for (int s = 0; s < kW4AStages; ++s) {
      mbar_init(&a_full[s], 8);
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<C::kTmemCols>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d = tmem_base;            // columns [0, TN)
  const uint32_t tmem_a = tmem_base + TN;       // columns [TN, TN + 128)

  if (warp == 0) {
    if (elect_one()) {
      const uint64_t pol_w = policy_evict_first(), pol_x = policy_evict_last();
      for (int i = 0; i < nkb; ++i) {
        const int s = i % C::kStages;
        mbar_wait(&empty_bar[s], ((i / C::kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], C::kStageBytes);
        unsigned char* qs = smem + s * C::kStageBytes;
        tma_load_2d_hint(qs, &tmap_q, n0, (kb0 + i) * (kW4TileK / 8), &full_bar[s], pol_w);
        tma_load_2d_hint(qs + C::kQBytes, &tmap_x, (kb0 + i) * kW4TileK, t0, &full_bar[s], pol_x);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_f16_f32acc(kW4TileM, TN);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % C::kStages, as = i % kW4AStages;
      mbar_wait(&full_bar[s], (i / C::kStages) & 1);
      mbar_wait(&a_full[as], (i / kW4AStages) & 1);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint64_t bdesc = umma_desc_kmajor_sw128(smem_u32(smem + s * C::kStageBytes + C::kQBytes));
#pragma unroll
        for (int k = 0; k < kW4TileK / 16; ++k)
          umma_f16_ts(tmem_d, tmem_a + as * kW4AColsPerStage + k * 8, bdesc + (uint64_t)(k * 2), idesc, (i | k) != 0);
        umma_commit(&empty_bar[s]);
        umma_commit(&a_empty[as]);
        if (i == nkb - 1) umma_commit(tmem_full);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ dequant warps, then epilogue
    const int dw = warp - 2;            // 0..7
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int half = dw >> 2;           // which 32-wide k half of the k-block
    const int m = quarter * 32 + lane;  // feature row within the tile
    const int n = n0 + m;
    const bool n_ok = n < N;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    int cur_group = -1;
    __half2 sc2 = __floats2half2_rn(0.f, 0.f), z1024 = sc2, z64 = sc2;
    for (int i = 0; i < nkb; ++i) {
      const int s = i % C::kStages, as = i % kW4AStages;
      const int k_first = (kb0 + i) * kW4TileK + half * 32;
      const int grp = groupsize > 0 ? k_first / groupsize : 0;
      if (grp != cur_group) {
        cur_group = grp;
        if (n_ok) {
          const __half sv = scales[(size_t)grp * N + n];
          const int zp = ((qzeros[(size_t)grp * (N >> 3) + (n >> 3)] >> ((n & 7) * 4)) & 15) + 1;
          sc2 = __half2half2(sv);
          z1024 = __half2half2(__int2half_rn(1024 + zp));
          z64 = __half2half2(__int2half_rn(-(64 + zp)));
        }
      }
      mbar_wait(&full_bar[s], (i / C::kStages) & 1);
      const uint32_t* qs = reinterpret_cast<const uint32_t*>(smem + s * C::kStageBytes);
      uint32_t w[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) w[r] = qs[(half * 4 + r) * kW4TileM + m];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
      uint32_t v[16];
#pragma unroll
      for (int r = 0; r < 4; ++r) dequant_word(w[r], z1024, z64, sc2, &v[r * 4]);
      mbar_wait(&a_empty[as], ((i / kW4AStages) & 1) ^ 1);
      tcgen05_fence_after();
      tmem_st_32x32b_x16(tmem_a + lane_base + as * kW4AColsPerStage + half * 16, v);
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[as]);
    }

    if (nkb > 0) {
      mbar_wait(tmem_full, 0);
      tcgen05_fence_after();
    }
    const float bv = (bias && n_ok) ? __half2float(bias[n]) : 0.f;
#pragma unroll 1
    for (int c = half * 16; c < TN; c += 32) {
      uint32_t v[16];
      if (nkb > 0) {
        tmem_ld_32x32b_x16(tmem_d + lane_base + c, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0;
      }
      if (n_ok) {
        if (n_splits == 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int t = t0 + c + j;
            if (t < T) y[(size_t)t * N + n] = __float2half_rn(__uint_as_float(v[j]) + bv);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int t = t0 + c + j;
            if (t < T) partial[((size_t)split * T + t) * N + n] = __uint_as_float(v[j]);
          }
        }
      }
    }
  }

  if (n_splits > 1) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const int tile = blockIdx.y * gridDim.x + blockIdx.x;
      const int prev = atomicAdd(&counters[tile], 1);
      s_is_last = prev == n_splits - 1;
      if (s_is_last) counters[tile] = 0;
    }
    __syncthreads();
    if (s_is_last) {
      __threadfence();
      const int t_hi = min(T, t0 + TN);
      for (int idx = threadIdx.x; idx < (t_hi - t0) * kW4TileM; idx += kW4Threads) {
        const int t = t0 + idx / kW4TileM, n = n0 + idx % kW4TileM;
        if (n < N) {
          float acc = 0.f;
          for (int s = 0; s < n_splits; ++s) acc += __ldcg(&partial[((size_t)s * T + t) * N + n]);
          if (bias) acc += __half2float(bias[n]);
          y[(size_t)t * N + n] = __float2half_rn(acc);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<C::kTmemCols>(tmem_base);
}

// in-place nibble re-order of qweight [K/8][N]: (k0..k7) -> low half k0 k2 k4 k6, high half k1 k3 k5 k7
__global__ void gptq_repack_kernel(uint32_t* __restrict__ qweight, int64_t n_words, int inverse) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t w = qweight[i];
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int pos = (k & 1) * 4 + (k >> 1);  // nibble position of k in the repacked word
      if (!inverse) r |= ((w >> (4 * k)) & 15u) << (4 * pos);
      else r |= ((w >> (4 * pos)) & 15u) << (4 * k);
    }
    qweight[i] = r;
  }
}

}  // namespace b200

using namespace b200;

static int pick_tn_w4(int64_t T) { return T <= 16 ? 16 : T <= 32 ? 32 : T <= 64 ? 64 : T <= 128 ? 128 : 256; }

extern "C" int b200_gptq_repack(void* qweight, int64_t K, int64_t N, int inverse, void* stream) {
  if (K % 8 != 0) { b200_set_last_error("gptq_repack: K % 8 != 0"); return B200_ERR_ARG; }
  const int64_t n_words = K / 8 * N;
  if (n_words == 0) return B200_OK;
  int64_t blocks = (n_words + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  gptq_repack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((uint32_t*)qweight, n_words, inverse);
  B200_CHECK_LAUNCH();
  b200_count_launches(1);
  return B200_OK;
}

template <int TN>
static int launch_gemm_w4(const CUtensorMap* mq, const CUtensorMap* mx, const void* qzeros, const void* scales, void* y,
                          void* workspace, const void* bias, int T, int N, int K, int groupsize, cudaStream_t st) {
  using C = GemmW4Cfg<TN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_w4a16_kernel<TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) { b200_set_last_error(cudaGetErrorString(e)); return B200_ERR_CUDA; }
    configured = true;
  }
  const int n_tiles_n = (N + kW4TileM - 1) / kW4TileM, n_tiles_t = (T + TN - 1) / TN;
  const int n_kblocks = (K + kW4TileK - 1) / kW4TileK;
  int splits = workspace ? b200_pick_splits(n_tiles_n * n_tiles_t, n_kblocks) : 1;
  if ((int64_t)n_tiles_n * n_tiles_t * 4 > kW4CounterBytes) splits = 1;
  const int per = (n_kblocks + splits - 1) / splits;
  splits = (n_kblocks + per - 1) / per;
  int* counters = (int*)workspace;
  float* partial = workspace ? (float*)((char*)workspace + kW4CounterBytes) : nullptr;
  dim3 grid(n_tiles_n, n_tiles_t, splits);
  b200_timing_mark(B200_TIME_GEMM_W4A16, 0, st);
  gemm_w4a16_kernel<TN><<<grid, kW4Threads, C::kSmemBytes, st>>>(*mq, *mx, (const int32_t*)qzeros, (const __half*)scales, (__half*)y,
                                                                 partial, counters, (const __half*)bias, T, N, n_kblocks, per,
                                                                 groupsize);
  b200_timing_mark(B200_TIME_GEMM_W4A16, 1, st);
  B200_CHECK_LAUNCH();
  b200_count_launches(1);
  return B200_OK;
}

// qweight must have been passed through b200_gptq_repack once.  groupsize: multiple of 32, or <= 0 for one group.
// workspace as for b200_gemm_f16 (b200_gemm_workspace_bytes).
extern "C" int b200_gemm_w4a16(const void* x, const void* qweight_repacked, const void* qzeros, const void* scales,
                               const void* bias, void* y, int64_t T, int64_t N, int64_t K, int groupsize, void* workspace,
                               void* stream) {
  if (T == 0 || N == 0) return B200_OK;
  if (K % 64 != 0 || N % 8 != 0 || (groupsize > 0 && groupsize % 32 != 0)) {
    b200_set_last_error("gemm_w4a16: need K % 64 == 0, N % 8 == 0, groupsize % 32 == 0");
    return B200_ERR_ARG;
  }
  const int TN = pick_tn_w4(T);
  const CUtensorMap* mq = get_tmap_2d(qweight_repacked, K / 8, N, N, kW4TileK / 8, kW4TileM, TmapDtype::kI32, TmapSwizzle::kNone);
  const CUtensorMap* mx = get_tmap_2d(x, T, K, K, TN, kW4TileK, TmapDtype::kF16, TmapSwizzle::k128B);
  if (!mq || !mx) return B200_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  switch (TN) {
    case 16: return launch_gemm_w4<16>(mq, mx, qzeros, scales, y, workspace, bias, (int)T, (int)N, (int)K, groupsize, st);
    case 32: return launch_gemm_w4<32>(mq, mx, qzeros, scales, y, workspace, bias, (int)T, (int)N, (int)K, groupsize, st);
    case 64: return launch_gemm_w4<64>(mq, mx, qzeros, scales, y, workspace, bias, (int)T, (int)N, (int)K, groupsize, st);
    case 128: return launch_gemm_w4<128>(mq, mx, qzeros, scales, y, workspace, bias, (int)T, (int)N, (int)K, groupsize, st);
    default: return launch_gemm_w4<256>(mq, mx, qzeros, scales, y, workspace, bias, (int)T, (int)N, (int)K, groupsize, st);
  }
}
