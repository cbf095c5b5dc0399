// fp16 linear  Y[T,N] = X[T,K] . W[N,K]^T  (fp32 accumulate, fp16 out) on tcgen05 tensor cores.
//
// Replaces the cuBLAS HGEMM behind `FastLinear.forward` / `TensorParallelHead.forward`
// (/root/reference/server/text_generation_server/utils/layers.py:110-111, 257-262).
//
// Swap-AB: the 128 output features of a weight tile are the UMMA M dimension (A operand, K-major, streamed
// from HBM once by TMA with 128B swizzle), the step's tokens are the UMMA N dimension (B operand, TN = 16..256),
// the fp32 accumulator D[128 x TN] lives in TMEM.  Decode (T <= 64) is weight-streaming / HBM-bound: split-K
// spreads the K range over CTAs so that all 148 SMs pull weights; partials go to an fp32 workspace and the last
// CTA of each tile reduces them in fixed order (deterministic).
// Warp roles: 0 = TMA producer, 1 = MMA issuer (one elected thread), 2..5 = epilogue (TMEM -> registers -> HBM).
#include "common.cuh"
#include "tmap.cuh"

namespace b200 {

constexpr int kGemmThreads = 192;
constexpr int kTileM = 128;  // features per tile
constexpr int kTileK = 64;   // fp16 elements per k-block = one 128-byte swizzle row

template <int TN>
struct GemmF16Cfg {
  static constexpr int kABytes = kTileM * kTileK * 2;
  static constexpr int kBBytes = TN * kTileK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = TN <= 64 ? 8 : (TN == 128 ? 6 : 4);
  static constexpr int kTmemCols = TN < 32 ? 32 : TN;
  static constexpr int kSmemBytes = kStages * kStageBytes + (2 * kStages + 1) * 8 + 16 + 1024;
};

template <int TN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, __half* __restrict__ y,
                float* __restrict__ partial, int* __restrict__ counters, const __half* __restrict__ bias, int T, int N,
                int n_kblocks, int kblocks_per_split) {
  using C = GemmF16Cfg<TN>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tmem_full = empty_bar + C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  __shared__ int s_is_last;

  const int warp = warp_id(), lane = lane_id();
  const int n0 = blockIdx.x * kTileM, t0 = blockIdx.y * TN, split = blockIdx.z, n_splits = gridDim.z;
  const int kb0 = split * kblocks_per_split;
  const int kb1 = min(n_kblocks, kb0 + kblocks_per_split);
  const int nkb = kb1 - kb0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<C::kTmemCols>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      const uint64_t pol_w = policy_evict_first(), pol_x = policy_evict_last();
      for (int i = 0; i < nkb; ++i) {
        const int s = i % C::kStages;
        mbar_wait(&empty_bar[s], ((i / C::kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], C::kStageBytes);
        unsigned char* a = smem + s * C::kStageBytes;
        tma_load_2d_hint(a, &tmap_w, (kb0 + i) * kTileK, n0, &full_bar[s], pol_w);
        tma_load_2d_hint(a + C::kABytes, &tmap_x, (kb0 + i) * kTileK, t0, &full_bar[s], pol_x);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_f16_f32acc(kTileM, TN);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % C::kStages;
      mbar_wait(&full_bar[s], (i / C::kStages) & 1);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(smem + s * C::kStageBytes);
        const uint64_t adesc = umma_desc_kmajor_sw128(a_addr);
        const uint64_t bdesc = umma_desc_kmajor_sw128(a_addr + C::kABytes);
#pragma unroll
        for (int k = 0; k < kTileK / 16; ++k)
          umma_f16_ss(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (i | k) != 0);
        umma_commit(&empty_bar[s]);
        if (i == nkb - 1) umma_commit(tmem_full);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue: thread = one output feature
    const int quarter = warp & 3;
    const int n = n0 + quarter * 32 + lane;
    if (nkb > 0) {
      mbar_wait(tmem_full, 0);
      tcgen05_fence_after();
    }
    const float bv = (bias && n < N) ? __half2float(bias[n]) : 0.f;
#pragma unroll 1
    for (int c = 0; c < TN; c += 16) {
      uint32_t v[16];
      if (nkb > 0) {
        tmem_ld_32x32b_x16(tmem_d + ((uint32_t)(quarter * 32) << 16) + c, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0;
      }
      if (n < N) {
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
    // last-arriving CTA of this (feature tile, token tile) reduces the partials in split order
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const int tile = blockIdx.y * gridDim.x + blockIdx.x;
      const int prev = atomicAdd(&counters[tile], 1);
      s_is_last = prev == n_splits - 1;
      if (s_is_last) counters[tile] = 0;  // re-armed for the next launch (graph replay safe)
    }
    __syncthreads();
    if (s_is_last) {
      __threadfence();
      const int t_hi = min(T, t0 + TN);
      if ((N & 3) == 0) {
        // float4 per thread, all split loads of an element in flight together (latency-bound otherwise)
        for (int idx = threadIdx.x; idx < (t_hi - t0) * (kTileM / 4); idx += kGemmThreads) {
          const int t = t0 + idx / (kTileM / 4), n = n0 + (idx % (kTileM / 4)) * 4;
          if (n < N) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int s = 0; s < n_splits; ++s) {
              const float4 v = __ldcg(reinterpret_cast<const float4*>(&partial[((size_t)s * T + t) * N + n]));
              acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            if (bias) {
              acc.x += __half2float(bias[n]); acc.y += __half2float(bias[n + 1]);
              acc.z += __half2float(bias[n + 2]); acc.w += __half2float(bias[n + 3]);
            }
            uint2 o;
            o.x = pack_half2(acc.x, acc.y);
            o.y = pack_half2(acc.z, acc.w);
            *reinterpret_cast<uint2*>(&y[(size_t)t * N + n]) = o;
          }
        }
      } else {
        for (int idx = threadIdx.x; idx < (t_hi - t0) * kTileM; idx += kGemmThreads) {
          const int t = t0 + idx / kTileM, n = n0 + idx % kTileM;
          if (n < N) {
            float acc = 0.f;
            for (int s = 0; s < n_splits; ++s) acc += __ldcg(&partial[((size_t)s * T + t) * N + n]);
            if (bias) acc += __half2float(bias[n]);
            y[(size_t)t * N + n] = __float2half_rn(acc);
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<C::kTmemCols>(tmem_d);
}

}  // namespace b200

using namespace b200;

// choose split-K so that the grid covers the 148 SMs while each split keeps >= 4 k-blocks
int b200_pick_splits(int n_tiles, int n_kblocks) {
  if (n_tiles >= 120) return 1;
  int best = 1;
  for (int s = 2; s <= 16; ++s) {
    if (n_kblocks / s < 4) break;
    best = s;
    if (n_tiles * s >= 148) break;
  }
  return best;
}

static int pick_tn(int64_t T) { return T <= 16 ? 16 : T <= 32 ? 32 : T <= 64 ? 64 : T <= 128 ? 128 : 256; }

// Workspace layout shared by both GEMMs: [tile counters, kCounterBytes][fp32 split-K partials].
constexpr int64_t kCounterBytes = 64 * 1024;

int64_t b200_w4_partial_bytes(int64_t T, int64_t N, int64_t K);  // gemm_w4a16.cu (stream-K partials)

extern "C" int64_t b200_gemm_workspace_bytes(int64_t T, int64_t N, int64_t K) {
  const int TN = pick_tn(T);
  const int n_tiles = (int)(((N + kTileM - 1) / kTileM) * ((T + TN - 1) / TN));
  const int splits = b200_pick_splits(n_tiles, (int)((K + kTileK - 1) / kTileK));
  const int64_t f16 = splits > 1 ? (int64_t)splits * T * N * 4 : 0;
  const int64_t w4 = b200_w4_partial_bytes(T, N, K);
  return kCounterBytes + (f16 > w4 ? f16 : w4);
}

// upper bound of b200_gemm_workspace_bytes over every T (split-K only happens while the tile grid is small)
extern "C" int64_t b200_gemm_workspace_bytes_max(int64_t N, int64_t K) {
  int64_t best = kCounterBytes;
  for (int64_t T = 1; T <= 32768; T = T < 256 ? T + 1 : T + 256) {
    const int64_t b = b200_gemm_workspace_bytes(T, N, K);
    if (b > best) best = b;
  }
  return best;
}

template <int TN>
static int launch_gemm_f16(const CUtensorMap* mw, const CUtensorMap* mx, void* y, void* workspace, const void* bias, int T, int N,
                           int K, cudaStream_t st) {
  using C = GemmF16Cfg<TN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_f16_kernel<TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) { b200_set_last_error(cudaGetErrorString(e)); return B200_ERR_CUDA; }
    configured = true;
  }
  const int n_tiles_n = (N + kTileM - 1) / kTileM, n_tiles_t = (T + TN - 1) / TN;
  const int n_kblocks = (K + kTileK - 1) / kTileK;
  int splits = workspace ? b200_pick_splits(n_tiles_n * n_tiles_t, n_kblocks) : 1;
  if (n_tiles_n * n_tiles_t * 4 > kCounterBytes) splits = 1;
  const int per = (n_kblocks + splits - 1) / splits;
  splits = (n_kblocks + per - 1) / per;
  int* counters = (int*)workspace;
  float* partial = workspace ? (float*)((char*)workspace + kCounterBytes) : nullptr;
  dim3 grid(n_tiles_n, n_tiles_t, splits);
  b200_timing_mark(B200_TIME_GEMM_F16, 0, st);
  gemm_f16_kernel<TN><<<grid, kGemmThreads, C::kSmemBytes, st>>>(*mw, *mx, (__half*)y, partial, counters, (const __half*)bias, T, N,
                                                                  n_kblocks, per);
  b200_timing_mark(B200_TIME_GEMM_F16, 1, st);
  B200_CHECK_LAUNCH();
  b200_count_launches(1);
  return B200_OK;
}

// workspace: >= b200_gemm_workspace_bytes(T, N, K) bytes whose first 64 KiB (tile counters) were zeroed once by the
// caller, or NULL (no split-K).
extern "C" int b200_gemm_f16(const void* x, const void* w, const void* bias, void* y, int64_t T, int64_t N, int64_t K,
                             void* workspace, void* stream) {
  if (T == 0 || N == 0) return B200_OK;
  if (K % 8 != 0 || K < 64) { b200_set_last_error("gemm_f16: need K % 8 == 0 and K >= 64"); return B200_ERR_ARG; }
  const int TN = pick_tn(T);
  const CUtensorMap* mw = get_tmap_2d(w, N, K, K, kTileM, kTileK, TmapDtype::kF16, TmapSwizzle::k128B);
  const CUtensorMap* mx = get_tmap_2d(x, T, K, K, TN, kTileK, TmapDtype::kF16, TmapSwizzle::k128B);
  if (!mw || !mx) return B200_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  switch (TN) {
    case 16: return launch_gemm_f16<16>(mw, mx, y, workspace, bias, (int)T, (int)N, (int)K, st);
    case 32: return launch_gemm_f16<32>(mw, mx, y, workspace, bias, (int)T, (int)N, (int)K, st);
    case 64: return launch_gemm_f16<64>(mw, mx, y, workspace, bias, (int)T, (int)N, (int)K, st);
    case 128: return launch_gemm_f16<128>(mw, mx, y, workspace, bias, (int)T, (int)N, (int)K, st);
    default: return launch_gemm_f16<256>(mw, mx, y, workspace, bias, (int)T, (int)N, (int)K, st);
  }
}
