// fp16 linear  Y[T,N] = X[T,K] . W[N,K]^T  (fp32 accumulate, fp16 out) on tcgen05 tensor cores.
//
// Replaces the cuBLAS HGEMM behind `FastLinear.forward` / `TensorParallelHead.forward`
// (/root/reference/server/text_generation_server/utils/layers.py:110-111, 257-262).
//
// Swap-AB: the 128 output features of a weight tile are the UMMA M dimension (A operand, K-major, streamed
// from HBM once by TMA with 128B swizzle), the step's tokens are the UMMA N dimension (B operand, TN = 16..256),
// the fp32 accumulator D[128 x TN] lives in TMEM.  Decode (T <= 256) is weight-streaming / HBM-bound: the flattened
// (feature tile, 64-wide k-block) unit space is cut into equal contiguous ranges, one per SM (stream-K), so all 148 SMs
// pull the same number of weight bytes; a tile that straddles CTAs gets its fp32 partials summed in contributor order by
// the last CTA to finish (deterministic).  PDL: the weight ring is filled before griddepcontrol.wait.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (one elected thread), 2..5 = epilogue (TMEM -> registers -> HBM).
#include "common.cuh"
#include "tmap.cuh"
#include "../../include/b200_tgis.h"

#include <cstdlib>

namespace b200 {

constexpr int kGemmThreads = 192;
constexpr int kTileM = 128;  // features per tile
constexpr int kTileK = 64;   // fp16 elements per k-block = one 128-byte swizzle row
constexpr int64_t kCounterBytes = 64 * 1024;

template <int TN>
struct GemmF16Cfg {
  static constexpr int kABytes = kTileM * kTileK * 2;
  static constexpr int kBBytes = TN * kTileK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = TN <= 64 ? 8 : (TN == 128 ? 6 : 4);
  static constexpr int kTmemCols = TN < 32 ? 32 : TN;
  static constexpr int kSmemBytes = kStages * kStageBytes + (2 * kStages + 2) * 8 + 16 + 1024;
};

struct F16Params {
  __half* y;
  float* partial;  // [token tile][feature tile][contributor][TN][128] fp32
  int* counters;   // [token tile][feature tile]
  const __half* bias;
  int T, N;
  int nkb;            // 64-wide k-blocks per tile
  int n_tiles_n;      // feature tiles
  int units_per_cta;  // contiguous (tile, k-block) units per CTA
  int total_units;    // per token tile
  int max_contrib;    // partial slots per tile
  int defer;          // 1: leave every tile segment as an fp32 partial for the consumer kernel (B200SplitK), no fix-up
};

template <int TN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, const F16Params p) {
  using C = GemmF16Cfg<TN>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tmem_full = empty_bar + C::kStages;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  __shared__ int s_is_last;

  pdl_launch_dependents();
  const int warp = warp_id(), lane = lane_id();
  const int t0 = blockIdx.y * TN;
  const int u0 = blockIdx.x * p.units_per_cta;
  const int u1 = min(p.total_units, u0 + p.units_per_cta);
  const int n_units = u1 - u0;
  const int tile0 = u0 / p.nkb, kb0 = u0 - tile0 * p.nkb;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 4);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<C::kTmemCols>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      const uint64_t pol_w = policy_evict_first(), pol_x = policy_evict_last();
      // PDL: the weights do not depend on the previous kernel — fill the ring with weight tiles first, then wait for
      // the producer of x and add the activation tiles
      const int npre = min(n_units, C::kStages);
      int tile = tile0, kb = kb0;
      for (int i = 0; i < npre; ++i) {
        mbar_arrive_expect_tx(&full_bar[i], C::kStageBytes);
        tma_load_2d_hint(smem + i * C::kStageBytes, &tmap_w, kb * kTileK, tile * kTileM, &full_bar[i], pol_w);
        if (++kb == p.nkb) { kb = 0; ++tile; }
      }
      pdl_wait();
      int kbx = kb0;
      for (int i = 0; i < npre; ++i) {
        tma_load_2d_hint(smem + i * C::kStageBytes + C::kABytes, &tmap_x, kbx * kTileK, t0, &full_bar[i], pol_x);
        if (++kbx == p.nkb) kbx = 0;
      }
      for (int i = npre; i < n_units; ++i) {
        const int s = i % C::kStages;
        mbar_wait(&empty_bar[s], ((i / C::kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], C::kStageBytes);
        unsigned char* a = smem + s * C::kStageBytes;
        tma_load_2d_hint(a, &tmap_w, kb * kTileK, tile * kTileM, &full_bar[s], pol_w);
        tma_load_2d_hint(a + C::kABytes, &tmap_x, kb * kTileK, t0, &full_bar[s], pol_x);
        if (++kb == p.nkb) { kb = 0; ++tile; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_f16_f32acc(kTileM, TN);
    int seg = 0, kb = kb0;
    for (int i = 0; i < n_units; ++i) {
      const bool seg_first = (i == 0) || kb == 0;
      const bool seg_last = (i == n_units - 1) || kb == p.nkb - 1;
      const int s = i % C::kStages;
      if (seg_first && seg > 0) mbar_wait(tmem_empty, (seg - 1) & 1);  // epilogue drained the previous segment's D
      mbar_wait(&full_bar[s], (i / C::kStages) & 1);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(smem + s * C::kStageBytes);
        const uint64_t adesc = umma_desc_kmajor_sw128(a_addr);
        const uint64_t bdesc = umma_desc_kmajor_sw128(a_addr + C::kABytes);
#pragma unroll
        for (int k = 0; k < kTileK / 16; ++k)
          umma_f16_ss(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (!seg_first || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[s]);
        if (seg_last) umma_commit(tmem_full);
      }
      __syncwarp();
      if (seg_last) ++seg;
      if (++kb == p.nkb) kb = 0;
    }
  } else {
    // ------------------------------------------------------------------ epilogue: thread = one output feature
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const int etid = threadIdx.x - 64;  // 0..127
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    pdl_wait();  // outputs / bias / stream-K workspace belong to the stream order
    int tile = tile0, kb = kb0, seg = 0;
    for (int i = 0; i < n_units;) {
      const int seg_len = min(p.nkb - kb, n_units - i);
      const int n = tile * kTileM + m;
      const bool n_ok = n < p.N;
      const int c_first = (tile * p.nkb) / p.units_per_cta;
      const int c_last = ((tile + 1) * p.nkb - 1) / p.units_per_cta;
      const int n_contrib = c_last - c_first + 1;
      const int my_contrib = (int)blockIdx.x - c_first;
      const int tix = blockIdx.y * p.n_tiles_n + tile;
      mbar_wait(tmem_full, seg & 1);
      tcgen05_fence_after();
      const float bv = (p.bias && n_ok) ? __half2float(p.bias[n]) : 0.f;
      float* part = p.partial + ((size_t)tix * p.max_contrib + my_contrib) * (TN * kTileM);
#pragma unroll 1
      for (int c = 0; c < TN; c += 16) {
        uint32_t d[16];
        tmem_ld_32x32b_x16(tmem_d + lane_base + c, d);
        tmem_ld_wait();
        if (n_contrib == 1 && !p.defer) {
          if (n_ok) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int t = t0 + c + j;
              if (t < p.T) p.y[(size_t)t * p.N + n] = __float2half_rn(__uint_as_float(d[j]) + bv);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) part[(c + j) * kTileM + m] = __uint_as_float(d[j]);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
      if (n_contrib > 1 && !p.defer) {
        // last-arriving contributor sums the slots in contributor order (deterministic); release / acquire through
        // thread 0's gpu-scope fences around the counter, ordered with the other threads by the named barrier
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (etid == 0) {
          __threadfence();
          const int prev = atomicAdd(&p.counters[tix], 1);
          s_is_last = prev == n_contrib - 1;
          if (s_is_last) {
            p.counters[tix] = 0;  // re-armed for the next launch (graph replay safe)
            __threadfence();
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (s_is_last) {
          const float* base = p.partial + (size_t)tix * p.max_contrib * (TN * kTileM);
          const int n_vec = min(p.T - t0, TN) * (kTileM / 4);
          const bool vec_ok = (p.N & 3) == 0;
          for (int idx0 = etid; idx0 < n_vec; idx0 += 256) {
            float4 acc[2];
            float4 ld[2][4];
#pragma unroll
            for (int e = 0; e < 2; ++e) acc[e] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c0 = 0; c0 < n_contrib; c0 += 4) {
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int idx = idx0 + e * 128;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                  ld[e][cc] = (idx < n_vec && c0 + cc < n_contrib)
                                  ? __ldcg(reinterpret_cast<const float4*>(&base[(size_t)(c0 + cc) * (TN * kTileM) + idx * 4]))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                }
              }
#pragma unroll
              for (int e = 0; e < 2; ++e) {
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                  acc[e].x += ld[e][cc].x; acc[e].y += ld[e][cc].y; acc[e].z += ld[e][cc].z; acc[e].w += ld[e][cc].w;
                }
              }
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int idx = idx0 + e * 128;
              const int tt = idx / (kTileM / 4), mm = (idx % (kTileM / 4)) * 4;
              const int nn = tile * kTileM + mm;
              if (idx < n_vec) {
                const float a4[4] = {acc[e].x, acc[e].y, acc[e].z, acc[e].w};
                if (vec_ok && nn + 3 < p.N) {
                  float b4[4] = {0.f, 0.f, 0.f, 0.f};
                  if (p.bias) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) b4[q] = __half2float(p.bias[nn + q]);
                  }
                  uint2 o;
                  o.x = pack_half2(a4[0] + b4[0], a4[1] + b4[1]);
                  o.y = pack_half2(a4[2] + b4[2], a4[3] + b4[3]);
                  *reinterpret_cast<uint2*>(&p.y[(size_t)(t0 + tt) * p.N + nn]) = o;
                } else {
#pragma unroll
                  for (int q = 0; q < 4; ++q)
                    if (nn + q < p.N)
                      p.y[(size_t)(t0 + tt) * p.N + nn + q] = __float2half_rn(a4[q] + (p.bias ? __half2float(p.bias[nn + q]) : 0.f));
                }
              }
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // s_is_last is reused by the next segment
      }
      i += seg_len;
      kb = 0;
      ++tile;
      ++seg;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<C::kTmemCols>(tmem_d);
}

struct F16Plan {
  int TN, nkb, n_tiles_n, n_tiles_t, units_per_cta, n_ctas, max_contrib;
};

static int f16_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

static F16Plan plan_f16(int64_t T, int64_t N, int64_t K, int sms, bool defer = false) {
  F16Plan pl;
  pl.TN = T <= 16 ? 16 : T <= 32 ? 32 : T <= 64 ? 64 : T <= 128 ? 128 : 256;
  pl.nkb = (int)((K + kTileK - 1) / kTileK);
  pl.n_tiles_n = (int)((N + kTileM - 1) / kTileM);
  pl.n_tiles_t = (int)((T + pl.TN - 1) / pl.TN);
  const int total = pl.n_tiles_n * pl.nkb;
  if (pl.n_tiles_t > 1) {
    pl.units_per_cta = pl.nkb;  // whole tiles (prefill: plenty of tiles)
  } else {
    const int ctas = total < sms ? total : sms;
    pl.units_per_cta = (total + ctas - 1) / ctas;
    if (pl.units_per_cta < 4 && pl.nkb >= 4) pl.units_per_cta = 4;
    // prefer the largest aligned cut nkb / d (d = 8, 4, 2, 1) that fits the SM count: a CTA that straddles two tiles pays two
    // fix-ups (measured on B200, round 2: 17.4 -> 14.4 us at 64 x 4096 x 4096, 47.7 -> 33.9 us at T = 256)
    if (!defer) {
      for (int d = 8; d >= 1; d >>= 1) {
        if (pl.nkb % d == 0 && (int64_t)pl.n_tiles_n * d <= sms && pl.nkb / d >= 4) { pl.units_per_cta = pl.nkb / d; break; }
      }
    }
  }
  pl.n_ctas = (total + pl.units_per_cta - 1) / pl.units_per_cta;
  pl.max_contrib = (pl.nkb + pl.units_per_cta - 1) / pl.units_per_cta + 1;
  if (pl.units_per_cta % pl.nkb == 0) pl.max_contrib = 1;
  return pl;
}

}  // namespace b200

using namespace b200;

int64_t b200_w4_partial_bytes(int64_t T, int64_t N, int64_t K);  // gemm_w4a16.cu (stream-K partials)
void b200_w4_plan_debug(int64_t T, int64_t N, int64_t K, int sms, int32_t* out);

// tests (no GPU needed): the stream-K plan of one GEMM launch.  kind 0 = fp16 weights, 1 = int4.
// out[8] = {token tile, k-blocks per tile, feature tiles (int4: super-tiles), token tiles, units per CTA, CTAs,
//           contributor slots per tile in the workspace, feature tiles per unit (int4: 2)}
extern "C" int b200_debug_gemm_plan(int kind, int64_t T, int64_t N, int64_t K, int sms, int32_t* out) {
  if (!out || T <= 0 || N <= 0 || K <= 0 || sms <= 0 || (kind != 0 && kind != 1)) {
    b200_set_last_error("debug_gemm_plan: bad arguments");
    return B200_ERR_ARG;
  }
  if (kind == 1) {
    b200_w4_plan_debug(T, N, K, sms, out);
    return B200_OK;
  }
  const F16Plan pl = plan_f16(T, N, K, sms);
  const int32_t v[8] = {pl.TN, pl.nkb, pl.n_tiles_n, pl.n_tiles_t, pl.units_per_cta, pl.n_ctas, pl.max_contrib, 1};
  for (int i = 0; i < 8; ++i) out[i] = v[i];
  return B200_OK;
}

extern "C" int64_t b200_gemm_workspace_bytes(int64_t T, int64_t N, int64_t K) {
  const F16Plan pl = plan_f16(T, N, K, 148);
  int64_t f16 = pl.max_contrib > 1 ? (int64_t)pl.n_tiles_t * pl.n_tiles_n * pl.max_contrib * pl.TN * kTileM * 4 : 0;
  if (pl.n_tiles_t == 1) {  // deferred reduction (one token tile) leaves a partial for every tile, even un-split ones
    const F16Plan pd = plan_f16(T, N, K, 148, true);
    const int64_t fd = (int64_t)pd.n_tiles_n * pd.max_contrib * pd.TN * kTileM * 4;
    if (fd > f16) f16 = fd;
  }
  const int64_t w4 = b200_w4_partial_bytes(T, N, K);
  return kCounterBytes + (f16 > w4 ? f16 : w4);
}

// upper bound of b200_gemm_workspace_bytes over every T (stream-K only happens while there is a single token tile)
extern "C" int64_t b200_gemm_workspace_bytes_max(int64_t N, int64_t K) {
  int64_t best = kCounterBytes;
  for (int64_t T : {1, 16, 17, 32, 33, 64, 65, 128, 129, 256}) {
    const int64_t b = b200_gemm_workspace_bytes(T, N, K);
    if (b > best) best = b;
  }
  return best;
}

template <int TN>
static int launch_gemm_f16(const CUtensorMap* mw, const CUtensorMap* mx, void* y, void* workspace, const void* bias, int T, int N,
                           const F16Plan& pl, int defer, cudaStream_t st) {
  using C = GemmF16Cfg<TN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_f16_kernel<TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) { b200_set_last_error(cudaGetErrorString(e)); return B200_ERR_CUDA; }
    configured = true;
  }
  F16Params p;
  p.y = (__half*)y;
  p.counters = (int*)workspace;
  p.partial = workspace ? (float*)((char*)workspace + kCounterBytes) : nullptr;
  p.bias = (const __half*)bias;
  p.T = T;
  p.N = N;
  p.nkb = pl.nkb;
  p.n_tiles_n = pl.n_tiles_n;
  p.units_per_cta = pl.units_per_cta;
  p.total_units = pl.n_tiles_n * pl.nkb;
  p.max_contrib = pl.max_contrib;
  p.defer = defer;
  dim3 grid(pl.n_ctas, pl.n_tiles_t, 1);
  b200_timing_mark(B200_TIME_GEMM_F16, 0, st);
  B200_LAUNCH(gemm_f16_kernel<TN>, grid, dim3(kGemmThreads), (size_t)C::kSmemBytes, st, *mw, *mx, p);
  b200_timing_mark(B200_TIME_GEMM_F16, 1, st);
  b200_count_launches(1);
  return B200_OK;
}

// workspace: >= b200_gemm_workspace_bytes(T, N, K) bytes whose first 64 KiB (tile counters) were zeroed once by the
// caller, or NULL (every CTA then takes whole tiles: no stream-K).
// splitk != NULL: deferred reduction (see b200_gemm_w4a16_deferred): y is not written.
static int gemm_f16_impl(const void* x, const void* w, const void* bias, void* y, int64_t T, int64_t N, int64_t K, void* workspace,
                         B200SplitK* splitk, void* stream) {
  if (T == 0 || N == 0) return B200_OK;
  if (K % 8 != 0 || K < 64) { b200_set_last_error("gemm_f16: need K % 8 == 0 and K >= 64"); return B200_ERR_ARG; }
  const bool defer = splitk != nullptr;
  if (defer && (!workspace || T > 256 || N % 8 != 0)) {
    b200_set_last_error("gemm_f16_deferred: needs a workspace, T <= 256 (one token tile) and N % 8 == 0");
    return B200_ERR_UNSUPPORTED;
  }
  F16Plan pl = plan_f16(T, N, K, f16_num_sms(), defer);
  if (!defer && pl.max_contrib > 1 && (!workspace || (int64_t)pl.n_tiles_n * pl.n_tiles_t * 4 > kCounterBytes)) {
    pl.units_per_cta = pl.nkb;
    pl.n_ctas = pl.n_tiles_n;
    pl.max_contrib = 1;
  }
  const CUtensorMap* mw = get_tmap_2d(w, N, K, K, kTileM, kTileK, TmapDtype::kF16, TmapSwizzle::k128B);
  const CUtensorMap* mx = get_tmap_2d(x, T, K, K, pl.TN, kTileK, TmapDtype::kF16, TmapSwizzle::k128B);
  if (!mw || !mx) return B200_ERR_CUDA;
  if (defer) {
    splitk->partial = (const float*)((const char*)workspace + kCounterBytes);
    splitk->bias = bias;
    splitk->tiles_per_unit = 1;
    splitk->tn = pl.TN;
    splitk->nkb = pl.nkb;
    splitk->units_per_cta = pl.units_per_cta;
    splitk->max_contrib = pl.max_contrib;
    splitk->half_tiles = 0;
    splitk->N = (int32_t)N;
    splitk->T = (int32_t)T;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int df = defer ? 1 : 0;
  switch (pl.TN) {
    case 16: return launch_gemm_f16<16>(mw, mx, y, workspace, bias, (int)T, (int)N, pl, df, st);
    case 32: return launch_gemm_f16<32>(mw, mx, y, workspace, bias, (int)T, (int)N, pl, df, st);
    case 64: return launch_gemm_f16<64>(mw, mx, y, workspace, bias, (int)T, (int)N, pl, df, st);
    case 128: return launch_gemm_f16<128>(mw, mx, y, workspace, bias, (int)T, (int)N, pl, df, st);
    default: return launch_gemm_f16<256>(mw, mx, y, workspace, bias, (int)T, (int)N, pl, df, st);
  }
}
extern "C" int b200_gemm_f16(const void* x, const void* w, const void* bias, void* y, int64_t T, int64_t N, int64_t K,
                             void* workspace, void* stream) {
  return gemm_f16_impl(x, w, bias, y, T, N, K, workspace, nullptr, stream);
}
extern "C" int b200_gemm_f16_deferred(const void* x, const void* w, const void* bias, int64_t T, int64_t N, int64_t K, void* workspace,
                                      B200SplitK* splitk, void* stream) {
  if (!splitk) { b200_set_last_error("gemm_f16_deferred: splitk is NULL"); return B200_ERR_ARG; }
  return gemm_f16_impl(x, w, bias, nullptr, T, N, K, workspace, splitk, stream);
}
