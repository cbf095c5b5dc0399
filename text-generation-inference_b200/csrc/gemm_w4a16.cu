// GPTQ int4 linear  Y[T,N] = X[T,K] . dequant(Wq)[K,N]  on tcgen05 tensor cores, weights de-quantised in-kernel
// straight into TMEM (no fp16 scratch matrix, for any T).
//
// Replaces `exllamav2_kernels.make_q_matrix` / `gemm_half_q_half`
// (/root/reference/server/text_generation_server/utils/gptq/exllamav2.py:14-62,139-144), including its M > 50
// "dequantise everything to temp_dq, then cuBLAS" branch (:87), and follows the dequant formula of record
// utils/gptq/quant_linear.py:184-192:  W[k,n] = fp16( scales[g,n] * (q[k,n] - (qzeros[g,n] + 1)) ).
//
// Swap-AB like gemm_f16.cu: 128 output features = UMMA M.  Work unit = (feature tile, 128-wide k-block).
//   warp 0       TMA, weight ring (12 deep): packed int4 tile [16 words x 128 features] (8 KB) + its scale / zero rows.
//                Deep because only these bytes come from HBM: ~100 KB in flight per SM covers the DRAM latency.
//   warp 2       TMA, activation ring (4 deep): [TN x 128] fp16 as two 128B-swizzled [TN x 64] sub-tiles (L2 hits)
//   warps 3..18  dequant, two groups of 8 warps on alternate units: thread = one feature row, 64 of the 128 k per warp;
//                LOP3 nibble-pair extraction with the 0x6400 magic bias, exact zero-point subtraction (HADD2 / HFMA2 x 1/16),
//                one HMUL2 by the group scale -> 32 packed fp16 pairs -> tcgen05.st into a 4-deep ring of A-operand tiles in
//                TMEM (64 columns each); the next unit's words are fetched while the stores retire
//   warp 1       one thread issues 8 x tcgen05.mma.kind::f16 per unit (A from TMEM, B from shared memory, D fp32 in TMEM)
//   warps 3..18  epilogue at the end of a tile segment (the group that owns its last unit): tcgen05.ld -> fp16 -> HBM
// Decode (T <= 128) is weight-streaming / HBM-bound: the flattened (tile, k-block) unit space is cut into equal
// contiguous ranges, one per SM (stream-K).  A tile that straddles CTAs gets its fp32 partials summed in contributor
// order by the last CTA to finish (deterministic, unlike the reference kernel's fp16 atomicAdd across K slices).
// `b200_gptq_repack` re-orders the 8 nibbles of every qweight word once at load time (k0 k2 k4 k6 | k1 k3 k5 k7)
// so that one LOP3 yields a (k, k+1) half2 pair; scales / qzeros stay in the checkpoint layout.
//
// mbarrier rule used throughout: every thread that waits on a barrier waits on every phase of it, in order (a parity
// wait that skipped a phase could be satisfied by the phase before).  Hence even ring depths (a stage always belongs
// to the same dequant group) and the "observe only" wait on tmem_full by the group that does not own a segment.
#include "common.cuh"
#include "tmap.cuh"

namespace b200 {

constexpr int kW4Threads = 19 * 32;  // W-TMA warp, MMA warp, X-TMA warp, 2 groups of 8 dequant/epilogue warps
constexpr int kW4FirstDqWarp = 3;
constexpr int kW4TileM = 128;
constexpr int kW4BlockK = 128;
constexpr int kW4AStages = 4;         // TMEM A-operand ring
constexpr int kW4AColsPerStage = 64;  // 128 fp16 per row = 64 x 32-bit columns
constexpr int kW4MaxGroupRows = 4;    // groupsize >= 32
constexpr int64_t kW4CounterBytes = 64 * 1024;

template <int TN>
struct GemmW4Cfg {
  static constexpr int kQBytes = (kW4BlockK / 8) * kW4TileM * 4;        // 8192
  static constexpr int kSBytes = kW4MaxGroupRows * kW4TileM * 2;        // 1024
  static constexpr int kZBytes = kW4MaxGroupRows * (kW4TileM / 8) * 4;  // 256
  static constexpr int kWStageBytes = kQBytes + kSBytes + kZBytes;      // 9472 = 74 * 128
  static constexpr int kXSubBytes = TN * 128;                           // one [TN x 64] fp16 sub-tile
  static constexpr int kXStageBytes = 2 * kXSubBytes;
  static constexpr int kXStages = 4;
  static constexpr int kWStages = TN <= 64 ? 12 : 8;
  static constexpr int kTmemCols = 512;
  static constexpr int kNumBars = 2 * kWStages + 2 * kXStages + 2 * kW4AStages + 2;
  static constexpr int kSmemBytes = kXStages * kXStageBytes + kWStages * kWStageBytes + kNumBars * 8 + 16 + 1024;
  static_assert(kWStages % 2 == 0 && kWStages >= 4, "weight ring: even depth, >= 4 (dequant warps prefetch two units ahead)");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(b), "r"(c));  // (a & b) | c
  return r;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint16_t lds_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}

// one repacked word (8 weights along k) -> 4 half2 (k,k+1) pairs, each fp16(scale * (q - zero))
__device__ __forceinline__ void dequant_word(uint32_t w, __half2 z1024, __half2 z64, __half2 scale, uint32_t* out) {
  const uint32_t kMagic = 0x64006400u;  // half2(1024, 1024)
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

struct W4Params {
  __half* y;
  float* partial;   // [token tile][feature tile][contributor][TN][128] fp32
  int* counters;    // [token tile][feature tile]
  const __half* bias;
  int T, N;
  int nkb;            // 128-wide k-blocks per tile
  int n_tiles_n;      // feature tiles
  int units_per_cta;  // contiguous (tile, k-block) units per CTA
  int total_units;    // per token tile
  int max_contrib;    // partial slots per tile
  int groupsize;      // > 0
  int group_rows;     // scale/zero rows per k-block = max(1, 128 / groupsize)
  unsigned long long* trace;  // debug: per-CTA phase timestamps (globaltimer ns), NULL in production
  int debug_flags;            // debug timing experiments (results invalid): 1 = skip dequant math, 2 = skip tcgen05.st
};

template <int TN>
__global__ void __launch_bounds__(kW4Threads, 1)
gemm_w4a16_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                  const __grid_constant__ CUtensorMap tmap_s, const __grid_constant__ CUtensorMap tmap_z, const W4Params p) {
  using C = GemmW4Cfg<TN>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* x_ring = smem;                                     // 1024-aligned stages (128B swizzle)
  unsigned char* w_ring = smem + C::kXStages * C::kXStageBytes;     // 128-aligned stages
  uint64_t* full_w = reinterpret_cast<uint64_t*>(w_ring + C::kWStages * C::kWStageBytes);
  uint64_t* empty_w = full_w + C::kWStages;
  uint64_t* full_x = empty_w + C::kWStages;
  uint64_t* empty_x = full_x + C::kXStages;
  uint64_t* a_full = empty_x + C::kXStages;
  uint64_t* a_empty = a_full + kW4AStages;
  uint64_t* tmem_full = a_empty + kW4AStages;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  __shared__ int s_is_last[2];

  pdl_launch_dependents();
  const int warp = warp_id(), lane = lane_id();
  const int t0 = blockIdx.y * TN;
  const int u0 = blockIdx.x * p.units_per_cta;
  const int u1 = min(p.total_units, u0 + p.units_per_cta);
  const int n_units = u1 - u0;
  const int tile0 = u0 / p.nkb, kb0 = u0 - tile0 * p.nkb;
#define W4_TRACE(id)                                                                                   \
  do {                                                                                                 \
    if (p.trace && threadIdx.x == kW4FirstDqWarp * 32) {                                               \
      unsigned long long _t;                                                                           \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                                           \
      p.trace[blockIdx.x * 64 + (id)] = _t;                                                            \
    }                                                                                                  \
  } while (0)
  W4_TRACE(0);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_s);
    tma_prefetch_desc(&tmap_z);
    for (int s = 0; s < C::kWStages; ++s) {
      mbar_init(&full_w[s], 1);
      mbar_init(&empty_w[s], 8);  // the 8 dequant warps of the group that owns the stage
    }
    for (int s = 0; s < C::kXStages; ++s) {
      mbar_init(&full_x[s], 1);
      mbar_init(&empty_x[s], 1);  // MMA commit
    }
    for (int s = 0; s < kW4AStages; ++s) {
      mbar_init(&a_full[s], 8);
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 8);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<C::kTmemCols>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  W4_TRACE(1);
  const uint32_t tmem_d = tmem_base;        // columns [0, TN)
  const uint32_t tmem_a = tmem_base + 256;  // columns [256, 512): 4 stages x 64

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: weights (HBM stream)
    if (elect_one()) {
      const uint64_t pol_w = policy_evict_first();
      const uint32_t tx = C::kQBytes + p.group_rows * (kW4TileM * 2 + (kW4TileM / 8) * 4);
      int tile = tile0, kb = kb0, s = 0, ph = 0;
      for (int i = 0; i < n_units; ++i) {
        mbar_wait(&empty_w[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_w[s], tx);
        unsigned char* st = w_ring + s * C::kWStageBytes;
        const int n0 = tile * kW4TileM, k0 = kb * kW4BlockK;
        const int g0 = k0 / p.groupsize;
        tma_load_2d_hint(st, &tmap_q, n0, k0 / 8, &full_w[s], pol_w);
        tma_load_2d_hint(st + C::kQBytes, &tmap_s, n0, g0, &full_w[s], pol_w);
        tma_load_2d_hint(st + C::kQBytes + C::kSBytes, &tmap_z, n0 / 8, g0, &full_w[s], pol_w);
        if (++kb == p.nkb) { kb = 0; ++tile; }
        if (++s == C::kWStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ TMA producer: activations (L2 resident)
    if (elect_one()) {
      const uint64_t pol_x = policy_evict_last();
      int kb = kb0;
      pdl_wait();  // x is the previous kernel's output; the weight ring (warp 0) is already streaming
      for (int i = 0; i < n_units; ++i) {
        const int s = i % C::kXStages;
        mbar_wait(&empty_x[s], ((i / C::kXStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_x[s], C::kXStageBytes);
        unsigned char* st = x_ring + s * C::kXStageBytes;
        tma_load_2d_hint(st, &tmap_x, kb * kW4BlockK, t0, &full_x[s], pol_x);
        tma_load_2d_hint(st + C::kXSubBytes, &tmap_x, kb * kW4BlockK + 64, t0, &full_x[s], pol_x);
        if (++kb == p.nkb) kb = 0;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_f16_f32acc(kW4TileM, TN);
    int seg = 0, kb = kb0;
    for (int i = 0; i < n_units; ++i) {
      const bool seg_first = (i == 0) || kb == 0;
      const bool seg_last = (i == n_units - 1) || kb == p.nkb - 1;
      const int s = i % C::kXStages, as = i % kW4AStages;
      if (seg_first && seg > 0) mbar_wait(tmem_empty, (seg - 1) & 1);  // epilogue drained the previous segment's D
      mbar_wait(&full_x[s], (i / C::kXStages) & 1);
      mbar_wait(&a_full[as], (i / kW4AStages) & 1);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t xb = smem_u32(x_ring + s * C::kXStageBytes);
#pragma unroll
        for (int k = 0; k < kW4BlockK / 16; ++k) {
          const uint64_t bdesc = umma_desc_kmajor_sw128(xb + (k >> 2) * C::kXSubBytes) + (uint64_t)((k & 3) * 2);
          umma_f16_ts(tmem_d, tmem_a + as * kW4AColsPerStage + k * 8, bdesc, idesc, (!seg_first || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_x[s]);
        umma_commit(&a_empty[as]);
        if (seg_last) umma_commit(tmem_full);
      }
      __syncwarp();
      if (seg_last) ++seg;
      if (++kb == p.nkb) kb = 0;
    }
  } else {
    // ------------------------------------------------------------------ dequant warps, epilogue at segment ends
    const int dw = warp - kW4FirstDqWarp;  // 0..15
    const int group = dw >> 3;             // which alternate units
    const int half = (dw & 7) >> 2;        // which 64-wide k half of the k-block
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int m = quarter * 32 + lane;     // feature row within the tile
    const int gtid = threadIdx.x - kW4FirstDqWarp * 32 - group * 256;  // 0..255 within the group
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t w_base = smem_u32(w_ring);
    const uint32_t q_off = (uint32_t)((half * 8) * kW4TileM + m) * 4;
    uint32_t s_off[2], z_off[2];
#pragma unroll
    for (int sub = 0; sub < 2; ++sub) {
      // the two 32-k sub-chunks of this warp's half may sit in different groups when groupsize < 64
      const int grow = (half * 64 + sub * 32) / p.groupsize;  // row inside the stage's scale/zero tile
      s_off[sub] = C::kQBytes + (uint32_t)(grow * kW4TileM + m) * 2;
      z_off[sub] = C::kQBytes + C::kSBytes + (uint32_t)(grow * (kW4TileM / 8) + (m >> 3)) * 4;
    }
    const int z_shift = (m & 7) * 4;
    uint32_t w[8];
    uint16_t sraw[2];
    uint32_t zraw[2];
    int ws = group, wph = 0;  // weight-ring stage / phase of this group's next unit
    int trace_load = 0;
    auto load_unit = [&]() {
      mbar_wait(&full_w[ws], wph);
      if (trace_load) W4_TRACE(trace_load);  // weight tile landed
      const uint32_t st = w_base + ws * C::kWStageBytes;
#pragma unroll
      for (int r = 0; r < 8; ++r) w[r] = lds_u32(st + q_off + r * (kW4TileM * 4));
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        sraw[sub] = lds_u16(st + s_off[sub]);
        zraw[sub] = lds_u32(st + z_off[sub]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_w[ws]);
      ws += 2;
      if (ws >= C::kWStages) { ws -= C::kWStages; wph ^= 1; }
    };
    if (group < n_units) load_unit();
    if (group == 0) W4_TRACE(2);
    int seg = 0, kb = kb0, tile = tile0;
    for (int i = 0; i < n_units; ++i) {
      const bool seg_last = (i == n_units - 1) || kb == p.nkb - 1;
      if ((i & 1) == group) {
        const int as = i & (kW4AStages - 1);
        uint32_t v[32];
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          const int zp = ((zraw[sub] >> z_shift) & 15) + 1;
          const __half2 sc2 = __half2half2(__ushort_as_half(sraw[sub]));
          const __half2 z1024 = __half2half2(__ushort_as_half((unsigned short)(0x6400 + zp)));       // 1024 + zp, exact
          const __half2 z64 = __half2half2(__ushort_as_half((unsigned short)(0xD400 + (zp << 4))));  // -(64 + zp), exact
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            if (p.debug_flags & 1) {
              v[(sub * 4 + r) * 4] = v[(sub * 4 + r) * 4 + 1] = v[(sub * 4 + r) * 4 + 2] = v[(sub * 4 + r) * 4 + 3] = w[sub * 4 + r] & 0x03ff03ffu;
            } else {
              dequant_word(w[sub * 4 + r], z1024, z64, sc2, &v[(sub * 4 + r) * 4]);
            }
          }
        }
        if (i == 8 || i == 10) W4_TRACE(40 + (i - 8) * 4);  // dequant done
        mbar_wait(&a_empty[as], ((i / kW4AStages) & 1) ^ 1);
        if (i == 8 || i == 10) W4_TRACE(41 + (i - 8) * 4);  // A stage free
        tcgen05_fence_after();
        const uint32_t ta = tmem_a + lane_base + as * kW4AColsPerStage + half * 32;
        if (!(p.debug_flags & 2)) {
          tmem_st_32x32b_x16(ta, reinterpret_cast<const uint32_t(&)[16]>(v[0]));
          tmem_st_32x32b_x16(ta + 16, reinterpret_cast<const uint32_t(&)[16]>(v[16]));
        } else if (v[3] == 0x12345u) {
          tmem_st_32x32b_x16(ta, reinterpret_cast<const uint32_t(&)[16]>(v[0]));  // keeps v alive
        }
        trace_load = (i == 8) ? 52 : (i == 10) ? 53 : 0;
        if (i + 2 < n_units) load_unit();  // overlaps the TMEM store latency
        if (i == 8 || i == 10) W4_TRACE(42 + (i - 8) * 4);  // next unit's words in registers
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[as]);
        if (i == 8 || i == 10) W4_TRACE(43 + (i - 8) * 4);  // stores retired, handed to the MMA warp
        if (i == 0) W4_TRACE(3);

        if (seg_last) {
          // -------------------------------------------------------------- epilogue of this tile segment (this group)
          pdl_wait();  // outputs / bias / stream-K workspace belong to the stream order
          const int n = tile * kW4TileM + m;
          const bool n_ok = n < p.N;
          // contributors of this tile: CTAs whose unit range intersects [tile*nkb, (tile+1)*nkb)
          const int c_first = (tile * p.nkb) / p.units_per_cta;
          const int c_last = ((tile + 1) * p.nkb - 1) / p.units_per_cta;
          const int n_contrib = c_last - c_first + 1;
          const int my_contrib = (int)blockIdx.x - c_first;
          const int tix = blockIdx.y * p.n_tiles_n + tile;
          if (group == 0) W4_TRACE(4 + 4 * (seg & 7));
          mbar_wait(tmem_full, seg & 1);
          if (group == 0) W4_TRACE(5 + 4 * (seg & 7));
          tcgen05_fence_after();
          const float bv = (p.bias && n_ok) ? __half2float(p.bias[n]) : 0.f;
          float* part = p.partial + ((size_t)tix * p.max_contrib + my_contrib) * (TN * kW4TileM);
#pragma unroll 1
          for (int c = half * 16; c < TN; c += 32) {
            uint32_t d[16];
            tmem_ld_32x32b_x16(tmem_d + lane_base + c, d);
            tmem_ld_wait();
            if (n_contrib == 1) {
              if (n_ok) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int t = t0 + c + j;
                  if (t < p.T) p.y[(size_t)t * p.N + n] = __float2half_rn(__uint_as_float(d[j]) + bv);
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) part[(c + j) * kW4TileM + m] = __uint_as_float(d[j]);
            }
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty);
          if (group == 0) W4_TRACE(6 + 4 * (seg & 7));
          if (n_contrib > 1) {
            // last-arriving contributor sums the slots in contributor order (deterministic).  Release: the group
            // barrier orders every thread's partial stores before thread 0's gpu-scope fence + counter increment;
            // acquire: thread 0's fence after observing the count, then the barrier, then .cg loads.
            asm volatile("bar.sync %0, 256;" ::"r"(2 + group) : "memory");
            if (gtid == 0) {
              __threadfence();
              const int prev = atomicAdd(&p.counters[tix], 1);
              s_is_last[group] = prev == n_contrib - 1;
              if (s_is_last[group]) {
                p.counters[tix] = 0;  // re-armed for the next launch (graph replay safe)
                __threadfence();
              }
            }
            asm volatile("bar.sync %0, 256;" ::"r"(2 + group) : "memory");
            if (s_is_last[group]) {
              const float* base = p.partial + (size_t)tix * p.max_contrib * (TN * kW4TileM);
              const int n_vec = min(p.T - t0, TN) * (kW4TileM / 4);
              // float4 per thread; every contributor's load of two elements is in flight before the first add
              for (int idx0 = gtid; idx0 < n_vec; idx0 += 512) {
                float4 acc[2];
                float4 ld[2][4];
#pragma unroll
                for (int e = 0; e < 2; ++e) acc[e] = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int c0 = 0; c0 < n_contrib; c0 += 4) {
#pragma unroll
                  for (int e = 0; e < 2; ++e) {
                    const int idx = idx0 + e * 256;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                      ld[e][cc] = (idx < n_vec && c0 + cc < n_contrib)
                                      ? __ldcg(reinterpret_cast<const float4*>(&base[(size_t)(c0 + cc) * (TN * kW4TileM) + idx * 4]))
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
                  const int idx = idx0 + e * 256;
                  const int tt = idx / (kW4TileM / 4), mm = (idx % (kW4TileM / 4)) * 4;
                  const int nn = tile * kW4TileM + mm;
                  if (idx < n_vec && nn < p.N) {  // N % 32 == 0: a float4 never straddles N
                    if (p.bias) {
                      acc[e].x += __half2float(p.bias[nn]); acc[e].y += __half2float(p.bias[nn + 1]);
                      acc[e].z += __half2float(p.bias[nn + 2]); acc[e].w += __half2float(p.bias[nn + 3]);
                    }
                    uint2 o;
                    o.x = pack_half2(acc[e].x, acc[e].y);
                    o.y = pack_half2(acc[e].z, acc[e].w);
                    *reinterpret_cast<uint2*>(&p.y[(size_t)(t0 + tt) * p.N + nn]) = o;
                  }
                }
              }
            }
            asm volatile("bar.sync %0, 256;" ::"r"(2 + group) : "memory");  // s_is_last is reused by the next segment
          }
          if (group == 0) W4_TRACE(7 + 4 * (seg & 7));
        }
      } else if (seg_last) {
        // the other group owns this segment's epilogue; still observe the phase (see the mbarrier rule above)
        mbar_wait(tmem_full, seg & 1);
      }
      if (seg_last) ++seg;
      if (++kb == p.nkb) { kb = 0; ++tile; }
    }
  }

  W4_TRACE(63);
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

static unsigned long long* g_w4_trace = nullptr;
static int g_w4_debug_flags = 0;

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

struct W4Plan {
  int TN, nkb, n_tiles_n, n_tiles_t, units_per_cta, n_ctas, max_contrib;
};

static W4Plan plan_w4(int64_t T, int64_t N, int64_t K, int sms) {
  W4Plan pl;
  pl.TN = T <= 16 ? 16 : T <= 32 ? 32 : T <= 64 ? 64 : 128;
  pl.nkb = (int)((K + kW4BlockK - 1) / kW4BlockK);
  pl.n_tiles_n = (int)((N + kW4TileM - 1) / kW4TileM);
  pl.n_tiles_t = (int)((T + pl.TN - 1) / pl.TN);
  const int total = pl.n_tiles_n * pl.nkb;
  if (pl.n_tiles_t > 1) {
    pl.units_per_cta = pl.nkb;  // whole tiles (prefill: plenty of tiles)
  } else {
    const int ctas = total < sms ? total : sms;
    pl.units_per_cta = (total + ctas - 1) / ctas;
    if (pl.units_per_cta < 2 && pl.nkb >= 2) pl.units_per_cta = 2;
  }
  pl.n_ctas = (total + pl.units_per_cta - 1) / pl.units_per_cta;
  pl.max_contrib = (pl.nkb + pl.units_per_cta - 1) / pl.units_per_cta + 1;
  if (pl.units_per_cta % pl.nkb == 0) pl.max_contrib = 1;
  return pl;
}

}  // namespace b200

using namespace b200;

// debug: device buffer of [n_ctas][64] uint64 receiving per-CTA phase timestamps of the next int4 GEMM launches
extern "C" void b200_debug_w4_trace(void* device_buffer) { g_w4_trace = (unsigned long long*)device_buffer; }
extern "C" void b200_debug_w4_flags(int flags) { g_w4_debug_flags = flags; }

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

// bytes of split-K partials the int4 GEMM may write for this shape (the tile counters sit in the first 64 KiB)
int64_t b200_w4_partial_bytes(int64_t T, int64_t N, int64_t K) {
  const W4Plan pl = plan_w4(T, N, K, 148);
  if (pl.max_contrib <= 1) return 0;
  return (int64_t)pl.n_tiles_t * pl.n_tiles_n * pl.max_contrib * pl.TN * kW4TileM * 4;
}

template <int TN>
static int launch_gemm_w4(const CUtensorMap* mq, const CUtensorMap* mx, const CUtensorMap* ms, const CUtensorMap* mz, void* y,
                          void* workspace, const void* bias, int T, int N, const W4Plan& pl, int groupsize, cudaStream_t st) {
  using C = GemmW4Cfg<TN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_w4a16_kernel<TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) { b200_set_last_error(cudaGetErrorString(e)); return B200_ERR_CUDA; }
    configured = true;
  }
  W4Params p;
  p.y = (__half*)y;
  p.counters = (int*)workspace;
  p.partial = workspace ? (float*)((char*)workspace + kW4CounterBytes) : nullptr;
  p.bias = (const __half*)bias;
  p.T = T;
  p.N = N;
  p.nkb = pl.nkb;
  p.n_tiles_n = pl.n_tiles_n;
  p.units_per_cta = pl.units_per_cta;
  p.total_units = pl.n_tiles_n * pl.nkb;
  p.max_contrib = pl.max_contrib;
  p.groupsize = groupsize;
  p.group_rows = groupsize >= kW4BlockK ? 1 : kW4BlockK / groupsize;
  p.trace = g_w4_trace;
  p.debug_flags = g_w4_debug_flags;
  dim3 grid(pl.n_ctas, pl.n_tiles_t, 1);
  b200_timing_mark(B200_TIME_GEMM_W4A16, 0, st);
  B200_LAUNCH(gemm_w4a16_kernel<TN>, grid, dim3(kW4Threads), (size_t)C::kSmemBytes, st, *mq, *mx, *ms, *mz, p);
  b200_timing_mark(B200_TIME_GEMM_W4A16, 1, st);
  b200_count_launches(1);
  return B200_OK;
}

// qweight must have been passed through b200_gptq_repack once.  groupsize: multiple of 32, or <= 0 for one group.
// workspace as for b200_gemm_f16 (b200_gemm_workspace_bytes); without it every CTA takes whole tiles (no stream-K).
extern "C" int b200_gemm_w4a16(const void* x, const void* qweight_repacked, const void* qzeros, const void* scales,
                               const void* bias, void* y, int64_t T, int64_t N, int64_t K, int groupsize, void* workspace,
                               void* stream) {
  if (T == 0 || N == 0) return B200_OK;
  if (K % 64 != 0 || N % 32 != 0 || (groupsize > 0 && groupsize % 32 != 0)) {  // exllamav2.py:118-119 asserts the same
    b200_set_last_error("gemm_w4a16: need K % 64 == 0, N % 32 == 0, groupsize % 32 == 0");
    return B200_ERR_ARG;
  }
  if (groupsize <= 0) groupsize = (int)((K + kW4BlockK - 1) / kW4BlockK * kW4BlockK);  // one group: always row 0
  if (groupsize > kW4BlockK && groupsize % kW4BlockK != 0) {
    b200_set_last_error("gemm_w4a16: groupsize above 128 must be a multiple of 128");
    return B200_ERR_UNSUPPORTED;
  }
  W4Plan pl = plan_w4(T, N, K, num_sms());
  if (pl.max_contrib > 1 && (!workspace || (int64_t)pl.n_tiles_n * pl.n_tiles_t * 4 > kW4CounterBytes)) {
    pl.units_per_cta = pl.nkb;  // whole tiles per CTA
    pl.n_ctas = pl.n_tiles_n;
    pl.max_contrib = 1;
  }
  const int64_t G = (K + groupsize - 1) / groupsize;
  const int grows = groupsize >= kW4BlockK ? 1 : kW4BlockK / groupsize;
  const CUtensorMap* mq = get_tmap_2d(qweight_repacked, K / 8, N, N, kW4BlockK / 8, kW4TileM, TmapDtype::kI32, TmapSwizzle::kNone);
  const CUtensorMap* mx = get_tmap_2d(x, T, K, K, pl.TN, 64, TmapDtype::kF16, TmapSwizzle::k128B);
  const CUtensorMap* ms = get_tmap_2d(scales, G, N, N, grows, kW4TileM, TmapDtype::kF16, TmapSwizzle::kNone);
  const CUtensorMap* mz = get_tmap_2d(qzeros, G, N / 8, N / 8, grows, kW4TileM / 8, TmapDtype::kI32, TmapSwizzle::kNone);
  if (!mq || !mx || !ms || !mz) return B200_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  switch (pl.TN) {
    case 16: return launch_gemm_w4<16>(mq, mx, ms, mz, y, workspace, bias, (int)T, (int)N, pl, groupsize, st);
    case 32: return launch_gemm_w4<32>(mq, mx, ms, mz, y, workspace, bias, (int)T, (int)N, pl, groupsize, st);
    case 64: return launch_gemm_w4<64>(mq, mx, ms, mz, y, workspace, bias, (int)T, (int)N, pl, groupsize, st);
    default: return launch_gemm_w4<128>(mq, mx, ms, mz, y, workspace, bias, (int)T, (int)N, pl, groupsize, st);
  }
}
