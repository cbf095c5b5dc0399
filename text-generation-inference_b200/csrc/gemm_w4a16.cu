// GPTQ int4 linear  Y[T,N] = X[T,K] . dequant(Wq)[K,N]  on tcgen05 tensor cores, weights de-quantised in-kernel
// straight into TMEM (no fp16 scratch matrix, for any T).
//
// Replaces `exllamav2_kernels.make_q_matrix` / `gemm_half_q_half`
// (/root/reference/server/text_generation_server/utils/gptq/exllamav2.py:14-62,139-144), including its M > 50
// "dequantise everything to temp_dq, then cuBLAS" branch (:87), and follows the dequant formula of record
// utils/gptq/quant_linear.py:184-192:  W[k,n] = fp16( scales[g,n] * (q[k,n] - (qzeros[g,n] + 1)) ).
//
// Weight format.  `b200_gptq_pack` (the make_q_matrix analogue, once per linear at load time) turns the checkpoint
// tensors into a stream of self-contained UNIT RECORDS, one per (128-feature tile, 128-wide k-block):
//     words  [4 chunks][128 features][4 x u32]   8 KB   chunk c = k 32c..32c+31 of the block; inside a word the 8 nibbles
//                                                        are ordered k0 k2 k4 k6 | k1 k3 k5 k7 so one LOP3 yields a (k, k+1) pair
//     meta   [group_rows][128 features] u32      512 B per row: fp16 scale | (zero + 1) << 16, group_rows = max(1, 128 / groupsize)
// ordered (super-tile = pair of feature tiles, k-block, tile of the pair).  A CTA's work is a contiguous range of
// SUPER-UNITS (super-tile, k-block) == a contiguous byte range of HBM: one cp.async.bulk (TMA) per unit, no tensor map, and
// the shared-memory image is exactly what the dequant threads want (LDS.128, conflict free).  The two units of a
// super-unit share one activation tile (at T = 64 the activations are twice the weight bytes on the L2 -> SM path).
//
// Swap-AB like gemm_f16.cu: 128 output features = UMMA M, tokens = UMMA N (TN = 16..128), fp32 accumulators in TMEM.
// 24 warps, one CTA per SM.  The SM sub-partition arbiter favours high warp ids, so the pacing roles sit on top:
//   warp 22       TMA: unit records -> 12..16-deep (8 at TN = 128) weight ring; starts before the CTA set-up barrier.
//                 Only these bytes come from HBM.
//   warp 23       TMA: one activation tile [TN x 128] fp16 (two 128B-swizzled sub-tiles) per super-unit -> 6-deep ring
//                 (L2 hits); also watches the tiles land and counts them into s_ready
//   warps 4..19   dequant: 4 teams x 4 warps; team = unit index mod 4, warp = TMEM lane quarter, thread = one feature row x
//                 all 128 k of the unit (16 words).  Per 8 weights: 4 LOP3 + SHF (magic-number fp16: 1024 + q / 64 + q),
//                 4 HADD2 (exact q - zero), 4 HMUL2 (x scale: bit-identical to the formula of record) -> tcgen05.st into a
//                 ring of A-operand tiles in TMEM (64 columns each), counted into s_ready.  Teams run on their own clocks.
//   warps 20, 21  MMA issue, alternate super-units: poll s_ready (one shared-memory word per super-unit), wait for the
//                 turn (s_issued), 16 x tcgen05.mma.kind::f16 (A from TMEM, B from shared memory), ONE tcgen05.commit that
//                 frees the activation stage and both A stages
//   warps 0..3    epilogue: tcgen05.ld of a finished super-tile, then fp16 store / fused SiLU(gate) * up store / fp32
//                 partial store when the super-tile is shared with other CTAs; they sleep while they wait
// Decode (T <= 128) is weight-streaming / HBM-bound: the (super-tile, k-block) space is cut into contiguous ranges, one
// per CTA, either equal shares of a super-tile or the balanced stream-K cut, whichever plan_w4's cost model prefers.
// Super-tiles shared by several CTAs are finished after the main loop by ALL their contributors: each sums one slice over
// the contributors' fp32 partials in contributor order (deterministic, unlike the reference kernel's fp16 atomicAdd across
// K slices; a reduce-scatter, so the tail costs one partial's worth of L2 reads per CTA however many CTAs share it).
// The contributors wait for one another: all CTAs of the grid are co-resident (grid <= #SMs, one CTA per SM).
//
// mbarrier rule used throughout: a thread that waits on a barrier observes every phase of it in order, or an earlier
// observation implies the skipped phase completed (see the ring-depth static_asserts).  Polls whose result is wanted
// later use test_wait: try_wait may suspend the thread until a time-out while the phase is pending.
#include "common.cuh"
#include "tmap.cuh"
#include "../../include/b200_tgis.h"

#include <cstdlib>
#include <string>

namespace b200 {

constexpr int kW4TileM = 128;
constexpr int kW4BlockK = 128;
constexpr int kW4WordBytes = (kW4BlockK / 8) * kW4TileM * 4;  // 8192
constexpr int kW4MetaRowBytes = kW4TileM * 4;                 // 512
constexpr int kW4MaxGroupRows = 4;                            // groupsize >= 32
constexpr int kW4RecMaxBytes = kW4WordBytes + kW4MaxGroupRows * kW4MetaRowBytes;  // 10240 = shared-memory stage stride
// Warp roles.  The SM sub-partition arbiter favours the higher warp id, so the three single-thread roles that pace the
// pipeline (MMA issue, the two TMA producers) get the highest ids, the 16 issue-bound dequant warps sit in the middle and
// the mostly-waiting epilogue warps get the lowest (and sleep while they wait).
constexpr int kW4Teams = 4;
constexpr int kW4EpWarps = 4;                                   // warps 0..3
constexpr int kW4FirstDqWarp = kW4EpWarps;                      // warps 4..19
constexpr int kW4MmaWarp = kW4FirstDqWarp + 4 * kW4Teams;       // 20, 21: two MMA warps taking alternate super-units
constexpr int kW4WProdWarp = kW4MmaWarp + 2;                    // 22
constexpr int kW4XProdWarp = kW4MmaWarp + 3;                    // 23
constexpr int kW4Threads = (kW4XProdWarp + 1) * 32;             // 768
constexpr int kW4AColsPerStage = 64;  // 128 fp16 per row = 64 x 32-bit columns
constexpr int64_t kW4CounterBytes = 64 * 1024;

constexpr int kW4R = 2;  // feature tiles per super-tile: consecutive units (super-tile, k-block, r = 0..1) share one activation tile

template <int TN>
struct GemmW4Cfg {
  static constexpr int kSuRing = TN >= 128 ? 2 : 3;            // super-units whose A tiles fit in TMEM at once
  static constexpr int kAStages = kSuRing * kW4R;              // TMEM A-operand ring (per unit)
  static constexpr int kDBufs = TN >= 64 ? 1 : 2;              // accumulator sets (one set = kW4R tiles x TN columns)
  static constexpr int kDCols = kDBufs * kW4R * TN;
  static constexpr int kABase = kDCols < 64 ? 64 : kDCols;
  static constexpr int kTmemCols = 512;
  static constexpr int kXSubBytes = TN * 128;                  // one [TN x 64] fp16 sub-tile
  static constexpr int kXStageBytes = 2 * kXSubBytes;
  static constexpr int kXStages = TN >= 128 ? 4 : 6;           // activation ring (per super-unit = kW4R units)
  static constexpr int kWStages = TN >= 128 ? 8 : (TN == 64 ? 12 : 16);
  static constexpr int kNumBars = 2 * kWStages + kXStages + kSuRing + 4;
  static constexpr int kSmemBytes = kXStages * kXStageBytes + kWStages * kW4RecMaxBytes + kNumBars * 8 + 16 + 1024;
  static_assert(kABase + kAStages * kW4AColsPerStage <= kTmemCols, "TMEM budget");
  // a team observes every phase of the weight stages it uses only if a stage always belongs to the same team
  static_assert(kWStages % kW4Teams == 0, "weight ring depth must be a multiple of the team count");
  // a team at super-unit j has seen super-unit j - 2 - kSuRing complete, which implies j - 2 kSuRing did (no phase aliasing)
  static_assert(kSuRing >= 2, "A ring depth");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(b), "r"(c));  // (a & b) | c
  return r;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void lds_v4(uint32_t addr, uint32_t* v) {
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
}
__device__ __forceinline__ uint32_t hadd2_u(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t hmul2_u(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t lds_acquire_cta(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_cta_shared_inc(uint32_t* p) {
  asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(smem_u32(p)) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// non-blocking poll (mbarrier.try_wait may suspend the thread until a time-out when the phase is still pending)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// wait of a warp that has nothing else to do for a long time: sleep between polls so it does not take issue slots
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(128);
    if (++spins > (1u << 23)) { __trap(); }
  }
}

// one packed word (8 weights along k) -> 4 half2 (k, k+1) pairs, each fp16(scale * (q - zero)):
//   low nibbles:  0x6400 | q      = 1024 + q   (ulp 1);     + (-(1024 + zero)) is exact
//   high nibbles: 0x5400 | q << 4 = 64 + q     (ulp 1/16);  + (-(64 + zero))   is exact
// then one rounding in the multiplication by the scale, as in fp16(scale * (q - zero)).
__device__ __forceinline__ void dequant_word(uint32_t w, uint32_t nz1024, uint32_t nz64, uint32_t scale2, uint32_t* out) {
  const uint32_t w8 = w >> 8;
  const uint32_t q0 = lop3_and_or(w, 0x000f000fu, 0x64006400u);   // (k0, k1)
  const uint32_t q1 = lop3_and_or(w, 0x00f000f0u, 0x54005400u);   // (k2, k3)
  const uint32_t q2 = lop3_and_or(w8, 0x000f000fu, 0x64006400u);  // (k4, k5)
  const uint32_t q3 = lop3_and_or(w8, 0x00f000f0u, 0x54005400u);  // (k6, k7)
  out[0] = hmul2_u(hadd2_u(q0, nz1024), scale2);
  out[1] = hmul2_u(hadd2_u(q1, nz64), scale2);
  out[2] = hmul2_u(hadd2_u(q2, nz1024), scale2);
  out[3] = hmul2_u(hadd2_u(q3, nz64), scale2);
}
// meta word (fp16 scale | (zero + 1) << 16) -> the three half2 constants of dequant_word
__device__ __forceinline__ void dequant_consts(uint32_t rec, uint32_t& scale2, uint32_t& nz1024, uint32_t& nz64) {
  scale2 = prmt(rec, rec, 0x1010);             // fp16 scale in both halves
  const uint32_t z2 = prmt(rec, rec, 0x3232);  // zero + 1 (1..16) in both halves
  nz1024 = 0xE400E400u + z2;                   // -(1024 + zero), exact
  nz64 = 0xD400D400u + (z2 << 4);              // -(64 + zero), exact
}

// fp16(fp16(silu(fp16 g)) * fp16 u) from the two fp32 accumulators: exactly what the separate linear (fp16 output) followed
// by silu_mul_kernel (elementwise.cu; torch fp16 SiLU = fp32 x / (1 + exp(-x)), one rounding, then an fp16 multiply) gives
__device__ __forceinline__ __half silu_mul_f16(float gate_acc, float up_acc) {
  const float g = __half2float(__float2half_rn(gate_acc));
  const __half a = __float2half_rn(g / (1.f + expf(-g)));
  return __hmul(a, __float2half_rn(up_acc));
}

struct W4Params {
  __half* y;
  float* partial;   // [token tile][super-tile][contributor][kW4R][TN][128] fp32
  int* counters;    // [token tile][super-tile][2]: partials written / slices reduced
  const __half* bias;
  const unsigned char* packed;  // unit records, (super-tile, k-block, r) order
  int T, N;
  int nkb;          // 128-wide k-blocks per tile
  int n_super;      // super-tiles (pairs of 128-feature tiles)
  int su_per_cta;   // contiguous (super-tile, k-block) super-units per CTA
  int total_su;     // per token tile
  int max_contrib;  // partial slots per super-tile
  int rec_bytes;    // 8192 + group_rows * 512
  int half_tiles;   // 0: super-tile s = feature tiles (2s, 2s+1); > 0 ("gate|up" layout): tiles (s, s + half_tiles)
  int act;          // 1: y[t, 128 s + m] = fp16(silu(fp16 gate)) * fp16 up  (needs the gate|up layout), row stride ldy
  int ldy;          // output row stride in halves
  int persist;      // 1 (several token tiles, prefill): grid.x persistent CTAs take whole (token tile, super-tile) items c, c + grid, ...
  int n_items;      //    n_token_tiles * n_super items, token tile major (concurrent CTAs share an activation tile in L2)
  int defer;        // 1: every super-tile segment is left as an fp32 partial in the workspace and the NEXT kernel of the stream sums
                    //    them (B200SplitK consumers: rmsnorm / rope / SiLU*up / all-reduce); no counters, no waiting on peer CTAs
  unsigned long long* trace;  // debug: per-CTA phase timestamps (globaltimer ns), NULL in production
};

// kGR = meta rows per unit record (1, 2 or 4 = 128 / groupsize, 1 for groupsize >= 128).
template <int TN, int kGR>
__global__ void __launch_bounds__(kW4Threads, 1)
gemm_w4a16_kernel(const __grid_constant__ CUtensorMap tmap_x, const W4Params p) {
  using C = GemmW4Cfg<TN>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* x_ring = smem;                                 // 1024-aligned stages (128B swizzle)
  unsigned char* w_ring = smem + C::kXStages * C::kXStageBytes; // 1024-aligned stages
  uint64_t* full_w = reinterpret_cast<uint64_t*>(w_ring + C::kWStages * kW4RecMaxBytes);
  uint64_t* empty_w = full_w + C::kWStages;
  uint64_t* full_x = empty_w + C::kWStages;
  uint64_t* su_done = full_x + C::kXStages;     // [kSuRing] MMA commit of super-unit j -> barrier j % kSuRing: its x stage and its A stages are free
  uint64_t* tmem_full = su_done + C::kSuRing;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  // Inputs of super-unit j (its activation tile + the 8 team-warp quarters of its two A tiles) are counted in
  // s_ready[j % 8]; the MMA warp polls that word with plain shared-memory loads.  (In the warp that issues tcgen05.mma
  // every dependent wait costs > 100 cycles during which the tensor pipe drains - tools/ubench/umma_rate.cu - so it gets
  // exactly one per 16 MMAs.)  Monotonic: super-unit j is ready at 9 * (j / 8 + 1).
  __shared__ uint32_t s_ready[8];
  __shared__ uint32_t s_issued;  // super-units whose MMAs have been issued (hand-over between the two MMA warps)
  __shared__ int s_fix[2][4];  // super-tiles this CTA shares with others: {counter index, super-tile, contributors, my index}
  __shared__ int s_nfix;

  pdl_launch_dependents();
  const int warp = warp_id(), lane = lane_id();
  // Two ways to cut the work.  Decode (one token tile): a contiguous range of super-units of the (super-tile, k-block) space.
  // Prefill (persist): whole items = all k-blocks of one super-tile for one token tile; CTA c takes items c, c + grid, ... so that
  // launch, set-up and pipeline ramp are paid once per SM instead of once per item (they cost as much as ~30 k-blocks of main loop:
  // the gate_up projection ran at 43 % of the tensor peak against 67 % for the 3.5 x deeper down projection, ncu r2).
  const bool persist = p.persist != 0;
  const int n_items_cta = persist ? (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int t0 = persist ? 0 : blockIdx.y * TN;
  const int su0 = persist ? 0 : blockIdx.x * p.su_per_cta;
  const int su1 = persist ? n_items_cta * p.nkb : min(p.total_su, su0 + p.su_per_cta);
  const int n_su = su1 - su0;
  const int n_units = n_su * kW4R;
  const int kb0 = persist ? 0 : su0 % p.nkb;
  auto item_of_seg = [&](int seg) { return (int)blockIdx.x + seg * (int)gridDim.x; };
  auto t0_of_seg = [&](int seg) { return persist ? (item_of_seg(seg) / p.n_super) * TN : t0; };
  auto unit_src = [&](int i) -> const unsigned char* {  // HBM address of this CTA's i-th unit record
    if (!persist) return p.packed + ((size_t)su0 * kW4R + i) * p.rec_bytes;
    const int per_item = p.nkb * kW4R;
    const int seg = i / per_item;
    return p.packed + ((size_t)(item_of_seg(seg) % p.n_super) * per_item + (i - seg * per_item)) * p.rec_bytes;
  };
#define W4_TRACE(id, who)                                                                              \
  do {                                                                                                 \
    if (p.trace && threadIdx.x == (who)) {                                                             \
      unsigned long long _t;                                                                           \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                                           \
      p.trace[blockIdx.x * 64 + (id)] = _t;                                                            \
    }                                                                                                  \
  } while (0)
  W4_TRACE(0, 0);

  // The weight stream does not depend on anything: the W producer initialises its own barriers and issues the first ring
  // of unit records before the CTA-wide set-up barrier (TMEM allocation, the other barriers), so HBM latency overlaps it.
  int w_issued = 0;
  if (warp == kW4WProdWarp) {
    if (elect_one()) {
      for (int s = 0; s < C::kWStages; ++s) {
        mbar_init(&full_w[s], 1);
        mbar_init(&empty_w[s], 4);  // the 4 warps of the team that owns the stage
      }
      mbar_fence_init();
      const uint64_t pol_w = persist ? policy_evict_last() : policy_evict_first();  // prefill re-reads the weights once per token tile
      const int n0 = min(n_units, C::kWStages);
      for (int i = 0; i < n0; ++i) {
        mbar_arrive_expect_tx(&full_w[i], p.rec_bytes);
        tma_bulk_g2s_hint(w_ring + i * kW4RecMaxBytes, unit_src(i), p.rec_bytes, &full_w[i], pol_w);
      }
    }
    w_issued = min(n_units, C::kWStages);
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < C::kXStages; ++s) mbar_init(&full_x[s], 1);
    for (int b = 0; b < C::kSuRing; ++b) mbar_init(&su_done[b], 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], kW4EpWarps);
    }
    for (int i = 0; i < 8; ++i) s_ready[i] = 0;
    s_issued = 0;
    s_nfix = 0;
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<C::kTmemCols>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d = tmem_base;               // accumulator sets
  const uint32_t tmem_a = tmem_base + C::kABase;   // A-operand ring
  W4_TRACE(1, 0);

  if (warp == kW4WProdWarp) {
    // ------------------------------------------------------------------ TMA producer: unit records (HBM stream); the first
    // ring was issued above
    if (elect_one()) {
      const uint64_t pol_w = persist ? policy_evict_last() : policy_evict_first();
      int s = 0, ph = 0;  // stage 0 again: wait for its first release
      for (int i = w_issued; i < n_units; ++i) {
        mbar_wait(&empty_w[s], ph);
        mbar_arrive_expect_tx(&full_w[s], p.rec_bytes);
        tma_bulk_g2s_hint(w_ring + s * kW4RecMaxBytes, unit_src(i), p.rec_bytes, &full_w[s], pol_w);
        if (++s == C::kWStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == kW4XProdWarp) {
    // ------------------------------------------------------------------ TMA producer: activations (L2 resident),
    // one [TN x 128] tile per super-unit; the same thread watches the tiles land and counts them into s_ready
    if (elect_one()) {
      const uint64_t pol_x = policy_evict_last();
      int kb = kb0, seg_x = 0, t0x = t0_of_seg(0);
      auto load_x = [&](int s) {
        mbar_arrive_expect_tx(&full_x[s], C::kXStageBytes);
        unsigned char* st = x_ring + s * C::kXStageBytes;
        tma_load_2d_hint(st, &tmap_x, kb * kW4BlockK, t0x, &full_x[s], pol_x);
        tma_load_2d_hint(st + C::kXSubBytes, &tmap_x, kb * kW4BlockK + 64, t0x, &full_x[s], pol_x);
        if (++kb == p.nkb) {
          kb = 0;
          t0x = t0_of_seg(++seg_x);  // persist: the next item may belong to another token tile
        }
      };
      pdl_wait();  // x is the previous kernel's output; the weight ring is already streaming
      for (int j = 0; j < C::kXStages && j < n_su; ++j) load_x(j);
      // two independent duties, both polled without blocking: count landed tiles into s_ready (in order), and refill the
      // stage of super-unit j with the tile of super-unit j + kXStages once the MMAs of j have completed
      int nj = 0, ns = 0, nph = 0;  // next tile to report: index, stage, parity
      int rj = 0, rs = 0, rb = 0, rph = 0;  // next super-unit whose stage is refilled: index, stage, su_done barrier, parity
      const int n_refill = n_su - C::kXStages;
      uint32_t idle = 0;
      while (nj < n_su) {
        bool progress = false;
        if (mbar_test_wait(&full_x[ns], nph)) {
          red_release_cta_shared_inc(&s_ready[nj & 7]);
          ++nj;
          if (++ns == C::kXStages) { ns = 0; nph ^= 1; }
          progress = true;
        }
        if (rj < n_refill && rj < nj && mbar_test_wait(&su_done[rb], rph)) {
          load_x(rs);
          ++rj;
          if (++rs == C::kXStages) rs = 0;
          if (++rb == C::kSuRing) { rb = 0; rph ^= 1; }
          progress = true;
        }
        if (!progress) {
          __nanosleep(20);
          if (++idle > (1u << 26)) __trap();
        }
      }
    }
  } else if (warp == kW4MmaWarp || warp == kW4MmaWarp + 1) {
    // ------------------------------------------------------------------ MMA issuers.  tcgen05.mma issue is in order and the
    // queue is shallow, so a warp that has just issued 16 MMAs cannot hide its own bookkeeping (barrier polls, commit)
    // behind them: the tensor pipe would idle ~40 % of the time (tools/ubench/umma_rate.cu).  Two warps take alternate
    // super-units; while one is blocked issuing, the other has already found its inputs ready and only waits for its turn
    // (s_issued), so MMAs reach the pipe back to back and in super-unit order (summation order stays fixed).
    // Warp-uniform loops, one elected thread issues: keeps the tcgen05 operands in uniform registers.
    const int me = warp - kW4MmaWarp;
    constexpr uint32_t idesc = umma_idesc_f16_f32acc(kW4TileM, TN);
    const uint64_t bdesc0 = umma_desc_kmajor_sw128(smem_u32(x_ring));
    for (int j = me; j < n_su; j += 2) {
      const int kbj = kb0 + j;            // k-block index counted from the CTA's first super-tile
      const int kb = kbj % p.nkb, seg = kbj / p.nkb;
      const bool seg_first = (j == 0) || kb == 0;
      const bool seg_last = (j == n_su - 1) || kb == p.nkb - 1;
      const int buf = C::kDBufs == 2 ? (seg & 1) : 0;
      const int sx = j % C::kXStages, sj = j % C::kSuRing;
      {
        const uint32_t target = 9u * (uint32_t)((j >> 3) + 1);
        uint32_t spins = 0;
        while (lds_acquire_cta(&s_ready[j & 7]) < target) {
          if (++spins > (1u << 26)) __trap();
        }
      }
      if (seg_first) {  // the epilogue drained this accumulator set's previous super-tile
        if (C::kDBufs == 2) mbar_wait(&tmem_empty[buf], ((seg >> 1) & 1) ^ 1);
        else mbar_wait(&tmem_empty[0], (seg & 1) ^ 1);
      }
      {
        uint32_t spins = 0;
        while (lds_acquire_cta(&s_issued) < (uint32_t)j) {  // my turn
          if (++spins > (1u << 26)) __trap();
        }
      }
      tcgen05_fence_after();
      if (elect_one()) {
        const uint64_t bdesc = bdesc0 + (uint64_t)((sx * C::kXStageBytes) >> 4);
#pragma unroll
        for (int r = 0; r < kW4R; ++r) {
          const uint32_t d = tmem_d + (buf * kW4R + r) * TN, a = tmem_a + (sj * kW4R + r) * kW4AColsPerStage;
#pragma unroll
          for (int k = 0; k < kW4BlockK / 16; ++k)
            umma_f16_ts(d, a + k * 8, bdesc + (uint64_t)((k >> 2) * (C::kXSubBytes >> 4) + (k & 3) * 2), idesc, (!seg_first || k > 0) ? 1u : 0u);
        }
        asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(&s_issued)), "r"((uint32_t)j + 1) : "memory");
        umma_commit(&su_done[sj]);
        if (seg_last) umma_commit(&tmem_full[buf]);
      }
      __syncwarp();
    }
    W4_TRACE(6, kW4MmaWarp * 32);
  } else if (warp >= kW4FirstDqWarp) {
    // ------------------------------------------------------------------ dequant teams
    const int team = (warp - kW4FirstDqWarp) >> 2;  // unit index mod 4
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    const int m = quarter * 32 + lane;              // feature row within the tile
    const uint32_t a_row = tmem_a + ((uint32_t)(quarter * 32) << 16);
    const uint32_t w_thread = smem_u32(w_ring) + m * 16;
    const uint32_t meta_thread = smem_u32(w_ring) + kW4WordBytes + m * 4;
    uint32_t wq[16], mrec[kGR];
    int ws = team, wph = 0;   // weight-ring stage / parity of the unit being loaded
    // unit i (super-unit j = i / 2) lives in A stage i % kAStages and may be written once the MMAs of super-unit
    // j - kSuRing have completed: barrier su_done[j % kSuRing], phase j / kSuRing - 1
    auto load_unit = [&](bool landed) {
      if (!landed) mbar_wait(&full_w[ws], wph);
      const uint32_t st = ws * kW4RecMaxBytes;
#pragma unroll
      for (int c = 0; c < 4; ++c) lds_v4(w_thread + st + c * (kW4TileM * 16), &wq[c * 4]);
#pragma unroll
      for (int r = 0; r < kGR; ++r) mrec[r] = lds_u32(meta_thread + st + r * kW4MetaRowBytes);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_w[ws]);
      ws += kW4Teams;
      if (ws >= C::kWStages) { ws -= C::kWStages; wph ^= 1; }
    };
    if (team < n_units) load_unit(false);
    if (team == 0) W4_TRACE(2, kW4FirstDqWarp * 32);
    int n = 0;
    for (int i = team; i < n_units; i += kW4Teams, ++n) {
      const bool more = i + kW4Teams < n_units;
      // barrier polls are issued well before their result is needed
      const int j = i >> 1, jr = j % C::kSuRing;
      const uint32_t done_par = (uint32_t)(j / C::kSuRing - 1) & 1u;
      const uint32_t a_dst = a_row + (i % C::kAStages) * kW4AColsPerStage;
      const bool a_free = j < C::kSuRing || mbar_test_wait(&su_done[jr], done_par);  // looked at after chunk 0
      const bool w_landed = more && mbar_test_wait(&full_w[ws], wph);                // looked at after chunk 3
      uint32_t scale2, nz1024, nz64;
      dequant_consts(mrec[0], scale2, nz1024, nz64);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        // chunk c = k 32c..32c+31 of the block; its meta row is c * kGR / 4
        if (c > 0 && (c * kGR) % 4 == 0) dequant_consts(mrec[c * kGR / 4], scale2, nz1024, nz64);
        uint32_t v[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) dequant_word(wq[c * 4 + j], nz1024, nz64, scale2, &v[j * 4]);
        if (c == 0) {
          if (!a_free) mbar_wait(&su_done[jr], done_par);  // the MMAs that read this A stage have completed
          tcgen05_fence_after();
        }
        tmem_st_32x32b_x16(a_dst + c * 16, v);
      }
      if (more) load_unit(w_landed);  // next unit's words: the shared-memory reads overlap the TMEM store latency
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) red_release_cta_shared_inc(&s_ready[(i >> 1) & 7]);
    }
    if (team == 0) W4_TRACE(4, kW4FirstDqWarp * 32);
  } else {
    // ------------------------------------------------------------------ epilogue warps: one pass per super-tile segment
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    int seg = 0, nfix = 0;
    pdl_wait();  // outputs / bias / stream-K workspace belong to the stream order
    for (int u = su0; u < su1; ++seg) {
      const int sup = persist ? item_of_seg(seg) % p.n_super : u / p.nkb;
      const int seg_end = persist ? u + p.nkb : min(su1, (sup + 1) * p.nkb);
      const int t0 = t0_of_seg(seg);  // shadows the CTA-wide value: persistent CTAs change token tile between items
      const int buf = C::kDBufs == 2 ? (seg & 1) : 0;
      // contributors of this super-tile: CTAs whose range intersects [sup*nkb, (sup+1)*nkb); a persistent CTA owns whole items
      const int c_first = persist ? (int)blockIdx.x : (sup * p.nkb) / p.su_per_cta;
      const int c_last = persist ? (int)blockIdx.x : ((sup + 1) * p.nkb - 1) / p.su_per_cta;
      const int n_contrib = c_last - c_first + 1;
      const int my_contrib = (int)blockIdx.x - c_first;
      const int tix = blockIdx.y * p.n_super + sup;
      float* part = p.partial + ((size_t)tix * p.max_contrib + my_contrib) * (kW4R * TN * kW4TileM);
      mbar_wait_relaxed(&tmem_full[buf], C::kDBufs == 2 ? ((seg >> 1) & 1) : (seg & 1));
      tcgen05_fence_after();
      W4_TRACE(8 + (seg & 3) * 2, 0);
      {
        // both tiles of the super-tile, kCh token columns at a time
        const int n0 = (p.half_tiles ? sup : sup * kW4R) * kW4TileM + m;
        const int n1 = (p.half_tiles ? sup + p.half_tiles : sup * kW4R + 1) * kW4TileM + m;
        const bool ok0 = n0 < p.N, ok1 = n1 < p.N;
        const bool direct = n_contrib == 1 && !p.defer;
        const float bv0 = (p.bias && ok0 && direct) ? __half2float(p.bias[n0]) : 0.f;
        const float bv1 = (p.bias && ok1 && direct) ? __half2float(p.bias[n1]) : 0.f;
        constexpr int kCh = TN < 32 ? TN : 32;  // columns fetched per tcgen05.wait::ld
#pragma unroll 1
        for (int c = 0; c < TN; c += kCh) {
          uint32_t d0[kCh], d1[kCh];
#pragma unroll
          for (int q = 0; q < kCh / 16; ++q) {
            tmem_ld_32x32b_x16(tmem_d + lane_base + (buf * kW4R + 0) * TN + c + q * 16, reinterpret_cast<uint32_t(&)[16]>(d0[q * 16]));
            tmem_ld_32x32b_x16(tmem_d + lane_base + (buf * kW4R + 1) * TN + c + q * 16, reinterpret_cast<uint32_t(&)[16]>(d1[q * 16]));
          }
          tmem_ld_wait();
          if (!direct) {
#pragma unroll
            for (int j = 0; j < kCh; ++j) {
              if (t0 + c + j < p.T) {
                part[(c + j) * kW4TileM + m] = __uint_as_float(d0[j]);
                part[(TN + c + j) * kW4TileM + m] = __uint_as_float(d1[j]);
              }
            }
          } else if (p.act) {
            if (ok0) {
#pragma unroll
              for (int j = 0; j < kCh; ++j) {
                const int t = t0 + c + j;
                if (t < p.T) p.y[(size_t)t * p.ldy + n0] = silu_mul_f16(__uint_as_float(d0[j]) + bv0, __uint_as_float(d1[j]) + bv1);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < kCh; ++j) {
              const int t = t0 + c + j;
              if (t < p.T) {
                if (ok0) p.y[(size_t)t * p.ldy + n0] = __float2half_rn(__uint_as_float(d0[j]) + bv0);
                if (ok1) p.y[(size_t)t * p.ldy + n1] = __float2half_rn(__uint_as_float(d1[j]) + bv1);
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      W4_TRACE(9 + (seg & 3) * 2, 0);
      if (n_contrib > 1 && !p.defer) {
        // release: the epilogue barrier orders every thread's partial stores before thread 0's gpu-scope fence + count
        asm volatile("bar.sync 1, %0;" ::"n"(kW4EpWarps * 32) : "memory");
        if (threadIdx.x == 0) {
          __threadfence();
          atomicAdd(&p.counters[2 * tix], 1);
          s_fix[nfix][0] = tix;
          s_fix[nfix][1] = sup;
          s_fix[nfix][2] = n_contrib;
          s_fix[nfix][3] = my_contrib;
        }
        ++nfix;
      }
      u = seg_end;
    }
    if (threadIdx.x == 0) s_nfix = nfix;
    W4_TRACE(5, 0);
  }

  // -------------------------------------------------------------------- shared super-tiles: every contributor reduces a slice
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<C::kTmemCols>(tmem_base);
  const int nfix = s_nfix;
  if (nfix > 0) pdl_wait();
  W4_TRACE(40, 0);
  for (int f = 0; f < nfix; ++f) {
    const int tix = s_fix[f][0], sup = s_fix[f][1], n_contrib = s_fix[f][2], my_contrib = s_fix[f][3];
    if (threadIdx.x == 0) {
      uint32_t spins = 0;
      while (ld_acquire_gpu(&p.counters[2 * tix]) < n_contrib) {  // all contributors are co-resident (grid <= #SMs, 1 CTA/SM)
        __nanosleep(64);
        if (++spins > (1u << 24)) __trap();
      }
    }
    __syncthreads();
    W4_TRACE(41 + f * 3, 0);
    const float* base = p.partial + (size_t)tix * p.max_contrib * (kW4R * TN * kW4TileM);
    const int rows = min(p.T - t0, TN);
    const int n_vec = kW4R * rows * (kW4TileM / 4);  // float4 elements: [r][row][32]
    const int per = (n_vec + n_contrib - 1) / n_contrib;
    const int hi = min(n_vec, (my_contrib + 1) * per);
    constexpr size_t kPartStride = (size_t)kW4R * TN * kW4TileM;
    if (!p.act) {
      for (int idx = my_contrib * per + (int)threadIdx.x; idx < hi; idx += kW4Threads) {
        const int r = idx / (rows * (kW4TileM / 4));
        const int rem = idx - r * rows * (kW4TileM / 4);
        const int tt = rem / (kW4TileM / 4), mm = (rem % (kW4TileM / 4)) * 4;
        const size_t off = (size_t)(r * TN + tt) * kW4TileM + mm;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c0 = 0; c0 < n_contrib; c0 += 4) {  // contributor order: deterministic
          float4 ld[4];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
            ld[cc] = (c0 + cc < n_contrib) ? __ldcg(reinterpret_cast<const float4*>(&base[(size_t)(c0 + cc) * kPartStride + off]))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) { acc.x += ld[cc].x; acc.y += ld[cc].y; acc.z += ld[cc].z; acc.w += ld[cc].w; }
        }
        const int nn = (p.half_tiles ? sup + r * p.half_tiles : sup * kW4R + r) * kW4TileM + mm;
        if (nn < p.N) {  // N % 32 == 0: a float4 never straddles N
          if (p.bias) {
            acc.x += __half2float(p.bias[nn]); acc.y += __half2float(p.bias[nn + 1]);
            acc.z += __half2float(p.bias[nn + 2]); acc.w += __half2float(p.bias[nn + 3]);
          }
          uint2 o;
          o.x = pack_half2(acc.x, acc.y);
          o.y = pack_half2(acc.z, acc.w);
          *reinterpret_cast<uint2*>(&p.y[(size_t)(t0 + tt) * p.ldy + nn]) = o;
        }
      }
    } else {
      // fused SiLU(gate) * up: an element needs both tiles of the super-tile; slices over [row][32 float4]
      const int n_vec_a = rows * (kW4TileM / 4);
      const int per_a = (n_vec_a + n_contrib - 1) / n_contrib;
      const int hi_a = min(n_vec_a, (my_contrib + 1) * per_a);
      for (int idx = my_contrib * per_a + (int)threadIdx.x; idx < hi_a; idx += kW4Threads) {
        const int tt = idx / (kW4TileM / 4), mm = (idx % (kW4TileM / 4)) * 4;
        const size_t off = (size_t)tt * kW4TileM + mm;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f), u = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c0 = 0; c0 < n_contrib; c0 += 2) {  // contributor order: deterministic
          float4 lg[2], lu[2];
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const bool in = c0 + cc < n_contrib;
            lg[cc] = in ? __ldcg(reinterpret_cast<const float4*>(&base[(size_t)(c0 + cc) * kPartStride + off])) : make_float4(0.f, 0.f, 0.f, 0.f);
            lu[cc] = in ? __ldcg(reinterpret_cast<const float4*>(&base[(size_t)(c0 + cc) * kPartStride + (size_t)TN * kW4TileM + off]))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            g.x += lg[cc].x; g.y += lg[cc].y; g.z += lg[cc].z; g.w += lg[cc].w;
            u.x += lu[cc].x; u.y += lu[cc].y; u.z += lu[cc].z; u.w += lu[cc].w;
          }
        }
        const int nn = sup * kW4TileM + mm;  // column of the [T, N / 2] output
        const int ng = nn, nu = nn + p.half_tiles * kW4TileM;
        if (p.bias) {
          g.x += __half2float(p.bias[ng]); g.y += __half2float(p.bias[ng + 1]); g.z += __half2float(p.bias[ng + 2]); g.w += __half2float(p.bias[ng + 3]);
          u.x += __half2float(p.bias[nu]); u.y += __half2float(p.bias[nu + 1]); u.z += __half2float(p.bias[nu + 2]); u.w += __half2float(p.bias[nu + 3]);
        }
        __half o[4] = {silu_mul_f16(g.x, u.x), silu_mul_f16(g.y, u.y), silu_mul_f16(g.z, u.z), silu_mul_f16(g.w, u.w)};
        *reinterpret_cast<uint2*>(&p.y[(size_t)(t0 + tt) * p.ldy + nn]) = *reinterpret_cast<uint2*>(o);
      }
    }
    W4_TRACE(42 + f * 3, 0);
    __syncthreads();
    W4_TRACE(43 + f * 3, 0);
    if (threadIdx.x == 0) {
      // the last contributor to finish its slice re-arms both counters for the next launch (graph replay safe)
      const int prev = atomicAdd(&p.counters[2 * tix + 1], 1);
      if (prev == n_contrib - 1) {
        p.counters[2 * tix] = 0;
        p.counters[2 * tix + 1] = 0;
        __threadfence();
      }
    }
  }
  W4_TRACE(63, 0);
}

// checkpoint tensors -> unit records.  One thread per output u32.
__global__ void gptq_pack_kernel(const uint32_t* __restrict__ qweight, const uint32_t* __restrict__ qzeros,
                                 const __half* __restrict__ scales, uint32_t* __restrict__ packed, int64_t K, int64_t N, int groupsize,
                                 int group_rows, int nkb, int64_t n_words_total, int half_tiles, const int32_t* __restrict__ row_perm) {
  const int rec_words = (kW4WordBytes + group_rows * kW4MetaRowBytes) / 4;
  const int64_t G = (K + groupsize - 1) / groupsize;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n_words_total; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t rec = o / rec_words;          // record index in (super-tile, k-block, r) order
    const int r = (int)(o - rec * rec_words);
    const int64_t sup = rec / ((int64_t)nkb * kW4R);
    const int kb = (int)((rec / kW4R) % nkb);
    const int64_t tile = half_tiles ? sup + (rec % kW4R) * half_tiles : sup * kW4R + rec % kW4R;
    uint32_t out = 0;
    if (r < kW4WordBytes / 4) {
      const int c = r / (kW4TileM * 4), m = (r / 4) % kW4TileM, j = r % 4;
      const int64_t n = tile * kW4TileM + m;
      const int64_t kw = (int64_t)kb * (kW4BlockK / 8) + c * 4 + j;  // checkpoint word row: k = 8 kw .. 8 kw + 7
      if (n < N && kw * 8 < K) {
        if (!row_perm) {
          const uint32_t w = qweight[kw * N + n];
#pragma unroll
          for (int k = 0; k < 8; ++k) out |= ((w >> (4 * k)) & 15u) << (4 * ((k & 1) * 4 + (k >> 1)));
        } else {
          // act-order: packed row k' holds checkpoint row row_perm[k'] (rows sorted by group)
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int64_t src = row_perm[kw * 8 + k];
            const uint32_t q = (qweight[(src >> 3) * N + n] >> (4 * (src & 7))) & 15u;
            out |= q << (4 * ((k & 1) * 4 + (k >> 1)));
          }
        }
      }
    } else {
      const int mr = r - kW4WordBytes / 4;
      const int row = mr / kW4TileM, m = mr % kW4TileM;
      const int64_t n = tile * kW4TileM + m;
      const int64_t k0 = (int64_t)kb * kW4BlockK + (int64_t)row * (kW4BlockK / group_rows);
      out = 1u << 16;  // padding: scale 0, zero 1
      if (n < N && k0 < K) {
        const int64_t g = min(k0 / groupsize, G - 1);
        const uint32_t z = ((qzeros[g * (N / 8) + n / 8] >> (4 * (n % 8))) & 15u) + 1u;  // the +1 of quant_linear.py:185
        out = (uint32_t)__half_as_ushort(scales[g * N + n]) | (z << 16);
      }
    }
    packed[o] = out;
  }
}

static unsigned long long* g_w4_trace = nullptr;

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// k-blocks per tile.  Rounded up to a multiple of 8 (empty records; the activation TMA zero-fills past K) when that adds
// at most 5 % of traffic: it lets a super-tile be cut into 2 / 4 / 8 equal CTA shares (Llama-2 down_proj: 86 -> 88).
static int w4_nkb(int64_t K) {
  const int nkb = (int)((K + kW4BlockK - 1) / kW4BlockK);
  const int up = (nkb + 7) / 8 * 8;
  return (up - nkb) * 20 <= nkb ? up : nkb;
}

struct W4Plan {
  int TN, nkb, n_super, n_tiles_t, su_per_cta, n_ctas, max_contrib;
};

// Decode (one token tile): how many super-units per CTA.  Candidates are the balanced stream-K cut (every SM the same
// number of units, but most CTAs then straddle two super-tiles: two fix-ups at the tail) and the aligned cuts nkb / d
// (a CTA = 1/d of one super-tile: one fix-up with d contributors, or none for d = 1, at the price of idle SMs when
// n_super * d is not close to a multiple of the SM count).  Costs in microseconds fitted to tools/sweep_w4_su.py on B200
// (cold weights): 0.35 per unit in the main loop, ~1 for a direct epilogue, 4.5-6 for one split-K fix-up (partial store,
// gpu-scope fence, waiting for the slowest contributor, L2-latency-bound slice reduction), ~10 when CTAs straddle.
static W4Plan plan_w4(int64_t T, int64_t N, int64_t K, int sms, bool defer = false) {
  W4Plan pl;
  pl.TN = T <= 16 ? 16 : T <= 32 ? 32 : T <= 64 ? 64 : 128;
  pl.nkb = w4_nkb(K);
  pl.n_super = (int)((N + kW4R * kW4TileM - 1) / (kW4R * kW4TileM));
  pl.n_tiles_t = (int)((T + pl.TN - 1) / pl.TN);
  const int total = pl.n_super * pl.nkb;
  if (pl.n_tiles_t > 1) {
    pl.su_per_cta = pl.nkb;  // whole super-tiles (prefill: plenty of tiles)
  } else {
    const int ctas = total < sms ? total : sms;
    int best = (total + ctas - 1) / ctas;
    if (best < 2 && pl.nkb >= 2) best = 2;  // one unit per team at least
    // deferred reduction (the consumer kernel sums the partials): no in-kernel fix-up, a CTA that straddles two super-tiles only
    // pays a mid-loop accumulator drain (~1 us)
    auto cost = [&](int su, bool aligned, int d) {
      const float fix = defer ? (aligned ? 0.f : 1.0f) : (aligned ? (d == 1 ? 1.0f : 4.5f + 0.2f * d) : 10.0f);
      return 0.35f * kW4R * su + fix;
    };
    float best_cost = cost(best, pl.nkb % best == 0, pl.nkb / best);
    for (int d = 1; d <= 8; d *= 2) {
      if (pl.nkb % d != 0) break;
      const int su = pl.nkb / d;
      if ((int64_t)pl.n_super * d > sms || su < 2) continue;
      const float c = cost(su, true, d);
      if (c < best_cost) { best_cost = c; best = su; }
    }
    pl.su_per_cta = best;
    static int env_su = -1;  // experiments: B200_W4_SU forces the cut
    if (env_su < 0) {
      const char* e = getenv("B200_W4_SU");
      env_su = e ? atoi(e) : 0;
    }
    if (env_su > 0 && (total + env_su - 1) / env_su <= sms) pl.su_per_cta = env_su;
  }
  pl.n_ctas = (total + pl.su_per_cta - 1) / pl.su_per_cta;
  pl.max_contrib = (pl.nkb + pl.su_per_cta - 1) / pl.su_per_cta + 1;
  if (pl.su_per_cta % pl.nkb == 0) pl.max_contrib = 1;
  return pl;
}

static int w4_group_rows(int64_t K, int* groupsize) {
  if (*groupsize <= 0) *groupsize = (int)((K + kW4BlockK - 1) / kW4BlockK * kW4BlockK);  // one group
  return *groupsize >= kW4BlockK ? 1 : kW4BlockK / *groupsize;
}

}  // namespace b200

using namespace b200;

// debug: device buffer of [n_ctas][64] uint64 receiving per-CTA phase timestamps of the next int4 GEMM launches
extern "C" void b200_debug_w4_trace(void* device_buffer) { g_w4_trace = (unsigned long long*)device_buffer; }
extern "C" void b200_debug_w4_flags(int) {}

static bool w4_check_shape(int64_t N, int64_t K, int groupsize, const char* who) {
  if (K % 32 != 0 || N % 32 != 0 || K <= 0 || N <= 0) {  // exllamav2.py:118-119 asserts the same
    b200_set_last_error((std::string(who) + ": need K % 32 == 0 and N % 32 == 0").c_str());
    return false;
  }
  if (groupsize > 0 && (groupsize % 32 != 0 || (groupsize > kW4BlockK && groupsize % kW4BlockK != 0) ||
                        (groupsize < kW4BlockK && kW4BlockK % groupsize != 0))) {
    b200_set_last_error((std::string(who) + ": groupsize must be 32, 64, a multiple of 128, or <= 0 (one group)").c_str());
    return false;
  }
  return true;
}

extern "C" int64_t b200_gptq_packed_bytes(int64_t K, int64_t N, int groupsize) {
  if (!w4_check_shape(N, K, groupsize, "gptq_packed_bytes")) return B200_ERR_ARG;
  const int gr = w4_group_rows(K, &groupsize);
  const int64_t nkb = w4_nkb(K), ns = (N + kW4R * kW4TileM - 1) / (kW4R * kW4TileM);
  return ns * kW4R * nkb * (kW4WordBytes + gr * kW4MetaRowBytes);  // the last super-tile is padded with empty tiles
}

// layout 0: super-tile s = feature tiles (2s, 2s+1).  layout 1 ("gate|up", for a fused [gate; up] projection with
// N = 2 I, I % 128 == 0): super-tile s = tiles (s, s + I/128), i.e. gate feature n sits next to up feature n, which is what
// lets the GEMM apply SiLU(gate) * up in its epilogue (b200_gemm_w4a16_ex act = 1).
static int w4_half_tiles(int64_t N, int layout, const char* who) {
  if (layout == 0) return 0;
  if (layout != 1 || N % (2 * kW4TileM) != 0) {
    b200_set_last_error((std::string(who) + ": the gate|up layout needs N % 256 == 0").c_str());
    return -1;
  }
  return (int)(N / (2 * kW4TileM));
}

// qweight int32 [K/8, N], qzeros int32 [ceil(K/g), N/8], scales fp16 [ceil(K/g), N] (checkpoint layout, left untouched)
// -> packed (b200_gptq_packed_bytes bytes, 16-byte aligned).  row_perm == NULL: groups are k // groupsize (trivial g_idx).
// Act-order checkpoints (g_idx not sorted; exllamav2.py:31-48 builds q_perm for them): row_perm[k'] = the checkpoint row
// stored at packed row k', a stable arg-sort of g_idx, so that packed rows k' // groupsize share a group again; the GEMM
// is then fed x[:, row_perm] (b200_permute_columns).
extern "C" int b200_gptq_pack_ex(const void* qweight, const void* qzeros, const void* scales, const int32_t* row_perm, void* packed,
                                 int64_t K, int64_t N, int groupsize, int layout, void* stream) {
  if (!w4_check_shape(N, K, groupsize, "gptq_pack")) return B200_ERR_ARG;
  if (((uintptr_t)packed & 15) != 0) { b200_set_last_error("gptq_pack: packed buffer must be 16-byte aligned"); return B200_ERR_ARG; }
  const int half_tiles = w4_half_tiles(N, layout, "gptq_pack");
  if (half_tiles < 0) return B200_ERR_ARG;
  const int gr = w4_group_rows(K, &groupsize);
  const int nkb = w4_nkb(K);
  const int64_t n_words = b200_gptq_packed_bytes(K, N, groupsize) / 4;
  int64_t blocks = (n_words + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gptq_pack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint32_t*)qweight, (const uint32_t*)qzeros,
                                                                        (const __half*)scales, (uint32_t*)packed, K, N, groupsize, gr,
                                                                        nkb, n_words, half_tiles, row_perm);
  B200_CHECK_LAUNCH();
  b200_count_launches(1);
  return B200_OK;
}
extern "C" int b200_gptq_pack(const void* qweight, const void* qzeros, const void* scales, void* packed, int64_t K, int64_t N,
                              int groupsize, void* stream) {
  return b200_gptq_pack_ex(qweight, qzeros, scales, nullptr, packed, K, N, groupsize, 0, stream);
}

// bytes of split-K partials the int4 GEMM may write for this shape (the tile counters sit in the first 64 KiB)
// The plan a launch uses: the stream-K cut needs the workspace (partials + one counter pair per super-tile) and every CTA
// resident at once (the fix-up spins on its peers); otherwise each CTA takes whole super-tiles.
static W4Plan launch_plan_w4(int64_t T, int64_t N, int64_t K, int sms, bool has_workspace, bool defer = false) {
  W4Plan pl = plan_w4(T, N, K, sms, defer);
  if (defer) return pl;  // the caller checked: one token tile, workspace present; no counters, no co-residency requirement
  if (pl.max_contrib > 1 && (!has_workspace || (int64_t)pl.n_super * pl.n_tiles_t * 8 > kW4CounterBytes || pl.n_ctas > sms)) {
    pl.su_per_cta = pl.nkb;
    pl.n_ctas = pl.n_super;
    pl.max_contrib = 1;
  }
  return pl;
}

// tests (no GPU needed): out = {TN, k-blocks, super-tiles, token tiles, super-units per CTA, CTAs, contributor slots, R}
void b200_w4_plan_debug(int64_t T, int64_t N, int64_t K, int sms, int32_t* out) {
  const W4Plan pl = launch_plan_w4(T, N, K, sms, true);
  const int32_t v[8] = {pl.TN, pl.nkb, pl.n_super, pl.n_tiles_t, pl.su_per_cta, pl.n_ctas, pl.max_contrib, kW4R};
  for (int i = 0; i < 8; ++i) out[i] = v[i];
}

int64_t b200_w4_partial_bytes(int64_t T, int64_t N, int64_t K) {
  const W4Plan pl = plan_w4(T, N, K, num_sms());
  int64_t need = pl.max_contrib <= 1 ? 0 : (int64_t)pl.n_tiles_t * pl.n_super * pl.max_contrib * kW4R * pl.TN * kW4TileM * 4;
  if (pl.n_tiles_t == 1) {  // deferred reduction: every super-tile has at least one partial slot
    const W4Plan pd = plan_w4(T, N, K, num_sms(), true);
    const int64_t nd = (int64_t)pd.n_super * pd.max_contrib * kW4R * pd.TN * kW4TileM * 4;
    if (nd > need) need = nd;
  }
  return need;
}

template <int TN, int kGR>
static int launch_gemm_w4(const CUtensorMap* mx, const void* packed, void* y, void* workspace, const void* bias, int T, int N,
                          const W4Plan& pl, int half_tiles, int act, int defer, cudaStream_t st) {
  using C = GemmW4Cfg<TN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_w4a16_kernel<TN, kGR>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) { b200_set_last_error(cudaGetErrorString(e)); return B200_ERR_CUDA; }
    configured = true;
  }
  W4Params p;
  p.y = (__half*)y;
  p.counters = (int*)workspace;
  p.partial = workspace ? (float*)((char*)workspace + kW4CounterBytes) : nullptr;
  p.bias = (const __half*)bias;
  p.packed = (const unsigned char*)packed;
  p.T = T;
  p.N = N;
  p.nkb = pl.nkb;
  p.n_super = pl.n_super;
  p.su_per_cta = pl.su_per_cta;
  p.total_su = pl.n_super * pl.nkb;
  p.max_contrib = pl.max_contrib;
  p.rec_bytes = kW4WordBytes + kGR * kW4MetaRowBytes;
  p.half_tiles = half_tiles;
  p.act = act;
  p.ldy = act ? N / 2 : N;
  p.defer = defer;
  p.trace = g_w4_trace;
  dim3 grid(pl.n_ctas, pl.n_tiles_t, 1);
  p.persist = 0;
  p.n_items = 0;
  if (pl.n_tiles_t > 1) {  // prefill: persistent CTAs over (token tile, super-tile) items
    p.persist = 1;
    p.n_items = pl.n_super * pl.n_tiles_t;
    const int sms = num_sms();
    grid = dim3(p.n_items < sms ? p.n_items : sms, 1, 1);
  }
  b200_timing_mark(B200_TIME_GEMM_W4A16, 0, st);
  constexpr auto kernel = gemm_w4a16_kernel<TN, kGR>;
  B200_LAUNCH_AS("gemm_w4a16_kernel", kernel, grid, dim3(kW4Threads), (size_t)C::kSmemBytes, st, *mx, p);
  b200_timing_mark(B200_TIME_GEMM_W4A16, 1, st);
  b200_count_launches(1);
  return B200_OK;
}

// packed: output of b200_gptq_pack[_ex] for the same (K, N, groupsize, layout).
// act = 1 (layout 1 only): y [T, N/2] = SiLU(x Wgate) * (x Wup), the LlamaMLP activation (flash_llama_modeling.py:332-335) fused
// into the projection.  workspace as for b200_gemm_f16 (b200_gemm_workspace_bytes); without it every CTA takes whole
// super-tiles (no stream-K).
// splitk != NULL: deferred reduction - nothing is written to y; *splitk describes the fp32 partials left in the workspace.
static int gemm_w4a16_impl(const void* x, const void* packed, const void* bias, void* y, int64_t T, int64_t N, int64_t K, int groupsize,
                           int layout, int act, void* workspace, B200SplitK* splitk, void* stream) {
  if (T == 0 || N == 0) return B200_OK;
  if (!w4_check_shape(N, K, groupsize, "gemm_w4a16")) return B200_ERR_ARG;
  const int half_tiles = w4_half_tiles(N, layout, "gemm_w4a16");
  if (half_tiles < 0) return B200_ERR_ARG;
  if (act != 0 && (act != 1 || layout != 1)) { b200_set_last_error("gemm_w4a16: act = 1 needs the gate|up layout"); return B200_ERR_ARG; }
  const bool defer = splitk != nullptr;
  if (defer && (!workspace || T > 128)) {
    b200_set_last_error("gemm_w4a16_deferred: needs a workspace and T <= 128 (one token tile)");
    return B200_ERR_UNSUPPORTED;
  }
  const int gr = w4_group_rows(K, &groupsize);
  const W4Plan pl = launch_plan_w4(T, N, K, num_sms(), workspace != nullptr, defer);
  const CUtensorMap* mx = get_tmap_2d(x, T, K, K, pl.TN, 64, TmapDtype::kF16, TmapSwizzle::k128B);
  if (!mx) return B200_ERR_CUDA;
  if (defer) {
    splitk->partial = (const float*)((const char*)workspace + kW4CounterBytes);
    splitk->bias = bias;
    splitk->tiles_per_unit = kW4R;
    splitk->tn = pl.TN;
    splitk->nkb = pl.nkb;
    splitk->units_per_cta = pl.su_per_cta;
    splitk->max_contrib = pl.max_contrib;
    splitk->half_tiles = half_tiles;
    splitk->N = (int32_t)N;
    splitk->T = (int32_t)T;
  }
  cudaStream_t st = (cudaStream_t)stream;
#define W4_DISPATCH(TNV)                                                                                                           \
  (gr == 1 ? launch_gemm_w4<TNV, 1>(mx, packed, y, workspace, bias, (int)T, (int)N, pl, half_tiles, act, defer ? 1 : 0, st)          \
           : gr == 2 ? launch_gemm_w4<TNV, 2>(mx, packed, y, workspace, bias, (int)T, (int)N, pl, half_tiles, act, defer ? 1 : 0, st) \
                     : launch_gemm_w4<TNV, 4>(mx, packed, y, workspace, bias, (int)T, (int)N, pl, half_tiles, act, defer ? 1 : 0, st))
  switch (pl.TN) {
    case 16: return W4_DISPATCH(16);
    case 32: return W4_DISPATCH(32);
    case 64: return W4_DISPATCH(64);
    default: return W4_DISPATCH(128);
  }
#undef W4_DISPATCH
}
extern "C" int b200_gemm_w4a16_ex(const void* x, const void* packed, const void* bias, void* y, int64_t T, int64_t N, int64_t K,
                                  int groupsize, int layout, int act, void* workspace, void* stream) {
  return gemm_w4a16_impl(x, packed, bias, y, T, N, K, groupsize, layout, act, workspace, nullptr, stream);
}
extern "C" int b200_gemm_w4a16(const void* x, const void* packed, const void* bias, void* y, int64_t T, int64_t N, int64_t K,
                               int groupsize, void* workspace, void* stream) {
  return gemm_w4a16_impl(x, packed, bias, y, T, N, K, groupsize, 0, 0, workspace, nullptr, stream);
}
// Deferred split-K reduction: the GEMM leaves fp32 partials in `workspace` and fills *splitk; the next kernel of the stream
// (b200_rmsnorm_residual_splitk, b200_rope_kv_write_paged_splitk, b200_splitk_silu_mul, b200_splitk_reduce, the fused all-reduce)
// sums them in contributor order - bit-identical to b200_gemm_w4a16_ex's own fix-up, without its grid-wide wait.  The bias is
// applied by the consumer.  T <= 128.
extern "C" int b200_gemm_w4a16_deferred(const void* x, const void* packed, const void* bias, int64_t T, int64_t N, int64_t K,
                                        int groupsize, int layout, void* workspace, B200SplitK* splitk, void* stream) {
  if (!splitk) { b200_set_last_error("gemm_w4a16_deferred: splitk is NULL"); return B200_ERR_ARG; }
  return gemm_w4a16_impl(x, packed, bias, nullptr, T, N, K, groupsize, layout, 0, workspace, splitk, stream);
}
