// Fused next-token chooser: one kernel turns a row of fp16 logits into the chosen token id (+ its log-probability and rank),
// applying per-request parameters.
//
// Replaces, for one step of a batch, the torch op chain of HeterogeneousNextTokenChooser.__call__
// (/root/reference/server/text_generation_server/utils/tokens.py:238-271): the min_new_tokens EOS mask (:242-246) and the length
// penalty (:247-252), HeterogeneousRepetitionPenaltyLogitsProcessor (utils/logits_process.py:93-143), the temperature, top-k and
// top-p warpers (:146-176, :241-317, :179-238), Greedy / per-request Sampling (tokens.py:30-46, 49-78) and the log_softmax /
// rank of get_token_info (:388-425).  Rows with typical-p or top-n details stay on the host-side torch path.
//
// One CTA per row, the row is streamed from L2 a few times (B x V fp16 = 16 MB at bs 64, V 128k):
//   pass 0  warp: EOS mask, length penalty, repetition penalty (seen tokens = a V-bit map in shared memory built from the
//           row's history), temperature - in fp16 arithmetic like the torch ops on fp16 scores - into a scratch row; row max
//   top-k   two histogram passes (high byte, then low byte of the order-preserving 16-bit key of the fp16 value) give the
//           k-th largest VALUE; everything below it is dropped, ties with it stay (as `scores < kth` does)
//   top-p   two histogram passes over the survivors with probability MASS per bin: from the smallest value up, values are
//           dropped while their cumulative probability stays <= 1 - p (the most probable token always stays)
//   choose  arg-max of the surviving values (greedy rows), or of value + Gumbel noise (sampling rows: the same distribution as
//           the reference's `softmax(scores) / Exp(1)` arg-max), noise from Philox4x32-10 keyed by the request's seed and a
//           per-row draw counter kept on the device (CUDA-graph replayable); then log-softmax of the survivors at the chosen
//           id and its rank
// Determinism (tensor-parallel shards choose in lock step, sharded_client.rs:38-48): counts are integer atomics and masses are
// accumulated in 2^-40 fixed point, so no result depends on the order of atomic updates.
// Where this differs from the torch chain: values equal to a cut-off are kept or dropped as a group (torch.sort orders ties
// arbitrarily and cuts inside the group), and probability sums are fp32 / fixed point instead of fp16 cumsum.
#include "common.cuh"
#include "../../include/b200_tgis.h"

namespace b200 {

constexpr int kChThreads = 1024;
constexpr int kChMaxVocab = 131072;  // bitmap of seen tokens: 16 KB of shared memory
constexpr double kChFix = 1099511627776.0;  // 2^40

__device__ __forceinline__ uint32_t half_key(__half h) {  // order-preserving 16-bit key: larger value -> larger key
  const uint32_t b = __half_as_ushort(h);
  return (b & 0x8000u) ? (~b & 0xFFFFu) : (b | 0x8000u);
}
__device__ __forceinline__ __half key_half(uint32_t k) {
  const uint32_t b = (k & 0x8000u) ? (k & 0x7FFFu) : (~k & 0xFFFFu);
  return __ushort_as_half((unsigned short)b);
}

// Philox4x32-10 (Salmon et al.), counter = (c0, c1, c2, c3), key = (k0, k1)
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

template <typename T>
__device__ __forceinline__ T block_reduce_max(T v, T* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T ov = __shfl_xor_sync(0xffffffffu, v, o);
    v = ov > v ? ov : v;
  }
  __syncthreads();
  if (lane_id() == 0) scratch[warp_id()] = v;
  __syncthreads();
  v = scratch[lane_id()];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T ov = __shfl_xor_sync(0xffffffffu, v, o);
    v = ov > v ? ov : v;
  }
  return v;  // every thread holds the block maximum (32 warps: one scratch entry per lane)
}

struct ChooserArgs {
  const __half* logits;
  __half* warped;  // scratch [B, V]
  int64_t ld, V;
  const float* temperature;
  const int32_t* top_k;
  const float* top_p;
  const float* rep_penalty;
  const int64_t* history;
  int64_t history_stride;
  const int64_t* position_ids;
  int B;
  int history_len_bias;
  int64_t rep_exclude_id;
  const int64_t* banned_ids;
  const float* lp_factor;
  int64_t eos_id;
  const uint64_t* seeds;
  int64_t* counters;
  int64_t* next_ids;
  float* logprobs;
  int32_t* ranks;
};

__global__ void __launch_bounds__(kChThreads, 1) chooser_kernel(const ChooserArgs a) {
  __shared__ uint32_t s_seen[kChMaxVocab / 32];
  __shared__ uint32_t s_cnt[256];
  __shared__ unsigned long long s_mass[256];
  __shared__ float s_redf[32];
  __shared__ int s_redi[32];
  __shared__ unsigned long long s_redu[32];
  __shared__ uint32_t s_sel[4];
  __shared__ unsigned long long s_acc;

  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x, tid = threadIdx.x;
  const int V = (int)a.V;
  const __half* in = a.logits + (size_t)row * a.ld;
  __half* w = a.warped + (size_t)row * a.V;
  const float temp = a.temperature ? a.temperature[row] : 0.f;
  const bool sample = temp != 0.f;
  const float tdiv = sample ? temp : 1.f;
  const int k_top = a.top_k ? a.top_k[row] : 0;
  const float p_top = a.top_p ? a.top_p[row] : 1.f;
  const float pen = a.rep_penalty ? a.rep_penalty[row] : 1.f;
  const int ban = a.banned_ids ? (int)a.banned_ids[row] : -1;
  const float lpf = a.lp_factor ? a.lp_factor[row] : 0.f;

  // ---- seen-token bitmap (repetition penalty): the history every row sees is all_input_ids[:, :max_seqlen] (tokens.py:256,
  // flash_causal_lm.py:525-527), pads of shorter rows included
  const bool use_pen = pen != 1.f && a.history != nullptr;
  if (use_pen) {
    for (int i = tid; i < kChMaxVocab / 32; i += kChThreads) s_seen[i] = 0;
    int s_len = 0;
    for (int b = 0; b < a.B; ++b) s_len = max(s_len, (int)a.position_ids[b] + a.history_len_bias);
    s_len = min(s_len, (int)a.history_stride);
    __syncthreads();
    const int64_t* hist = a.history + (size_t)row * a.history_stride;
    for (int i = tid; i < s_len; i += kChThreads) {
      const int64_t t = hist[i];
      if (t >= 0 && t < V) atomicOr(&s_seen[t >> 5], 1u << (t & 31));
    }
    __syncthreads();
  }

  // ---- pass 0: warp the row (fp16 arithmetic with one rounding per op, like the torch ops on fp16 scores), find the maximum
  float mx = -INFINITY;
  for (int i = tid; i < V; i += kChThreads) {
    __half h = in[i];
    if (i == ban) h = __float2half_rn(-INFINITY);
    else if (lpf != 0.f && i == (int)a.eos_id) {
      const float e = __half2float(h);
      h = __float2half_rn(e + __half2float(__float2half_rn(fabsf(e) * lpf)));
    }
    if (use_pen && (s_seen[i >> 5] >> (i & 31) & 1u) && !(a.rep_exclude_id == i && a.B != 1)) {
      const float x = __half2float(h);
      h = __float2half_rn(x < 0.f ? x * pen : x / pen);
    }
    if (tdiv != 1.f) h = __float2half_rn(__half2float(h) / tdiv);
    w[i] = h;
    mx = fmaxf(mx, __half2float(h));
  }
  mx = block_reduce_max<float>(mx, s_redf);
  __syncthreads();  // the row is re-read by other threads below

  // ---- top-k: the k-th largest value (two-level histogram over the 16-bit key)
  uint32_t key_min = 0;  // survivors have key >= key_min
  if (k_top > 0 && k_top < V) {
    uint32_t prefix = 0;
    uint32_t need = (uint32_t)k_top;
    for (int level = 0; level < 2; ++level) {
      for (int i = tid; i < 256; i += kChThreads) s_cnt[i] = 0;
      __syncthreads();
      for (int i = tid; i < V; i += kChThreads) {
        const uint32_t key = half_key(w[i]);
        if (level == 0) atomicAdd(&s_cnt[key >> 8], 1u);
        else if ((key >> 8) == prefix) atomicAdd(&s_cnt[key & 255], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        uint32_t cum = 0;
        int b = 255;
        for (; b > 0; --b) {
          if (cum + s_cnt[b] >= need) break;
          cum += s_cnt[b];
        }
        s_sel[0] = (uint32_t)b;
        s_sel[1] = need - cum;  // still needed inside the selected bin
      }
      __syncthreads();
      if (level == 0) { prefix = s_sel[0]; need = s_sel[1]; }
      else key_min = (prefix << 8) | s_sel[0];
      __syncthreads();
    }
  }

  // ---- top-p over the survivors
  if (p_top < 1.f) {
    uint32_t prefix = 0;
    unsigned long long cum_below = 0, thr = 0;
    for (int level = 0; level < 2; ++level) {
      for (int i = tid; i < 256; i += kChThreads) s_mass[i] = 0;
      __syncthreads();
      for (int i = tid; i < V; i += kChThreads) {
        const __half h = w[i];
        const uint32_t key = half_key(h);
        if (key < key_min) continue;
        if (level == 1 && (key >> 8) != prefix) continue;
        const unsigned long long m = (unsigned long long)((double)__expf(__half2float(h) - mx) * kChFix);
        atomicAdd(&s_mass[level == 0 ? (key >> 8) : (key & 255)], m);
      }
      __syncthreads();
      if (tid == 0) {
        if (level == 0) {
          unsigned long long z = 0;
          for (int b = 0; b < 256; ++b) z += s_mass[b];
          thr = (unsigned long long)((double)(1.f - p_top) * (double)z);  // mass that may be dropped
          s_acc = thr;
        }
        const unsigned long long limit = s_acc;
        unsigned long long cum = level == 0 ? 0 : cum_below;
        int b = 0;
        for (; b < 255; ++b) {  // ascending: drop bins while the cumulative mass stays <= limit
          if (cum + s_mass[b] > limit) break;
          cum += s_mass[b];
        }
        s_sel[0] = (uint32_t)b;
        s_redu[0] = cum;
      }
      __syncthreads();
      if (level == 0) { prefix = s_sel[0]; cum_below = s_redu[0]; }
      else {
        const uint32_t key_p = (prefix << 8) | s_sel[0];
        if (key_p > key_min) key_min = key_p;
      }
      __syncthreads();
    }
    // the most probable token always survives (min_tokens_to_keep = 1, logits_process.py:215)
    const uint32_t key_mx = half_key(__float2half_rn(mx));
    if (key_min > key_mx) key_min = key_mx;
  }

  // ---- choose: arg-max of value (+ Gumbel noise on sampling rows) over the survivors; total surviving mass for the log-softmax
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  unsigned long long mass = 0;
  uint2 pkey = make_uint2(0, 0);
  uint32_t draw = 0;
  if (sample) {
    const uint64_t seed = a.seeds ? a.seeds[row] : 0;
    pkey = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    draw = (uint32_t)a.counters[row];
  }
  for (int i0 = tid * 4; i0 < V; i0 += kChThreads * 4) {
    uint4 rnd = make_uint4(0, 0, 0, 0);
    if (sample) rnd = philox4x32(make_uint4((uint32_t)(i0 >> 2), draw, (uint32_t)row, 0x5eedu), pkey);
    const uint32_t r[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j;
      if (i >= V) break;
      const __half h = w[i];
      if (half_key(h) < key_min) continue;
      const float x = __half2float(h);
      if (x == -INFINITY) continue;
      mass += (unsigned long long)((double)__expf(x - mx) * kChFix);
      float score = x;
      if (sample) {
        const float u = ((float)(r[j] >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0, 1)
        score = x - __logf(-__logf(u));                                       // Gumbel(0, 1)
      }
      if (score > best || (score == best && i < best_i)) { best = score; best_i = i; }
    }
  }
  // block arg-max (value, then lowest index) and mass sum
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    mass += __shfl_xor_sync(0xffffffffu, mass, o);
  }
  __syncthreads();
  if (lane_id() == 0) { s_redf[warp_id()] = best; s_redi[warp_id()] = best_i; s_redu[warp_id()] = mass; }
  __syncthreads();
  if (warp_id() == 0) {
    best = s_redf[lane_id()];
    best_i = s_redi[lane_id()];
    mass = s_redu[lane_id()];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
      mass += __shfl_xor_sync(0xffffffffu, mass, o);
    }
    if (lane_id() == 0) {
      if (best_i == 0x7fffffff) best_i = 0;
      s_sel[2] = (uint32_t)best_i;
      s_acc = mass;
      a.next_ids[row] = best_i;
      if (sample) a.counters[row] = (int64_t)draw + 1;
    }
  }
  __syncthreads();
  const int chosen = (int)s_sel[2];
  const float xc = __half2float(w[chosen]);
  if (a.logprobs && tid == 0) a.logprobs[row] = (xc - mx) - logf((float)((double)s_acc / kChFix));
  if (a.ranks) {  // 1 + number of surviving-or-not values strictly greater than the chosen one (tokens.py:423-424 on the warped scores)
    int cnt = 0;
    const uint32_t kc = half_key(w[chosen]);
    for (int i = tid; i < V; i += kChThreads) {
      const uint32_t key = half_key(w[i]);
      cnt += (key >= key_min && key > kc) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    __syncthreads();
    if (lane_id() == 0) s_redi[warp_id()] = cnt;
    __syncthreads();
    if (warp_id() == 0) {
      cnt = s_redi[lane_id()];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (lane_id() == 0) a.ranks[row] = cnt + 1;
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_choose_tokens(const B200ChooserParams* p, void* stream) {
  if (!p || !p->logits || !p->warped_scratch || !p->next_ids || p->B < 0) {
    b200_set_last_error("choose_tokens: logits, warped_scratch and next_ids are required");
    return B200_ERR_ARG;
  }
  if (p->B == 0) return B200_OK;
  if (p->V <= 0 || p->V > kChMaxVocab || p->ld < p->V) { b200_set_last_error("choose_tokens: need 0 < V <= 131072 and ld >= V"); return B200_ERR_ARG; }
  if (p->rep_penalty && (!p->history || !p->position_ids)) {
    b200_set_last_error("choose_tokens: the repetition penalty needs history and position_ids");
    return B200_ERR_ARG;
  }
  if (p->temperature && !p->counters) { b200_set_last_error("choose_tokens: sampling rows need draw counters"); return B200_ERR_ARG; }
  ChooserArgs a;
  a.logits = (const __half*)p->logits;
  a.warped = (__half*)p->warped_scratch;
  a.ld = p->ld;
  a.V = p->V;
  a.temperature = p->temperature;
  a.top_k = p->top_k;
  a.top_p = p->top_p;
  a.rep_penalty = p->rep_penalty;
  a.history = p->history;
  a.history_stride = p->history_stride;
  a.position_ids = p->position_ids;
  a.B = p->B;
  a.history_len_bias = p->history_len_bias;
  a.rep_exclude_id = p->rep_exclude_id;
  a.banned_ids = p->banned_ids;
  a.lp_factor = p->length_penalty_factor;
  a.eos_id = p->eos_id;
  a.seeds = p->seeds;
  a.counters = p->counters;
  a.next_ids = p->next_ids;
  a.logprobs = p->logprobs;
  a.ranks = p->ranks;
  B200_LAUNCH(chooser_kernel, dim3((unsigned)p->B), dim3(kChThreads), 0, (cudaStream_t)stream, a);
  b200_count_launches(1);
  return B200_OK;
}
