// HBM-bound row kernels of the decode step: fused residual-add + RMSNorm, RoPE + paged KV write,
// SiLU*mul, vocabulary-parallel embedding gather, greedy arg-max.
//
// Reference ops replaced (paths under /root/reference/server/text_generation_server/):
//   rmsnorm_residual ......... dropout_layer_norm.dropout_add_ln_fwd, models/custom_modeling/flash_llama_modeling.py:132-148
//   rope_kv_write_paged ...... rotary_emb.apply_rotary (utils/layers.py:466-472) + KV append
//                              (flash_llama_modeling.py:268,282) / fms-extras reshape_and_cache (paged_llama_modeling.py:250)
//   silu_mul ................. flash_llama_modeling.py:332-335
//   embedding ................ TensorParallelEmbedding.forward, utils/layers.py:346-357
//   argmax ................... Greedy, utils/tokens.py:44-46
#include "common.cuh"
#include "splitk.cuh"

#include <string>

namespace b200 {

// ------------------------------------------------------------------------------------------------
// residual-add + RMSNorm.  One CTA per token row; 8 halves (16 B) per thread per iteration.
// fp32 statistics on the un-rounded sum.
// Optional split-K input: `h` may be `n_parts` fp32 partial matrices [n_parts][T][H] (gemm split-K
// workspace); they are summed and rounded to fp16 first, which is exactly the GEMM's fp16 output.
// ------------------------------------------------------------------------------------------------
template <int kThreads, bool kSplitK>
__global__ void __launch_bounds__(kThreads) rmsnorm_residual_kernel(const __half* __restrict__ h,
                                                                      const __half* __restrict__ residual,
                                                                      const __half* __restrict__ gamma,
                                                                      __half* __restrict__ normed, __half* __restrict__ res_out,
                                                                      int H, float eps, const B200SplitK parts) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* xs = reinterpret_cast<float*>(smem_raw);  // [H] fp32 copy of the summed row
  __shared__ float red[32];
  const int row = blockIdx.x;
  const __half* hp = h + (size_t)row * H;
  const __half* rp = residual ? residual + (size_t)row * H : nullptr;
  float ss = 0.f;
  for (int i = threadIdx.x * 8; i < H; i += kThreads * 8) {
    float x[8];
    if constexpr (kSplitK) {
      // h is a deferred GEMM output: sum its fp32 partials and round to fp16 first, which is exactly the GEMM's fp16 result
      const float4 a = splitk_sum4(parts, row, i), b = splitk_sum4(parts, row, i + 4);
      const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = __half2float(__float2half_rn(v[j]));
    } else {
      uint4 hv = *reinterpret_cast<const uint4*>(hp + i);
      const __half2* h2 = reinterpret_cast<const __half2*>(&hv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(h2[j]);
        x[2 * j] = f.x;
        x[2 * j + 1] = f.y;
      }
    }
    if (rp) {
      uint4 rv = *reinterpret_cast<const uint4*>(rp + i);
      const __half2* r2 = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(r2[j]);
        x[2 * j] += f.x;
        x[2 * j + 1] += f.y;
      }
      uint4 ov;
      __half2* o2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
      for (int j = 0; j < 4; ++j) o2[j] = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
      *reinterpret_cast<uint4*>(res_out + (size_t)row * H + i) = ov;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xs[i + j] = x[j];
      ss += x[j] * x[j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane_id() == 0) red[warp_id()] = ss;
  __syncthreads();
  if (warp_id() == 0) {
    float v = lane_id() < kThreads / 32 ? red[lane_id()] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane_id() == 0) red[0] = v;
  }
  __syncthreads();
  const float rstd = rsqrtf(red[0] / (float)H + eps);
  for (int i = threadIdx.x * 8; i < H; i += kThreads * 8) {
    uint4 gv = *reinterpret_cast<const uint4*>(gamma + i);
    const __half2* g2 = reinterpret_cast<const __half2*>(&gv);
    uint4 ov;
    __half2* o2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 g = __half22float2(g2[j]);
      o2[j] = __floats2half2_rn(xs[i + 2 * j] * rstd * g.x, xs[i + 2 * j + 1] * rstd * g.y);
    }
    *reinterpret_cast<uint4*>(normed + (size_t)row * H + i) = ov;
  }
}

// ------------------------------------------------------------------------------------------------
// Deferred split-K consumers without another op to fuse into: materialise the fp16 GEMM output, and the LlamaMLP
// activation SiLU(gate) * up of a gate|up projection (same arithmetic as silu_mul_kernel on the two fp16 outputs).
// grid (N / 4 / 256, T); one float4 of columns per thread.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ __half splitk_silu_mul_f16(float gate_acc, float up_acc) {
  const float g = __half2float(__float2half_rn(gate_acc));
  const __half a = __float2half_rn(g / (1.f + expf(-g)));
  return __hmul(a, __float2half_rn(up_acc));
}
__global__ void __launch_bounds__(256) splitk_reduce_kernel(__half* __restrict__ y, const B200SplitK parts) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.y;
  const int n = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (n >= parts.N) return;
  const float4 a = splitk_sum4(parts, t, n);
  uint2 o;
  o.x = pack_half2(a.x, a.y);
  o.y = pack_half2(a.z, a.w);
  *reinterpret_cast<uint2*>(y + (size_t)t * parts.N + n) = o;
}
__global__ void __launch_bounds__(256) splitk_silu_mul_kernel(__half* __restrict__ out, const B200SplitK parts) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.y;
  const int I = parts.N / 2;
  const int n = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (n >= I) return;
  const float4 g = splitk_sum4(parts, t, n), u = splitk_sum4(parts, t, n + I);
  __half o[4] = {splitk_silu_mul_f16(g.x, u.x), splitk_silu_mul_f16(g.y, u.y), splitk_silu_mul_f16(g.z, u.z),
                 splitk_silu_mul_f16(g.w, u.w)};
  *reinterpret_cast<uint2*>(out + (size_t)t * I + n) = *reinterpret_cast<uint2*>(o);
}

// ------------------------------------------------------------------------------------------------
// RoPE (half-split pairs, fp32 math, fp16 tables gathered by position) applied in place to q and k of
// the fused qkv activation, plus the scatter of k and v into the paged pool.
// Pool layout per layer: K,V [num_blocks][n_kv][16 tokens][d] fp16, 16-byte chunks XOR-swizzled with
// (token & 7) so a page-head tile lands bank-conflict-free in shared memory with one bulk copy.
// grid = (T, n_heads + 2*n_kv); block = d/2 threads... one thread per rotation pair.
// ------------------------------------------------------------------------------------------------
// kSplitK: the fused QKV projection arrives as deferred split-K partials (`parts`); the sums, rounded to fp16, are the GEMM
// output, and qkv receives the final activation.
template <int kHeadDim, bool kSplitK>
__global__ void rope_kv_write_kernel(__half* __restrict__ qkv, const __half* __restrict__ cos_t,
                                     const __half* __restrict__ sin_t, const int64_t* __restrict__ position_ids,
                                     const int64_t* __restrict__ slot_mapping, __half* __restrict__ k_pool,
                                     __half* __restrict__ v_pool, int n_heads, int n_kv, int rot_half, const B200SplitK parts) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x;
  const int head = blockIdx.y;  // [0,h): q, [h,h+kv): k, [h+kv, h+2kv): v
  const int j = threadIdx.x;    // 0..d/2-1: this thread owns elements j and j + d/2 of the head row
  __half* base = qkv + ((size_t)t * (n_heads + 2 * n_kv) + head) * kHeadDim;
  const bool is_v = head >= n_heads + n_kv;
  __half lo, hi;
  if constexpr (kSplitK) {
    lo = __float2half_rn(splitk_sum1(parts, t, head * kHeadDim + j));
    hi = __float2half_rn(splitk_sum1(parts, t, head * kHeadDim + j + kHeadDim / 2));
    if (is_v) {  // v is not rotated: it only has to reach qkv (prefill-style readers) and the pool
      base[j] = lo;
      base[j + kHeadDim / 2] = hi;
    }
  } else {
    lo = base[j];
    hi = base[j + kHeadDim / 2];
  }
  if (!is_v) {
    // rotary_dim = cos.shape[-1] * 2 = 2 rot_half (utils/layers.py:467-469): pairs (x[i], x[i + rot_half]), i < rot_half;
    // elements from 2 rot_half on pass through (GPT-NeoX rotary_pct < 1).  Full rotation (Llama): rot_half == d/2 and
    // the pair is exactly this thread's (lo, hi).
    const int64_t pos = position_ids[t];
    if (rot_half == kHeadDim / 2) {
      const float c = __half2float(cos_t[pos * rot_half + j]);
      const float s = __half2float(sin_t[pos * rot_half + j]);
      const float x1 = __half2float(lo), x2 = __half2float(hi);
      lo = __float2half_rn(x1 * c - x2 * s);
      hi = __float2half_rn(x1 * s + x2 * c);
    } else {
      __shared__ __half row[kHeadDim];
      row[j] = lo;
      row[j + kHeadDim / 2] = hi;
      __syncthreads();
      if (j < rot_half) {
        const float c = __half2float(cos_t[pos * rot_half + j]);
        const float s = __half2float(sin_t[pos * rot_half + j]);
        const float x1 = __half2float(row[j]), x2 = __half2float(row[j + rot_half]);
        row[j] = __float2half_rn(x1 * c - x2 * s);
        row[j + rot_half] = __float2half_rn(x1 * s + x2 * c);
      }
      __syncthreads();
      lo = row[j];
      hi = row[j + kHeadDim / 2];
    }
    base[j] = lo;
    base[j + kHeadDim / 2] = hi;
  }
  if (head >= n_heads) {
    const int64_t slot = slot_mapping[t];
    if (slot < 0) return;  // padding token
    const int64_t blk = slot / kPageTokens;
    const int tok = (int)(slot % kPageTokens);
    const int hk = is_v ? head - n_heads - n_kv : head - n_heads;
    __half* pool = is_v ? v_pool : k_pool;
    unsigned char* tile = reinterpret_cast<unsigned char*>(pool + ((size_t)blk * n_kv + hk) * kPageTokens * kHeadDim);
    const int e_lo = j, e_hi = j + kHeadDim / 2;
    *reinterpret_cast<__half*>(tile + kv_swizzled_chunk_offset<kHeadDim>(tok, e_lo >> 3) + (e_lo & 7) * 2) = lo;
    *reinterpret_cast<__half*>(tile + kv_swizzled_chunk_offset<kHeadDim>(tok, e_hi >> 3) + (e_hi & 7) * 2) = hi;
  }
}

// Prefill-sized form of the same op (full rotation, plain fp16 input): 16-byte accesses, one (token, head) row per D/16 threads,
// grid-stride.  The per-(token, head) CTAs above are launch-bound at T = 65 536 (3 M CTAs of 64 threads: 4.9 ms per layer under ncu,
// as long as the QKV projection); this form moves the same 1.7 GB at HBM speed.
template <int kHeadDim>
__global__ void __launch_bounds__(256) rope_kv_write_vec_kernel(__half* __restrict__ qkv, const __half* __restrict__ cos_t,
                                                                const __half* __restrict__ sin_t, const int64_t* __restrict__ position_ids,
                                                                const int64_t* __restrict__ slot_mapping, __half* __restrict__ k_pool,
                                                                __half* __restrict__ v_pool, int n_heads, int n_kv, int64_t n_rows) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int kLanes = kHeadDim / 16;  // threads per (token, head) row: each owns 8 halves of the low half and the matching 8 of the high half
  const int n_all = n_heads + 2 * n_kv;
  const int lane = threadIdx.x % kLanes;
  for (int64_t r = (int64_t)blockIdx.x * (256 / kLanes) + threadIdx.x / kLanes; r < n_rows; r += (int64_t)gridDim.x * (256 / kLanes)) {
    const int64_t t = r / n_all;
    const int head = (int)(r - t * n_all);
    __half* base = qkv + (size_t)r * kHeadDim;
    const bool is_v = head >= n_heads + n_kv;
    uint4 lo = *reinterpret_cast<const uint4*>(base + lane * 8);
    uint4 hi = *reinterpret_cast<const uint4*>(base + kHeadDim / 2 + lane * 8);
    if (!is_v) {
      const int64_t pos = position_ids[t];
      const uint4 cv = *reinterpret_cast<const uint4*>(cos_t + pos * (kHeadDim / 2) + lane * 8);
      const uint4 sv = *reinterpret_cast<const uint4*>(sin_t + pos * (kHeadDim / 2) + lane * 8);
      __half* l = reinterpret_cast<__half*>(&lo);
      __half* h = reinterpret_cast<__half*>(&hi);
      const __half* c = reinterpret_cast<const __half*>(&cv);
      const __half* sn = reinterpret_cast<const __half*>(&sv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x1 = __half2float(l[i]), x2 = __half2float(h[i]), cc = __half2float(c[i]), ss = __half2float(sn[i]);
        l[i] = __float2half_rn(x1 * cc - x2 * ss);
        h[i] = __float2half_rn(x1 * ss + x2 * cc);
      }
      *reinterpret_cast<uint4*>(base + lane * 8) = lo;
      *reinterpret_cast<uint4*>(base + kHeadDim / 2 + lane * 8) = hi;
    }
    if (head >= n_heads) {
      const int64_t slot = slot_mapping[t];
      if (slot < 0) continue;  // padding token
      const int64_t blk = slot / kPageTokens;
      const int tok = (int)(slot % kPageTokens);
      const int hk = is_v ? head - n_heads - n_kv : head - n_heads;
      __half* pool = is_v ? v_pool : k_pool;
      unsigned char* tile = reinterpret_cast<unsigned char*>(pool + ((size_t)blk * n_kv + hk) * kPageTokens * kHeadDim);
      *reinterpret_cast<uint4*>(tile + kv_swizzled_chunk_offset<kHeadDim>(tok, lane)) = lo;
      *reinterpret_cast<uint4*>(tile + kv_swizzled_chunk_offset<kHeadDim>(tok, kHeadDim / 16 + lane)) = hi;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// residual-add + LayerNorm (FastLayerNorm, utils/layers.py:360-392 -> dropout_layer_norm.dropout_add_ln_fwd):
// x = h + residual in fp32, residual_out = fp16(x), normed = fp16((x - mean) * rstd * gamma + beta), fp32 statistics.
// One CTA per token row.
// ------------------------------------------------------------------------------------------------
template <int kThreads>
__global__ void __launch_bounds__(kThreads) layernorm_residual_kernel(const __half* __restrict__ h, const __half* __restrict__ residual,
                                                                        const __half* __restrict__ gamma, const __half* __restrict__ beta,
                                                                        __half* __restrict__ normed, __half* __restrict__ res_out, int H,
                                                                        float eps) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* xs = reinterpret_cast<float*>(smem_raw);  // [H] fp32 copy of the summed row
  __shared__ float red[32];
  const int row = blockIdx.x;
  const __half* hp = h + (size_t)row * H;
  const __half* rp = residual ? residual + (size_t)row * H : nullptr;
  auto block_sum = [&](float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();  // red[] reuse
    if (lane_id() == 0) red[warp_id()] = v;
    __syncthreads();
    float t = lane_id() < kThreads / 32 ? red[lane_id()] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
  };
  float sum = 0.f;
  for (int i = threadIdx.x * 2; i < H; i += kThreads * 2) {
    float2 x = __half22float2(*reinterpret_cast<const __half2*>(hp + i));
    if (rp) {
      const float2 r = __half22float2(*reinterpret_cast<const __half2*>(rp + i));
      x.x += r.x;
      x.y += r.y;
      if (res_out) *reinterpret_cast<__half2*>(res_out + (size_t)row * H + i) = __floats2half2_rn(x.x, x.y);
    }
    xs[i] = x.x;
    xs[i + 1] = x.y;
    sum += x.x + x.y;
  }
  const float mean = block_sum(sum) / H;
  float sq = 0.f;
  for (int i = threadIdx.x; i < H; i += kThreads) {
    const float d = xs[i] - mean;
    sq += d * d;
  }
  const float rstd = rsqrtf(block_sum(sq) / H + eps);
  for (int i = threadIdx.x * 2; i < H; i += kThreads * 2) {
    const float2 g = __half22float2(*reinterpret_cast<const __half2*>(gamma + i));
    const float2 b = beta ? __half22float2(*reinterpret_cast<const __half2*>(beta + i)) : make_float2(0.f, 0.f);
    *reinterpret_cast<__half2*>(normed + (size_t)row * H + i) =
        __floats2half2_rn((xs[i] - mean) * rstd * g.x + b.x, (xs[i + 1] - mean) * rstd * g.y + b.y);
  }
}

// ------------------------------------------------------------------------------------------------
// out = gelu(x) on fp16 (fp32 math, one rounding: torch's fp16 GELU); tanh approximation for gelu_fast / gelu_pytorch_tanh
// (flash_neox_modeling.py:186-196)
// ------------------------------------------------------------------------------------------------
__global__ void gelu_kernel(const __half* __restrict__ x, __half* __restrict__ out, int64_t n2, int approximate_tanh) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 v = __half22float2(reinterpret_cast<const __half2*>(x)[i]);
    float2 r;
    if (approximate_tanh) {
      const float k0 = 0.7978845608028654f, k1 = 0.044715f;
      r.x = 0.5f * v.x * (1.f + tanhf(k0 * (v.x + k1 * v.x * v.x * v.x)));
      r.y = 0.5f * v.y * (1.f + tanhf(k0 * (v.y + k1 * v.y * v.y * v.y)));
    } else {
      r.x = 0.5f * v.x * (1.f + erff(v.x * 0.7071067811865476f));
      r.y = 0.5f * v.y * (1.f + erff(v.y * 0.7071067811865476f));
    }
    reinterpret_cast<__half2*>(out)[i] = __floats2half2_rn(r.x, r.y);
  }
}

// ------------------------------------------------------------------------------------------------
// out[t, i] = fp16( fp16(silu(g)) * u ),  gate_up [T, 2, I]
// ------------------------------------------------------------------------------------------------
__global__ void silu_mul_kernel(const __half* __restrict__ gate_up, __half* __restrict__ out, int64_t I, int64_t total8) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total8; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t per_row = I / 8;
    const int64_t t = idx / per_row;
    const int64_t i = (idx % per_row) * 8;
    uint4 gv = *reinterpret_cast<const uint4*>(gate_up + t * 2 * I + i);
    uint4 uv = *reinterpret_cast<const uint4*>(gate_up + t * 2 * I + I + i);
    const __half2* g2 = reinterpret_cast<const __half2*>(&gv);
    const __half2* u2 = reinterpret_cast<const __half2*>(&uv);
    uint4 ov;
    __half2* o2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 g = __half22float2(g2[j]);
      // torch fp16 SiLU: fp32 x / (1 + exp(-x)), one rounding; then an fp16 multiply
      __half2 a = __floats2half2_rn(g.x / (1.f + expf(-g.x)), g.y / (1.f + expf(-g.y)));
      o2[j] = __hmul2(a, u2[j]);
    }
    *reinterpret_cast<uint4*>(out + t * I + i) = ov;
  }
}

// ------------------------------------------------------------------------------------------------
// out[t] = (vocab_start <= id < vocab_start + rows) ? table[id - vocab_start] : 0
// ------------------------------------------------------------------------------------------------
__global__ void embedding_kernel(const __half* __restrict__ table, const int64_t* __restrict__ ids, __half* __restrict__ out,
                                 int H, int64_t vocab_start, int64_t rows) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x;
  const int64_t id = ids[t] - vocab_start;
  const bool ok = id >= 0 && id < rows;
  for (int i = threadIdx.x * 8; i < H; i += blockDim.x * 8) {
    uint4 v = ok ? *reinterpret_cast<const uint4*>(table + id * H + i) : make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(out + (size_t)t * H + i) = v;
  }
}

// ------------------------------------------------------------------------------------------------
// greedy arg-max over fp16 logits (first index wins ties, like torch.argmax on a row scan)
// ------------------------------------------------------------------------------------------------
constexpr int kArgmaxThreads = 1024;
__global__ void __launch_bounds__(kArgmaxThreads) argmax_kernel(const __half* __restrict__ logits, int64_t* __restrict__ out, int64_t V,
                                                                 int64_t ld, const int64_t* __restrict__ banned) {
  pdl_launch_dependents();
  pdl_wait();
  const __half* row = logits + (size_t)blockIdx.x * ld;
  // banned[row] >= 0: that token's score counts as -inf (min_new_tokens EOS mask, utils/tokens.py:244-246)
  const int ban = banned ? (int)banned[blockIdx.x] : -1;
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  // 16-byte loads need rows that start on a 16-byte boundary; otherwise (e.g. a vocabulary of 32001 after an added pad
  // token) the whole row takes the scalar loop below
  const bool vec = (ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  const int V8 = vec ? (int)(V & ~7LL) : 0;
  // a thread visits its indices in increasing order, so a strict '>' keeps the first maximum
  for (int i = threadIdx.x * 8; i < V8; i += kArgmaxThreads * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(row + i);
    const __half* hv = reinterpret_cast<const __half*>(&v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float f = (i + j == ban) ? -INFINITY : __half2float(hv[j]);
      if (f > best) { best = f; best_i = i + j; }
    }
  }
  for (int i = V8 + threadIdx.x; i < (int)V; i += kArgmaxThreads) {
    const float f = (i == ban) ? -INFINITY : __half2float(row[i]);
    if (f > best || (f == best && i < best_i)) { best = f; best_i = i; }
  }
  __shared__ float sv[kArgmaxThreads / 32];
  __shared__ int si[kArgmaxThreads / 32];
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
    if (lane_id() == 0) out[blockIdx.x] = best_i == 0x7fffffff ? 0 : best_i;
  }
}

// ------------------------------------------------------------------------------------------------
// Masked softmax over rows of attention scores: the one CUDA kernel of the reference tree,
// server/custom_kernels/custom_kernels/fused_attention_cuda.cu:28-107 (and its BLOOM twin): cast to fp32, positions with
// mask != 0 are excluded, softmax over the rest in fp32, masked positions give 0, an all-masked row gives zeros (:95-99),
// cast back.  One warp per row, 16-byte loads, shuffle reductions: no shared-memory atomics and no kv_length <= 4096 limit
// (the reference needs kv / 4 <= 1024 threads, bloom_modeling.py:386-389).
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }

template <typename T>
__global__ void __launch_bounds__(256) masked_softmax_kernel(const T* __restrict__ scores, const uint8_t* __restrict__ mask,
                                                             T* __restrict__ out, int64_t rows, int64_t kv) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + warp_id();
  if (row >= rows) return;
  const T* sp = scores + row * kv;
  const uint8_t* mp = mask + row * kv;
  T* op = out + row * kv;
  const int lane = lane_id();
  float mx = -INFINITY;
  for (int64_t i = lane; i < kv; i += 32)
    if (mp[i] == 0) mx = fmaxf(mx, to_f32(sp[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int64_t i = lane; i < kv; i += 32)
    if (mp[i] == 0) sum += expf(to_f32(sp[i]) - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  for (int64_t i = lane; i < kv; i += 32)
    op[i] = from_f32<T>((mp[i] == 0 && sum > 0.f) ? expf(to_f32(sp[i]) - mx) / sum : 0.f);
}

// ------------------------------------------------------------------------------------------------
// out[t, k'] = x[t, perm[k']]   (act-order GPTQ: activations follow the row order of the packed weight)
// ------------------------------------------------------------------------------------------------
__global__ void permute_columns_kernel(const __half* __restrict__ x, const int32_t* __restrict__ perm, __half* __restrict__ out,
                                       int64_t K) {
  pdl_launch_dependents();
  pdl_wait();
  const __half* xr = x + (size_t)blockIdx.x * K;
  __half* orow = out + (size_t)blockIdx.x * K;
  for (int64_t k = threadIdx.x * 2; k < K; k += blockDim.x * 2) {  // K % 2 == 0
    __half2 v;
    v.x = xr[perm[k]];
    v.y = xr[perm[k + 1]];
    *reinterpret_cast<__half2*>(orow + k) = v;
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_rmsnorm_residual(const void* h, const void* residual, const void* gamma, void* normed_out,
                                     void* residual_out, int64_t T, int64_t H, float eps, void* stream) {
  if (T == 0) return B200_OK;
  if (H % 8 != 0 || H > 16384 || !h || !gamma || !normed_out || (residual && !residual_out)) {
    b200_set_last_error("rmsnorm_residual: need H % 8 == 0, H <= 16384, non-null h/gamma/out");
    return B200_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)H * sizeof(float);
  // 512 threads like the split-K and the fused all-reduce variants: the same thread -> element mapping and reduction tree, so the
  // three produce bit-identical statistics for the same row
  constexpr auto kernel = rmsnorm_residual_kernel<512, false>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  B200_LAUNCH_AS("rmsnorm_residual_kernel", kernel, dim3((unsigned)T), dim3(512), smem, st, (const __half*)h, (const __half*)residual,
                 (const __half*)gamma, (__half*)normed_out, (__half*)residual_out, (int)H, eps, B200SplitK{});
  b200_count_launches(1);
  return B200_OK;
}

static bool splitk_ok(const B200SplitK* p, const char* who) {
  if (!p || !p->partial || p->T <= 0 || p->T > p->tn || p->N <= 0 || p->N % 8 != 0 || p->nkb <= 0 || p->units_per_cta <= 0 ||
      p->max_contrib <= 0 || p->tiles_per_unit <= 0) {
    b200_set_last_error((std::string(who) + ": not a valid B200SplitK (it comes from a *_deferred GEMM)").c_str());
    return false;
  }
  return true;
}

extern "C" int b200_rmsnorm_residual_splitk(const B200SplitK* h_parts, const void* residual, const void* gamma, void* normed_out,
                                            void* residual_out, float eps, void* stream) {
  if (!splitk_ok(h_parts, "rmsnorm_residual_splitk")) return B200_ERR_ARG;
  const int64_t T = h_parts->T, H = h_parts->N;
  if (H > 16384 || h_parts->half_tiles != 0 || !gamma || !normed_out || !residual || !residual_out) {
    b200_set_last_error("rmsnorm_residual_splitk: need H <= 16384, the plain layout, non-null residual/gamma/outputs");
    return B200_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)H * sizeof(float);
  constexpr auto kernel = rmsnorm_residual_kernel<512, true>;  // more threads: the partial loads are L2-latency bound
  if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  B200_LAUNCH_AS("rmsnorm_residual_kernel<splitk>", kernel, dim3((unsigned)T), dim3(512), smem, st, (const __half*)nullptr,
                 (const __half*)residual, (const __half*)gamma, (__half*)normed_out, (__half*)residual_out, (int)H, eps, *h_parts);
  b200_count_launches(1);
  return B200_OK;
}

// y[T, N] fp16 = sum of the partials (+ bias): what the GEMM's own fix-up would have written
extern "C" int b200_splitk_reduce(const B200SplitK* parts, void* y, void* stream) {
  if (!splitk_ok(parts, "splitk_reduce") || !y) return B200_ERR_ARG;
  const int vec = parts->N / 4;
  B200_LAUNCH(splitk_reduce_kernel, dim3((unsigned)((vec + 255) / 256), (unsigned)parts->T), dim3(256), 0, (cudaStream_t)stream,
              (__half*)y, *parts);
  b200_count_launches(1);
  return B200_OK;
}

extern "C" int b200_splitk_silu_mul(const B200SplitK* parts, void* out, void* stream) {
  if (!splitk_ok(parts, "splitk_silu_mul") || !out) return B200_ERR_ARG;
  if (parts->half_tiles > 0 && parts->N != parts->half_tiles * 256) {
    b200_set_last_error("splitk_silu_mul: N does not match the gate|up layout");
    return B200_ERR_ARG;
  }
  const int vec = parts->N / 8;  // I / 4 output vectors per row
  B200_LAUNCH(splitk_silu_mul_kernel, dim3((unsigned)((vec + 255) / 256), (unsigned)parts->T), dim3(256), 0, (cudaStream_t)stream,
              (__half*)out, *parts);
  b200_count_launches(1);
  return B200_OK;
}

template <bool kSplitK>
static int rope_kv_write_impl(void* qkv, const void* cos, const void* sin, const int64_t* position_ids, const int64_t* slot_mapping,
                              void* k_pool, void* v_pool, int64_t T, int n_heads, int n_kv_heads, int head_dim, int rotary_dim,
                              const B200SplitK& parts, void* stream) {
  if (T == 0) return B200_OK;
  if (rotary_dim <= 0 || rotary_dim > head_dim || rotary_dim % 2 != 0) {
    b200_set_last_error("rope_kv_write_paged: rotary_dim must be even and in (0, head_dim]");
    return B200_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if constexpr (!kSplitK) {
    if (T >= 256 && rotary_dim == head_dim && (head_dim == 128 || head_dim == 64) && ((uintptr_t)qkv & 15) == 0 &&
        ((uintptr_t)cos & 15) == 0 && ((uintptr_t)sin & 15) == 0) {  // prefill-sized: the vectorised grid-stride form
      const int64_t n_rows = T * (n_heads + 2 * n_kv_heads);
      const int per_block = 256 / (head_dim / 16);
      int64_t blocks = (n_rows + per_block - 1) / per_block;
      if (blocks > 148 * 32) blocks = 148 * 32;
      if (head_dim == 128) {
        B200_LAUNCH(rope_kv_write_vec_kernel<128>, dim3((unsigned)blocks), dim3(256), 0, st, (__half*)qkv, (const __half*)cos,
                    (const __half*)sin, position_ids, slot_mapping, (__half*)k_pool, (__half*)v_pool, n_heads, n_kv_heads, n_rows);
      } else {
        B200_LAUNCH(rope_kv_write_vec_kernel<64>, dim3((unsigned)blocks), dim3(256), 0, st, (__half*)qkv, (const __half*)cos,
                    (const __half*)sin, position_ids, slot_mapping, (__half*)k_pool, (__half*)v_pool, n_heads, n_kv_heads, n_rows);
      }
      b200_count_launches(1);
      return B200_OK;
    }
  }
  dim3 grid((unsigned)T, n_heads + 2 * n_kv_heads);
  if (head_dim == 128) {
    constexpr auto kernel = rope_kv_write_kernel<128, kSplitK>;
    B200_LAUNCH_AS(kSplitK ? "rope_kv_write_kernel<splitk>" : "rope_kv_write_kernel", kernel, grid, dim3(64), 0, st, (__half*)qkv,
                   (const __half*)cos, (const __half*)sin, position_ids, slot_mapping, (__half*)k_pool, (__half*)v_pool, n_heads,
                   n_kv_heads, rotary_dim / 2, parts);
  } else if (head_dim == 64) {
    constexpr auto kernel = rope_kv_write_kernel<64, kSplitK>;
    B200_LAUNCH_AS(kSplitK ? "rope_kv_write_kernel<splitk>" : "rope_kv_write_kernel", kernel, grid, dim3(32), 0, st, (__half*)qkv,
                   (const __half*)cos, (const __half*)sin, position_ids, slot_mapping, (__half*)k_pool, (__half*)v_pool, n_heads,
                   n_kv_heads, rotary_dim / 2, parts);
  } else {
    b200_set_last_error("rope_kv_write_paged: head_dim must be 64 or 128");
    return B200_ERR_UNSUPPORTED;
  }
  b200_count_launches(1);
  return B200_OK;
}
extern "C" int b200_rope_kv_write_paged_ex(void* qkv, const void* cos, const void* sin, const int64_t* position_ids,
                                           const int64_t* slot_mapping, void* k_pool, void* v_pool, int64_t T, int n_heads,
                                           int n_kv_heads, int head_dim, int rotary_dim, void* stream) {
  return rope_kv_write_impl<false>(qkv, cos, sin, position_ids, slot_mapping, k_pool, v_pool, T, n_heads, n_kv_heads, head_dim,
                                   rotary_dim, B200SplitK{}, stream);
}
extern "C" int b200_rope_kv_write_paged_splitk(const B200SplitK* qkv_parts, void* qkv_out, const void* cos, const void* sin,
                                               const int64_t* position_ids, const int64_t* slot_mapping, void* k_pool, void* v_pool,
                                               int n_heads, int n_kv_heads, int head_dim, void* stream) {
  if (!splitk_ok(qkv_parts, "rope_kv_write_paged_splitk") || !qkv_out) return B200_ERR_ARG;
  if (qkv_parts->half_tiles != 0 || qkv_parts->N != (n_heads + 2 * n_kv_heads) * head_dim) {
    b200_set_last_error("rope_kv_write_paged_splitk: the partials are not a [T, (n_heads + 2 n_kv) d] projection in the plain layout");
    return B200_ERR_ARG;
  }
  return rope_kv_write_impl<true>(qkv_out, cos, sin, position_ids, slot_mapping, k_pool, v_pool, qkv_parts->T, n_heads, n_kv_heads,
                                  head_dim, head_dim, *qkv_parts, stream);
}
extern "C" int b200_rope_kv_write_paged(void* qkv, const void* cos, const void* sin, const int64_t* position_ids,
                                        const int64_t* slot_mapping, void* k_pool, void* v_pool, int64_t T, int n_heads,
                                        int n_kv_heads, int head_dim, void* stream) {
  return b200_rope_kv_write_paged_ex(qkv, cos, sin, position_ids, slot_mapping, k_pool, v_pool, T, n_heads, n_kv_heads, head_dim,
                                     head_dim, stream);
}

extern "C" int b200_layernorm_residual(const void* h, const void* residual, const void* gamma, const void* beta, void* normed_out,
                                       void* residual_out, int64_t T, int64_t H, float eps, void* stream) {
  if (T == 0) return B200_OK;
  if (H % 2 != 0 || H * 4 > 200 * 1024) { b200_set_last_error("layernorm_residual: need H % 2 == 0 and H <= 51200"); return B200_ERR_ARG; }
  constexpr int kThreads = 256;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(layernorm_residual_kernel<kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    configured = true;
  }
  B200_LAUNCH(layernorm_residual_kernel<kThreads>, dim3((unsigned)T), dim3(kThreads), (size_t)H * 4, (cudaStream_t)stream,
              (const __half*)h, (const __half*)residual, (const __half*)gamma, (const __half*)beta, (__half*)normed_out,
              (__half*)residual_out, (int)H, eps);
  b200_count_launches(1);
  return B200_OK;
}

extern "C" int b200_gelu(const void* x, void* out, int64_t n, int approximate_tanh, void* stream) {
  if (n == 0) return B200_OK;
  if (n % 2 != 0) { b200_set_last_error("gelu: n % 2 != 0"); return B200_ERR_ARG; }
  int64_t blocks = (n / 2 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  B200_LAUNCH(gelu_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const __half*)x, (__half*)out, n / 2,
              approximate_tanh);
  b200_count_launches(1);
  return B200_OK;
}

extern "C" int b200_silu_mul(const void* gate_up, void* out, int64_t T, int64_t I, void* stream) {
  if (T == 0) return B200_OK;
  if (I % 8 != 0) { b200_set_last_error("silu_mul: I % 8 != 0"); return B200_ERR_ARG; }
  const int64_t total8 = T * I / 8;
  int64_t blocks = (total8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  B200_LAUNCH(silu_mul_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const __half*)gate_up, (__half*)out, I, total8);
  b200_count_launches(1);
  return B200_OK;
}

extern "C" int b200_masked_softmax(const void* scores, const void* mask, void* out, int64_t rows, int64_t kv, int is_fp32, void* stream) {
  if (rows == 0 || kv == 0) return B200_OK;
  const unsigned blocks = (unsigned)((rows + 7) / 8);
  if (is_fp32) {
    B200_LAUNCH(masked_softmax_kernel<float>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (const float*)scores, (const uint8_t*)mask,
                (float*)out, rows, kv);
  } else {
    B200_LAUNCH(masked_softmax_kernel<__half>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (const __half*)scores,
                (const uint8_t*)mask, (__half*)out, rows, kv);
  }
  b200_count_launches(1);
  return B200_OK;
}

extern "C" int b200_permute_columns(const void* x, const int32_t* perm, void* out, int64_t T, int64_t K, void* stream) {
  if (T == 0 || K == 0) return B200_OK;
  if (K % 2 != 0) { b200_set_last_error("permute_columns: K % 2 != 0"); return B200_ERR_ARG; }
  B200_LAUNCH(permute_columns_kernel, dim3((unsigned)T), dim3(256), 0, (cudaStream_t)stream, (const __half*)x, perm, (__half*)out, K);
  b200_count_launches(1);
  return B200_OK;
}

extern "C" int b200_embedding(const void* table, const int64_t* ids, void* out, int64_t T, int64_t H, int64_t vocab_start,
                              int64_t vocab_rows, void* stream) {
  if (T == 0) return B200_OK;
  if (H % 8 != 0) { b200_set_last_error("embedding: H % 8 != 0"); return B200_ERR_ARG; }
  B200_LAUNCH(embedding_kernel, dim3((unsigned)T), dim3(128), 0, (cudaStream_t)stream, (const __half*)table, ids, (__half*)out, (int)H,
              vocab_start, vocab_rows);
  b200_count_launches(1);
  return B200_OK;
}

extern "C" int b200_argmax(const void* logits, int64_t* out_ids, int64_t B, int64_t V, int64_t ld, const int64_t* banned_ids,
                           void* stream) {
  if (B == 0) return B200_OK;
  if (ld < V) { b200_set_last_error("argmax: row stride smaller than the row"); return B200_ERR_ARG; }
  B200_LAUNCH(argmax_kernel, dim3((unsigned)B), dim3(kArgmaxThreads), 0, (cudaStream_t)stream, (const __half*)logits, out_ids, V, ld, banned_ids);
  b200_count_launches(1);
  return B200_OK;
}
