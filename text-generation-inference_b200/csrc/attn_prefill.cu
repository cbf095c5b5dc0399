// Varlen causal prefill attention over the fresh (un-paged) q/k/v of the step.
//
// Replaces: utils/flash_attn.py:43-127 `attention(q, k, v, cu_seqlens, max_s, softmax_scale)` in its prefill form
// (flash_llama_modeling.py:271-278; flash_attn_2_cuda.varlen_fwd, causal) of /root/reference/server/text_generation_server.
//
// 64 query rows per CTA (4 warps x 16 rows), 64-key tiles double-buffered with cp.async into XOR-swizzled shared
// memory, m16n8k16 HMMA with fp32 accumulators, fp32 online softmax, P rounded to fp16 before P.V (flash-attn semantics).
// Round-1 kernel: correct and tensor-core based; the tcgen05/TMEM version is the planned successor.
#include "common.cuh"

namespace b200 {

constexpr int kBM = 64, kBN = 64;

template <int D>
__device__ __forceinline__ uint32_t sw_off(int row, int chunk) {
  return row * (D * 2) + ((chunk ^ (row & 7)) << 4);
}

template <int D>
__device__ __forceinline__ void load_tile_async(unsigned char* dst, const __half* src, int64_t token_stride, int row0, int n_rows_valid) {
  // 64 rows x D halves; 128 threads; chunk = 16 B
  constexpr int kChunksPerRow = D / 8;
  for (int idx = threadIdx.x; idx < kBN * kChunksPerRow; idx += 128) {
    const int r = idx / kChunksPerRow, c = idx % kChunksPerRow;
    const bool ok = r < n_rows_valid;
    const __half* p = src + (int64_t)(row0 + (ok ? r : 0)) * token_stride + c * 8;
    cp_async_16(dst + sw_off<D>(r, c), p, ok);
  }
}

template <int D>
__global__ void __launch_bounds__(128)
attn_prefill_varlen_kernel(const __half* __restrict__ q, int64_t q_stride, const __half* __restrict__ k, int64_t k_stride,
                           const __half* __restrict__ v, int64_t v_stride, const int32_t* __restrict__ cu_seqlens,
                           __half* __restrict__ out, int64_t out_stride, int n_heads, int n_kv, float scale_log2, int causal) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int kTileBytes = kBN * D * 2;
  unsigned char* sQ = smem;
  unsigned char* sK = smem + kTileBytes;       // 2 buffers
  unsigned char* sV = sK + 2 * kTileBytes;     // 2 buffers

  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.z, head = blockIdx.y;
  const int seq0 = cu_seqlens[b], L = cu_seqlens[b + 1] - seq0;
  const int q0 = blockIdx.x * kBM;
  if (q0 >= L) return;
  const int hk = head / (n_heads / n_kv);
  const int warp = warp_id(), lane = lane_id(), g = lane >> 2, tig = lane & 3, mi = lane >> 3, r8 = lane & 7;
  const int nq = min(kBM, L - q0);
  const int k_end = causal ? min(L, q0 + kBM) : L;
  const int n_kt = (k_end + kBN - 1) / kBN;

  const __half* qp = q + (int64_t)seq0 * q_stride + head * D;
  const __half* kp = k + (int64_t)seq0 * k_stride + hk * D;
  const __half* vp = v + (int64_t)seq0 * v_stride + hk * D;

  load_tile_async<D>(sQ, qp, q_stride, q0, nq);
  load_tile_async<D>(sK, kp, k_stride, 0, min(kBN, k_end));
  load_tile_async<D>(sV, vp, v_stride, 0, min(kBN, k_end));
  cp_async_commit();

  uint32_t qf[D / 16][4];
  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -1.0e30f, m1 = -1.0e30f, l0 = 0.f, l1 = 0.f;

  for (int kt = 0; kt < n_kt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < n_kt) {
      const int nk = min(kBN, k_end - (kt + 1) * kBN);
      load_tile_async<D>(sK + (buf ^ 1) * kTileBytes, kp, k_stride, (kt + 1) * kBN, nk);
      load_tile_async<D>(sV + (buf ^ 1) * kTileBytes, vp, v_stride, (kt + 1) * kBN, nk);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kt == 0) {
      // Q fragments via ldmatrix: matrices (rows 0-7,c), (rows 8-15,c), (rows 0-7,c+1), (rows 8-15,c+1)
      const uint32_t qb = smem_u32(sQ);
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks) {
        const int row = warp * 16 + (mi & 1) * 8 + r8;
        ldmatrix_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], qb + sw_off<D>(row, ks * 2 + (mi >> 1)));
      }
    }
    const uint32_t kb = smem_u32(sK + buf * kTileBytes), vb = smem_u32(sV + buf * kTileBytes);
    float sc[kBN / 8][4];
#pragma unroll
    for (int nt = 0; nt < kBN / 8; ++nt) {
      sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
      const int row = nt * 8 + r8;
#pragma unroll
      for (int kc = 0; kc < D / 32; ++kc) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(b0, b1, b2, b3, kb + sw_off<D>(row, kc * 4 + mi));
        mma_m16n8k16_f16f32(sc[nt], qf[kc * 2], b0, b1);
        mma_m16n8k16_f16f32(sc[nt], qf[kc * 2 + 1], b2, b3);
      }
    }
    // masking: key index j valid if j < k_end and (!causal or j <= query index)
    const int qi0 = q0 + warp * 16 + g, qi1 = qi0 + 8;
    float mx0 = -1.0e30f, mx1 = -1.0e30f;
#pragma unroll
    for (int nt = 0; nt < kBN / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = kt * kBN + nt * 8 + tig * 2 + e;
        const bool in = j < k_end;
        const bool ok0 = in && (!causal || j <= qi0);
        const bool ok1 = in && (!causal || j <= qi1);
        sc[nt][e] = ok0 ? sc[nt][e] * scale_log2 : -INFINITY;
        sc[nt][2 + e] = ok1 ? sc[nt][2 + e] * scale_log2 : -INFINITY;
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
    for (int nt = 0; nt < kBN / 8; ++nt) {
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
#pragma unroll
    for (int i = 0; i < D / 8; ++i) { o[i][0] *= a0; o[i][1] *= a0; o[i][2] *= a1; o[i][3] *= a1; }
#pragma unroll
    for (int kk = 0; kk < kBN / 16; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_half2(sc[2 * kk][0], sc[2 * kk][1]);
      pa[1] = pack_half2(sc[2 * kk][2], sc[2 * kk][3]);
      pa[2] = pack_half2(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
      pa[3] = pack_half2(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
#pragma unroll
      for (int dc = 0; dc < D / 16; ++dc) {
        uint32_t v0, v1, v2, v3;
        const int row = kk * 16 + (mi & 1) * 8 + r8;
        ldmatrix_x4_trans(v0, v1, v2, v3, vb + sw_off<D>(row, 2 * dc + (mi >> 1)));
        mma_m16n8k16_f16f32(o[2 * dc], pa, v0, v1);
        mma_m16n8k16_f16f32(o[2 * dc + 1], pa, v2, v3);
      }
    }
    __syncthreads();  // buffer `buf` is refilled by the next iteration's prefetch
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
  const int r0 = warp * 16 + g, r1 = r0 + 8;
  __half* op = out + (int64_t)seq0 * out_stride + head * D;
#pragma unroll
  for (int nt = 0; nt < D / 8; ++nt) {
    const int col = nt * 8 + tig * 2;
    if (r0 < nq) *reinterpret_cast<__half2*>(op + (int64_t)(q0 + r0) * out_stride + col) = __floats2half2_rn(o[nt][0] * inv0, o[nt][1] * inv0);
    if (r1 < nq) *reinterpret_cast<__half2*>(op + (int64_t)(q0 + r1) * out_stride + col) = __floats2half2_rn(o[nt][2] * inv1, o[nt][3] * inv1);
  }
}

}  // namespace b200

using namespace b200;

template <int D>
static int launch_prefill(const void* q, int64_t qs, const void* k, int64_t ks, const void* v, int64_t vs, const int32_t* cu,
                          void* out, int64_t os, int B, int max_s, int n_heads, int n_kv, float scale, int causal, cudaStream_t st) {
  constexpr int kSmem = 5 * kBN * D * 2;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_prefill_varlen_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) { b200_set_last_error(cudaGetErrorString(e)); return B200_ERR_CUDA; }
    configured = true;
  }
  dim3 grid((max_s + kBM - 1) / kBM, n_heads, B);
  B200_LAUNCH(attn_prefill_varlen_kernel<D>, grid, dim3(128), (size_t)kSmem, st, (const __half*)q, qs, (const __half*)k, ks, (const __half*)v,
              vs, cu, (__half*)out, os, n_heads, n_kv, scale * 1.4426950408889634f, causal);
  b200_count_launches(1);
  return B200_OK;
}

extern "C" int b200_attn_prefill_varlen(const void* q, int64_t q_token_stride, const void* k, int64_t k_token_stride, const void* v,
                                        int64_t v_token_stride, const int32_t* cu_seqlens, void* out, int64_t out_token_stride,
                                        int B, int max_s, int n_heads, int n_kv_heads, int head_dim, float softmax_scale,
                                        int causal, void* stream) {
  if (B == 0 || max_s == 0) return B200_OK;
  if (n_kv_heads <= 0 || n_heads % n_kv_heads != 0) { b200_set_last_error("attn_prefill_varlen: bad head counts"); return B200_ERR_ARG; }
  if ((q_token_stride | k_token_stride | v_token_stride | out_token_stride) & 7) {
    b200_set_last_error("attn_prefill_varlen: token strides must be multiples of 8 halves (16 B)");
    return B200_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (head_dim == 128)
    return launch_prefill<128>(q, q_token_stride, k, k_token_stride, v, v_token_stride, cu_seqlens, out, out_token_stride, B, max_s,
                               n_heads, n_kv_heads, softmax_scale, causal, st);
  if (head_dim == 64)
    return launch_prefill<64>(q, q_token_stride, k, k_token_stride, v, v_token_stride, cu_seqlens, out, out_token_stride, B, max_s,
                              n_heads, n_kv_heads, softmax_scale, causal, st);
  b200_set_last_error("attn_prefill_varlen: head_dim must be 64 or 128");
  return B200_ERR_UNSUPPORTED;
}
