/* b200_tgis.h — C ABI of the B200-native TGIS decode hot path (libb200_tgis.so).
 *
 * Conventions (SURVEY.md §8b "Op / C-ABI"):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says "host"
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*): no host synchronisation, no
 *     allocation — outputs and workspaces are caller-provided; all entry points are CUDA-graph capturable
 *   - returns 0 (B200_OK) or a negative error code; b200_last_error() gives the message (thread-local)
 *   - fp16 tensors are row-major `__half`; token-major activations [T, features]
 * Each entry point names the reference interface it replaces (paths under
 * /root/reference/server/text_generation_server/).
 */
#ifndef B200_TGIS_H_
#define B200_TGIS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_ARG (-1)
#define B200_ERR_CUDA (-2)
#define B200_ERR_UNSUPPORTED (-3)
#define B200_ERR_NOMEM (-4)

#define B200_KV_PAGE_TOKENS 16 /* models/paged_causal_lm.py:308 */

int b200_abi_version(void);
const char* b200_last_error(void);
const char* b200_cuda_peek_error(void); /* debug: pending CUDA runtime error string, not cleared */
void b200_debug_w4_flags(int flags); /* debug timing experiments (results invalid): 1 no x loads, 2 no MMAs, 4 no weight loads, 8 no dequant math */
void b200_debug_w4_trace(void* device_buffer); /* debug: [n_ctas][64] u64 phase timestamps of int4 GEMM launches; NULL = off */
int b200_debug_gemm_plan(int kind, int64_t T, int64_t N, int64_t K, int sms, int32_t* out8); /* tests, host only: stream-K plan of a launch; kind 0 fp16, 1 int4; out = {token tile, k-blocks, feature (super-)tiles, token tiles, units per CTA, CTAs, contributor slots, tiles per unit} */
/* debug: per-kernel device timeline of a step.  While a buffer of `capacity` uint64 slots is attached, every kernel this
 * library enqueues is followed by a one-thread kernel storing %globaltimer (ns) into the next slot (slot 0: _begin);
 * _names writes the newline-separated kernel names in launch order and returns the count.  Serialises the stream. */
void b200_debug_step_trace(void* device_buffer /* NULL = off */, int capacity);
void b200_debug_step_trace_begin(void* stream);
int b200_debug_step_trace_names(char* out /* host */, int64_t capacity);
/* number of kernels this library has enqueued in this process (bench.py reports the delta as "gpu_launches") */
int64_t b200_launch_count(void);

/* ---- per-kernel device timing (bench.py roofline): CUDA events on the launching stream around every launch of one
 * kernel family while a handle is attached.  collect() synchronises and returns the launch count. */
#define B200_TIME_ATTN_DECODE 1
#define B200_TIME_GEMM_W4A16 2
#define B200_TIME_GEMM_F16 3
void* b200_timing_create(int max_launches);
void b200_timing_destroy(void* timing);
void b200_timing_attach(void* timing /* NULL detaches */, int which);
int b200_timing_collect(void* timing, float* total_ms /* host */);

/* ---- fused residual-add + RMSNorm ---------------------------------------------------------------------
 * replaces dropout_layer_norm.dropout_add_ln_fwd(h, residual, gamma, None x5, 0.0, eps, 1.0, 0, None, False, True)
 * (models/custom_modeling/flash_llama_modeling.py:132-148).  residual may be NULL (first layer, :149-150; then
 * residual_out is not written and the caller aliases it to h). */
int b200_rmsnorm_residual(const void* h, const void* residual, const void* gamma, void* normed_out, void* residual_out,
                          int64_t T, int64_t H, float eps, void* stream);

/* ---- RoPE + paged KV write ----------------------------------------------------------------------------
 * replaces rotary_emb.apply_rotary(x1, x2, cos, sin, x1, x2, False) on q and k (utils/layers.py:466-472,
 * flash_llama_modeling.py:262-263) and the KV append (flash_llama_modeling.py:268,282; fms-extras
 * reshape_and_cache via paged_llama_modeling.py:250).
 * qkv [T, (n_heads + 2 n_kv) * d] is rotated in place (q and k parts); k and v are scattered into the pools at
 * slot_mapping[t] = block * 16 + offset (negative = skip).  cos/sin: fp16 tables [max_pos, d/2] indexed by
 * position_ids (utils/layers.py:453-464).  Pool layout: [num_blocks][n_kv][16][d] fp16, 16-byte chunks XOR-swizzled
 * by (token & 7) (DESIGN.md "KV page layout"). */
int b200_rope_kv_write_paged(void* qkv, const void* cos, const void* sin, const int64_t* position_ids,
                             const int64_t* slot_mapping, void* k_pool, void* v_pool, int64_t T, int n_heads, int n_kv_heads,
                             int head_dim, void* stream);

/* same with a partial rotation (GPT-NeoX rotary_pct < 1): rotary_dim = 2 * cos.shape[-1] (utils/layers.py:467-469), tables
 * [max_pos, rotary_dim/2]; elements from rotary_dim on pass through. */
int b200_rope_kv_write_paged_ex(void* qkv, const void* cos, const void* sin, const int64_t* position_ids,
                                const int64_t* slot_mapping, void* k_pool, void* v_pool, int64_t T, int n_heads, int n_kv_heads,
                                int head_dim, int rotary_dim, void* stream);

/* ---- fused residual-add + LayerNorm: FastLayerNorm.forward (utils/layers.py:360-392 ->
 * dropout_layer_norm.dropout_add_ln_fwd(h, residual, gamma, beta, ..., eps, 1.0, 0, None, False, False)), the norm of the
 * GPT-NeoX family (flash_neox_modeling.py:202-281).  residual / beta / residual_out may be NULL. */
int b200_layernorm_residual(const void* h, const void* residual, const void* gamma, const void* beta, void* normed_out,
                            void* residual_out, int64_t T, int64_t H, float eps, void* stream);

/* ---- GELU on n fp16 values (FlashMLP.act, flash_neox_modeling.py:186-196); approximate_tanh for gelu_fast / gelu_pytorch_tanh */
int b200_gelu(const void* x, void* out, int64_t n, int approximate_tanh, void* stream);

/* ---- masked softmax over rows of attention scores [rows, kv] (fp16, or fp32 when is_fp32): replaces
 * forward_masked_softmax_kernel of server/custom_kernels/custom_kernels/fused_attention_cuda.cu:28-107 (and the identical
 * kernel of fused_bloom_attention_cuda.cu), the non-flash BLOOM / GPT-NeoX attention (bloom_modeling.py:394,
 * neox_modeling.py:214).  mask: bool [rows, kv], non-zero = excluded; fp32 softmax over the rest, excluded positions and
 * all-excluded rows give 0.  No kv_length limit (the reference kernel stops at 4096). */
int b200_masked_softmax(const void* scores, const void* mask, void* out, int64_t rows, int64_t kv, int is_fp32, void* stream);

/* ---- SiLU(gate) * up   (flash_llama_modeling.py:332-335); gate_up [T, 2, I] -> out [T, I] */
int b200_silu_mul(const void* gate_up, void* out, int64_t T, int64_t I, void* stream);

/* ---- vocabulary-parallel embedding gather (TensorParallelEmbedding.forward, utils/layers.py:346-357):
 * out[t] = table[ids[t] - vocab_start] if in [0, vocab_rows) else 0 */
int b200_embedding(const void* table, const int64_t* ids, void* out, int64_t T, int64_t H, int64_t vocab_start,
                   int64_t vocab_rows, void* stream);

/* ---- greedy arg-max over fp16 logits rows (Greedy, utils/tokens.py:44-46); ld = row stride in halves.
 * banned_ids (optional, [B]): token whose score counts as -inf for that row, or -1 — the min_new_tokens EOS mask
 * `scores[idx, eos] = -inf` (utils/tokens.py:244-246) folded into the arg-max. */
int b200_argmax(const void* logits, int64_t* out_ids, int64_t B, int64_t V, int64_t ld, const int64_t* banned_ids, void* stream);

/* ---- fused next-token chooser ------------------------------------------------------------------------------
 * One launch per step replaces HeterogeneousNextTokenChooser.__call__ (utils/tokens.py:238-271) for batches without
 * typical-p: min_new_tokens EOS mask, length penalty, repetition penalty (utils/logits_process.py:93-143), temperature,
 * top-k, top-p (:146-317), greedy arg-max or seeded sampling per row (tokens.py:30-78), and the chosen token's log-softmax
 * and rank over the warped scores (tokens.py:388-425).  Per-row arrays are device pointers, NULL = off for every row.
 * Deterministic (integer / fixed-point accumulation): tensor-parallel shards choose identically.  CUDA-graph capturable:
 * the sampling draw counters live on the device. */
typedef struct {
  const void* logits;      /* fp16 [B, ld] */
  void* warped_scratch;    /* fp16 [B, V] workspace; holds the warped scores afterwards */
  int64_t ld, V;           /* V <= 131072 */
  int32_t B;
  int32_t history_len_bias;      /* see position_ids */
  const float* temperature;      /* [B]; 0 = greedy row */
  const int32_t* top_k;          /* [B]; 0 = off */
  const float* top_p;            /* [B]; 1 = off */
  const float* rep_penalty;      /* [B]; 1 = off; needs history + position_ids */
  const int64_t* history;        /* [B, history_stride] token ids (FlashCausalLMBatch.all_input_ids_tensor) */
  int64_t history_stride;
  const int64_t* position_ids;   /* [B]: every row sees history[:, :max(position_ids) + history_len_bias], the reference's
                                    all_input_ids_tensor[:, :max_seqlen] (flash_causal_lm.py:525-527) */
  int64_t rep_exclude_id;        /* token exempt from the repetition penalty when B != 1 (pad == eos case, logits_process.py:98-118), or -1 */
  const int64_t* banned_ids;     /* [B]: id masked to -inf (min_new_tokens, tokens.py:242-246) or -1 */
  const float* length_penalty_factor; /* [B]: pow(decay, steps past start) - 1 for the EOS score (tokens.py:247-252), 0 = off */
  int64_t eos_id;
  const uint64_t* seeds;         /* [B] Philox key of sampling rows */
  int64_t* counters;             /* [B] draws made so far (advanced by the kernel) */
  int64_t* next_ids;             /* out [B] */
  float* logprobs;               /* out [B] or NULL */
  int32_t* ranks;                /* out [B] or NULL */
} B200ChooserParams;
int b200_choose_tokens(const B200ChooserParams* params /* host */, void* stream);

/* ---- decode attention over the paged KV pool ----------------------------------------------------------
 * replaces attention(q, layer_past[:,0], layer_past[:,1], cu_seqlens, max_s, scale, cu_seqlens_q, 1, False)
 * (utils/flash_attn.py:43-127 <- flash_llama_modeling.py:285-295) and fms-extras paged_attention
 * (paged_llama_modeling.py:267).  q: one token per sequence, head-major [n_heads][d] at q + b*q_token_stride (halves).
 * block_table[b][i] = pool block of the i-th 16-token page; context_lens[b] includes the token just written.
 * out[b] at out + b*out_token_stride.  workspace >= b200_attn_decode_workspace_bytes(...). */
int64_t b200_attn_decode_workspace_bytes(int B, int n_heads, int head_dim, int max_context_len);
int b200_attn_decode_paged(const void* q, int64_t q_token_stride, const void* k_pool, const void* v_pool,
                           const int32_t* block_table, int64_t block_table_stride, const int32_t* context_lens, void* out,
                           int64_t out_token_stride, void* workspace, int64_t workspace_bytes, int B, int n_heads,
                           int n_kv_heads, int head_dim, int max_context_len, float softmax_scale, void* stream);

/* ---- varlen prefill attention -------------------------------------------------------------------------
 * replaces attention(q, k, v, cu_seqlens, max_s, softmax_scale) (utils/flash_attn.py:43-127 <-
 * flash_llama_modeling.py:271-278; flash_attn_2_cuda.varlen_fwd).  Strides in halves between tokens. */
int b200_attn_prefill_varlen(const void* q, int64_t q_token_stride, const void* k, int64_t k_token_stride, const void* v,
                             int64_t v_token_stride, const int32_t* cu_seqlens, void* out, int64_t out_token_stride, int B,
                             int max_s, int n_heads, int n_kv_heads, int head_dim, float softmax_scale, int causal, void* stream);

/* ---- prefill / chunked-prefill attention over the paged pool (tcgen05 + TMA) ----------------------------
 * Same reference op as b200_attn_prefill_varlen, but K and V come from the block pool, which the step's own tokens have
 * already been appended to (b200_rope_kv_write_paged): the queries of a sequence may follow a cached context
 * (paged_llama_modeling.py:253-264 attends a prefill over the fresh K/V only; an add-on prefill after a prompt prefix, a chunked
 * prefill or a mixed prefill + decode step need this form).  Sequence b owns query tokens cu_seqlens_q[b] .. cu_seqlens_q[b+1]
 * of q (T_q tokens, head-major [n_heads][d] at q + t * q_token_stride); context_lens[b] counts its tokens in the pool including
 * them; query i sits at absolute position context_lens[b] - n_q(b) + i and attends keys 0 .. that position.  num_blocks = blocks
 * of the pool (its first dimension); max_q = the largest n_q. */
int b200_attn_prefill_paged(const void* q, int64_t q_token_stride, int64_t T_q, const void* k_pool, const void* v_pool,
                            int64_t num_blocks, const int32_t* block_table, int64_t block_table_stride, const int32_t* context_lens,
                            const int32_t* cu_seqlens_q, void* out, int64_t out_token_stride, int B, int max_q, int n_heads,
                            int n_kv_heads, int head_dim, float softmax_scale, void* stream);

/* ---- linears ------------------------------------------------------------------------------------------
 * workspace: >= b200_gemm_workspace_bytes(T, N, K) bytes; its first 64 KiB must have been zeroed once (split-K tile
 * counters; the kernels re-arm them).  NULL disables split-K. */
int64_t b200_gemm_workspace_bytes(int64_t T, int64_t N, int64_t K);
int64_t b200_gemm_workspace_bytes_max(int64_t N, int64_t K); /* max over all T: size a persistent workspace with this */

/* y[T,N] = x[T,K] . w[N,K]^T (+ bias[N]);  replaces F.linear in FastLinear.forward (utils/layers.py:110-111) and
 * torch.mm in TensorParallelHead.forward (:257-262). */
int b200_gemm_f16(const void* x, const void* w, const void* bias, void* y, int64_t T, int64_t N, int64_t K, void* workspace,
                  void* stream);

/* one-time conversion of the GPTQ checkpoint tensors of one linear into the kernel's streaming layout ("unit records":
 * one contiguous 8 KB + meta block per (128-feature tile, 128-wide k-block)); replaces
 * exllamav2_kernels.make_q_matrix (utils/gptq/exllamav2.py:23-62), `packed` plays the role of its q_handle.
 * qweight int32 [K/8, N], qzeros int32 [ceil(K/g), N/8], scales fp16 [ceil(K/g), N] in checkpoint layout (read only; the
 * caller may free them afterwards).  groupsize: 32, 64, a multiple of 128, or <= 0 for one group; groups are
 * k // groupsize (trivial g_idx; act-order is rejected by the host wrapper).  K % 32 == 0, N % 32 == 0
 * (exllamav2.py:118-119).  packed: b200_gptq_packed_bytes(K, N, groupsize) bytes, 16-byte aligned. */
int64_t b200_gptq_packed_bytes(int64_t K, int64_t N, int groupsize);
int b200_gptq_pack(const void* qweight, const void* qzeros, const void* scales, void* packed, int64_t K, int64_t N,
                   int groupsize, void* stream);
/* layout 0 = b200_gptq_pack.  layout 1 ("gate|up", N = 2 I with I % 128 == 0): the records of gate feature tile s and of
 * up feature tile s are adjacent, for the fused [gate; up] projection of LlamaMLP (flash_llama_modeling.py:315-335). */
#define B200_W4_LAYOUT_PLAIN 0
#define B200_W4_LAYOUT_GATE_UP 1
/* row_perm (int32 [K], device) or NULL: act-order checkpoints (non-trivial g_idx; exllamav2.py:31-48 q_perm): a stable
 * arg-sort of g_idx.  Packed row k' = checkpoint row row_perm[k']; feed the GEMM x[:, row_perm] (b200_permute_columns). */
int b200_gptq_pack_ex(const void* qweight, const void* qzeros, const void* scales, const int32_t* row_perm, void* packed,
                      int64_t K, int64_t N, int groupsize, int layout, void* stream);
/* out[t, k'] = x[t, perm[k']], fp16 [T, K] */
int b200_permute_columns(const void* x, const int32_t* perm, void* out, int64_t T, int64_t K, void* stream);

/* y[T,N] = x[T,K] . dequant(packed) (+ bias);  replaces exllamav2_kernels.gemm_half_q_half
 * (utils/gptq/exllamav2.py:14-20) for every T (no dequant-to-scratch branch, cf. :87). */
int b200_gemm_w4a16(const void* x, const void* packed, const void* bias, void* y, int64_t T, int64_t N, int64_t K,
                    int groupsize, void* workspace, void* stream);
/* same for a weight packed with `layout`; act = 1 (layout 1 only) additionally replaces `act(gu[:, 0]) * gu[:, 1]`
 * (flash_llama_modeling.py:332-335, with b200_silu_mul's arithmetic): y is [T, N/2] = SiLU(x Wgate) * (x Wup). */
int b200_gemm_w4a16_ex(const void* x, const void* packed, const void* bias, void* y, int64_t T, int64_t N, int64_t K,
                       int groupsize, int layout, int act, void* workspace, void* stream);

/* ---- deferred split-K reduction ------------------------------------------------------------------------
 * At decode sizes every SM streams a K-slice of the weights, so an output tile is the sum of several CTAs' fp32 partials.
 * Instead of finishing that sum inside the GEMM (a grid-wide wait on the slowest contributor + a second pass, 5-8 us per
 * launch), the *_deferred GEMMs leave the partials in the workspace and describe them with a B200SplitK; the consumer that
 * follows in the stream anyway (residual + RMSNorm, RoPE + KV write, SiLU * up, the tensor-parallel all-reduce) adds them up
 * in contributor order as it reads its input - bit-identical to the in-GEMM fix-up.  A B200SplitK is valid until the next GEMM
 * that uses the same workspace.  Element (t, n), plain layout: tile = n / 128, unit u = tile / tiles_per_unit,
 * r = tile % tiles_per_unit, contributors c = 0 .. c_last - c_first with c_first = u nkb / units_per_cta,
 * c_last = ((u + 1) nkb - 1) / units_per_cta:  partial[(((u max_contrib + c) tiles_per_unit + r) tn + t) 128 + n % 128].
 * gate|up layout (half_tiles > 0): unit u holds tiles (u, u + half_tiles) as r = 0, 1. */
typedef struct {
  const float* partial;
  const void* bias; /* fp16 [N] or NULL: added to the sum by the consumer */
  int32_t tiles_per_unit, tn, nkb, units_per_cta, max_contrib, half_tiles, N, T;
} B200SplitK;
int b200_gemm_w4a16_deferred(const void* x, const void* packed, const void* bias, int64_t T, int64_t N, int64_t K, int groupsize,
                             int layout, void* workspace, B200SplitK* splitk /* host, out */, void* stream); /* T <= 128 */
int b200_gemm_f16_deferred(const void* x, const void* w, const void* bias, int64_t T, int64_t N, int64_t K, void* workspace,
                           B200SplitK* splitk /* host, out */, void* stream); /* T <= 256 */
/* consumers.  y[T, N] fp16 = the GEMM's output (plain layout) */
int b200_splitk_reduce(const B200SplitK* parts, void* y, void* stream);
/* b200_rmsnorm_residual with h = the deferred GEMM output [T, H] (o_proj / down_proj, H = parts->N) */
int b200_rmsnorm_residual_splitk(const B200SplitK* h_parts, const void* residual, const void* gamma, void* normed_out,
                                 void* residual_out, float eps, void* stream);
/* b200_rope_kv_write_paged with qkv = the deferred fused QKV projection; qkv_out [T, (n_heads + 2 n_kv) d] receives the final
 * activation (q and k rotated) */
int b200_rope_kv_write_paged_splitk(const B200SplitK* qkv_parts, void* qkv_out, const void* cos, const void* sin,
                                    const int64_t* position_ids, const int64_t* slot_mapping, void* k_pool, void* v_pool,
                                    int n_heads, int n_kv_heads, int head_dim, void* stream);
/* out [T, N/2] = SiLU(gate) * up of a deferred fused [gate; up] projection (either weight layout), b200_silu_mul's arithmetic */
int b200_splitk_silu_mul(const B200SplitK* gate_up_parts, void* out, void* stream);

/* ---- paged KV block allocator (host) + per-step bookkeeping (device) -------------------------------------
 * replaces fms-extras PagedKVCacheManager block bookkeeping (models/paged_causal_lm.py:338-353,
 * utils/paged.py:92-134; block size 16).  Block ids index the pools' first dimension. */
void* b200_kv_alloc_create(int32_t num_blocks);
void b200_kv_alloc_destroy(void* allocator);
int32_t b200_kv_alloc_num_free(void* allocator);
int b200_kv_alloc_take(void* allocator, int32_t n, int32_t* out_ids /* host */);
int b200_kv_alloc_release(void* allocator, const int32_t* ids /* host */, int32_t n);
/* start of a decode step, per sequence b: pos = context_lens[b]; position_ids[b] = pos;
 * slot_mapping[b] = block_table[b][pos/16]*16 + pos%16; context_lens[b] = pos+1; input_ids[b] = next_ids[b]
 * (next_ids/input_ids may be NULL).  context_lens[b] < 0 marks a padding row (slot -1). */
int b200_decode_advance(const int32_t* block_table, int64_t block_table_stride, int32_t* context_lens, int64_t* position_ids,
                        int64_t* slot_mapping, const int64_t* next_ids, int64_t* input_ids, int B, void* stream);

/* ---- FlashLlama step runtime ---------------------------------------------------------------------------
 * One call enqueues a whole prefill / decode step of the Llama graph; replaces the Python op sequence of
 * FlashLlamaForCausalLM.forward (models/custom_modeling/flash_llama_modeling.py:425-540).  All pointers device
 * memory owned by the caller (weights: the model; scratch: the host runtime). */
/* ---- one-shot all-reduce over NVLink peer memory (replaces torch.distributed.all_reduce at the tensor-parallel
 * layer boundary, utils/layers.py:303-306, :343-345, for decode-sized messages).  create -> exchange the IPC handles of all
 * ranks (any transport) -> connect -> allreduce (CUDA-graph capturable; every rank must issue the same calls). */
int b200_p2p_handle_bytes(void);
int b200_p2p_create(int64_t max_bytes, int world, int rank, void** ctx_out, void* handle_out);
int b200_p2p_connect(void* ctx, const void* handles /* [world][b200_p2p_handle_bytes()] */);
int64_t b200_p2p_max_bytes(void* ctx);
int b200_p2p_allreduce_f16(void* ctx, void* data /* fp16 [n], in place, sums in rank order with fp32 accumulation */, int64_t n, void* stream);
void b200_p2p_destroy(void* ctx);
/* A window serves ONE of the three kernels (allreduce_f16 / allreduce_rmsnorm / argmax): their bookkeeping is per block index.
 * Fused layer boundary: all-reduce of this rank's partial hidden state (h fp16 [T, H], or h_parts = a deferred row-parallel GEMM)
 * + residual add + RMSNorm (utils/layers.py:318-322 followed by flash_llama_modeling.py:132-148).  residual NULL (first layer):
 * residual_out = the reduced hidden state.  Row t belongs to rank t % world: every rank receives every row of normed_out, but
 * residual is read and residual_out written only for the rows the rank owns (the residual stream of a row lives on its owner;
 * route every boundary of a step through this call).  T <= 2048; the window must have been created with
 * max_bytes >= 2048 * H * 2. */
int b200_p2p_allreduce_rmsnorm(void* ctx, const void* h, const B200SplitK* h_parts, const void* residual, const void* gamma,
                               void* normed_out, void* residual_out, int64_t T, int64_t H, float eps, void* stream);
/* greedy ids of a vocabulary-sharded head without gathering logits (replaces utils/layers.py:249-269 + utils/tokens.py:44-46 for
 * all-greedy batches): local arg-max, (value, index) pairs over NVLink, ties to the lowest global id.  B <= 256. */
int b200_p2p_argmax(void* ctx, const void* logits, int64_t* out_ids, int64_t B, int64_t V_local, int64_t ld,
                    const int64_t* banned_ids /* global ids or -1 */, void* stream);

typedef struct {
  const void* weight;  /* fp16 [N, K], or NULL when GPTQ */
  const void* qweight; /* GPTQ: the b200_gptq_pack output for this linear, or NULL */
  const void* perm;    /* GPTQ act-order: int32 [K] row permutation the weight was packed with (x is gathered by it), or NULL */
  const void* _unused; /* reserved */
  const void* bias;    /* fp16 [N] or NULL */
  int64_t N, K;
  int32_t groupsize;
  int32_t layout; /* B200_W4_LAYOUT_* of `qweight` */
} B200Linear;

typedef struct {
  const void* input_ln; /* fp16 [H] */
  const void* post_ln;
  B200Linear qkv;     /* column-parallel: [ (h + 2 h_kv) d / tp , H ]   (flash_llama_modeling.py:220-231) */
  B200Linear o;       /* row-parallel:    [ H, h d / tp ] */
  B200Linear gate_up; /* column-parallel: [ 2 I / tp, H ], gate rows then up rows (:315-321) */
  B200Linear down;    /* row-parallel:    [ H, I / tp ] */
} B200LlamaLayer;

typedef struct {
  int32_t n_layers, hidden_size, n_heads, n_kv_heads, head_dim; /* heads are per-rank counts */
  int32_t tp_size, tp_rank, _pad;
  float rms_eps, softmax_scale;
  const B200LlamaLayer* layers; /* host array [n_layers] */
  const void* embed;            /* fp16 [vocab_rows, H]: this rank's rows of the vocab-parallel table */
  int64_t vocab_start, vocab_rows;
  const void* final_norm;
  const void* lm_head; /* fp16 [vocab_rows_head, H] */
  int64_t vocab_rows_head;
  const void* rope_cos; /* fp16 [max_pos, d/2] (utils/layers.py:436-451) */
  const void* rope_sin;
} B200LlamaWeights;

typedef struct {
  int64_t T;          /* tokens in this step (decode: == B) */
  int32_t B;          /* sequences */
  int32_t is_prefill; /* 1: varlen causal attention over cu_seqlens; 0: paged decode attention */
  int32_t max_s;      /* max sequence length in the batch (decode: max context incl. the new token) */
  int32_t _pad;
  const int64_t* input_ids;    /* [T] */
  const int64_t* position_ids; /* [T] */
  const int64_t* slot_mapping; /* [T] block*16+offset, <0 = padding row */
  const int32_t* cu_seqlens;   /* [B+1] (prefill) */
  const int32_t* block_table;  /* [B, block_table_stride] (decode) */
  int64_t block_table_stride;
  const int32_t* context_lens; /* [B] (decode) */
  void* kv_pool;               /* base of [n_layers][2][num_blocks][n_kv][16][d] fp16 */
  int64_t kv_layer_stride_bytes, kv_v_offset_bytes;
  /* scratch, fp16 unless noted */
  void* hidden;   /* [T, H] in: embeddings (or caller-provided inputs_embeds); out: last block output */
  void* residual; /* [T, H] */
  void* normed;   /* [T, H] */
  void* qkv;      /* [T, (h + 2 h_kv) d] */
  void* attn_out; /* [T, h d] */
  void* gate_up;  /* [T, 2 I / tp] */
  void* act;      /* [T, I / tp] */
  void* perm_x;   /* [T, max K] scratch for the gathered activations of act-order GPTQ linears, or NULL if there are none */
  void* attn_ws;  /* b200_attn_decode_workspace_bytes */
  int64_t attn_ws_bytes;
  void* gemm_ws; /* b200_gemm_workspace_bytes (max over the model's linears), first 64 KiB zeroed once */
  /* head */
  const int64_t* head_rows; /* optional [n_head_rows] token rows to project (lm_head_indices); NULL = all T */
  int64_t n_head_rows;
  void* head_in;     /* [n_head_rows, H] scratch when head_rows != NULL */
  void* logits;      /* [rows, vocab_rows_head] fp16 */
  int64_t* next_ids; /* optional [rows] greedy ids (sharded head: needs p2p_argmax) */
  const int64_t* banned_ids; /* optional [rows], see b200_argmax */
  /* decode-sized steps (T <= 128): */
  int32_t defer_splitk; /* 1: linears leave split-K partials to their consumer kernels (B200SplitK) */
  int32_t _pad2;
  void* p2p_norm;   /* tensor parallel: b200_p2p_create window for b200_p2p_allreduce_rmsnorm (the layer boundary runs inside the
                       step, no host-side collective), or NULL: the caller all-reduces `hidden` between the block calls */
  void* p2p_argmax; /* tensor parallel: window for b200_p2p_argmax (sharded head -> next_ids), or NULL */
  /* prefill steps: */
  int64_t kv_num_blocks; /* blocks of the pool: > 0 lets a prefill step attend through the pool (b200_attn_prefill_paged; needs
                            block_table + context_lens, context_lens counting the step's tokens), 0 = b200_attn_prefill_varlen */
  int32_t max_q;         /* largest number of query tokens of a sequence in this step (0: max_s) */
  int32_t _pad3;
} B200LlamaStep;

int b200_llama_embed(const B200LlamaWeights* w, const B200LlamaStep* s, void* stream);
int b200_llama_attn_block(const B200LlamaWeights* w, const B200LlamaStep* s, int layer, void* stream);
int b200_llama_mlp_block(const B200LlamaWeights* w, const B200LlamaStep* s, int layer, void* stream);
int b200_llama_head(const B200LlamaWeights* w, const B200LlamaStep* s, void* stream);
int b200_llama_step(const B200LlamaWeights* w, const B200LlamaStep* s, void* stream); /* tp_size == 1, or p2p_norm set */

#ifdef __cplusplus
}
#endif
#endif /* B200_TGIS_H_ */
