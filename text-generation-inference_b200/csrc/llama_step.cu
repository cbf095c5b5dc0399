// C++ step runtime for the FlashLlama graph: one C call enqueues a whole decode (or prefill) step on a stream, so the
// Python host pays one FFI call per step and the sequence is CUDA-graph capturable.
//
// Restates the op order of FlashLlamaModel.forward / FlashLlamaLayer.forward / FlashLlamaAttention.forward / LlamaMLP.forward
// (/root/reference/server/text_generation_server/models/custom_modeling/flash_llama_modeling.py:425-497, 356-389, 240-297,
// 332-335) and FlashLlamaForCausalLM.forward (:514-540) over the kernels of this library.
//
// Two things move between kernels at decode sizes (T <= 128), without changing any result bit:
//   * deferred split-K (s->defer_splitk): a linear leaves its K-slices' fp32 partials in the workspace and the kernel that
//     consumes its output anyway (residual + RMSNorm, RoPE + KV write, SiLU * up) sums them while it reads - the GEMM no longer
//     waits for its slowest CTA and re-reads the partials itself (include/b200_tgis.h "B200SplitK");
//   * the tensor-parallel layer boundary (s->p2p_norm): all-reduce over NVLink peer memory + residual + RMSNorm in one
//     kernel (b200_p2p_allreduce_rmsnorm), fed directly by the row-parallel linear's partials.  The whole sharded step is then
//     one C call like the single-rank one; without a window the caller all-reduces `hidden` (NCCL) between the block calls.
#include "common.cuh"
#include "../../include/b200_tgis.h"

// act-order GPTQ: the weight rows were packed in group order, the activations follow (exllamav2 q_perm)
static int gathered(const B200Linear* L, const B200LlamaStep* s, const void** x, int64_t T, void* stream) {
  if (!L->qweight || !L->perm) return B200_OK;
  if (!s->perm_x) { b200_set_last_error("llama_step: act-order linear but no perm_x scratch"); return B200_ERR_ARG; }
  const int st = b200_permute_columns(*x, (const int32_t*)L->perm, s->perm_x, T, L->K, stream);
  *x = s->perm_x;
  return st;
}

static int linear(const B200Linear* L, const B200LlamaStep* s, const void* x, void* y, int64_t T, void* stream) {
  if (L->qweight) {
    const int st = gathered(L, s, &x, T, stream);
    if (st != B200_OK) return st;
    return b200_gemm_w4a16_ex(x, L->qweight, L->bias, y, T, L->N, L->K, L->groupsize, L->layout, 0, s->gemm_ws, stream);
  }
  return b200_gemm_f16(x, L->weight, L->bias, y, T, L->N, L->K, s->gemm_ws, stream);
}

// same linear with its split-K reduction left to the consumer of *parts (the next kernel enqueued on the stream)
static int linear_deferred(const B200Linear* L, const B200LlamaStep* s, const void* x, B200SplitK* parts, int64_t T, void* stream) {
  if (L->qweight) {
    const int st = gathered(L, s, &x, T, stream);
    if (st != B200_OK) return st;
    return b200_gemm_w4a16_deferred(x, L->qweight, L->bias, T, L->N, L->K, L->groupsize, L->layout, s->gemm_ws, parts, stream);
  }
  return b200_gemm_f16_deferred(x, L->weight, L->bias, T, L->N, L->K, s->gemm_ws, parts, stream);
}

static bool deferring(const B200LlamaStep* s) { return s->defer_splitk && s->gemm_ws && s->T <= 128 && !s->is_prefill; }
// the fused NVLink boundary takes steps of up to 2048 token rows (decode, and prefill chunks of continuous batching); anything
// larger is all-reduced by the caller between the block calls
static bool fused_boundary(const B200LlamaWeights* w, const B200LlamaStep* s) { return w->tp_size > 1 && s->p2p_norm && s->T <= 2048; }

#define RUN(expr)            \
  do {                       \
    int _s = (expr);         \
    if (_s != B200_OK) return _s; \
  } while (0)

// The hidden state entering a norm: the fp16 tensor s->hidden, or the deferred partials of the row-parallel linear that produced it.
struct HiddenIn {
  bool deferred = false;
  B200SplitK parts = {};
};

// (all-reduce +) residual add + RMSNorm: hidden (+ residual) -> normed, residual.  `first`: residual = None (flash_llama_modeling.py:149-150).
static int boundary_norm(const B200LlamaWeights* w, const B200LlamaStep* s, const HiddenIn& in, const void* gamma, bool first, void* stream) {
  const int64_t T = s->T, H = w->hidden_size;
  if (fused_boundary(w, s))
    return b200_p2p_allreduce_rmsnorm(s->p2p_norm, in.deferred ? nullptr : s->hidden, in.deferred ? &in.parts : nullptr,
                                      first ? nullptr : s->residual, gamma, s->normed, s->residual, T, H, w->rms_eps, stream);
  if (in.deferred) {
    if (w->tp_size > 1) { b200_set_last_error("llama_step: deferred row-parallel linear without the fused boundary"); return B200_ERR_ARG; }
    return b200_rmsnorm_residual_splitk(&in.parts, s->residual, gamma, s->normed, s->residual, w->rms_eps, stream);
  }
  if (first) {
    RUN(b200_rmsnorm_residual(s->hidden, nullptr, gamma, s->normed, nullptr, T, H, w->rms_eps, stream));
    if (cudaMemcpyAsync(s->residual, s->hidden, (size_t)T * H * 2, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) != cudaSuccess) {
      b200_set_last_error("llama_step: residual copy failed");
      return B200_ERR_CUDA;
    }
    return B200_OK;
  }
  return b200_rmsnorm_residual(s->hidden, s->residual, gamma, s->normed, s->residual, T, H, w->rms_eps, stream);
}

// A row-parallel linear (o_proj, down_proj): its output is the next boundary's input.  Deferred when that boundary can sum the
// partials (single rank, or the fused NVLink boundary); materialised into s->hidden otherwise (the caller's NCCL all-reduce).
static int row_linear(const B200LlamaWeights* w, const B200LlamaStep* s, const B200Linear* L, const void* x, HiddenIn* out, void* stream) {
  const bool can_defer = deferring(s) && (w->tp_size == 1 || fused_boundary(w, s));
  out->deferred = false;
  if (can_defer) {
    RUN(linear_deferred(L, s, x, &out->parts, s->T, stream));
    out->deferred = true;
    return B200_OK;
  }
  return linear(L, s, x, s->hidden, s->T, stream);
}

extern "C" int b200_llama_embed(const B200LlamaWeights* w, const B200LlamaStep* s, void* stream) {
  return b200_embedding(w->embed, s->input_ids, s->hidden, s->T, w->hidden_size, w->vocab_start, w->vocab_rows, stream);
}

// normed -> qkv -> rope + KV write -> attention -> o_proj
static int attn_body(const B200LlamaWeights* w, const B200LlamaStep* s, int layer, HiddenIn* out, void* stream) {
  const int64_t T = s->T;
  const int d = w->head_dim;
  const B200LlamaLayer* L = &w->layers[layer];
  char* k_pool = (char*)s->kv_pool + (size_t)layer * s->kv_layer_stride_bytes;
  char* v_pool = k_pool + s->kv_v_offset_bytes;
  if (deferring(s)) {
    B200SplitK qkv_parts;
    RUN(linear_deferred(&L->qkv, s, s->normed, &qkv_parts, T, stream));
    RUN(b200_rope_kv_write_paged_splitk(&qkv_parts, s->qkv, w->rope_cos, w->rope_sin, s->position_ids, s->slot_mapping, k_pool, v_pool,
                                        w->n_heads, w->n_kv_heads, d, stream));
  } else {
    RUN(linear(&L->qkv, s, s->normed, s->qkv, T, stream));
    RUN(b200_rope_kv_write_paged(s->qkv, w->rope_cos, w->rope_sin, s->position_ids, s->slot_mapping, k_pool, v_pool, T, w->n_heads,
                                 w->n_kv_heads, d, stream));
  }
  const int64_t qkv_stride = (int64_t)(w->n_heads + 2 * w->n_kv_heads) * d;
  if (s->is_prefill && s->kv_num_blocks > 0 && s->block_table && s->context_lens) {
    // through the pool (the step's K / V were just appended): tcgen05 + TMA, and the queries may follow a cached context
    RUN(b200_attn_prefill_paged(s->qkv, qkv_stride, T, k_pool, v_pool, s->kv_num_blocks, s->block_table, s->block_table_stride,
                                s->context_lens, s->cu_seqlens, s->attn_out, (int64_t)w->n_heads * d, s->B,
                                s->max_q > 0 ? s->max_q : s->max_s, w->n_heads, w->n_kv_heads, d, w->softmax_scale, stream));
  } else if (s->is_prefill) {
    const __half* q = (const __half*)s->qkv;
    RUN(b200_attn_prefill_varlen(q, qkv_stride, q + (int64_t)w->n_heads * d, qkv_stride,
                                 q + (int64_t)(w->n_heads + w->n_kv_heads) * d, qkv_stride, s->cu_seqlens, s->attn_out,
                                 (int64_t)w->n_heads * d, s->B, s->max_s, w->n_heads, w->n_kv_heads, d, w->softmax_scale, 1, stream));
  } else {
    RUN(b200_attn_decode_paged(s->qkv, qkv_stride, k_pool, v_pool, s->block_table, s->block_table_stride, s->context_lens,
                               s->attn_out, (int64_t)w->n_heads * d, s->attn_ws, s->attn_ws_bytes, s->B, w->n_heads, w->n_kv_heads,
                               d, s->max_s, w->softmax_scale, stream));
  }
  return row_linear(w, s, &L->o, s->attn_out, out, stream);
}

// normed -> gate_up -> silu * up -> down
static int mlp_body(const B200LlamaWeights* w, const B200LlamaStep* s, int layer, HiddenIn* out, void* stream) {
  const int64_t T = s->T;
  const B200LlamaLayer* L = &w->layers[layer];
  if (deferring(s)) {
    // the [T, 2 I] intermediate never exists: SiLU(gate) * up reads the projection's partials (either weight layout)
    B200SplitK gu_parts;
    RUN(linear_deferred(&L->gate_up, s, s->normed, &gu_parts, T, stream));
    RUN(b200_splitk_silu_mul(&gu_parts, s->act, stream));
  } else if (L->gate_up.qweight && L->gate_up.layout == B200_W4_LAYOUT_GATE_UP) {
    // int4 gate|up layout: SiLU(gate) * up is the GEMM's epilogue
    const void* x = s->normed;
    RUN(gathered(&L->gate_up, s, &x, T, stream));
    RUN(b200_gemm_w4a16_ex(x, L->gate_up.qweight, L->gate_up.bias, s->act, T, L->gate_up.N, L->gate_up.K, L->gate_up.groupsize,
                           B200_W4_LAYOUT_GATE_UP, 1, s->gemm_ws, stream));
  } else {
    RUN(linear(&L->gate_up, s, s->normed, s->gate_up, T, stream));
    RUN(b200_silu_mul(s->gate_up, s->act, T, L->gate_up.N / 2, stream));
  }
  return row_linear(w, s, &L->down, s->act, out, stream);
}

// hidden (+ residual) -> input_layernorm -> qkv -> rope + KV write -> attention -> o_proj -> hidden (rank-partial when TP).
// Block-wise entry point for callers that all-reduce `hidden` themselves between the blocks (NCCL): everything is materialised.
extern "C" int b200_llama_attn_block(const B200LlamaWeights* w, const B200LlamaStep* s, int layer, void* stream) {
  B200LlamaStep plain = *s;
  plain.p2p_norm = nullptr;  // the caller owns the boundary
  HiddenIn in, out;
  RUN(boundary_norm(w, &plain, in, w->layers[layer].input_ln, layer == 0, stream));
  RUN(attn_body(w, &plain, layer, &out, stream));
  if (out.deferred) RUN(b200_splitk_reduce(&out.parts, s->hidden, stream));
  return B200_OK;
}

// hidden + residual -> post_attention_layernorm -> gate_up -> silu*mul -> down -> hidden (rank-partial when TP)
extern "C" int b200_llama_mlp_block(const B200LlamaWeights* w, const B200LlamaStep* s, int layer, void* stream) {
  B200LlamaStep plain = *s;
  plain.p2p_norm = nullptr;
  HiddenIn in, out;
  RUN(boundary_norm(w, &plain, in, w->layers[layer].post_ln, false, stream));
  RUN(mlp_body(w, &plain, layer, &out, stream));
  if (out.deferred) RUN(b200_splitk_reduce(&out.parts, s->hidden, stream));
  return B200_OK;
}

// (gather rows) -> lm_head -> logits [n_rows, vocab_rows]; optional greedy ids.  s->normed holds the final norm's output.
static int head_body(const B200LlamaWeights* w, const B200LlamaStep* s, void* stream) {
  const int64_t T = s->T;
  const void* x = s->normed;
  int64_t rows = T;
  if (s->head_rows) {  // lm_head_indices (flash_llama_modeling.py:537-538)
    RUN(b200_embedding(s->normed, s->head_rows, s->head_in, s->n_head_rows, w->hidden_size, 0, T, stream));
    x = s->head_in;
    rows = s->n_head_rows;
  }
  B200Linear head = {};
  head.weight = w->lm_head;
  head.N = w->vocab_rows_head;
  head.K = w->hidden_size;
  RUN(linear(&head, s, x, s->logits, rows, stream));
  if (s->next_ids) {
    if (w->tp_size > 1) {
      if (!s->p2p_argmax) { b200_set_last_error("llama_head: greedy ids of a sharded head need the p2p_argmax window"); return B200_ERR_ARG; }
      RUN(b200_p2p_argmax(s->p2p_argmax, s->logits, s->next_ids, rows, w->vocab_rows_head, w->vocab_rows_head, s->banned_ids, stream));
    } else {
      RUN(b200_argmax(s->logits, s->next_ids, rows, w->vocab_rows_head, w->vocab_rows_head, s->banned_ids, stream));
    }
  }
  return B200_OK;
}

// final norm -> head, for callers that drive the blocks themselves (hidden is materialised and already all-reduced)
extern "C" int b200_llama_head(const B200LlamaWeights* w, const B200LlamaStep* s, void* stream) {
  B200LlamaStep plain = *s;
  plain.p2p_norm = nullptr;
  HiddenIn in;
  RUN(boundary_norm(w, &plain, in, w->final_norm, false, stream));
  return head_body(w, s, stream);
}

// whole step in one call: single rank, or tensor parallel with the fused NVLink boundary (no host-side collective)
extern "C" int b200_llama_step(const B200LlamaWeights* w, const B200LlamaStep* s, void* stream) {
  if (w->tp_size != 1 && !fused_boundary(w, s)) {
    b200_set_last_error("llama_step: tensor-parallel ranks without a p2p_norm window (or with T > 2048) drive attn/mlp blocks separately "
                        "(all-reduce in between)");
    return B200_ERR_UNSUPPORTED;
  }
  if (s->T == 0) return B200_OK;
  RUN(b200_llama_embed(w, s, stream));
  HiddenIn h;  // embeddings: a plain fp16 tensor (rank-partial when TP: the vocab-parallel lookup, utils/layers.py:346-357)
  for (int l = 0; l < w->n_layers; ++l) {
    RUN(boundary_norm(w, s, h, w->layers[l].input_ln, l == 0, stream));
    RUN(attn_body(w, s, l, &h, stream));
    RUN(boundary_norm(w, s, h, w->layers[l].post_ln, false, stream));
    RUN(mlp_body(w, s, l, &h, stream));
  }
  RUN(boundary_norm(w, s, h, w->final_norm, false, stream));
  return head_body(w, s, stream);
}
