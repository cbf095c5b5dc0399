// C++ step runtime for the FlashLlama graph: one C call enqueues a whole decode (or prefill) step on a stream, so the
// Python host pays one FFI call per step (per half-layer when tensor-parallel) and the sequence is CUDA-graph capturable.
//
// Restates the op order of FlashLlamaModel.forward / FlashLlamaLayer.forward / FlashLlamaAttention.forward / LlamaMLP.forward
// (/root/reference/server/text_generation_server/models/custom_modeling/flash_llama_modeling.py:425-497, 356-389, 240-297,
// 332-335) and FlashLlamaForCausalLM.forward (:514-540) over the kernels of this library.
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

#define RUN(expr)            \
  do {                       \
    int _s = (expr);         \
    if (_s != B200_OK) return _s; \
  } while (0)

extern "C" int b200_llama_embed(const B200LlamaWeights* w, const B200LlamaStep* s, void* stream) {
  return b200_embedding(w->embed, s->input_ids, s->hidden, s->T, w->hidden_size, w->vocab_start, w->vocab_rows, stream);
}

// hidden (+ residual) -> input_layernorm -> qkv -> rope + KV write -> attention -> o_proj -> hidden (rank-partial when TP)
extern "C" int b200_llama_attn_block(const B200LlamaWeights* w, const B200LlamaStep* s, int layer, void* stream) {
  const int64_t T = s->T;
  const int d = w->head_dim;
  const B200LlamaLayer* L = &w->layers[layer];
  // first layer: residual = None -> residual_out aliases hidden (flash_llama_modeling.py:149-150)
  if (layer == 0) {
    RUN(b200_rmsnorm_residual(s->hidden, nullptr, L->input_ln, s->normed, nullptr, T, w->hidden_size, w->rms_eps, stream));
    if (cudaMemcpyAsync(s->residual, s->hidden, (size_t)T * w->hidden_size * 2, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) !=
        cudaSuccess) {
      b200_set_last_error("llama_attn_block: residual copy failed");
      return B200_ERR_CUDA;
    }
  } else {
    RUN(b200_rmsnorm_residual(s->hidden, s->residual, L->input_ln, s->normed, s->residual, T, w->hidden_size, w->rms_eps, stream));
  }
  RUN(linear(&L->qkv, s, s->normed, s->qkv, T, stream));
  char* k_pool = (char*)s->kv_pool + (size_t)layer * s->kv_layer_stride_bytes;
  char* v_pool = k_pool + s->kv_v_offset_bytes;
  RUN(b200_rope_kv_write_paged(s->qkv, w->rope_cos, w->rope_sin, s->position_ids, s->slot_mapping, k_pool, v_pool, T, w->n_heads,
                               w->n_kv_heads, d, stream));
  const int64_t qkv_stride = (int64_t)(w->n_heads + 2 * w->n_kv_heads) * d;
  if (s->is_prefill) {
    const __half* q = (const __half*)s->qkv;
    RUN(b200_attn_prefill_varlen(q, qkv_stride, q + (int64_t)w->n_heads * d, qkv_stride,
                                 q + (int64_t)(w->n_heads + w->n_kv_heads) * d, qkv_stride, s->cu_seqlens, s->attn_out,
                                 (int64_t)w->n_heads * d, s->B, s->max_s, w->n_heads, w->n_kv_heads, d, w->softmax_scale, 1, stream));
  } else {
    RUN(b200_attn_decode_paged(s->qkv, qkv_stride, k_pool, v_pool, s->block_table, s->block_table_stride, s->context_lens,
                               s->attn_out, (int64_t)w->n_heads * d, s->attn_ws, s->attn_ws_bytes, s->B, w->n_heads, w->n_kv_heads,
                               d, s->max_s, w->softmax_scale, stream));
  }
  RUN(linear(&L->o, s, s->attn_out, s->hidden, T, stream));
  return B200_OK;
}

// hidden + residual -> post_attention_layernorm -> gate_up -> silu*mul -> down -> hidden (rank-partial when TP)
extern "C" int b200_llama_mlp_block(const B200LlamaWeights* w, const B200LlamaStep* s, int layer, void* stream) {
  const int64_t T = s->T;
  const B200LlamaLayer* L = &w->layers[layer];
  RUN(b200_rmsnorm_residual(s->hidden, s->residual, L->post_ln, s->normed, s->residual, T, w->hidden_size, w->rms_eps, stream));
  if (L->gate_up.qweight && L->gate_up.layout == B200_W4_LAYOUT_GATE_UP) {
    // int4 gate|up layout: SiLU(gate) * up is the GEMM's epilogue, the [T, 2 I] intermediate never exists
    const void* x = s->normed;
    RUN(gathered(&L->gate_up, s, &x, T, stream));
    RUN(b200_gemm_w4a16_ex(x, L->gate_up.qweight, L->gate_up.bias, s->act, T, L->gate_up.N, L->gate_up.K, L->gate_up.groupsize,
                           B200_W4_LAYOUT_GATE_UP, 1, s->gemm_ws, stream));
  } else {
    RUN(linear(&L->gate_up, s, s->normed, s->gate_up, T, stream));
    RUN(b200_silu_mul(s->gate_up, s->act, T, L->gate_up.N / 2, stream));
  }
  RUN(linear(&L->down, s, s->act, s->hidden, T, stream));
  return B200_OK;
}

// final norm -> (gather rows) -> lm_head -> logits [n_rows, vocab_rows]; optional greedy ids
extern "C" int b200_llama_head(const B200LlamaWeights* w, const B200LlamaStep* s, void* stream) {
  const int64_t T = s->T;
  RUN(b200_rmsnorm_residual(s->hidden, s->residual, w->final_norm, s->normed, s->residual, T, w->hidden_size, w->rms_eps, stream));
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
  if (s->next_ids) RUN(b200_argmax(s->logits, s->next_ids, rows, w->vocab_rows_head, w->vocab_rows_head, s->banned_ids, stream));
  return B200_OK;
}

// whole step, single rank (no collectives)
extern "C" int b200_llama_step(const B200LlamaWeights* w, const B200LlamaStep* s, void* stream) {
  if (w->tp_size != 1) {
    b200_set_last_error("llama_step: tensor-parallel ranks drive attn/mlp blocks separately (all-reduce in between)");
    return B200_ERR_UNSUPPORTED;
  }
  if (s->T == 0) return B200_OK;
  RUN(b200_llama_embed(w, s, stream));
  for (int l = 0; l < w->n_layers; ++l) {
    RUN(b200_llama_attn_block(w, s, l, stream));
    RUN(b200_llama_mlp_block(w, s, l, stream));
  }
  return b200_llama_head(w, s, stream);
}
