// Paged KV block allocator (host side, C++) and the device-side per-step bookkeeping kernel.
//
// Replaces the un-vendored fms-extras `PagedKVCacheManager` block bookkeeping used by
// /root/reference/server/text_generation_server/models/paged_causal_lm.py:338-353 and utils/paged.py:92-134
// (allocate_tokens / free_sequences; block size 16).  The allocator is a LIFO free list of block ids; the
// per-sequence block lists live in the Python manager (utils/paged.py of this package).
#include "common.cuh"

#include <mutex>
#include <vector>

struct B200KvAllocator {
  std::vector<int32_t> free_list;
  std::vector<uint8_t> in_use;  // per block: handed out and not yet released
  int32_t num_blocks;
  std::mutex mu;
};

extern "C" void* b200_kv_alloc_create(int32_t num_blocks) {
  if (num_blocks <= 0) { b200_set_last_error("kv_alloc_create: num_blocks must be positive"); return nullptr; }
  auto* a = new B200KvAllocator();
  a->num_blocks = num_blocks;
  a->free_list.reserve(num_blocks);
  a->in_use.assign(num_blocks, 0);
  for (int32_t i = num_blocks - 1; i >= 0; --i) a->free_list.push_back(i);  // block 0 is handed out first
  return a;
}

extern "C" void b200_kv_alloc_destroy(void* h) { delete static_cast<B200KvAllocator*>(h); }

extern "C" int32_t b200_kv_alloc_num_free(void* h) {
  auto* a = static_cast<B200KvAllocator*>(h);
  std::lock_guard<std::mutex> lock(a->mu);
  return (int32_t)a->free_list.size();
}

// Takes n blocks (all or nothing).  Returns B200_ERR_NOMEM when fewer than n are free.
extern "C" int b200_kv_alloc_take(void* h, int32_t n, int32_t* out_ids /* host */) {
  auto* a = static_cast<B200KvAllocator*>(h);
  std::lock_guard<std::mutex> lock(a->mu);
  if (n < 0 || (size_t)n > a->free_list.size()) {
    b200_set_last_error("kv_alloc_take: out of KV cache blocks");
    return B200_ERR_NOMEM;
  }
  for (int32_t i = 0; i < n; ++i) {
    out_ids[i] = a->free_list.back();
    a->free_list.pop_back();
    a->in_use[out_ids[i]] = 1;
  }
  return B200_OK;
}

extern "C" int b200_kv_alloc_release(void* h, const int32_t* ids /* host */, int32_t n) {
  auto* a = static_cast<B200KvAllocator*>(h);
  std::lock_guard<std::mutex> lock(a->mu);
  // validate everything before touching the list: a bad call leaves the allocator unchanged
  for (int32_t i = 0; i < n; ++i) {
    if (ids[i] < 0 || ids[i] >= a->num_blocks) { b200_set_last_error("kv_alloc_release: bad block id"); return B200_ERR_ARG; }
    if (a->in_use[ids[i]] != 1) {  // free already, or listed twice in this call (marked 2 below)
      for (int32_t j = 0; j < i; ++j) a->in_use[ids[j]] = 1;
      b200_set_last_error("kv_alloc_release: double free");
      return B200_ERR_ARG;
    }
    a->in_use[ids[i]] = 2;
  }
  for (int32_t i = 0; i < n; ++i) {
    a->in_use[ids[i]] = 0;
    a->free_list.push_back(ids[i]);
  }
  return B200_OK;
}

namespace b200 {
// Start-of-decode-step bookkeeping, one thread per sequence:
//   pos = context_lens[b] (tokens already cached) -> position_ids[b] = pos, slot_mapping[b] = block(pos)*16 + pos%16,
//   context_lens[b] = pos + 1 (the attention sees the token written this step: flash_llama_modeling.py:282-295),
//   input_ids[b] = next_ids[b] (device-to-device chaining of the greedy token).
// Rows with context_lens < 0 are padding: slot -1, context stays negative (attention treats <= 0 as empty).
__global__ void decode_advance_kernel(const int32_t* __restrict__ block_table, int64_t bt_stride, int32_t* __restrict__ context_lens,
                                      int64_t* __restrict__ position_ids, int64_t* __restrict__ slot_mapping,
                                      const int64_t* __restrict__ next_ids, int64_t* __restrict__ input_ids, int B) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int pos = context_lens[b];
  if (pos < 0) {
    slot_mapping[b] = -1;
    position_ids[b] = 0;
    if (input_ids) input_ids[b] = 0;
    return;
  }
  const int blk = block_table[(int64_t)b * bt_stride + pos / kPageTokens];
  slot_mapping[b] = (int64_t)blk * kPageTokens + pos % kPageTokens;
  position_ids[b] = pos;
  context_lens[b] = pos + 1;
  if (next_ids && input_ids) input_ids[b] = next_ids[b];
}
}  // namespace b200

extern "C" int b200_decode_advance(const int32_t* block_table, int64_t block_table_stride, int32_t* context_lens,
                                   int64_t* position_ids, int64_t* slot_mapping, const int64_t* next_ids, int64_t* input_ids,
                                   int B, void* stream) {
  if (B == 0) return B200_OK;
  B200_LAUNCH(b200::decode_advance_kernel, dim3((B + 127) / 128), dim3(128), 0, (cudaStream_t)stream, block_table, block_table_stride,
              context_lens, position_ids, slot_mapping, next_ids, input_ids, B);
  b200_count_launches(1);
  return B200_OK;
}
