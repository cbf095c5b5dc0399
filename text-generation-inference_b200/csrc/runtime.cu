// Library-level runtime pieces behind the C ABI: last-error string, TMA tensor-map cache, ABI version.
#include "common.cuh"
#include "tmap.cuh"

#include <cudaTypedefs.h>
#include <map>
#include <mutex>
#include <cstring>
#include <string>
#include <tuple>

static thread_local std::string g_last_error;

void b200_set_last_error(const char* msg) { g_last_error = msg ? msg : ""; }

#include <atomic>
static std::atomic<long long> g_launches{0};
void b200_count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" int64_t b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

bool b200_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_PDL");
    v = (e && (e[0] == '0' || e[0] == 'f' || e[0] == 'F')) ? 0 : 1;
  }
  return v == 1;
}

extern "C" const char* b200_last_error(void) { return g_last_error.c_str(); }
// debug aid: pending CUDA runtime error of this library's runtime instance (does not clear it)
extern "C" const char* b200_cuda_peek_error(void) { return cudaGetErrorString(cudaPeekAtLastError()); }
extern "C" int b200_abi_version(void) { return 1; }

namespace b200 {

static PFN_cuTensorMapEncodeTiled_v12000 resolve_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

const CUtensorMap* get_tmap_2d(const void* ptr, uint64_t rows, uint64_t cols, uint64_t row_stride_elems, uint32_t box_rows,
                               uint32_t box_cols, TmapDtype dt, TmapSwizzle sw) {
  using Key = std::tuple<const void*, uint64_t, uint64_t, uint64_t, uint32_t, uint32_t, int, int>;
  static std::map<Key, CUtensorMap*> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  Key key{ptr, rows, cols, row_stride_elems, box_rows, box_cols, (int)dt, (int)sw};
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  auto encode = resolve_encode();
  if (!encode) {
    b200_set_last_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return nullptr;
  }
  const uint64_t esz = dt == TmapDtype::kF16 ? 2 : 4;
  CUtensorMap* m = nullptr;
  if (posix_memalign(reinterpret_cast<void**>(&m), 64, sizeof(CUtensorMap)) != 0) return nullptr;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_stride_elems * esz};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = encode(m, dt == TmapDtype::kF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_INT32, 2,
                      const_cast<void*>(ptr), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      sw == TmapSwizzle::k128B ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    free(m);
    char buf[160];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed: CUresult %d (rows %llu cols %llu box %u x %u)", (int)r,
             (unsigned long long)rows, (unsigned long long)cols, box_rows, box_cols);
    b200_set_last_error(buf);
    return nullptr;
  }
  cache.emplace(key, m);
  return m;
}

}  // namespace b200

// ---------------------------------------------------------------------------------------------------------
// Per-kernel device timing for bench.py's roofline line: CUDA events recorded on the launching stream around
// every launch of one chosen kernel family while a timing handle is attached.
// ---------------------------------------------------------------------------------------------------------
#include <vector>
struct B200Timing {
  std::vector<cudaEvent_t> ev;  // start/stop pairs
  int used = 0;
  int which = 0;
};
static B200Timing* g_timing = nullptr;

extern "C" void* b200_timing_create(int max_launches) {
  auto* t = new B200Timing();
  t->ev.resize((size_t)max_launches * 2);
  for (auto& e : t->ev)
    if (cudaEventCreate(&e) != cudaSuccess) { b200_set_last_error("timing_create: cudaEventCreate failed"); delete t; return nullptr; }
  return t;
}
extern "C" void b200_timing_destroy(void* h) {
  auto* t = static_cast<B200Timing*>(h);
  if (!t) return;
  if (g_timing == t) g_timing = nullptr;
  for (auto& e : t->ev) cudaEventDestroy(e);
  delete t;
}
// which: B200_TIME_ATTN_DECODE / _GEMM_W4A16 / _GEMM_F16; h == NULL detaches
extern "C" void b200_timing_attach(void* h, int which) {
  g_timing = static_cast<B200Timing*>(h);
  if (g_timing) { g_timing->which = which; g_timing->used = 0; }
}
// synchronises on the recorded events; returns the number of timed launches and their summed duration
extern "C" int b200_timing_collect(void* h, float* total_ms) {
  auto* t = static_cast<B200Timing*>(h);
  float sum = 0.f;
  for (int i = 0; i + 1 < t->used; i += 2) {
    float ms = 0.f;
    cudaError_t e1 = cudaEventSynchronize(t->ev[i + 1]);
    cudaError_t e2 = e1 == cudaSuccess ? cudaEventElapsedTime(&ms, t->ev[i], t->ev[i + 1]) : e1;
    if (e2 != cudaSuccess) {
      char buf[128];
      snprintf(buf, sizeof(buf), "timing_collect: pair %d failed: %s", i / 2, cudaGetErrorString(e2));
      b200_set_last_error(buf);
      cudaGetLastError();
      return -1;
    }
    sum += ms;
  }
  *total_ms = sum;
  return t->used / 2;
}
void b200_timing_mark(int which, int is_stop, cudaStream_t st) {
  B200Timing* t = g_timing;
  if (!t || t->which != which || t->used + (is_stop ? 0 : 2) > (int)t->ev.size()) return;
  if (!is_stop) { cudaEventRecord(t->ev[t->used], st); }
  else if (t->used + 1 < (int)t->ev.size()) { cudaEventRecord(t->ev[t->used + 1], st); t->used += 2; }
}

// ---------------------------------------------------------------------------------------------------------
// Step trace (debug): while a device buffer is attached, every B200_LAUNCH is followed by a one-thread kernel that
// stores %globaltimer into the next slot.  Consecutive differences = device time of each kernel including its launch gap,
// in stream order, L2 state as in production; the launch sequence can be captured into a CUDA graph and replayed.
// ---------------------------------------------------------------------------------------------------------
static unsigned long long* g_trace_buf = nullptr;
static int g_trace_cap = 0, g_trace_n = 0;
static std::string g_trace_names;

__global__ void b200_stamp_kernel(unsigned long long* out) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *out = t;
}

// device_buf: capacity uint64 slots (NULL detaches).  Slot 0 is stamped by b200_debug_step_trace_begin.
extern "C" void b200_debug_step_trace(void* device_buf, int capacity) {
  g_trace_buf = static_cast<unsigned long long*>(device_buf);
  g_trace_cap = device_buf ? capacity : 0;
  g_trace_n = 0;
  g_trace_names.clear();
}
extern "C" void b200_debug_step_trace_begin(void* stream) { b200_step_trace_stamp("<begin>", (cudaStream_t)stream); }
// newline-separated kernel names of the stamped launches so far, in order; returns how many
extern "C" int b200_debug_step_trace_names(char* out, int64_t cap) {
  if (out && cap > 0) {
    const size_t n = g_trace_names.size() < (size_t)cap - 1 ? g_trace_names.size() : (size_t)cap - 1;
    memcpy(out, g_trace_names.data(), n);
    out[n] = 0;
  }
  return g_trace_n;
}
void b200_step_trace_stamp(const char* kernel_name, cudaStream_t stream) {
  if (!g_trace_buf || g_trace_n >= g_trace_cap) return;
  b200_stamp_kernel<<<1, 1, 0, stream>>>(g_trace_buf + g_trace_n);
  ++g_trace_n;
  g_trace_names += kernel_name;
  g_trace_names += '\n';
}
