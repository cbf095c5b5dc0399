// Library-level runtime pieces behind the C ABI: last-error string, TMA tensor-map cache, ABI version.
#include "common.cuh"
#include "tmap.cuh"

#include <cudaTypedefs.h>
#include <map>
#include <mutex>
#include <string>
#include <tuple>

static thread_local std::string g_last_error;

void b200_set_last_error(const char* msg) { g_last_error = msg ? msg : ""; }

extern "C" const char* b200_last_error(void) { return g_last_error.c_str(); }
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
