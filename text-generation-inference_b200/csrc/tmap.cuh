// Host-side TMA tensor-map cache.  cuTensorMapEncodeTiled is resolved through cudaGetDriverEntryPoint so the
// library has no link-time dependency on libcuda (the CPU build box has none).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

enum class TmapSwizzle { kNone = 0, k128B = 1 };
enum class TmapDtype { kF16 = 0, kI32 = 1 };

// 2-D row-major tensor [rows][cols] (cols innermost), box [box_rows][box_cols].  Returns nullptr on failure
// (b200_last_error set).  Maps are cached by (ptr, shape, box, dtype, swizzle); pointers handed out stay valid
// for the life of the process.
const CUtensorMap* get_tmap_2d(const void* ptr, uint64_t rows, uint64_t cols, uint64_t row_stride_elems, uint32_t box_rows,
                               uint32_t box_cols, TmapDtype dt, TmapSwizzle sw);

}  // namespace b200
