// Host-side TMA tensor-map construction (cuTensorMapEncodeTiled through the runtime's driver entry point,
// so the library does not link libcuda).  bf16 elements, SWIZZLE_128B, zero fill for out-of-bounds boxes.
#pragma once
#include <cuda.h>

#include "common.h"

namespace dad {

// dtype: 0 = bf16, 1 = fp32
// elem_strides (optional, per dimension): traversal step of the box; ceil(box[i] / elem_strides[i]) elements are loaded
int make_tmap(CUtensorMap* m, int dtype, const void* base, int rank, const cuuint64_t* dims,
              const cuuint64_t* strides_bytes, const cuuint32_t* box, const cuuint32_t* elem_strides = nullptr);
inline int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims,
                          const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    return make_tmap(m, 0, base, rank, dims, strides_bytes, box);
}

}  // namespace dad
