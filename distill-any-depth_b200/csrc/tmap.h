// Host-side TMA tensor-map construction (cuTensorMapEncodeTiled through the runtime's driver entry point,
// so the library does not link libcuda).  bf16 elements, SWIZZLE_128B, zero fill for out-of-bounds boxes.
#pragma once
#include <cuda.h>

#include "common.h"

namespace dad {

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box);

}  // namespace dad
