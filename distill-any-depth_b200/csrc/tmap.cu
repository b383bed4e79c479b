#include "tmap.h"

namespace dad {

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
}  // namespace

int make_tmap(CUtensorMap* m, int dtype, const void* base, int rank, const cuuint64_t* dims,
              const cuuint64_t* strides_bytes, const cuuint32_t* box, const cuuint32_t* elem_strides) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_error(DAD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (elem_strides)
        for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
        return set_error(DAD_ERR_INVALID, "TMA base address %p not 16-byte aligned", base);
    for (int i = 0; i < rank - 1; ++i)
        if (strides_bytes[i] % 16 != 0)
            return set_error(DAD_ERR_INVALID, "TMA stride %llu not a multiple of 16 bytes",
                             (unsigned long long)strides_bytes[i]);
    CUresult r = fn(m, dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(DAD_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return DAD_OK;
}

}  // namespace dad
