// tma_maps.h -- host-side construction of CUtensorMap descriptors (cuTensorMapEncodeTiled through the runtime's driver entry
// point: no link-time dependency on libcuda) with a small per-thread cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>
#include <stdlib.h>

namespace ob {

typedef CUresult (*ob_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                       const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static ob_encode_tiled_fn encode_tiled_fn() {
    static ob_encode_tiled_fn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (ob_encode_tiled_fn)p;
    }();
    return fn;
}

// tensor maps are cached per (pointer, shape): a model re-launches with the same parents every stage
struct StageMapKey { const void *p; unsigned long long px, py, pz; int tw, th, esz; };
static bool stage_map(CUtensorMap *out, const void *ptr, cuuint64_t Px, cuuint64_t Py, cuuint64_t Pz, int tw, int th, int esz) {
    struct Entry { StageMapKey k; CUtensorMap m; };
    static thread_local std::vector<Entry> cache;
    for (const Entry &e : cache)
        if (e.k.p == ptr && e.k.px == Px && e.k.py == Py && e.k.pz == Pz && e.k.tw == tw && e.k.th == th && e.k.esz == esz) { *out = e.m; return true; }
    if (!encode_tiled_fn()) return false;
    const cuuint64_t dims[3] = {Px, Py, Pz};
    const cuuint64_t strides[2] = {Px * (cuuint64_t)esz, Px * Py * (cuuint64_t)esz};
    const cuuint32_t box[3] = {(cuuint32_t)tw, (cuuint32_t)th, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUtensorMap m;
    if (encode_tiled_fn()(&m, esz == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(ptr), dims, strides, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    if (cache.size() > 256) cache.clear();
    cache.push_back(Entry{StageMapKey{ptr, Px, Py, Pz, tw, th, esz}, m});
    *out = m;
    return true;
}

}  // namespace ob
