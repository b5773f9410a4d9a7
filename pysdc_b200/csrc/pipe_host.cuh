// Host-side helpers of the pipelined kernels: co-resident grid size, tensor-map encoding (driver entry point resolved
// through the runtime: no link-time dependency on libcuda).
#pragma once
#include "cg_pipe.cuh"

namespace sdcb200 {
namespace {

// co-resident grid of a pipelined kernel (2 CTAs per SM when they fit), cached per kernel
template <auto kernel>  // (a non-type parameter: kernels of equal signature must not share the cache)
int pipe_grid(size_t smem, int* out) {
    static int cached = 0;
    if (cached == 0) {
        SDC_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        SDC_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kPipeThreads, smem));
        if (per_sm < 1) return fail("pipe_grid", "pipelined solver kernel does not fit on an SM");
        if (per_sm > 2) per_sm = 2;
        cached = per_sm * sm_count();
    }
    *out = cached;
    return 0;
}

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int tensor_map_encoder(TensorMapEncodeFn* out) {
    static TensorMapEncodeFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        SDC_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (p == nullptr || q != cudaDriverEntryPointSuccess)
            return fail("tensor_map_encoder", "the CUDA driver does not provide cuTensorMapEncodeTiled");
        fn = reinterpret_cast<TensorMapEncodeFn>(p);
    }
    *out = fn;
    return 0;
}

// Tiled fp64 map over a field.  Dirichlet (walled) grids: 2-D {P, P} from element (0,0); 3-D {P, P, nz+2} starting ONE
// PLANE BELOW the field (guard = lower halo plane) up to and including plane nz (wall / upper halo plane).  Periodic
// (dense) grids: {n, n[, n]} from element (0,0[,0]); what lies across an edge is fetched by the wrap boxes.
int encode_field_map(CUtensorMap* map, const Geom& g, const double* field, int box_x, int box_y) {
    TensorMapEncodeFn enc = nullptr;
    if (int rc = tensor_map_encoder(&enc)) return rc;
    const cuuint32_t rank = (cuuint32_t)g.ndim;
    const bool below = g.ndim == 3 && !g.periodic;
    cuuint64_t dims[3] = {(cuuint64_t)g.P, (cuuint64_t)g.P, (cuuint64_t)(below ? g.nz + 2 : g.nz)};
    cuuint64_t strides[2] = {(cuuint64_t)g.sy * 8u, (cuuint64_t)g.sz * 8u};
    cuuint32_t box[3] = {(cuuint32_t)box_x, (cuuint32_t)box_y, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    void* base = const_cast<double*>(below ? field - g.sz : field);
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank, base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("encode_field_map", "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return 0;
}

// the maps of one field read with halo: the 18x68 box and, on periodic grids, its wrap row (1x68) and column pair (18x2)
int encode_halo_maps(CUtensorMap* maps_of_sys, int halo_slot, const Geom& g, const double* field) {
    if (int rc = encode_field_map(maps_of_sys + halo_slot, g, field, kPHX, kPHY)) return rc;
    if (g.periodic) {
        if (int rc = encode_field_map(maps_of_sys + halo_slot + kMapRowOf, g, field, kPHX, 1)) return rc;
        if (int rc = encode_field_map(maps_of_sys + halo_slot + kMapColOf, g, field, 2, kPHY)) return rc;
    }
    return 0;
}

int encode_system_maps(CUtensorMap* m, const Geom& g, const Sys& S) {
    if (int rc = encode_halo_maps(m, kMapRHalo, g, S.r)) return rc;
    if (int rc = encode_halo_maps(m, kMapPHalo, g, S.p)) return rc;
    if (int rc = encode_halo_maps(m, kMapQHalo, g, S.q)) return rc;
    if (int rc = encode_field_map(m + kMapRCentre, g, S.r, kPX, kPY)) return rc;
    if (int rc = encode_field_map(m + kMapXCentre, g, S.x, kPX, kPY)) return rc;
    if (S.z != nullptr)
        if (int rc = encode_halo_maps(m, kMapZHalo, g, S.z)) return rc;
    if (S.dvec != nullptr)
        if (int rc = encode_field_map(m + kMapDCentre, g, S.dvec, kPX, kPY)) return rc;
    return 0;
}


}  // namespace
}  // namespace sdcb200
