// Shared device/host helpers for libsdcb200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/sdc_b200.h"

namespace sdcb200 {

// ---------------------------------------------------------------------------------------------------------------------
// host-side error reporting
// ---------------------------------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(const char* where, const std::string& msg);
int fail_cuda(const char* where, cudaError_t e);

#define SDC_CUDA_OK(call)                                         \
    do {                                                          \
        cudaError_t _e = (call);                                  \
        if (_e != cudaSuccess) return fail_cuda(__func__, _e);    \
    } while (0)
#define SDC_REQUIRE(cond, msg)                                    \
    do {                                                          \
        if (!(cond)) return fail(__func__, msg);                  \
    } while (0)

int sm_count();

// ---------------------------------------------------------------------------------------------------------------------
// grid geometry of the walled layout (see include/sdc_b200.h)
// ---------------------------------------------------------------------------------------------------------------------
struct Geom {
    int ndim;       // 1..3
    int n;          // points per dimension
    int P;          // pitch = n + (n & 1)
    int periodic;   // 0 dirichlet-zero (walls), 1 periodic
    int nz;         // 3-D: planes owned along the slowest axis (= n unless the grid is slab-decomposed)
    int zhalo;      // 3-D slabs: planes -1 and nz are halo planes filled by the neighbours (no wrap in z)
    long long sy;   // stride of y = P
    long long sz;   // stride of z = P*P
    long long vol;  // doubles of one field (without the guard)
    long long owned;  // doubles spanning the owned planes (streaming loops)
};

inline Geom make_geom(int ndim, int n, int bc) {
    Geom g;
    g.ndim = ndim;
    g.n = n;
    g.P = n + (n & 1);
    g.periodic = bc == SDCB200_BC_PERIODIC;
    g.sy = g.P;
    g.sz = (long long)g.P * g.P;
    g.vol = g.P;
    for (int d = 1; d < ndim; ++d) g.vol *= g.P;
    g.nz = ndim == 3 ? n : 1;
    g.zhalo = 0;
    g.owned = ndim == 3 ? g.sz * g.nz : g.vol;
    return g;
}

// slab of a 3-D grid: nz owned planes of an n x n cross-section, halo planes at -1 (the guard) and nz
inline Geom make_slab_geom(int n, int nz, int bc) {
    Geom g = make_geom(3, n, bc);
    g.nz = nz;
    g.zhalo = 1;
    g.vol = g.sz * (nz + 1);
    g.owned = g.sz * nz;
    return g;
}

// ---------------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kThreads = 256;  // CTA size of the streaming / stencil kernels (block reductions read blockDim.x)

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum with a fixed reduction tree (deterministic for a fixed launch shape); result valid in ALL threads.
// `scratch` = 33 doubles of shared memory.  Safe to call back to back (leading barrier).
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = lane < (int)(blockDim.x >> 5) ? scratch[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}
__device__ __forceinline__ double block_max(double v, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = lane < (int)(blockDim.x >> 5) ? scratch[lane] : 0.0;
        t = warp_max(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// max over non-negative doubles via integer atomics (IEEE-754 ordering of non-negative values; NaN sorts above inf
// so a NaN residual surfaces instead of being dropped).  Order independent => deterministic.
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

}  // namespace sdcb200
