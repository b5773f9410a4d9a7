// General 1-D finite-difference operators applied as a Kronecker sum: the stencils of
// pySDC/helpers/problem_helper.py:4-80 - centred / upwind / forward / backward offsets within +-4 grid points - with the
// boundary treatment of :133-201: periodic wrap, or, on dirichlet-zero grids, the one-sided closure rows the reference
// derives for the points next to each boundary.  Shared by highorder.cu (order 4/6/8 Laplacians: eval_f, CG) and
// gmres.cu (restarted GMRES for any of these operators, advection included).
#pragma once
#include "cg_common.cuh"

namespace sdcb200 {

constexpr int kHoMaxH = 4;                 // order 8
constexpr int kHoMaxW = 2 * kHoMaxH + 1;   // closure rows carry order+1 coefficients

struct HoOp {
    int h;                       // half width: offsets -h .. h
    double cf[kHoMaxW];          // coefficient of offset k at cf[k + kHoMaxH], already scaled by coeff / dx^derivative
    double lo[kHoMaxH][kHoMaxW]; // dirichlet: row i (< h) acts on columns 0 .. order
    double hi[kHoMaxH][kHoMaxW]; // dirichlet: row n-1-i acts on columns n-1-order .. n-1
};

// (1-D operator along one axis) applied at index i of a line with element stride s; `line` points at index 0
__device__ __forceinline__ double ho_line(const HoOp& op, bool periodic, int n, const double* line, int i, long long s) {
    const int h = op.h;
    double acc = 0.0;
    if (periodic) {
        for (int k = -h; k <= h; ++k) {
            int j = i + k;
            j = j < 0 ? j + n : (j >= n ? j - n : j);
            acc = fma(op.cf[k + kHoMaxH], line[(long long)j * s], acc);
        }
        return acc;
    }
    const int w = 2 * h + 1;
    if (i < h) {
        for (int j = 0; j < w && j < n; ++j) acc = fma(op.lo[i][j], line[(long long)j * s], acc);
    } else if (i >= n - h) {
        const int r = n - 1 - i;
        for (int j = 0; j < w && j < n; ++j) acc = fma(op.hi[r][j], line[(long long)(n - w + j) * s], acc);
    } else {
        for (int k = -h; k <= h; ++k) acc = fma(op.cf[k + kHoMaxH], line[(long long)(i + k) * s], acc);
    }
    return acc;
}

// A u at grid point (x, y, z)
__device__ __forceinline__ double ho_apply(const HoOp& op, const Geom& g, const double* u, int x, int y, int z) {
    const bool per = g.periodic;
    double acc = ho_line(op, per, g.n, u + (long long)z * g.sz + (long long)y * g.sy, x, 1);
    if (g.ndim >= 2) acc += ho_line(op, per, g.n, u + (long long)z * g.sz + x, y, g.sy);
    if (g.ndim == 3) acc += ho_line(op, per, g.n, u + (long long)y * g.sy + x, z, g.sz);
    return acc;
}

// loop over the grid points of this thread: f(idx, x, y, z)
template <class F>
__device__ __forceinline__ void ho_points(const Geom& g, F&& f) {
    const long long npts = (long long)g.n * (g.ndim >= 2 ? g.n : 1) * (g.ndim == 3 ? g.n : 1);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < npts; t += stride) {
        const int x = (int)(t % g.n);
        const long long r = t / g.n;
        const int y = g.ndim >= 2 ? (int)(r % g.n) : 0;
        const int z = g.ndim == 3 ? (int)(r / g.n) : 0;
        f((long long)z * g.sz + (long long)y * g.sy + x, x, y, z);
    }
}

// general operator: coef[k + h] = coefficient of offset k; closure rows (dirichlet-zero only) of 2h+1 entries each
inline int fill_op_general(HoOp& op, int h, int bc, const double* coef, const double* lo, const double* hi) {
    SDC_REQUIRE(h >= 1 && h <= kHoMaxH, "stencil half width must be 1 .. 4");
    SDC_REQUIRE(coef != nullptr, "stencil coefficients missing");
    memset(&op, 0, sizeof(op));
    op.h = h;
    for (int k = -h; k <= h; ++k) op.cf[k + kHoMaxH] = coef[k + h];
    if (bc == SDCB200_BC_DIRICHLET) {
        SDC_REQUIRE(lo != nullptr && hi != nullptr, "dirichlet-zero grids need the closure rows");
        for (int i = 0; i < h; ++i)
            for (int j = 0; j <= 2 * h; ++j) {
                op.lo[i][j] = lo[i * (2 * h + 1) + j];
                op.hi[i][j] = hi[i * (2 * h + 1) + j];
            }
    }
    return 0;
}

inline bool ok8(const void* p) { return p != nullptr && (reinterpret_cast<size_t>(p) & 7u) == 0; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
constexpr int kMaxGrid = 148 * 8;

}  // namespace sdcb200
