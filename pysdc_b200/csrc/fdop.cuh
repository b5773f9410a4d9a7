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
// (H = op.h as a compile-time constant: the loops unroll and the coefficients become constant-bank operands)
template <int H>
__device__ __forceinline__ double ho_line(const HoOp& op, bool periodic, int n, const double* line, int i, long long s) {
    constexpr int h = H;
    double acc = 0.0;
    if (periodic) {
#pragma unroll
        for (int k = -h; k <= h; ++k) {
            int j = i + k;
            j = j < 0 ? j + n : (j >= n ? j - n : j);
            acc = fma(op.cf[k + kHoMaxH], line[(long long)j * s], acc);
        }
        return acc;
    }
    constexpr int w = 2 * h + 1;
    if (i < h) {
        for (int j = 0; j < w && j < n; ++j) acc = fma(op.lo[i][j], line[(long long)j * s], acc);
    } else if (i >= n - h) {
        const int r = n - 1 - i;
        for (int j = 0; j < w && j < n; ++j) acc = fma(op.hi[r][j], line[(long long)(n - w + j) * s], acc);
    } else {
#pragma unroll
        for (int k = -h; k <= h; ++k) acc = fma(op.cf[k + kHoMaxH], line[(long long)(i + k) * s], acc);
    }
    return acc;
}

// A u at grid point (x, y, z)
template <int H>
__device__ __forceinline__ double ho_apply(const HoOp& op, const Geom& g, const double* u, int x, int y, int z) {
    const bool per = g.periodic;
    double acc = ho_line<H>(op, per, g.n, u + (long long)z * g.sz + (long long)y * g.sy, x, 1);
    if (g.ndim >= 2) acc += ho_line<H>(op, per, g.n, u + (long long)z * g.sz + x, y, g.sy);
    if (g.ndim == 3) acc += ho_line<H>(op, per, g.n, u + (long long)y * g.sy + x, z, g.sz);
    return acc;
}

// loop over the grid points of this thread: f(idx, x, y, z).  The (x, y, z) of consecutive iterations differ by the
// grid stride; they are advanced with carries instead of 64-bit divisions per point.
template <class F>
__device__ __forceinline__ void ho_points(const Geom& g, F&& f) {
    const int n = g.n;
    const long long n2 = (long long)n * n;
    const long long npts = (long long)n * (g.ndim >= 2 ? n : 1) * (g.ndim == 3 ? n : 1);
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= npts) return;
    int x = (int)(t % n), y = (int)((t / n) % n), z = (int)(t / n2);
    const int sx = (int)(stride % n), sy = (int)((stride / n) % n), sz = (int)(stride / n2);
    for (; t < npts; t += stride) {
        f((long long)z * g.sz + (long long)y * g.sy + x, x, y, z);
        x += sx;
        if (x >= n) { x -= n; ++y; }
        y += sy;
        if (y >= n) { y -= n; ++z; }
        z += sz;
    }
}

// Streaming passes (axpy / dot-product sweeps without neighbours) run flat over the whole volume, walls included - every
// vector of these solvers has zero walls and the passes map zeros to zeros - four doubles per thread and step, all loads
// of a step issued before its first store: f(i, full) with i the first of four consecutive doubles, `full` false on a
// trailing half quad (1-D grids whose pitch is not a multiple of four).
struct Quad {
    double2 a, b;
};
__device__ __forceinline__ Quad ldq(const double* p, long long i, bool full) {
    Quad q;
    q.a = ld2(p + i);
    q.b = full ? ld2(p + i + 2) : make_double2(0.0, 0.0);
    return q;
}
__device__ __forceinline__ void stq(double* p, long long i, const Quad& q, bool full) {
    st2(p + i, q.a);
    if (full) st2(p + i + 2, q.b);
}
template <class F>
__device__ __forceinline__ void flat_quads(long long vol, F&& f) {
    const long long nq = (vol + 3) / 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nq; j += stride) f(4 * j, 4 * j + 4 <= vol);
}
// element-wise helpers on quads (component e = 0..3)
__device__ __forceinline__ double qe(const Quad& q, int e) { return e == 0 ? q.a.x : e == 1 ? q.a.y : e == 2 ? q.b.x : q.b.y; }
template <class G>
__device__ __forceinline__ Quad qmap(G&& g) {
    Quad o;
    o.a.x = g(0);
    o.a.y = g(1);
    o.b.x = g(2);
    o.b.y = g(3);
    return o;
}
// acc += sum_e x[e] * y[e], one fma after the other
__device__ __forceinline__ double qdot(const Quad& x, const Quad& y, double acc) {
    acc = fma(x.a.x, y.a.x, acc);
    acc = fma(x.a.y, y.a.y, acc);
    acc = fma(x.b.x, y.b.x, acc);
    return fma(x.b.y, y.b.y, acc);
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

// dispatch on the stencil half width: CALL is an expression / statement that uses the constant HW
#define SDC_DISPATCH_H(h, CALL)                        \
    switch (h) {                                       \
        case 1: { constexpr int HW = 1; CALL; } break; \
        case 2: { constexpr int HW = 2; CALL; } break; \
        case 3: { constexpr int HW = 3; CALL; } break; \
        default: { constexpr int HW = 4; CALL; } break; \
    }

inline bool ok8(const void* p) { return p != nullptr && (reinterpret_cast<size_t>(p) & 15u) == 0; }  // (double2 passes)
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
constexpr int kMaxGrid = 148 * 8;

}  // namespace sdcb200
