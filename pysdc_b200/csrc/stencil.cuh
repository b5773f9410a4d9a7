// Matrix-free order-2 Laplacian stencil on the walled layout (include/sdc_b200.h), shared by eval_f, the CG
// operator application and the Newton residual.
//
// Work decomposition: a "unit" is a tile of 64(x) x 8(y) points [x chunk_z planes in 3-D, marched in z with the centre
// column held in registers]; 1-D grids use 512-point segments.  One CTA (256 threads = 8 warps, one warp per tile
// row, one double2 per lane) processes a unit; x-neighbours come from warp shuffles, y-neighbours from L1/L2 (rows of
// the same tile are loaded by the neighbouring warps of the same CTA), z-neighbours from registers.  On Dirichlet
// grids every neighbour access is in-bounds and reads an exact zero at the boundary (walls / guard), so the inner
// loop has no boundary branches; periodic grids wrap indices explicitly.
//
// The field the stencil is applied to is described by a LOADER (pair(idx) -> double2, one(idx) -> double): a plain
// array, or an expression of several arrays evaluated on the fly (the CG search direction  r + beta*p_old, see cg.cu),
// which is what lets the direction update and the operator application share one pass over memory.
#pragma once
#include "common.cuh"

namespace sdcb200 {

constexpr int kTileX = 64;   // doubles per tile row (32 lanes x double2)
constexpr int kTileY = 8;    // rows per tile (one per warp)
constexpr int kSeg1D = 2 * kThreads;

struct Units {
    int nxt, nyt, nzc;
    int chunk_z;  // planes marched per unit in 3-D
    int per_field;
};

// `target_ctas` > 0: choose the z-chunk so that the units of one field fill whole waves of a persistent grid of that
// many CTAs with the longest possible march (fewer re-read halo planes); 0: fixed 32-plane chunks.
__host__ __device__ inline Units make_units(const Geom& g, int target_ctas = 0) {
    Units u;
    u.chunk_z = 1;
    if (g.ndim == 1) {
        u.nxt = (g.P + kSeg1D - 1) / kSeg1D;
        u.nyt = 1;
        u.nzc = 1;
    } else {
        u.nxt = (g.P + kTileX - 1) / kTileX;
        u.nyt = (g.n + kTileY - 1) / kTileY;
        u.nzc = 1;
        if (g.ndim == 3) {
            const int nz = g.nz;
            int chunk = 32;
            if (target_ctas > 0) {
                // candidates 128, 64, 32, 16: take the longest march that still leaves >= 4 rounds of units per CTA
                // and wastes < 8 % of the last round
                const int tiles = u.nxt * u.nyt;
                chunk = 16;
                for (int c = 128; c >= 16; c >>= 1) {
                    const long long units = (long long)tiles * ((nz + c - 1) / c);
                    const long long rounds = (units + target_ctas - 1) / target_ctas;
                    if (rounds >= 4 && (double)units / (double)(rounds * target_ctas) > 0.92) {
                        chunk = c;
                        break;
                    }
                }
            }
            u.chunk_z = chunk;
            u.nzc = (nz + chunk - 1) / chunk;
        }
    }
    u.per_field = u.nxt * u.nyt * u.nzc;
    return u;
}

// ---- loaders ---------------------------------------------------------------------------------------------------------
struct PlainLoader {
    const double* u;
    __device__ __forceinline__ double2 pair(long long idx) const { return ld2(u + idx); }
    __device__ __forceinline__ double one(long long idx) const { return u[idx]; }
};

// search direction of CG evaluated on the fly:  r + beta * p_old  with the rounding of scipy's  `p *= beta; p += r`
// (p_old == nullptr: first iteration, p = r)
struct DirectionLoader {
    const double* r;
    const double* p_old;
    double beta;
    __device__ __forceinline__ double2 pair(long long idx) const {
        const double2 rv = ld2(r + idx);
        if (p_old == nullptr) return rv;
        const double2 pv = ld2(p_old + idx);
        return make_double2(__dadd_rn(__dmul_rn(pv.x, beta), rv.x), __dadd_rn(__dmul_rn(pv.y, beta), rv.y));
    }
    __device__ __forceinline__ double one(long long idx) const {
        const double rv = r[idx];
        if (p_old == nullptr) return rv;
        return __dadd_rn(__dmul_rn(p_old[idx], beta), rv);
    }
};

struct NoHalo {
    __device__ __forceinline__ void operator()(long long, double2) const {}
};

// Visit every grid point of one unit.  f(idx, c, nb, v0, v1): idx = flat index of the pair (x, x+1), c = centre
// values, nb = sum of the 2*NDIM neighbours of each, v0/v1 = whether x / x+1 are grid points (false on the wall).
// The arrays behind `ld` must not be written by anybody while the phase that calls this runs.
// 3-D slabs (g.zhalo): planes -1 and nz hold the neighbouring slab's boundary planes (or zeros at the domain
// boundary); h(idx, c) is called with the loader's value on those two planes for the units that touch them, so that
// an on-the-fly field can be materialised there as well.
template <int NDIM, bool PER, class L, class F, class H = NoHalo>
__device__ __forceinline__ void stencil_unit_ld(const Geom& g, const Units& U, const L& ld, int unit, F&& f,
                                                H&& h = H()) {
    const int lane = threadIdx.x & 31;
    const int n = g.n, P = g.P;
    int x, y = 0, z0 = 0, z1 = 1;
    if constexpr (NDIM == 1) {
        x = unit * kSeg1D + 2 * (int)threadIdx.x;
    } else {
        const int tx = unit % U.nxt;
        const int ty = (unit / U.nxt) % U.nyt;
        x = tx * kTileX + 2 * lane;
        y = ty * kTileY + (threadIdx.x >> 5);
        if constexpr (NDIM == 3) {
            const int tz = unit / (U.nxt * U.nyt);
            z0 = tz * U.chunk_z;
            z1 = min(z0 + U.chunk_z, g.nz);
        }
        if (y >= n) return;  // warp-uniform: whole warp owns a row outside the grid
    }
    const bool inx = x < P;              // pair inside the padded row
    const bool v0 = x < n, v1 = x + 1 < n;
    const int xs = inx ? x : 0;          // clamp so that idle lanes still form valid addresses
    const long long row = (NDIM >= 2 ? (long long)y * g.sy : 0);
    // in-plane neighbour offsets (periodic wrap resolved once per unit)
    long long up = -g.sy, dn = g.sy;
    if constexpr (PER && NDIM >= 2) {
        if (y == 0) up = (long long)(n - 1) * g.sy;
        if (y == n - 1) dn = -(long long)(n - 1) * g.sy;
    }
    const bool need_left = (lane == 0);
    const bool need_right = (lane == 31) || (x + 2 >= P);
    const bool zwrap = PER && NDIM == 3 && !g.zhalo;

    long long idx = row + xs + (NDIM == 3 ? (long long)z0 * g.sz : 0);
    double2 c_prev = make_double2(0.0, 0.0), c_next = make_double2(0.0, 0.0), c_next2 = make_double2(0.0, 0.0);
    double2 c = ld.pair(idx);
    // plane z lives at offset plane_off(z) from plane z0 (periodic grids wrap; slabs and Dirichlet grids have real
    // planes at -1 and nz: halo / guard / wall)
    auto plane_off = [&](int z) -> long long {
        if (zwrap) z = z < 0 ? z + g.nz : (z >= g.nz ? z - g.nz : z);
        return (long long)(z - z0) * g.sz;
    };
    const long long base = idx;
    if constexpr (NDIM == 3) {
        c_prev = ld.pair(base + plane_off(z0 - 1));
        if (g.zhalo && z0 == 0 && inx) h(base + plane_off(-1), c_prev);
        c_next = ld.pair(base + plane_off(z0 + 1));
    }
    for (int z = z0; z < z1; ++z) {
        if constexpr (NDIM == 3) {
            // the plane after next is requested one step early: two DRAM-bound loads in flight per warp instead of one
            if (z + 1 < z1) c_next2 = ld.pair(base + plane_off(z + 2));
        }
        // x direction: shuffles inside the warp, two edge lanes load from the neighbouring tile / wrap around
        double left = __shfl_up_sync(0xffffffffu, c.y, 1);
        double right = __shfl_down_sync(0xffffffffu, c.x, 1);
        if (need_left) {
            if (PER && x == 0) left = ld.one(idx + (n - 1));
            else left = ld.one(idx - 1);  // Dirichlet x == 0: previous row's wall (or the guard) = 0
        }
        if (need_right) {
            if (x + 2 < P) right = ld.one(idx + 2);
            else right = PER ? ld.one(idx - xs) : 0.0;  // wrap to x = 0 / beyond the wall
        }
        double2 nb = make_double2(left + c.y, c.x + right);
        if constexpr (NDIM >= 2) {
            const double2 a = ld.pair(idx + up), b = ld.pair(idx + dn);
            nb.x += a.x + b.x;
            nb.y += a.y + b.y;
        }
        if constexpr (NDIM == 3) {
            nb.x += c_prev.x + c_next.x;
            nb.y += c_prev.y + c_next.y;
        }
        if (inx) f(idx, c, nb, v0, v1);
        if constexpr (NDIM == 3) {
            c_prev = c;
            c = c_next;
            c_next = c_next2;
            idx += g.sz;
        }
    }
    if constexpr (NDIM == 3) {
        if (g.zhalo && z1 == g.nz && inx) h(idx, c);  // c now holds plane nz
    }
}

// plain-array convenience overload
template <int NDIM, bool PER, class F>
__device__ __forceinline__ void stencil_unit(const Geom& g, const Units& U, const double* u, int unit, F&& f) {
    stencil_unit_ld<NDIM, PER>(g, U, PlainLoader{u}, unit, static_cast<F&&>(f));
}

}  // namespace sdcb200
