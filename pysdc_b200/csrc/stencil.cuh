// Matrix-free order-2 Laplacian stencil on the walled layout (include/sdc_b200.h), shared by eval_f, the CG
// operator application and the Newton residual.
//
// Work decomposition: a "unit" is a tile of 64(x) x 8(y) points [x ZC planes in 3-D, marched in z with the centre
// column held in registers]; 1-D grids use 512-point segments.  One CTA (256 threads = 8 warps, one warp per tile
// row, one double2 per lane) processes a unit; x-neighbours come from warp shuffles, y-neighbours from L1/L2 (rows of
// the same tile are loaded by the neighbouring warps of the same CTA), z-neighbours from registers.  On Dirichlet
// grids every neighbour access is in-bounds and reads an exact zero at the boundary (walls / guard), so the inner
// loop has no boundary branches; periodic grids wrap indices explicitly.
#pragma once
#include "common.cuh"

namespace sdcb200 {

constexpr int kTileX = 64;   // doubles per tile row (32 lanes x double2)
constexpr int kTileY = 8;    // rows per tile (one per warp)
constexpr int kChunkZ = 32;  // planes marched per unit in 3-D
constexpr int kSeg1D = 2 * kThreads;

struct Units {
    int nxt, nyt, nzc;
    int per_field;
};

__host__ __device__ inline Units make_units(const Geom& g) {
    Units u;
    if (g.ndim == 1) {
        u.nxt = (g.P + kSeg1D - 1) / kSeg1D;
        u.nyt = 1;
        u.nzc = 1;
    } else {
        u.nxt = (g.P + kTileX - 1) / kTileX;
        u.nyt = (g.n + kTileY - 1) / kTileY;
        u.nzc = g.ndim == 3 ? (g.n + kChunkZ - 1) / kChunkZ : 1;
    }
    u.per_field = u.nxt * u.nyt * u.nzc;
    return u;
}

// Visit every grid point of one unit.  f(idx, c, nb, v0, v1): idx = flat index of the pair (x, x+1), c = centre
// values, nb = sum of the 2*NDIM neighbours of each, v0/v1 = whether x / x+1 are grid points (false on the wall).
// `u` must not be written by anybody while the phase that calls this runs.
template <int NDIM, bool PER, class F>
__device__ __forceinline__ void stencil_unit(const Geom& g, const Units& U, const double* u, int unit, F&& f) {
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
            z0 = tz * kChunkZ;
            z1 = min(z0 + kChunkZ, n);
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

    long long idx = row + xs + (NDIM == 3 ? (long long)z0 * g.sz : 0);
    double2 c_prev = make_double2(0.0, 0.0), c_next = make_double2(0.0, 0.0);
    double2 c = ld2(u + idx);
    if constexpr (NDIM == 3) {
        long long below = -g.sz;
        if constexpr (PER) {
            if (z0 == 0) below = (long long)(n - 1) * g.sz;
        }
        c_prev = ld2(u + idx + below);
    }
    for (int z = z0; z < z1; ++z) {
        if constexpr (NDIM == 3) {
            long long above = g.sz;
            if constexpr (PER) {
                if (z == n - 1) above = -(long long)(n - 1) * g.sz;
            }
            c_next = ld2(u + idx + above);
        }
        // x direction: shuffles inside the warp, two edge lanes load from the neighbouring tile / wrap around
        double left = __shfl_up_sync(0xffffffffu, c.y, 1);
        double right = __shfl_down_sync(0xffffffffu, c.x, 1);
        if (need_left) {
            if (PER && x == 0) left = u[idx + (n - 1)];
            else left = u[idx - 1];  // Dirichlet x == 0: previous row's wall (or the guard) = 0
        }
        if (need_right) {
            if (x + 2 < P) right = u[idx + 2];
            else right = PER ? u[idx - xs] : 0.0;  // wrap to x = 0 / beyond the wall
        }
        double2 nb = make_double2(left + c.y, c.x + right);
        if constexpr (NDIM >= 2) {
            const double2 a = ld2(u + idx + up), b = ld2(u + idx + dn);
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
            idx += g.sz;
        }
    }
}

}  // namespace sdcb200
