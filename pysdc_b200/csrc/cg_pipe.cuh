// The two passes of a CG iteration on 2-D / 3-D Dirichlet grids as bulk-async (TMA) pipelines through shared memory.
//
// Why: the register-marching stencil (stencil.cuh) can only keep the loads of ONE plane per warp in flight, and only a
// third of those go to DRAM (the others are neighbour rows served by L1/L2) - about 32 KB per SM, where HBM3e needs
// 50-60 KB per SM in flight to stay busy (ncu: 50 % DRAM utilisation, long-scoreboard stalls).  Here a dedicated
// producer warp streams whole tile planes global -> shared with TMA tensor copies (cp.async.bulk.tensor, SASS
// UTMALDG: ONE request per 18x68 / 16x64 box - row-wise cp.async.bulk copies were measured to be limited by the TMA
// unit's request rate, ~70 cycles per request per SM), completion on mbarriers, kPipeStages planes deep, so the bytes
// in flight (2 CTAs x 2-3 planes x 19-26 KB per SM) no longer depend on registers or occupancy; 8 consumer warps read
// the staged planes from shared memory and never wait on DRAM.
//
//   tile      64 (x) x 16 (y) points of one z-plane (+ 1 halo row above/below, + 2 halo columns left/right so that
//             every row copy is 16-byte aligned); a unit marches chunk_z planes, z-neighbours live in registers
//   warp w    rows 2w and 2w+1 of the tile (adjacent rows: each is the other's y-neighbour), one double2 per lane;
//             x-neighbours by warp shuffle, tile-edge columns and the rows above/below from the staged plane
//   phase A   stage = rows of r and p_old;  p = r + beta*p_old is evaluated wherever the stencil needs it
//             (centre, row above, row below, edge columns) -> q = M p -> p.q ;  p is stored from registers
//   phase B   stage = rows of p (with halo), r, x;  r -= alpha*M p, x += alpha*p, r.r ;  r, x stored from registers
//   steps     a CTA walks its units plane by plane: planes z0-1 .. z1 of a unit are one step each (the first and
//             last only feed the z-neighbours); the producer runs up to kPipeStages-1 steps ahead, across unit and
//             system boundaries; a stage is released (empty mbarrier, one arrival per consumer warp) when the plane
//             after it has been processed
// Dirichlet only: halo reads either hit the zero walls / guard of the walled layout or fall outside the tensor map and
// are zero-filled by the TMA unit; no wrap-around copies.  Tensor maps (one per field and box shape) are encoded on the
// host per solve and passed as __grid_constant__ kernel parameters.
#pragma once
#include <cuda.h>

#include "cg_common.cuh"

namespace sdcb200 {

constexpr int kPX = 64;             // tile width in doubles
constexpr int kPY = 16;             // tile rows
constexpr int kPHX = kPX + 4;       // staged row with halo: cols 0,1 = x0-2, x0-1 | 2..65 tile | 66,67 = x0+64, x0+65
constexpr int kPHY = kPY + 2;       // staged rows with halo: row 0 = y0-1, rows 1..16 tile, row 17 = y0+16
constexpr int kPipeStages = 4;
constexpr int kPipeConsumers = 8;   // consumer warps
constexpr int kPipeThreads = 32 * (kPipeConsumers + 1);

constexpr int kHaloBoxBytes = kPHY * kPHX * 8;   // 9792
constexpr int kCentreBoxBytes = kPY * kPX * 8;   // 8192
constexpr int kHaloPad = (128 - kHaloBoxBytes % 128) % 128 / 8;  // doubles: every box starts 128-byte aligned in smem

struct StageA {
    double R[kPHY][kPHX];
    double pad0[kHaloPad];
    double P[kPHY][kPHX];
    double pad1[kHaloPad];
};
struct StageB {
    double P[kPHY][kPHX];
    double pad0[kHaloPad];
    double R[kPY][kPX];
    double X[kPY][kPX];
};
static_assert(offsetof(StageA, P) % 128 == 0 && offsetof(StageB, R) % 128 == 0 && offsetof(StageB, X) % 128 == 0,
              "TMA box destinations must be 128-byte aligned");

// tensor maps of one system: boxes with halo of r and both direction buffers, tile-only boxes of r and x
enum { kMapRHalo = 0, kMapPHalo, kMapQHalo, kMapRCentre, kMapXCentre, kMapZHalo, kMapsPerSys };
struct PipeMaps {
    CUtensorMap m[SDCB200_MAX_NODES][kMapsPerSys];
};
union alignas(128) PipeStage {
    StageA a;
    StageB b;
};
struct PipeSmem {
    PipeStage st[kPipeStages];
    unsigned long long full[kPipeStages];
    unsigned long long empty[kPipeStages];
    double wsum[SDCB200_MAX_NODES][kPipeConsumers];
    int act_list[SDCB200_MAX_NODES];
    int nact;
};

struct PUnits {
    int nxt, nyt, nzc, chunk_z, per_field;
};

// Tile grid of the pipelined passes; the z-chunk is the longest march whose units (pooled over the B systems) still
// fill whole waves of the persistent grid.
__host__ __device__ inline PUnits make_punits(const Geom& g, int B, int ctas) {
    PUnits u;
    u.nxt = (g.P + kPX - 1) / kPX;
    u.nyt = (g.n + kPY - 1) / kPY;
    u.chunk_z = 1;
    u.nzc = 1;
    if (g.ndim == 3) {
        const int tiles = u.nxt * u.nyt;
        int chunk = 16;
        for (int c = 128; c >= 16; c >>= 1) {
            const long long units = (long long)B * tiles * ((g.nz + c - 1) / c);
            const long long rounds = (units + ctas - 1) / ctas;
            if (rounds >= 3 && (double)units / (double)(rounds * ctas) > 0.92) {
                chunk = c;
                break;
            }
        }
        u.chunk_z = chunk;
        u.nzc = (g.nz + chunk - 1) / chunk;
    }
    u.per_field = u.nxt * u.nyt * u.nzc;
    return u;
}

// ---- PTX wrappers: mbarrier + bulk async copy ---------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// one TMA box: smem <- tensor map at element coordinates (x, y[, z]); out-of-range elements are zero-filled
__device__ __forceinline__ void tma_box_2d(void* dst_smem, const CUtensorMap* map, int x, int y, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst_smem)),
        "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_box_3d(void* dst_smem, const CUtensorMap* map, int x, int y, int z,
                                           unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
            "r"(smem_u32(dst_smem)),
        "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}
// order this thread's generic-proxy global accesses against async-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__device__ __forceinline__ void pipe_smem_init(PipeSmem& sm) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < kPipeStages; ++i) {
            mbar_init(&sm.full[i], 1);
            mbar_init(&sm.empty[i], kPipeConsumers);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}

// ---- walking the plane-steps of this CTA -----------------------------------------------------------------------------
struct StepCursor {
    int u;    // unit index within the system, strided by gridDim.x (the same CTA owns the same units of every system,
              // so a system's reduction tree does not depend on which other systems are still iterating)
    int a;    // position in the active-system list
    int x0, y0, z0, z1;
    int zp;   // plane of this step: z0-1 .. z1 in 3-D, 0 in 2-D
    bool valid;
};
template <int NDIM>
__device__ __forceinline__ void cursor_load_unit(StepCursor& c, const PUnits& U, int nact) {
    while (c.u >= U.per_field && c.a + 1 < nact) {  // next system
        c.u = blockIdx.x;
        ++c.a;
    }
    c.valid = c.u < U.per_field;
    if (!c.valid) return;
    int rem = c.u;
    const int tx = rem % U.nxt;
    rem /= U.nxt;
    const int ty = rem % U.nyt;
    const int tz = rem / U.nyt;
    c.x0 = tx * kPX;
    c.y0 = ty * kPY;
    c.z0 = 0;
    c.z1 = 1;
    c.zp = 0;
    if (NDIM == 3) {
        c.z0 = tz * U.chunk_z;
        c.zp = c.z0 - 1;
    }
}
template <int NDIM>
__device__ __forceinline__ void cursor_init(StepCursor& c, const PUnits& U, const Geom& g, int nact) {
    c.u = blockIdx.x;
    c.a = 0;
    cursor_load_unit<NDIM>(c, U, nact);
    if (NDIM == 3 && c.valid) c.z1 = min(c.z0 + U.chunk_z, g.nz);
}
template <int NDIM>
__device__ __forceinline__ void cursor_next(StepCursor& c, const PUnits& U, const Geom& g, int nact) {
    if (NDIM == 3 && c.zp < c.z1) {
        ++c.zp;
        return;
    }
    c.u += gridDim.x;
    cursor_load_unit<NDIM>(c, U, nact);
    if (NDIM == 3 && c.valid) c.z1 = min(c.z0 + U.chunk_z, g.nz);
}

// Producer (one lane): the TMA boxes of one plane-step.  `halo` boxes carry the ring around the tile, `centre` boxes
// the tile only and are skipped on the two planes that merely feed the z-neighbours.  3-D maps start one plane below
// the field (the guard / lower halo plane), hence z + 1.
template <int NDIM>
__device__ __forceinline__ void pipe_issue(const StepCursor& c, const CUtensorMap* const* halo, int nh, void* const* halo_dst,
                                           const CUtensorMap* const* centre, int nc, void* const* centre_dst,
                                           unsigned long long* bar) {
    const bool halo_plane = NDIM == 3 && (c.zp < c.z0 || c.zp >= c.z1);
    if (halo_plane) nc = 0;
    mbar_arrive_expect_tx(bar, (unsigned)(nh * kHaloBoxBytes + nc * kCentreBoxBytes));
    for (int f = 0; f < nh; ++f) {
        if (NDIM == 3) tma_box_3d(halo_dst[f], halo[f], c.x0 - 2, c.y0 - 1, c.zp + 1, bar);
        else tma_box_2d(halo_dst[f], halo[f], c.x0 - 2, c.y0 - 1, bar);
    }
    for (int f = 0; f < nc; ++f) {
        if (NDIM == 3) tma_box_3d(centre_dst[f], centre[f], c.x0, c.y0, c.zp + 1, bar);
        else tma_box_2d(centre_dst[f], centre[f], c.x0, c.y0, bar);
    }
}

__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ double2 dir2(double2 r, double2 p, double beta) {
    return make_double2(__dadd_rn(__dmul_rn(p.x, beta), r.x), __dadd_rn(__dmul_rn(p.y, beta), r.y));
}
__device__ __forceinline__ double dir1(double r, double p, double beta) { return __dadd_rn(__dmul_rn(p, beta), r); }

// per-system partial sums of one phase: warp partials -> shared -> (after the CTA barrier) CTA partial in global
__device__ __forceinline__ void pipe_flush(PipeSmem& sm, int a, double v) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sm.wsum[a][threadIdx.x >> 5] = v;
}
__device__ __forceinline__ void pipe_publish(PipeSmem& sm, double* partials, int slot) {
    __syncthreads();
    if ((int)threadIdx.x < sm.nact) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kPipeConsumers; ++w) t += sm.wsum[threadIdx.x][w];
        partials[(size_t)(slot * SDCB200_MAX_NODES + sm.act_list[threadIdx.x]) * gridDim.x + blockIdx.x] = t;
    }
}
__device__ __forceinline__ void pipe_begin(PipeSmem& sm) {
    if (threadIdx.x < SDCB200_MAX_NODES * kPipeConsumers) (&sm.wsum[0][0])[threadIdx.x] = 0.0;
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------------
// phase A:  p = r + beta p_old  (on the fly),  q = M p,  p.q      `first`: p = r, p_old is not read
// ---------------------------------------------------------------------------------------------------------------------
template <int NDIM>
__device__ void pipe_phase_a(const Geom& g, const PUnits& U, const Sys* s, const PipeMaps& maps, bool first, int cur,
                             const CgShared& sh, PipeSmem& sm, double* partials, unsigned& kstep) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nact = sm.nact;
    const int P = g.P, n = g.n;
    pipe_begin(sm);
    StepCursor c;
    cursor_init<NDIM>(c, U, g, nact);
    unsigned k = kstep;
    if (warp == kPipeConsumers) {
        // ---- producer ---------------------------------------------------------------------------------------------
        while (c.valid) {
            const unsigned stg = k % kPipeStages;
            if (k >= kPipeStages) mbar_wait(&sm.empty[stg], ((k / kPipeStages) - 1u) & 1u);
            if (lane == 0) {
                const CUtensorMap* mp = maps.m[sm.act_list[c.a]];
                StageA& A = sm.st[stg].a;
                // r (or the preconditioned residual z), p_old
                const CUtensorMap* hsrc[2] = {mp + (s[sm.act_list[c.a]].z != nullptr ? kMapZHalo : kMapRHalo),
                                              mp + (cur ? kMapPHalo : kMapQHalo)};
                void* hdst[2] = {A.R, A.P};
                pipe_issue<NDIM>(c, hsrc, first ? 1 : 2, hdst, nullptr, 0, nullptr, &sm.full[stg]);
            }
            ++k;
            cursor_next<NDIM>(c, U, g, nact);
        }
    } else {
        // ---- consumers --------------------------------------------------------------------------------------------
        const int ra = 2 * warp, col = 2 + 2 * lane;
        double2 cprev0 = make_double2(0.0, 0.0), cprev1 = cprev0, cc0 = cprev0, cc1 = cprev0;
        double pq = 0.0;
        int cur_a = -1;
        double beta = 0.0, m_diag = 0.0, m_off = 0.0;
        double* p_new = nullptr;
        while (c.valid) {
            if (c.a != cur_a) {
                if (cur_a >= 0) pipe_flush(sm, cur_a, pq);
                cur_a = c.a;
                pq = 0.0;
                const int b = sm.act_list[c.a];
                beta = sh.beta[b];
                m_diag = s[b].m_diag;
                m_off = s[b].m_off;
                p_new = cur ? s[b].q : s[b].p;
            }
            const unsigned stg = k % kPipeStages;
            mbar_wait(&sm.full[stg], (k / kPipeStages) & 1u);
            const StageA& A = sm.st[stg].a;
            const int x = c.x0 + 2 * lane, ya = c.y0 + ra;
            const bool inx = x < P;
            // search direction on the tile rows of plane zp
            double2 v0 = lds2(&A.R[1 + ra][col]), v1 = lds2(&A.R[2 + ra][col]);
            if (!first) {
                v0 = dir2(v0, lds2(&A.P[1 + ra][col]), beta);
                v1 = dir2(v1, lds2(&A.P[2 + ra][col]), beta);
            }
            if (NDIM == 2 || c.zp > c.z0) {
                // plane zc = zp-1 (3-D) / this plane (2-D) has all its neighbours now
                const StageA& A0 = NDIM == 3 ? sm.st[(k - 1u) % kPipeStages].a : A;
                const double2 ca = NDIM == 3 ? cc0 : v0, cb = NDIM == 3 ? cc1 : v1;
                double2 up = lds2(&A0.R[ra][col]), dn = lds2(&A0.R[ra + 3][col]);
                if (!first) {
                    up = dir2(up, lds2(&A0.P[ra][col]), beta);
                    dn = dir2(dn, lds2(&A0.P[ra + 3][col]), beta);
                }
                double la = __shfl_up_sync(0xffffffffu, ca.y, 1), ra_ = __shfl_down_sync(0xffffffffu, ca.x, 1);
                double lb = __shfl_up_sync(0xffffffffu, cb.y, 1), rb_ = __shfl_down_sync(0xffffffffu, cb.x, 1);
                if (lane == 0) {
                    la = A0.R[1 + ra][1];
                    lb = A0.R[2 + ra][1];
                    if (!first) {
                        la = dir1(la, A0.P[1 + ra][1], beta);
                        lb = dir1(lb, A0.P[2 + ra][1], beta);
                    }
                }
                if (lane == 31 || x + 2 >= P) {
                    if (x + 2 < P) {
                        ra_ = A0.R[1 + ra][kPX + 2];
                        rb_ = A0.R[2 + ra][kPX + 2];
                        if (!first) {
                            ra_ = dir1(ra_, A0.P[1 + ra][kPX + 2], beta);
                            rb_ = dir1(rb_, A0.P[2 + ra][kPX + 2], beta);
                        }
                    } else {
                        ra_ = rb_ = 0.0;  // beyond the wall
                    }
                }
                double2 nba = make_double2(la + ca.y, ca.x + ra_), nbb = make_double2(lb + cb.y, cb.x + rb_);
                nba.x += up.x + cb.x;
                nba.y += up.y + cb.y;
                nbb.x += ca.x + dn.x;
                nbb.y += ca.y + dn.y;
                if (NDIM == 3) {
                    nba.x += cprev0.x + v0.x;
                    nba.y += cprev0.y + v0.y;
                    nbb.x += cprev1.x + v1.x;
                    nbb.y += cprev1.y + v1.y;
                }
                const long long idx = (NDIM == 3 ? (long long)(c.zp - 1) * g.sz : 0) + (long long)ya * g.sy + x;
                const bool v0x = x < n, v1x = x + 1 < n;
                if (inx && ya < n) {
                    st2(p_new + idx, ca);  // wall column: r = p_old = 0 there, so the stored value is an exact zero
                    const double qx = v0x ? fma(m_off, nba.x, m_diag * ca.x) : 0.0;
                    const double qy = v1x ? fma(m_off, nba.y, m_diag * ca.y) : 0.0;
                    pq = fma(ca.x, qx, pq);
                    pq = fma(ca.y, qy, pq);
                }
                if (inx && ya + 1 < n) {
                    st2(p_new + idx + g.sy, cb);
                    const double qx = v0x ? fma(m_off, nbb.x, m_diag * cb.x) : 0.0;
                    const double qy = v1x ? fma(m_off, nbb.y, m_diag * cb.y) : 0.0;
                    pq = fma(cb.x, qx, pq);
                    pq = fma(cb.y, qy, pq);
                }
            }
            if (NDIM == 3 && g.zhalo && (c.zp < 0 || c.zp == g.nz)) {
                // slab halo planes: materialise p there too (bit-identical to what the neighbouring rank computes)
                const long long idx = (long long)c.zp * g.sz + (long long)ya * g.sy + x;
                if (inx && ya < n) st2(p_new + idx, v0);
                if (inx && ya + 1 < n) st2(p_new + idx + g.sy, v1);
            }
            __syncwarp();
            if (lane == 0) {
                if (NDIM == 3 && c.zp >= c.z0) mbar_arrive(&sm.empty[(k - 1u) % kPipeStages]);
                if (NDIM == 2 || c.zp == c.z1) mbar_arrive(&sm.empty[stg]);
            }
            cprev0 = cc0;
            cprev1 = cc1;
            cc0 = v0;
            cc1 = v1;
            ++k;
            cursor_next<NDIM>(c, U, g, nact);
        }
        if (cur_a >= 0) pipe_flush(sm, cur_a, pq);
    }
    kstep = k;
    pipe_publish(sm, partials, kSlotA);
}

// ---------------------------------------------------------------------------------------------------------------------
// phase B:  r -= alpha M p,  x += alpha p,  r.r
// ---------------------------------------------------------------------------------------------------------------------
template <int NDIM>
__device__ void pipe_phase_b(const Geom& g, const PUnits& U, const Sys* s, const PipeMaps& maps, int cur,
                             const CgShared& sh, PipeSmem& sm, double* partials, unsigned& kstep,
                             const SlabLink* link = nullptr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nact = sm.nact;
    const int P = g.P, n = g.n;
    pipe_begin(sm);
    StepCursor c;
    cursor_init<NDIM>(c, U, g, nact);
    unsigned k = kstep;
    if (warp == kPipeConsumers) {
        while (c.valid) {
            const unsigned stg = k % kPipeStages;
            if (k >= kPipeStages) mbar_wait(&sm.empty[stg], ((k / kPipeStages) - 1u) & 1u);
            if (lane == 0) {
                const CUtensorMap* mp = maps.m[sm.act_list[c.a]];
                StageB& Bq = sm.st[stg].b;
                const CUtensorMap* hsrc[1] = {mp + (cur ? kMapQHalo : kMapPHalo)};  // the new p
                void* hdst[1] = {Bq.P};
                const CUtensorMap* csrc[2] = {mp + kMapRCentre, mp + kMapXCentre};
                void* cdst[2] = {Bq.R, Bq.X};
                pipe_issue<NDIM>(c, hsrc, 1, hdst, csrc, 2, cdst, &sm.full[stg]);
            }
            ++k;
            cursor_next<NDIM>(c, U, g, nact);
        }
    } else {
        const int ra = 2 * warp, col = 2 + 2 * lane;
        double2 cprev0 = make_double2(0.0, 0.0), cprev1 = cprev0, cc0 = cprev0, cc1 = cprev0;
        double rr = 0.0;
        int cur_a = -1;
        double alpha = 0.0, m_diag = 0.0, m_off = 0.0;
        double *rp = nullptr, *xp = nullptr;
        double *r_lo = nullptr, *r_hi = nullptr;  // slab runs: the neighbours' halo planes of r
        while (c.valid) {
            if (c.a != cur_a) {
                if (cur_a >= 0) pipe_flush(sm, cur_a, rr);
                cur_a = c.a;
                rr = 0.0;
                const int b = sm.act_list[c.a];
                alpha = sh.alpha[b];
                m_diag = s[b].m_diag;
                m_off = s[b].m_off;
                rp = s[b].r;
                xp = s[b].x;
                if (link != nullptr) {
                    r_lo = link->has_lo ? link->lo_r_halo[b] : nullptr;
                    r_hi = link->has_hi ? link->hi_r_halo[b] : nullptr;
                }
            }
            const unsigned stg = k % kPipeStages;
            mbar_wait(&sm.full[stg], (k / kPipeStages) & 1u);
            const StageB& Bq = sm.st[stg].b;
            const int x = c.x0 + 2 * lane, ya = c.y0 + ra;
            const bool inx = x < P;
            const double2 v0 = lds2(&Bq.P[1 + ra][col]), v1 = lds2(&Bq.P[2 + ra][col]);
            if (NDIM == 2 || c.zp > c.z0) {
                const StageB& B0 = NDIM == 3 ? sm.st[(k - 1u) % kPipeStages].b : Bq;
                const double2 ca = NDIM == 3 ? cc0 : v0, cb = NDIM == 3 ? cc1 : v1;
                const double2 up = lds2(&B0.P[ra][col]), dn = lds2(&B0.P[ra + 3][col]);
                double la = __shfl_up_sync(0xffffffffu, ca.y, 1), ra_ = __shfl_down_sync(0xffffffffu, ca.x, 1);
                double lb = __shfl_up_sync(0xffffffffu, cb.y, 1), rb_ = __shfl_down_sync(0xffffffffu, cb.x, 1);
                if (lane == 0) {
                    la = B0.P[1 + ra][1];
                    lb = B0.P[2 + ra][1];
                }
                if (lane == 31 || x + 2 >= P) {
                    if (x + 2 < P) {
                        ra_ = B0.P[1 + ra][kPX + 2];
                        rb_ = B0.P[2 + ra][kPX + 2];
                    } else {
                        ra_ = rb_ = 0.0;
                    }
                }
                double2 nba = make_double2(la + ca.y, ca.x + ra_), nbb = make_double2(lb + cb.y, cb.x + rb_);
                nba.x += up.x + cb.x;
                nba.y += up.y + cb.y;
                nbb.x += ca.x + dn.x;
                nbb.y += ca.y + dn.y;
                if (NDIM == 3) {
                    nba.x += cprev0.x + v0.x;
                    nba.y += cprev0.y + v0.y;
                    nbb.x += cprev1.x + v1.x;
                    nbb.y += cprev1.y + v1.y;
                }
                const long long idx = (NDIM == 3 ? (long long)(c.zp - 1) * g.sz : 0) + (long long)ya * g.sy + x;
                const bool v0x = x < n, v1x = x + 1 < n;
                // slab boundary planes of r go straight into the neighbour's halo plane (peer memory over NVLink)
                double* push_lo = (NDIM == 3 && r_lo != nullptr && c.zp - 1 == 0) ? r_lo + (long long)ya * g.sy + x : nullptr;
                double* push_hi = (NDIM == 3 && r_hi != nullptr && c.zp == g.nz) ? r_hi + (long long)ya * g.sy + x : nullptr;
                if (inx && ya < n) {
                    double2 r = lds2(&B0.R[ra][2 * lane]), xv = lds2(&B0.X[ra][2 * lane]);
                    r.x = v0x ? __dsub_rn(r.x, __dmul_rn(alpha, fma(m_off, nba.x, m_diag * ca.x))) : 0.0;
                    r.y = v1x ? __dsub_rn(r.y, __dmul_rn(alpha, fma(m_off, nba.y, m_diag * ca.y))) : 0.0;
                    xv.x = __dadd_rn(xv.x, __dmul_rn(alpha, ca.x));
                    xv.y = __dadd_rn(xv.y, __dmul_rn(alpha, ca.y));
                    st2(rp + idx, r);
                    st2(xp + idx, xv);
                    if (push_lo != nullptr) st2(push_lo, r);
                    if (push_hi != nullptr) st2(push_hi, r);
                    rr = fma(r.x, r.x, rr);
                    rr = fma(r.y, r.y, rr);
                }
                if (inx && ya + 1 < n) {
                    double2 r = lds2(&B0.R[ra + 1][2 * lane]), xv = lds2(&B0.X[ra + 1][2 * lane]);
                    r.x = v0x ? __dsub_rn(r.x, __dmul_rn(alpha, fma(m_off, nbb.x, m_diag * cb.x))) : 0.0;
                    r.y = v1x ? __dsub_rn(r.y, __dmul_rn(alpha, fma(m_off, nbb.y, m_diag * cb.y))) : 0.0;
                    xv.x = __dadd_rn(xv.x, __dmul_rn(alpha, cb.x));
                    xv.y = __dadd_rn(xv.y, __dmul_rn(alpha, cb.y));
                    st2(rp + idx + g.sy, r);
                    st2(xp + idx + g.sy, xv);
                    if (push_lo != nullptr) st2(push_lo + g.sy, r);
                    if (push_hi != nullptr) st2(push_hi + g.sy, r);
                    rr = fma(r.x, r.x, rr);
                    rr = fma(r.y, r.y, rr);
                }
            }
            __syncwarp();
            if (lane == 0) {
                if (NDIM == 3 && c.zp >= c.z0) mbar_arrive(&sm.empty[(k - 1u) % kPipeStages]);
                if (NDIM == 2 || c.zp == c.z1) mbar_arrive(&sm.empty[stg]);
            }
            cprev0 = cc0;
            cprev1 = cc1;
            cc0 = v0;
            cc1 = v1;
            ++k;
            cursor_next<NDIM>(c, U, g, nact);
        }
        if (cur_a >= 0) pipe_flush(sm, cur_a, rr);
    }
    kstep = k;
    pipe_publish(sm, partials, kSlotB);
}

// ---------------------------------------------------------------------------------------------------------------------
// phase C (preconditioned runs):  z = pc_a r + pc_b M r,  r.z      stage = rows of r with the halo ring
// ---------------------------------------------------------------------------------------------------------------------
template <int NDIM>
__device__ void pipe_phase_c(const Geom& g, const PUnits& U, const Sys* s, const PipeMaps& maps, PipeSmem& sm,
                             double* partials, unsigned& kstep, const SlabLink* link = nullptr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nact = sm.nact;
    const int P = g.P, n = g.n;
    pipe_begin(sm);
    StepCursor c;
    cursor_init<NDIM>(c, U, g, nact);
    unsigned k = kstep;
    if (warp == kPipeConsumers) {
        while (c.valid) {
            const unsigned stg = k % kPipeStages;
            if (k >= kPipeStages) mbar_wait(&sm.empty[stg], ((k / kPipeStages) - 1u) & 1u);
            if (lane == 0) {
                const CUtensorMap* mp = maps.m[sm.act_list[c.a]];
                const CUtensorMap* hsrc[1] = {mp + kMapRHalo};
                void* hdst[1] = {sm.st[stg].a.R};
                pipe_issue<NDIM>(c, hsrc, 1, hdst, nullptr, 0, nullptr, &sm.full[stg]);
            }
            ++k;
            cursor_next<NDIM>(c, U, g, nact);
        }
    } else {
        const int ra = 2 * warp, col = 2 + 2 * lane;
        double2 cprev0 = make_double2(0.0, 0.0), cprev1 = cprev0, cc0 = cprev0, cc1 = cprev0;
        double rz = 0.0;
        int cur_a = -1;
        double m_diag = 0.0, m_off = 0.0, pa = 0.0, pb = 0.0;
        double* zp = nullptr;
        double *z_lo = nullptr, *z_hi = nullptr;
        while (c.valid) {
            if (c.a != cur_a) {
                if (cur_a >= 0) pipe_flush(sm, cur_a, rz);
                cur_a = c.a;
                rz = 0.0;
                const int b = sm.act_list[c.a];
                m_diag = s[b].m_diag;
                m_off = s[b].m_off;
                pa = s[b].pc_a;
                pb = s[b].pc_b;
                zp = s[b].z;
                if (link != nullptr) {
                    z_lo = link->has_lo ? link->lo_z_halo[b] : nullptr;
                    z_hi = link->has_hi ? link->hi_z_halo[b] : nullptr;
                }
            }
            const unsigned stg = k % kPipeStages;
            mbar_wait(&sm.full[stg], (k / kPipeStages) & 1u);
            const StageA& A = sm.st[stg].a;
            const int x = c.x0 + 2 * lane, ya = c.y0 + ra;
            const bool inx = x < P;
            const double2 v0 = lds2(&A.R[1 + ra][col]), v1 = lds2(&A.R[2 + ra][col]);
            if (NDIM == 2 || c.zp > c.z0) {
                const StageA& A0 = NDIM == 3 ? sm.st[(k - 1u) % kPipeStages].a : A;
                const double2 ca = NDIM == 3 ? cc0 : v0, cb = NDIM == 3 ? cc1 : v1;
                const double2 up = lds2(&A0.R[ra][col]), dn = lds2(&A0.R[ra + 3][col]);
                double la = __shfl_up_sync(0xffffffffu, ca.y, 1), ra_ = __shfl_down_sync(0xffffffffu, ca.x, 1);
                double lb = __shfl_up_sync(0xffffffffu, cb.y, 1), rb_ = __shfl_down_sync(0xffffffffu, cb.x, 1);
                if (lane == 0) {
                    la = A0.R[1 + ra][1];
                    lb = A0.R[2 + ra][1];
                }
                if (lane == 31 || x + 2 >= P) {
                    if (x + 2 < P) {
                        ra_ = A0.R[1 + ra][kPX + 2];
                        rb_ = A0.R[2 + ra][kPX + 2];
                    } else {
                        ra_ = rb_ = 0.0;
                    }
                }
                double2 nba = make_double2(la + ca.y, ca.x + ra_), nbb = make_double2(lb + cb.y, cb.x + rb_);
                nba.x += up.x + cb.x;
                nba.y += up.y + cb.y;
                nbb.x += ca.x + dn.x;
                nbb.y += ca.y + dn.y;
                if (NDIM == 3) {
                    nba.x += cprev0.x + v0.x;
                    nba.y += cprev0.y + v0.y;
                    nbb.x += cprev1.x + v1.x;
                    nbb.y += cprev1.y + v1.y;
                }
                const long long idx = (NDIM == 3 ? (long long)(c.zp - 1) * g.sz : 0) + (long long)ya * g.sy + x;
                const bool v0x = x < n, v1x = x + 1 < n;
                double* push_lo = (NDIM == 3 && z_lo != nullptr && c.zp - 1 == 0) ? z_lo + (long long)ya * g.sy + x : nullptr;
                double* push_hi = (NDIM == 3 && z_hi != nullptr && c.zp == g.nz) ? z_hi + (long long)ya * g.sy + x : nullptr;
                if (inx && ya < n) {
                    double2 z;
                    z.x = v0x ? fma(pb, fma(m_off, nba.x, m_diag * ca.x), pa * ca.x) : 0.0;
                    z.y = v1x ? fma(pb, fma(m_off, nba.y, m_diag * ca.y), pa * ca.y) : 0.0;
                    st2(zp + idx, z);
                    if (push_lo != nullptr) st2(push_lo, z);
                    if (push_hi != nullptr) st2(push_hi, z);
                    rz = fma(ca.x, z.x, rz);
                    rz = fma(ca.y, z.y, rz);
                }
                if (inx && ya + 1 < n) {
                    double2 z;
                    z.x = v0x ? fma(pb, fma(m_off, nbb.x, m_diag * cb.x), pa * cb.x) : 0.0;
                    z.y = v1x ? fma(pb, fma(m_off, nbb.y, m_diag * cb.y), pa * cb.y) : 0.0;
                    st2(zp + idx + g.sy, z);
                    if (push_lo != nullptr) st2(push_lo + g.sy, z);
                    if (push_hi != nullptr) st2(push_hi + g.sy, z);
                    rz = fma(cb.x, z.x, rz);
                    rz = fma(cb.y, z.y, rz);
                }
            }
            __syncwarp();
            if (lane == 0) {
                if (NDIM == 3 && c.zp >= c.z0) mbar_arrive(&sm.empty[(k - 1u) % kPipeStages]);
                if (NDIM == 2 || c.zp == c.z1) mbar_arrive(&sm.empty[stg]);
            }
            cprev0 = cc0;
            cprev1 = cc1;
            cc0 = v0;
            cc1 = v1;
            ++k;
            cursor_next<NDIM>(c, U, g, nact);
        }
        if (cur_a >= 0) pipe_flush(sm, cur_a, rz);
    }
    kstep = k;
    pipe_publish(sm, partials, kSlotC);
}

}  // namespace sdcb200
