// The stencil passes of the CG / Newton-CG solvers (and of eval_f) as bulk-async (TMA) pipelines through shared memory.
//
// Why: the register-marching stencil (stencil.cuh) can only keep the loads of ONE plane per warp in flight, and only a
// third of those go to DRAM (the others are neighbour rows served by L1/L2) - about 32 KB per SM, where HBM3e needs
// 50-60 KB per SM in flight to stay busy (ncu: 50 % DRAM utilisation, long-scoreboard stalls).  Here a dedicated
// producer warp streams whole tile planes global -> shared with TMA tensor copies (cp.async.bulk.tensor, SASS
// UTMALDG: ONE request per 18x68 / 16x64 box - row-wise cp.async.bulk copies were measured to be limited by the TMA
// unit's request rate, ~70 cycles per request per SM), completion on mbarriers, several planes deep, so the bytes in
// flight (2 CTAs x 2-3 planes x 19-36 KB per SM) no longer depend on registers or occupancy; 8 consumer warps read the
// staged planes from shared memory and never wait on DRAM.
//
//   tile      64 (x) x 16 (y) points of one z-plane (+ 1 halo row above/below, + 2 halo columns left/right so that
//             every row copy is 16-byte aligned); a unit marches chunk_z planes, z-neighbours live in registers
//   warp w    rows 2w and 2w+1 of the tile (adjacent rows: each is the other's y-neighbour), one double2 per lane;
//             x-neighbours by warp shuffle, tile-edge columns and the rows above/below from the staged plane
//   steps     a CTA walks its units plane by plane: planes z0-1 .. z1 of a unit are one step each (the first and
//             last only feed the z-neighbours); the producer runs up to kStages-1 steps ahead, across unit and
//             system boundaries; a stage is released (empty mbarrier, one arrival per consumer warp) when the plane
//             after it has been processed
//
// ONE pass template (pipe_pass) serves every phase; a phase only says which boxes a step needs, how the stencil operand
// is formed from them and what is done with (centre, neighbour sum) at a grid point:
//   phase A   boxes r (or the preconditioned z) and p_old with halo [+ the diagonal];  operand p = r + beta*p_old is
//             evaluated wherever the stencil needs it (centre, rows above/below, edge columns) -> q = M p -> p.q ;
//             p is stored from registers
//   phase B   boxes p with halo, r, x [+ the diagonal];  r -= alpha*M p, x += alpha*p, r.r ;  r, x stored
//   phase C   box r with halo;  z = pc_a r + pc_b M r (polynomial preconditioner), r.z
//   phase F   box u with halo [+ the forcing profile];  f = A u (+ reaction term / + profile*g(t)): eval_f
// Boundaries: on Dirichlet grids halo reads either hit the zero walls / guard of the walled layout or fall outside the
// tensor map and are zero-filled by the TMA unit.  On PERIODIC grids the same box is issued (zero-filled outside) and
// tiles on the domain edge fetch what lies across the edge with up to four small extra boxes per field - the row
// n-1 / 0 (68x1) and the column pair n-2,n-1 / 0,1 (2x18) - into a "wrap set" next to the box; the consumers of such a
// tile take the neighbours across the edge from there.  Variable diagonals (the Allen-Cahn Jacobian) come as a
// tile-only box.  Tensor maps are encoded on the host per solve and passed as __grid_constant__ kernel parameters.
#pragma once
#include <cuda.h>

#include "cg_common.cuh"

namespace sdcb200 {

constexpr int kPX = 64;             // tile width in doubles
#ifdef SDCB200_TALL_TILES           // A/B switch: 64x32 tiles, ONE CTA per SM with 16 consumer warps - measured equal
constexpr int kPY = 32;             // to the default within 1 % on configs 2, 3, 4 (profiles/r02/ab_tall_vs_short_tiles.txt)
#else
constexpr int kPY = 16;             // tile rows
#endif
constexpr int kPHX = kPX + 4;       // staged row with halo: cols 0,1 = x0-2, x0-1 | 2..65 tile | 66,67 = x0+64, x0+65
constexpr int kPHY = kPY + 2;       // staged rows with halo: row 0 = y0-1, rows 1..kPY tile, row kPY+1 = y0+kPY
constexpr int kPipeConsumers = kPY / 2;   // consumer warps: two adjacent rows each
constexpr int kPipeThreads = 32 * (kPipeConsumers + 1);

constexpr int kHaloBoxBytes = kPHY * kPHX * 8;   // 9792
constexpr int kCentreBoxBytes = kPY * kPX * 8;   // 8192
constexpr int kHaloSlot = (kHaloBoxBytes + 127) / 128 * 128;  // every box starts 128-byte aligned in smem
constexpr int kWrapRowBytes = kPHX * 8;          // 544: one row with halo columns
constexpr int kWrapColBytes = kPHY * 2 * 8;      // 288: a column pair over the rows with halo
constexpr int kWrapRowSlot = 640, kWrapColSlot = (kWrapColBytes + 127) / 128 * 128;
constexpr int kWrapSetBytes = 2 * kWrapRowSlot + 2 * kWrapColSlot;  // 2048

// what lies across the periodic edge of a tile: row y0-1 -> n-1 (T), row y0+16 -> 0 (B), columns x0-2,x0-1 -> n-2,n-1 (L),
// columns behind the last one -> 0,1 (R)
struct WrapSet {
    double rowT[kWrapRowSlot / 8];
    double rowB[kWrapRowSlot / 8];
    double colL[kWrapColSlot / 16][2];
    double colR[kWrapColSlot / 16][2];
};
static_assert(sizeof(WrapSet) == kWrapSetBytes, "wrap set layout");

enum PipePhase { kPhaseA = 0, kPhaseB = 1, kPhaseC = 2, kPhaseF = 3 };

// Shared-memory geometry of one kernel variant.  A stage holds up to two boxes with halo (+ their wrap sets on periodic
// grids) and up to three (four with a diagonal) tile-only boxes; it is sized for the largest phase of the variant.
// EVAL: the variant of the eval_f kernel - one box with halo and one tile-only box per stage, so the pipeline can be
// deeper (its single 10 KB box per step needs more steps in flight than the solver's 20-36 KB stages to keep HBM busy).
template <bool PER, bool DIAG, int EVAL = 0>
struct PipeCfg {
    static constexpr int kWrap = PER ? kWrapSetBytes : 0;
    static constexpr int kBytesA = 2 * kHaloSlot + 2 * kWrap + (DIAG ? kCentreBoxBytes : 0);
    static constexpr int kBytesB = kHaloSlot + kWrap + (2 + (DIAG ? 1 : 0)) * kCentreBoxBytes;
    // eval_f: EVAL == 1 with a tile-only box (forcing profile), EVAL == 2 without - twice the pipeline depth
    static constexpr int kBytesF = kHaloSlot + kWrap + (EVAL == 1 ? kCentreBoxBytes : 0);
    static constexpr int kStageBytes = EVAL ? kBytesF : (kBytesA > kBytesB ? kBytesA : kBytesB);
    // 2 CTAs per SM (16 consumer warps) must fit into 227 KB together with the static shared memory of the solver; the
    // kernels are held to 96 registers for that (5 warps of one SM sub-partition x 96 x 32 <= 16 K registers).
    // Measured alternative (-DSDCB200_ONE_CTA, profiles/r02/ab_one_vs_two_ctas.txt): ONE CTA per SM with 8 / 6 stages
    // is 8-10 % slower on every configuration - the consumers, not the bytes in flight, are what a second CTA adds.
#ifndef SDCB200_EVAL_STAGES
#define SDCB200_EVAL_STAGES 5
#endif
#ifndef SDCB200_SLIM_STAGES
#define SDCB200_SLIM_STAGES 7  // eval_f without the tile-only slot (experiment: scripts/gpu_r2p.sh)
#endif
#ifdef SDCB200_ONE_CTA
    static constexpr int kStages = EVAL == 2 ? (PER ? 8 : 10) : EVAL ? 5 : ((PER || DIAG) ? 6 : 8);
#else
    static constexpr int kStages = EVAL == 2 ? SDCB200_SLIM_STAGES : EVAL ? SDCB200_EVAL_STAGES : ((PER || DIAG) ? 3 : 4);
#endif
    // offsets inside a stage
    __host__ __device__ static constexpr int halo_off(int f) { return f * (kHaloSlot + kWrap); }
    __host__ __device__ static constexpr int wrap_off(int f) { return f * (kHaloSlot + kWrap) + kHaloSlot; }
    __host__ __device__ static constexpr int centre_off(int nh, int f) { return nh * (kHaloSlot + kWrap) + f * kCentreBoxBytes; }
};

constexpr int kMaxStages = 10;
struct PipeCtl {
    unsigned long long full[kMaxStages];
    unsigned long long empty[kMaxStages];
    double wsum[SDCB200_MAX_NODES][kPipeConsumers];
    int act_list[SDCB200_MAX_NODES];
    int nact;
};
template <bool PER, bool DIAG, int EVAL = 0>
struct PipeSmemT {
    using Cfg = PipeCfg<PER, DIAG, EVAL>;
    alignas(128) unsigned char st[Cfg::kStages][Cfg::kStageBytes];
    PipeCtl ctl;
};

// tensor maps of one system
enum {
    kMapRHalo = 0, kMapPHalo, kMapQHalo, kMapZHalo,          // 18x68 boxes
    kMapRCentre, kMapXCentre, kMapDCentre,                   // 16x64 boxes
    kMapRRow, kMapPRow, kMapQRow, kMapZRow,                  // periodic grids: 1x68 wrap rows
    kMapRCol, kMapPCol, kMapQCol, kMapZCol,                  //                 18x2 wrap column pairs
    kMapsPerSys
};
constexpr int kMapRowOf = kMapRRow - kMapRHalo;  // halo map index -> its wrap-row / wrap-column map
constexpr int kMapColOf = kMapRCol - kMapRHalo;
struct PipeMaps {
    CUtensorMap m[SDCB200_MAX_NODES][kMapsPerSys];
};

struct PUnits {
    int nxt, nyt, nzc, chunk_z, per_field;
};

// Tile grid of the pipelined passes.  The z-chunk trades re-read halo planes (2 per chunk and field read with halo)
// against how evenly the units fill the waves of the persistent grid: take the longest march among the candidates whose
// fill is within 2 % of the best one.  (Thin slabs - 64 planes per rank at 8 GPUs - give the same fill for every
// candidate, so the whole slab is marched in one go instead of re-reading 12.5 % halo planes with 16-plane chunks.)
__host__ __device__ inline PUnits make_punits(const Geom& g, int B, int ctas) {
    PUnits u;
    u.nxt = (g.P + kPX - 1) / kPX;
    u.nyt = (g.n + kPY - 1) / kPY;
    u.chunk_z = 1;
    u.nzc = 1;
    if (g.ndim == 3) {
        const int tiles = u.nxt * u.nyt;
        double best = 0.0;
        for (int c = 128; c >= 16; c >>= 1) {
            const long long units = (long long)B * tiles * ((g.nz + c - 1) / c);
            const long long rounds = (units + ctas - 1) / ctas;
            const double fill = (double)units / (double)(rounds * ctas);
            if (fill > best) best = fill;
        }
        int chunk = 16;
        for (int c = 128; c >= 16; c >>= 1) {
            const long long units = (long long)B * tiles * ((g.nz + c - 1) / c);
            const long long rounds = (units + ctas - 1) / ctas;
            if ((double)units / (double)(rounds * ctas) >= best - 0.02) {
                chunk = c;
                break;
            }
        }
        if (chunk > g.nz) chunk = g.nz < 16 ? 16 : ((g.nz + 15) / 16) * 16;
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
template <int NDIM>
__device__ __forceinline__ void tma_box(void* dst, const CUtensorMap* map, int x, int y, int z, unsigned long long* bar) {
    if (NDIM == 3) tma_box_3d(dst, map, x, y, z, bar);
    else tma_box_2d(dst, map, x, y, bar);
}
// order this thread's generic-proxy global accesses against async-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__device__ __forceinline__ void pipe_ctl_init(PipeCtl& ctl, int stages) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) {
            mbar_init(&ctl.full[i], 1);
            mbar_init(&ctl.empty[i], kPipeConsumers);
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

__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ double2 dir2(double2 r, double2 p, double beta) {
    return make_double2(__dadd_rn(__dmul_rn(p.x, beta), r.x), __dadd_rn(__dmul_rn(p.y, beta), r.y));
}
__device__ __forceinline__ double dir1(double r, double p, double beta) { return __dadd_rn(__dmul_rn(p, beta), r); }

// per-system partial sums of one phase: warp partials -> shared -> (after the CTA barrier) CTA partial in global
__device__ __forceinline__ void pipe_flush(PipeCtl& ctl, int a, double v) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) ctl.wsum[a][threadIdx.x >> 5] = v;
}
__device__ __forceinline__ void pipe_publish(PipeCtl& ctl, double* partials, int slot) {
    __syncthreads();
    if ((int)threadIdx.x < ctl.nact) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kPipeConsumers; ++w) t += ctl.wsum[threadIdx.x][w];
        partials[(size_t)(slot * SDCB200_MAX_NODES + ctl.act_list[threadIdx.x]) * gridDim.x + blockIdx.x] = t;
    }
}
__device__ __forceinline__ void pipe_begin(PipeCtl& ctl) {
    if (threadIdx.x < SDCB200_MAX_NODES * kPipeConsumers) (&ctl.wsum[0][0])[threadIdx.x] = 0.0;
    __syncthreads();
}

// eval_f as a phase: per-field arguments (no reductions)
struct EvalField {
    double* f;            // output  f = a_off * nb + a_diag * u  (+ reaction)
    double* f_expl;       // forced heat: profile * gt ; NULL otherwise
    double gt;
};
struct EvalPhase {
    double a_diag, a_off;
    double inv_eps2;      // Allen-Cahn reaction term  inv_eps2 * u * (1 - u^nu_exp) ; nu_exp == 0: none
    int nu_exp;
    int split;            // 1: the reaction term goes to f_expl instead of being added to f (allencahn_semiimplicit);
                          // 2: f = A u - inv_eps2 * u^(nu+1), f_expl = inv_eps2 * u (allencahn_semiimplicit_v2)
    EvalField e[SDCB200_MAX_NODES];
};

// run-time description of a pass (what differs between calls of the same phase)
struct PassArgs {
    bool first = false;   // phase A: p = r (p_old is not read)
    int cur = 0;          // which direction buffer is written (A) / read (B): (cur ? q : p) holds the new direction
    int slot = 0;         // partial-sum slot of the reduction
    const SlabLink* link = nullptr;
    const EvalPhase* ev = nullptr;  // phase F
};

// ---------------------------------------------------------------------------------------------------------------------
// the pass
// ---------------------------------------------------------------------------------------------------------------------
template <int NDIM, bool PER, bool DIAG, int PHASE, class SMEM>
__device__ void pipe_pass(const Geom& g, const PUnits& U, const Sys* s, const PipeMaps& maps, const CgShared& sh, SMEM& sm,
                          double* partials, unsigned& kstep, const PassArgs pa) {
    using Cfg = typename SMEM::Cfg;
    constexpr int kStages = Cfg::kStages;
    // boxes of a step of this phase
    constexpr int NH = PHASE == kPhaseA ? 2 : 1;                        // boxes with halo (A: r, p_old)
    constexpr int NCB = PHASE == kPhaseB ? 2 : 0;                       // tile-only boxes r, x
    constexpr bool kDiag = DIAG && (PHASE == kPhaseA || PHASE == kPhaseB);
    constexpr bool kProfile = PHASE == kPhaseF;                         // optional tile-only box: forcing profile
    constexpr int kDiagIdx = NCB;                                       // centre slot of the diagonal / profile
    PipeCtl& ctl = sm.ctl;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nact = ctl.nact;
    const int P = g.P, n = g.n;
    const int nh_run = (PHASE == kPhaseA && pa.first) ? 1 : NH;
    if (PHASE != kPhaseF) pipe_begin(ctl);
    StepCursor c;
    cursor_init<NDIM>(c, U, g, nact);
    unsigned k = kstep;
    auto halo_box = [&](unsigned stg, int f) { return reinterpret_cast<double(*)[kPHX]>(sm.st[stg] + Cfg::halo_off(f)); };
    auto wrap_set = [&](unsigned stg, int f) { return reinterpret_cast<WrapSet*>(sm.st[stg] + Cfg::wrap_off(f)); };
    auto centre_box = [&](unsigned stg, int f) { return reinterpret_cast<double(*)[kPX]>(sm.st[stg] + Cfg::centre_off(NH, f)); };

    if (warp == kPipeConsumers) {
        // ---- producer: the TMA boxes of one plane-step.  Boxes with halo carry the ring around the tile; tile-only
        // boxes are skipped on the two planes that merely feed the z-neighbours.  3-D Dirichlet maps start one plane
        // below the field (the guard / lower halo plane), hence z + 1; periodic 3-D maps start at plane 0 and the
        // plane index wraps.
        while (c.valid) {
            const unsigned stg = k % kStages;
            if (k >= (unsigned)kStages) mbar_wait(&ctl.empty[stg], ((k / kStages) - 1u) & 1u);
            if (lane == 0) {
                const int b = ctl.act_list[c.a];
                const CUtensorMap* mp = maps.m[b];
                int hmap[2];
                if (PHASE == kPhaseA) {
                    hmap[0] = s[b].z != nullptr ? kMapZHalo : kMapRHalo;  // r (or the preconditioned residual z)
                    hmap[1] = pa.cur ? kMapPHalo : kMapQHalo;             // p_old
                } else if (PHASE == kPhaseB) {
                    hmap[0] = pa.cur ? kMapQHalo : kMapPHalo;             // the new p
                } else if (PHASE == kPhaseC) {
                    hmap[0] = kMapRHalo;
                } else {
                    hmap[0] = kMapPHalo;  // phase F: the field f is evaluated on (encoded in map slot P)
                }
                const bool halo_plane = NDIM == 3 && (c.zp < c.z0 || c.zp >= c.z1);
                const bool profile = kProfile && pa.ev->e[b].f_expl != nullptr && !pa.ev->split;
                int ncentre = halo_plane ? 0 : NCB + (kDiag ? 1 : 0) + (profile ? 1 : 0);
                // which edges of the periodic domain does this tile touch?
                const bool wT = PER && c.y0 == 0, wB = PER && c.y0 + kPY >= n;
                const bool wL = PER && c.x0 == 0, wR = PER && c.x0 + kPX >= n;
                const unsigned wrap_bytes = (wT ? kWrapRowBytes : 0) + (wB ? kWrapRowBytes : 0) + (wL ? kWrapColBytes : 0) +
                                            (wR ? kWrapColBytes : 0);
                mbar_arrive_expect_tx(&ctl.full[stg],
                                      (unsigned)(nh_run * (kHaloBoxBytes + wrap_bytes) + ncentre * kCentreBoxBytes));
                int z = NDIM == 3 ? c.zp + 1 : 0;
                if (PER && NDIM == 3) z = c.zp < 0 ? c.zp + g.nz : (c.zp >= g.nz ? c.zp - g.nz : c.zp);
                for (int f = 0; f < nh_run; ++f) {
                    tma_box<NDIM>(halo_box(stg, f), mp + hmap[f], c.x0 - 2, c.y0 - 1, z, &ctl.full[stg]);
                    if (PER) {
                        WrapSet* w = wrap_set(stg, f);
                        const CUtensorMap* mrow = mp + hmap[f] + kMapRowOf;
                        const CUtensorMap* mcol = mp + hmap[f] + kMapColOf;
                        if (wT) tma_box<NDIM>(w->rowT, mrow, c.x0 - 2, n - 1, z, &ctl.full[stg]);
                        if (wB) tma_box<NDIM>(w->rowB, mrow, c.x0 - 2, 0, z, &ctl.full[stg]);
                        if (wL) tma_box<NDIM>(w->colL, mcol, n - 2, c.y0 - 1, z, &ctl.full[stg]);
                        if (wR) tma_box<NDIM>(w->colR, mcol, 0, c.y0 - 1, z, &ctl.full[stg]);
                    }
                }
                if (!halo_plane) {
                    if (PHASE == kPhaseB) {
                        tma_box<NDIM>(centre_box(stg, 0), mp + kMapRCentre, c.x0, c.y0, z, &ctl.full[stg]);
                        tma_box<NDIM>(centre_box(stg, 1), mp + kMapXCentre, c.x0, c.y0, z, &ctl.full[stg]);
                    }
                    if (kDiag || profile)
                        tma_box<NDIM>(centre_box(stg, kDiagIdx), mp + kMapDCentre, c.x0, c.y0, z, &ctl.full[stg]);
                }
            }
            ++k;
            cursor_next<NDIM>(c, U, g, nact);
        }
    } else {
        // ---- consumers --------------------------------------------------------------------------------------------
        const int ra = 2 * warp, col = 2 + 2 * lane;
        double2 cprev0 = make_double2(0.0, 0.0), cprev1 = cprev0, cc0 = cprev0, cc1 = cprev0;
        double acc = 0.0;
        int cur_a = -1;
        double beta = 0.0, alpha = 0.0, m_diag = 0.0, m_off = 0.0, pc_a = 0.0, pc_b = 0.0;
        double *out0 = nullptr, *out1 = nullptr;   // A: p_new | B: r, x | C: z | F: f, f_expl
        double *push_lo_base = nullptr, *push_hi_base = nullptr;  // slab runs: the neighbours' halo planes of r (B) / z (C)
        double gt = 0.0;
        while (c.valid) {
            if (c.a != cur_a) {
                if (cur_a >= 0 && PHASE != kPhaseF) pipe_flush(ctl, cur_a, acc);
                cur_a = c.a;
                acc = 0.0;
                const int b = ctl.act_list[c.a];
                if (PHASE == kPhaseF) {
                    m_diag = pa.ev->a_diag;
                    m_off = pa.ev->a_off;
                    out0 = pa.ev->e[b].f;
                    out1 = pa.ev->e[b].f_expl;
                    gt = pa.ev->e[b].gt;
                } else {
                    m_diag = s[b].m_diag;
                    m_off = s[b].m_off;
                }
                if (PHASE == kPhaseA) {
                    beta = sh.beta[b];
                    out0 = pa.cur ? s[b].q : s[b].p;
                } else if (PHASE == kPhaseB) {
                    alpha = sh.alpha[b];
                    out0 = s[b].r;
                    out1 = s[b].x;
                    if (pa.link != nullptr) {
                        push_lo_base = pa.link->has_lo ? pa.link->lo_r_halo[b] : nullptr;
                        push_hi_base = pa.link->has_hi ? pa.link->hi_r_halo[b] : nullptr;
                    }
                } else if (PHASE == kPhaseC) {
                    pc_a = s[b].pc_a;
                    pc_b = s[b].pc_b;
                    out0 = s[b].z;
                    if (pa.link != nullptr) {
                        push_lo_base = pa.link->has_lo ? pa.link->lo_z_halo[b] : nullptr;
                        push_hi_base = pa.link->has_hi ? pa.link->hi_z_halo[b] : nullptr;
                    }
                }
            }
            const unsigned stg = k % kStages;
            mbar_wait(&ctl.full[stg], (k / kStages) & 1u);
            const bool dir = PHASE == kPhaseA && !pa.first;
            // the stencil operand at a staged position: the staged field itself, or r + beta*p_old (phase A)
            auto val2 = [&](unsigned st, int row, int cl) {
                double2 v = lds2(&halo_box(st, 0)[row][cl]);
                if (dir) v = dir2(v, lds2(&halo_box(st, 1)[row][cl]), beta);
                return v;
            };
            auto val1 = [&](unsigned st, int row, int cl) {
                double v = halo_box(st, 0)[row][cl];
                if (dir) v = dir1(v, halo_box(st, 1)[row][cl], beta);
                return v;
            };
            const int x = c.x0 + 2 * lane, ya = c.y0 + ra;
            const bool inx = x < P;
            // operand on the two tile rows of this warp, plane zp
            const double2 v0 = val2(stg, 1 + ra, col), v1 = val2(stg, 2 + ra, col);
            if (NDIM == 2 || c.zp > c.z0) {
                // plane zc = zp-1 (3-D) / this plane (2-D) has all its neighbours now
                const unsigned s0 = NDIM == 3 ? (k - 1u) % kStages : stg;
                const double2 ca = NDIM == 3 ? cc0 : v0, cb = NDIM == 3 ? cc1 : v1;
                double2 up = val2(s0, ra, col), dn = val2(s0, ra + 3, col);
                double la = __shfl_up_sync(0xffffffffu, ca.y, 1), ra_ = __shfl_down_sync(0xffffffffu, ca.x, 1);
                double lb = __shfl_up_sync(0xffffffffu, cb.y, 1), rb_ = __shfl_down_sync(0xffffffffu, cb.x, 1);
                if (lane == 0) {
                    la = val1(s0, 1 + ra, 1);
                    lb = val1(s0, 2 + ra, 1);
                }
                if (lane == 31 || x + 2 >= P) {
                    if (x + 2 < P) {
                        ra_ = val1(s0, 1 + ra, kPX + 2);
                        rb_ = val1(s0, 2 + ra, kPX + 2);
                    } else {
                        ra_ = rb_ = 0.0;  // beyond the wall (periodic grids: replaced below)
                    }
                }
                if (PER) {
                    // neighbours across the periodic edge come from the wrap sets of this tile
                    const WrapSet* w0 = wrap_set(s0, 0);
                    const WrapSet* w1 = wrap_set(s0, 1);
                    if (ya == 0) {  // row above row 0 is row n-1
                        up = lds2(&w0->rowT[col]);
                        if (dir) up = dir2(up, lds2(&w1->rowT[col]), beta);
                    }
                    if (ya + 1 == n - 1) {  // row below row n-1 is row 0
                        dn = lds2(&w0->rowB[col]);
                        if (dir) dn = dir2(dn, lds2(&w1->rowB[col]), beta);
                    }
                    if (x == 0) {  // column left of column 0 is column n-1
                        la = w0->colL[1 + ra][1];
                        lb = w0->colL[2 + ra][1];
                        if (dir) {
                            la = dir1(la, w1->colL[1 + ra][1], beta);
                            lb = dir1(lb, w1->colL[2 + ra][1], beta);
                        }
                    }
                    if (x + 2 == n) {  // column right of column n-1 is column 0
                        ra_ = w0->colR[1 + ra][0];
                        rb_ = w0->colR[2 + ra][0];
                        if (dir) {
                            ra_ = dir1(ra_, w1->colR[1 + ra][0], beta);
                            rb_ = dir1(rb_, w1->colR[2 + ra][0], beta);
                        }
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
                // slab boundary planes of r / z go straight into the neighbour's halo plane (peer memory over NVLink)
                double* push_lo = (NDIM == 3 && push_lo_base != nullptr && c.zp - 1 == 0) ? push_lo_base + (long long)ya * g.sy + x : nullptr;
                double* push_hi = (NDIM == 3 && push_hi_base != nullptr && c.zp == g.nz) ? push_hi_base + (long long)ya * g.sy + x : nullptr;
#pragma unroll
                for (int h = 0; h < 2; ++h) {  // the warp's two rows
                    const double2 cv = h ? cb : ca, nb = h ? nbb : nba;
                    if (!(inx && ya + h < n)) continue;
                    const long long id = idx + (h ? g.sy : 0);
                    double2 d = make_double2(m_diag, m_diag);
                    if (kDiag) d = lds2(&centre_box(s0, kDiagIdx)[ra + h][2 * lane]);
                    // (M operand) at the two points; exact zero on the wall column of Dirichlet grids
                    const double mx = v0x ? fma(m_off, nb.x, d.x * cv.x) : 0.0;
                    const double my = v1x ? fma(m_off, nb.y, d.y * cv.y) : 0.0;
                    if (PHASE == kPhaseA) {
                        st2(out0 + id, cv);  // wall column: r = p_old = 0 there, so the stored value is an exact zero
                        acc = fma(cv.x, mx, acc);
                        acc = fma(cv.y, my, acc);
                    } else if (PHASE == kPhaseB) {
                        double2 r = lds2(&centre_box(s0, 0)[ra + h][2 * lane]), xv = lds2(&centre_box(s0, 1)[ra + h][2 * lane]);
                        r.x = v0x ? __dsub_rn(r.x, __dmul_rn(alpha, mx)) : 0.0;
                        r.y = v1x ? __dsub_rn(r.y, __dmul_rn(alpha, my)) : 0.0;
                        xv.x = __dadd_rn(xv.x, __dmul_rn(alpha, cv.x));
                        xv.y = __dadd_rn(xv.y, __dmul_rn(alpha, cv.y));
                        st2(out0 + id, r);
                        st2(out1 + id, xv);
                        if (push_lo != nullptr) st2(push_lo + (h ? g.sy : 0), r);
                        if (push_hi != nullptr) st2(push_hi + (h ? g.sy : 0), r);
                        acc = fma(r.x, r.x, acc);
                        acc = fma(r.y, r.y, acc);
                    } else if (PHASE == kPhaseC) {
                        double2 z;
                        z.x = v0x ? fma(pc_b, mx, pc_a * cv.x) : 0.0;
                        z.y = v1x ? fma(pc_b, my, pc_a * cv.y) : 0.0;
                        st2(out0 + id, z);
                        if (push_lo != nullptr) st2(push_lo + (h ? g.sy : 0), z);
                        if (push_hi != nullptr) st2(push_hi + (h ? g.sy : 0), z);
                        acc = fma(cv.x, z.x, acc);
                        acc = fma(cv.y, z.y, acc);
                    } else {
                        double2 f = make_double2(mx, my);
                        if (pa.ev->nu_exp > 0) {  // Allen-Cahn: A u + 1/eps^2 u (1 - u^nu), in the reference's order
                            const double ex = pa.ev->inv_eps2, ux = cv.x, uy = cv.y;
                            double px = ux, py = uy;
                            for (int i = 1; i < pa.ev->nu_exp; ++i) {
                                px = __dmul_rn(px, ux);
                                py = __dmul_rn(py, uy);
                            }
                            const double2 re = make_double2(__dmul_rn(__dmul_rn(ex, ux), __dsub_rn(1.0, px)),
                                                            __dmul_rn(__dmul_rn(ex, uy), __dsub_rn(1.0, py)));
                            if (pa.ev->split == 2) {  // AllenCahn_2D_FD.py:421-422
                                f.x = __dsub_rn(mx, __dmul_rn(ex, __dmul_rn(px, ux)));
                                f.y = __dsub_rn(my, __dmul_rn(ex, __dmul_rn(py, uy)));
                                st2(out1 + id, make_double2(__dmul_rn(ex, ux), __dmul_rn(ex, uy)));
                            } else if (pa.ev->split) {
                                st2(out1 + id, re);  // f.expl (AllenCahn_2D_FD.py:300)
                            } else {
                                f.x = __dadd_rn(mx, re.x);
                                f.y = __dadd_rn(my, re.y);
                            }
                        }
                        st2(out0 + id, f);
                        if (out1 != nullptr && !pa.ev->split) {
                            const double2 pr = lds2(&centre_box(s0, kDiagIdx)[ra + h][2 * lane]);
                            st2(out1 + id, make_double2(v0x ? __dmul_rn(pr.x, gt) : 0.0, v1x ? __dmul_rn(pr.y, gt) : 0.0));
                        }
                    }
                }
            }
            if (PHASE == kPhaseA && NDIM == 3 && g.zhalo && (c.zp < 0 || c.zp == g.nz)) {
                // slab halo planes: materialise p there too (bit-identical to what the neighbouring rank computes)
                const long long idx = (long long)c.zp * g.sz + (long long)ya * g.sy + x;
                if (inx && ya < n) st2(out0 + idx, v0);
                if (inx && ya + 1 < n) st2(out0 + idx + g.sy, v1);
            }
            __syncwarp();
            if (lane == 0) {
                if (NDIM == 3 && c.zp >= c.z0) mbar_arrive(&ctl.empty[(k - 1u) % kStages]);
                if (NDIM == 2 || c.zp == c.z1) mbar_arrive(&ctl.empty[stg]);
            }
            cprev0 = cc0;
            cprev1 = cc1;
            cc0 = v0;
            cc1 = v1;
            ++k;
            cursor_next<NDIM>(c, U, g, nact);
        }
        if (cur_a >= 0 && PHASE != kPhaseF) pipe_flush(ctl, cur_a, acc);
    }
    kstep = k;
    if (PHASE != kPhaseF) pipe_publish(ctl, partials, pa.slot);
}

}  // namespace sdcb200
