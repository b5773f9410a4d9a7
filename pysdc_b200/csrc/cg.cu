// K3: batched conjugate gradients for (I - factor*A) x = b, and K4: the Allen-Cahn Newton solver around it.
//
// One persistent cooperative launch runs the WHOLE solve: all CG iterations of all B node systems, the dot-product
// reductions (warp shuffle -> block -> fixed-order grid reduction, bitwise identical in every CTA so that control
// flow stays uniform), the per-system convergence tests and the iteration counters live on the device; the host
// never synchronises inside a solve.  Grid = (co-resident CTAs per SM) x (SM count), software grid barrier
// (release/acquire at gpu scope).  The recurrence follows scipy.sparse.linalg.cg (scipy 1.18.1, _isolve/iterative.py)
// statement by statement, including the unfused rounding of  p*=beta; p+=r;  x+=alpha*p;  r-=alpha*q, but is
// scheduled as two fused passes per iteration (see cg_collective).
#include "cg_common.cuh"
#include "cg_pipe.cuh"
#include "pipe_host.cuh"

namespace sdcb200 {
namespace {

// ---------------------------------------------------------------------------------------------------------------------
// the collective CG routine: every thread of the grid calls it with identical arguments.
//
// Two grid-wide phases per iteration, 8 field streams per system and iteration (a textbook CG moves 11):
//   phase A   p <- r + beta p_old   evaluated on the fly on every stencil point (tile + halo) from r and p_old, so the
//             direction update and the operator application share one pass:  reads r, p_old;  writes p;  p.(M p)
//             p is double-buffered (S.p / S.q alternate) because neighbouring tiles still read p_old's halo.
//   phase B   M p is evaluated again by the stencil (p is final now) instead of being stored and re-read:
//             r <- r - alpha M p,  x <- x + alpha p,  r.r :  reads p, r, x;  writes r, x.
// Rounding follows scipy:  p = fl(fl(p*beta) + r),  x = fl(x + fl(alpha*p)),  r = fl(r - fl(alpha*q)).
// ---------------------------------------------------------------------------------------------------------------------
template <int NDIM, bool PER>
__device__ void cg_collective(const Geom& g, int B, const Sys* s, double rtol, int maxiter, double* partials,
                              unsigned* bar, CgShared& sh) {
    const Units U = make_units(g, (int)gridDim.x);
    const long long n2 = g.owned / 2;
    const long long gtid = (long long)blockIdx.x * kThreads + threadIdx.x;
    const long long gstride = (long long)gridDim.x * kThreads;

    // ---- r = b - M x0, ||b||^2, ||r||^2 ------------------------------------------------------------------------------
    for (int b = 0; b < B; ++b) {
        const Sys& S = s[b];
        double bb = 0.0, rr = 0.0;
        for (int unit = blockIdx.x; unit < U.per_field; unit += gridDim.x) {
            stencil_unit<NDIM, PER>(g, U, S.x, unit, [&](long long idx, double2 c, double2 nb, bool v0, bool v1) {
                const double2 rhs = ld2(S.b + idx);
                double2 d = make_double2(S.m_diag, S.m_diag);
                if (S.dvec != nullptr) d = ld2(S.dvec + idx);
                double2 r;
                r.x = v0 ? rhs.x - fma(S.m_off, nb.x, d.x * c.x) : 0.0;
                r.y = v1 ? rhs.y - fma(S.m_off, nb.y, d.y * c.y) : 0.0;
                st2(S.r + idx, r);
                if (v0) bb = fma(rhs.x, rhs.x, bb);
                if (v1) bb = fma(rhs.y, rhs.y, bb);
                rr = fma(r.x, r.x, rr);
                rr = fma(r.y, r.y, rr);
            });
        }
        bb = block_sum(bb, sh.scratch);
        rr = block_sum(rr, sh.scratch);
        put_partial(partials, kSlotSetup0, b, bb);
        put_partial(partials, kSlotSetup1, b, rr);
    }
    grid_barrier(bar);
    for (int b = 0; b < B; ++b) {
        const double bb = grid_sum(partials, kSlotSetup0, b, sh.scratch);
        const double rr = grid_sum(partials, kSlotSetup1, b, sh.scratch);
        if (threadIdx.x == 0) {
            sh.bb[b] = bb;
            sh.rr[b] = rr;
            sh.iters[b] = 0;
            sh.rho_prev[b] = 1.0;
        }
    }
    if (threadIdx.x == 0) {
        unsigned act = 0;
        for (int b = 0; b < B; ++b)
            if (sh.bb[b] != 0.0) act |= 1u << b;  // scipy: ||b|| == 0 -> return b
        sh.active = act;
    }
    __syncthreads();
    // systems with a zero right-hand side: solution is b itself (all zeros)
    for (int b = 0; b < B; ++b) {
        if (sh.bb[b] == 0.0) {
            for (long long i = gtid; i < n2; i += gstride) st2(s[b].x + 2 * i, make_double2(0.0, 0.0));
        }
    }

    for (int it = 0;; ++it) {
        // ---- convergence test first (scipy: "if norm(r) < atol: return"), then the iteration budget ----------------
        if (threadIdx.x == 0) {
            unsigned act = sh.active;
            for (int b = 0; b < B; ++b) {
                if (!(act >> b & 1u)) continue;
                const double atol = rtol * sqrt(sh.bb[b]);
                if (sqrt(sh.rr[b]) < atol || it >= maxiter) {
                    act &= ~(1u << b);
                } else {
                    sh.beta[b] = it > 0 ? sh.rr[b] / sh.rho_prev[b] : 0.0;
                }
            }
            sh.active = act;
        }
        __syncthreads();
        const unsigned act = sh.active;
        if (act == 0) break;
        // all systems start together, so the parity of `it` tells which buffer holds p_old for every active one
        const int cur = it & 1;

        // ---- phase A: p = r + beta p_old on the fly, q = M p, p.q --------------------------------------------------------
        for (int b = 0; b < B; ++b) {
            if (!(act >> b & 1u)) continue;
            const Sys& S = s[b];
            double* p_new = cur ? S.q : S.p;
            const DirectionLoader dir{S.r, it == 0 ? nullptr : (cur ? S.p : S.q), sh.beta[b]};
            double pq = 0.0;
            for (int unit = blockIdx.x; unit < U.per_field; unit += gridDim.x) {
                stencil_unit_ld<NDIM, PER>(
                    g, U, dir, unit,
                    [&](long long idx, double2 c, double2 nb, bool v0, bool v1) {
                        st2(p_new + idx, c);  // walls: r = p_old = 0 there, so c is an exact zero
                        double2 d = make_double2(S.m_diag, S.m_diag);
                        if (S.dvec != nullptr) d = ld2(S.dvec + idx);
                        const double qx = v0 ? fma(S.m_off, nb.x, d.x * c.x) : 0.0;
                        const double qy = v1 ? fma(S.m_off, nb.y, d.y * c.y) : 0.0;
                        pq = fma(c.x, qx, pq);
                        pq = fma(c.y, qy, pq);
                    },
                    [&](long long idx, double2 c) { st2(p_new + idx, c); });
            }
            pq = block_sum(pq, sh.scratch);
            put_partial(partials, kSlotA, b, pq);
        }
        grid_barrier(bar);
        for (int b = 0; b < B; ++b) {
            if (!(act >> b & 1u)) continue;
            const double pq = grid_sum(partials, kSlotA, b, sh.scratch);
            if (threadIdx.x == 0) sh.alpha[b] = sh.rr[b] / pq;
        }
        __syncthreads();

        // ---- phase B: r -= alpha M p, x += alpha p, r.r ---------------------------------------------------------------
        for (int b = 0; b < B; ++b) {
            if (!(act >> b & 1u)) continue;
            const Sys& S = s[b];
            const double* p = cur ? S.q : S.p;
            const double alpha = sh.alpha[b];
            double rr = 0.0;
            for (int unit = blockIdx.x; unit < U.per_field; unit += gridDim.x) {
                stencil_unit<NDIM, PER>(g, U, p, unit, [&](long long idx, double2 c, double2 nb, bool v0, bool v1) {
                    double2 d = make_double2(S.m_diag, S.m_diag);
                    if (S.dvec != nullptr) d = ld2(S.dvec + idx);
                    double2 r = ld2(S.r + idx), x = ld2(S.x + idx);
                    r.x = v0 ? __dsub_rn(r.x, __dmul_rn(alpha, fma(S.m_off, nb.x, d.x * c.x))) : 0.0;
                    r.y = v1 ? __dsub_rn(r.y, __dmul_rn(alpha, fma(S.m_off, nb.y, d.y * c.y))) : 0.0;
                    x.x = __dadd_rn(x.x, __dmul_rn(alpha, c.x));
                    x.y = __dadd_rn(x.y, __dmul_rn(alpha, c.y));
                    st2(S.r + idx, r);
                    st2(S.x + idx, x);
                    rr = fma(r.x, r.x, rr);
                    rr = fma(r.y, r.y, rr);
                });
            }
            rr = block_sum(rr, sh.scratch);
            put_partial(partials, kSlotB, b, rr);
        }
        grid_barrier(bar);
        for (int b = 0; b < B; ++b) {
            if (!(act >> b & 1u)) continue;
            const double rr = grid_sum(partials, kSlotB, b, sh.scratch);
            if (threadIdx.x == 0) {
                sh.rho_prev[b] = sh.rr[b];
                sh.rr[b] = rr;
                sh.iters[b] += 1;  // scipy calls the callback once per completed iteration
            }
        }
        __syncthreads();
    }
}

template <int NDIM, bool PER>
__global__ void __launch_bounds__(kThreads) cg_kernel(const __grid_constant__ CgArgs a) {
    __shared__ CgShared sh;
    cg_collective<NDIM, PER>(a.g, a.B, a.s, a.rtol, a.maxiter, a.partials, a.bar, sh);
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.iters_out != nullptr)
        for (int b = 0; b < a.B; ++b) a.iters_out[b] += sh.iters[b];
}

// ---------------------------------------------------------------------------------------------------------------------
// the same solver with both passes of the iteration running as bulk-async pipelines (cg_pipe.cuh): 2-D / 3-D
// Dirichlet grids with a constant diagonal, i.e. the heat-equation node solves.  Launched with kPipeThreads threads
// (8 consumer warps + 1 producer warp); the set-up pass uses the register-marching stencil on the first 8 warps.
// ---------------------------------------------------------------------------------------------------------------------
// Sum the per-CTA partials of the listed (slot, system) pairs over the grid - and, on slab runs, over all ranks -
// leaving the results in sh.glob[i] (bit-identical on every thread of every CTA of every rank).
// `tl` (optional): ns accumulators of the first and the last CTA - [0] work since the previous synchronisation (the
// pass itself, pipeline fill / drain, scalar updates), [1] wait in the grid barrier, [2] summing the partials,
// [3] cross-rank exchange of the sums (slab runs).
template <bool SLAB>
__device__ __forceinline__ void reduce_all(double* partials, unsigned* bar, CgShared& sh, const SlabLink* link,
                                           unsigned long long& seq, const int* slots, const int* systems, int nv,
                                           unsigned long long* tl = nullptr) {
    const bool rec = tl != nullptr && threadIdx.x == 0;
    unsigned long long t0 = 0, t1 = 0, t2 = 0;
    if (rec) t0 = global_ns();
    if (SLAB) grid_barrier_sys(bar);
    else grid_barrier(bar);
    if (rec) t1 = global_ns();
    for (int i = 0; i < nv; ++i) {
        const double v = grid_sum(partials, slots[i], systems[i], sh.scratch);
        if (threadIdx.x == 0) (SLAB ? sh.loc : sh.glob)[i] = v;
    }
    if (rec) t2 = global_ns();
    if (SLAB) cross_rank_sum(*link, sh, nv, ++seq);
    else __syncthreads();
    if (rec) {
        const unsigned long long t3 = global_ns();
        unsigned long long* per_cta = tl + 16 + 4 * blockIdx.x;  // every CTA: its own four accumulators
        per_cta[0] += t0 - sh.t_last;
        per_cta[1] += t1 - t0;
        per_cta[2] += t2 - t1;
        per_cta[3] += t3 - t2;
        if (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) {
            unsigned long long* a = tl + (blockIdx.x == 0 ? 0 : 4);
            a[0] += t0 - sh.t_last;
            a[1] += t1 - t0;
            a[2] += t2 - t1;
            a[3] += t3 - t2;
        }
        sh.t_last = t3;
    }
}

// All threads of the grid (and, on slabs, of all ranks) call this with identical arguments.  `start_mask`: the systems
// that take part (bit b); the others are left untouched.  sh.rtol[b] holds each system's relative tolerance.
template <int NDIM, bool SLAB, bool PER, bool DIAG, class SMEM>
__device__ void cg_collective_pipe(const Geom& g, int B, unsigned start_mask, const Sys* s, const PipeMaps& maps,
                                   int maxiter, double* partials, unsigned* bar, CgShared& sh, SMEM& sm,
                                   const SlabLink* link, unsigned& kstep, unsigned long long* tl = nullptr) {
    if (threadIdx.x == 0) sh.t_last = global_ns();
    if (tl != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {  // the launch shape the timings belong to
        const PUnits pu = make_punits(g, 1, (int)gridDim.x);
        tl[8] = gridDim.x;
        tl[9] = (unsigned long long)pu.per_field;
        tl[10] = (unsigned long long)pu.chunk_z;
    }
    const Units U = make_units(g, (int)gridDim.x);
    const PUnits PU = make_punits(g, 1, (int)gridDim.x);
    const long long n2 = g.owned / 2;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gstride = (long long)gridDim.x * blockDim.x;
    PipeCtl& ctl = sm.ctl;
    unsigned long long seq = SLAB ? *link->seq : 0ull;
    __shared__ int r_slots[kMailVals], r_sys[kMailVals];
    __shared__ int r_count;

    // ---- r = b - M x0, ||b||^2, ||r||^2 ------------------------------------------------------------------------------
    for (int b = 0; b < B; ++b) {
        if (!(start_mask >> b & 1u)) continue;
        const Sys& S = s[b];
        double bb = 0.0, rr = 0.0;
        if (threadIdx.x < kThreads) {
            double* r_lo = SLAB && link->has_lo ? link->lo_r_halo[b] : nullptr;
            double* r_hi = SLAB && link->has_hi ? link->hi_r_halo[b] : nullptr;
            for (int unit = blockIdx.x; unit < U.per_field; unit += gridDim.x) {
                stencil_unit<NDIM, PER>(g, U, S.x, unit, [&](long long idx, double2 c, double2 nb, bool v0, bool v1) {
                    const double2 rhs = ld2(S.b + idx);
                    double2 d = make_double2(S.m_diag, S.m_diag);
                    if (DIAG) d = ld2(S.dvec + idx);
                    double2 r;
                    r.x = v0 ? rhs.x - fma(S.m_off, nb.x, d.x * c.x) : 0.0;
                    r.y = v1 ? rhs.y - fma(S.m_off, nb.y, d.y * c.y) : 0.0;
                    st2(S.r + idx, r);
                    if (SLAB) {
                        if (r_lo != nullptr && idx < g.sz) st2(r_lo + idx, r);
                        if (r_hi != nullptr && idx >= g.owned - g.sz) st2(r_hi + (idx - (g.owned - g.sz)), r);
                    }
                    if (v0) bb = fma(rhs.x, rhs.x, bb);
                    if (v1) bb = fma(rhs.y, rhs.y, bb);
                    rr = fma(r.x, r.x, rr);
                    rr = fma(r.y, r.y, rr);
                });
            }
        }
        bb = block_sum(bb, sh.scratch);
        rr = block_sum(rr, sh.scratch);
        put_partial(partials, kSlotSetup0, b, bb);
        put_partial(partials, kSlotSetup1, b, rr);
    }
    if (threadIdx.x == 0) {
        int nv = 0;
        for (int b = 0; b < B; ++b) {
            if (!(start_mask >> b & 1u)) continue;
            r_slots[nv] = kSlotSetup0;
            r_sys[nv++] = b;
            r_slots[nv] = kSlotSetup1;
            r_sys[nv++] = b;
        }
        r_count = nv;
    }
    __syncthreads();
    fence_proxy_async_global();
    reduce_all<SLAB>(partials, bar, sh, link, seq, r_slots, r_sys, r_count, tl);
    if (threadIdx.x == 0) {
        unsigned act = 0;
        for (int i = 0; i < r_count; i += 2) {
            const int b = r_sys[i];
            sh.bb[b] = sh.glob[i];
            sh.rr[b] = sh.glob[i + 1];
            sh.iters[b] = 0;
            sh.rho_prev[b] = 1.0;
            if (sh.bb[b] != 0.0) act |= 1u << b;  // scipy: ||b|| == 0 -> return b
        }
        sh.active = act;
    }
    __syncthreads();
    for (int b = 0; b < B; ++b) {
        if ((start_mask >> b & 1u) && sh.bb[b] == 0.0) {
            for (long long i = gtid; i < n2; i += gstride) st2(s[b].x + 2 * i, make_double2(0.0, 0.0));
        }
    }

    PassArgs pa;
    pa.link = link;
    // preconditioned runs: z = C(M) r for the systems that will iterate, r.z
    const bool pc = s[0].z != nullptr;
    if (pc) {
        if (threadIdx.x == 0) {
            int na = 0;
            for (int b = 0; b < B; ++b)
                if (sh.active >> b & 1u) {
                    ctl.act_list[na] = b;
                    r_sys[na] = b;
                    r_slots[na] = kSlotC;
                    ++na;
                }
            ctl.nact = na;
        }
        __syncthreads();
        if (ctl.nact > 0) {
            const int nc = ctl.nact;
            fence_proxy_async_global();
            pa.slot = kSlotC;
            pipe_pass<NDIM, PER, DIAG, kPhaseC>(g, PU, s, maps, sh, sm, partials, kstep, pa);
            fence_proxy_async_global();
            reduce_all<SLAB>(partials, bar, sh, link, seq, r_slots, r_sys, nc, tl);
            if ((int)threadIdx.x < nc) sh.rz[r_sys[threadIdx.x]] = sh.glob[threadIdx.x];
            __syncthreads();
        }
    } else {
        if ((int)threadIdx.x < B) sh.rz[threadIdx.x] = sh.rr[threadIdx.x];
        __syncthreads();
    }

    for (int it = 0;; ++it) {
        if (threadIdx.x == 0) {
            unsigned act = sh.active;
            int na = 0;
            for (int b = 0; b < B; ++b) {
                if (!(act >> b & 1u)) continue;
                const double atol = sh.rtol[b] * sqrt(sh.bb[b]);
                if (sqrt(sh.rr[b]) < atol || it >= maxiter) {
                    act &= ~(1u << b);
                } else {
                    sh.beta[b] = it > 0 ? sh.rz[b] / sh.rho_prev[b] : 0.0;
                    ctl.act_list[na] = b;
                    r_sys[na] = b;
                    ++na;
                }
            }
            sh.active = act;
            ctl.nact = na;
        }
        __syncthreads();
        const unsigned act = sh.active;
        if (act == 0) break;
        const int nact = ctl.nact;
        pa.cur = it & 1;  // p_new goes to (cur ? S.q : S.p), p_old is the other buffer
        pa.first = it == 0;

        // ---- phase A ---------------------------------------------------------------------------------------------------
        fence_proxy_async_global();
        pa.slot = kSlotA;
        pipe_pass<NDIM, PER, DIAG, kPhaseA>(g, PU, s, maps, sh, sm, partials, kstep, pa);
        if (threadIdx.x < nact) r_slots[threadIdx.x] = kSlotA;
        fence_proxy_async_global();
        reduce_all<SLAB>(partials, bar, sh, link, seq, r_slots, r_sys, nact, tl);
        if ((int)threadIdx.x < nact) sh.alpha[r_sys[threadIdx.x]] = sh.rz[r_sys[threadIdx.x]] / sh.glob[threadIdx.x];
        __syncthreads();

        // ---- phase B ---------------------------------------------------------------------------------------------------
        fence_proxy_async_global();
        pa.slot = kSlotB;
        pipe_pass<NDIM, PER, DIAG, kPhaseB>(g, PU, s, maps, sh, sm, partials, kstep, pa);
        if (threadIdx.x < nact) r_slots[threadIdx.x] = kSlotB;
        fence_proxy_async_global();
        reduce_all<SLAB>(partials, bar, sh, link, seq, r_slots, r_sys, nact, tl);
        if ((int)threadIdx.x < nact) {
            const int b = r_sys[threadIdx.x];
            sh.rr[b] = sh.glob[threadIdx.x];
            sh.iters[b] += 1;  // scipy calls the callback once per completed iteration
            if (!pc) {
                sh.rho_prev[b] = sh.rz[b];
                sh.rz[b] = sh.rr[b];
            }
        }
        __syncthreads();

        // ---- phase C (preconditioned runs): only for the systems that go on ------------------------------------------
        if (pc) {
            if (threadIdx.x == 0) {
                int na = 0;
                for (int b = 0; b < B; ++b) {
                    if (!(act >> b & 1u)) continue;
                    if (sqrt(sh.rr[b]) < sh.rtol[b] * sqrt(sh.bb[b]) || it + 1 >= maxiter) continue;
                    ctl.act_list[na] = b;
                    r_sys[na] = b;
                    r_slots[na] = kSlotC;
                    ++na;
                }
                ctl.nact = na;
            }
            __syncthreads();
            const int nc = ctl.nact;
            __syncthreads();  // everybody holds nc before thread 0 may rewrite the list at the top of the loop
            if (nc > 0) {
                fence_proxy_async_global();
                pa.slot = kSlotC;
                pipe_pass<NDIM, PER, DIAG, kPhaseC>(g, PU, s, maps, sh, sm, partials, kstep, pa);
                fence_proxy_async_global();
                reduce_all<SLAB>(partials, bar, sh, link, seq, r_slots, r_sys, nc, tl);
                if ((int)threadIdx.x < nc) {
                    const int b = r_sys[threadIdx.x];
                    sh.rho_prev[b] = sh.rz[b];
                    sh.rz[b] = sh.glob[threadIdx.x];
                }
                __syncthreads();
            }
        }
    }
    if (SLAB && blockIdx.x == 0 && threadIdx.x == 0) *link->seq = seq;
}

#ifdef SDCB200_ONE_CTA
#define SDCB200_SOLVER_MAXNREG 128
#else
#define SDCB200_SOLVER_MAXNREG 96  // two CTAs per SM (see PipeCfg)
#endif

struct PipeArgs {
    CgArgs cg;
    PipeMaps maps;
    SlabLink link;  // used by the SLAB instantiation only
};

template <int NDIM, bool SLAB, bool PER>
__global__ void __maxnreg__(SDCB200_SOLVER_MAXNREG) cg_pipe_kernel(const __grid_constant__ PipeArgs pa) {
    using Smem = PipeSmemT<PER, false>;
    extern __shared__ __align__(128) unsigned char pipe_smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(pipe_smem_raw);
    __shared__ CgShared sh;
    const CgArgs& a = pa.cg;
    pipe_ctl_init(sm.ctl, PipeCfg<PER, false>::kStages);
    if (threadIdx.x < SDCB200_MAX_NODES) sh.rtol[threadIdx.x] = a.rtol;
    __syncthreads();
    unsigned kstep = 0;
    cg_collective_pipe<NDIM, SLAB, PER, false>(a.g, a.B, (1u << a.B) - 1u, a.s, pa.maps, a.maxiter, a.partials, a.bar, sh,
                                                sm, SLAB ? &pa.link : nullptr, kstep, a.timeline);
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.iters_out != nullptr)
        for (int b = 0; b < a.B; ++b) a.iters_out[b] += sh.iters[b];
}

// u^k for small integer k the way numpy evaluates `u**nu` for nu = 2 (a multiplication); general k by repeated
// multiplication.
__device__ __forceinline__ double ipow(double u, int k) {
    double r = u;
    for (int i = 1; i < k; ++i) r = __dmul_rn(r, u);
    return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// K4: Allen-Cahn Newton (AllenCahn_2D_FD.py:137-205) for B node systems in one persistent launch.  Every Newton
// iteration evaluates the nonlinear residual g_b and the Jacobian diagonal d_b of the systems that are still iterating
// (register-marching pass: one of ~200 passes of a Newton step), then runs the pipelined CG on  (d_b I - factor_b a_off
// S) z_b = g_b  for those systems together - periodic wrap sets, diagonal as a tile-only box - and updates u_b -= z_b.
// Systems drop out of the Newton loop individually (residual below newton_tol or iteration budget spent), systems drop
// out of the inner CG individually; counters are per system, so a batched solve counts what separate solves count.
// ---------------------------------------------------------------------------------------------------------------------
struct NewtonPipeArgs {
    NewtonArgs nw;
    PipeMaps maps;
};

__global__ void __maxnreg__(SDCB200_SOLVER_MAXNREG) newton_pipe_kernel(const __grid_constant__ NewtonPipeArgs npa) {
    using Smem = PipeSmemT<true, true>;
    extern __shared__ __align__(128) unsigned char pipe_smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(pipe_smem_raw);
    __shared__ CgShared sh;
    __shared__ int s_newton[SDCB200_MAX_NODES], s_linear[SDCB200_MAX_NODES];
    __shared__ unsigned s_active;
    const NewtonArgs& a = npa.nw;
    const CgArgs& cg = a.cg;
    const Geom& g = cg.g;
    const int B = cg.B;
    const Units U = make_units(g);
    const long long n2 = g.vol / 2;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gstride = (long long)gridDim.x * blockDim.x;
    pipe_ctl_init(sm.ctl, PipeCfg<true, true>::kStages);
    if (threadIdx.x < SDCB200_MAX_NODES) {
        s_newton[threadIdx.x] = 0;
        s_linear[threadIdx.x] = 0;
        sh.rtol[threadIdx.x] = a.lin_tol;
    }
    if (threadIdx.x == 0) s_active = (1u << B) - 1u;
    __syncthreads();
    unsigned kstep = 0;
    for (;;) {
        const unsigned act = s_active;
        // g = u - factor*(A u + 1/eps^2 u (1 - u^nu)) - rhs ;  Jacobian diagonal ;  z = 0  (AllenCahn_2D_FD.py:170,183)
        for (int b = 0; b < B; ++b) {
            if (!(act >> b & 1u)) continue;
            const NewtonSys& N = a.ns[b];
            const Sys& S = cg.s[b];
            double* dvec = const_cast<double*>(S.dvec);
            double* gvec = const_cast<double*>(S.b);
            const double factor = N.factor;
            double gmax = 0.0;
            if (threadIdx.x < kThreads) {
                for (int unit = blockIdx.x; unit < U.per_field; unit += gridDim.x) {
                    stencil_unit<2, true>(g, U, N.u, unit, [&](long long idx, double2 c, double2 nb, bool, bool) {
                        const double2 rhs = ld2(N.rhs + idx);
                        double2 gv, dv;
                        const double cs[2] = {c.x, c.y}, nbs[2] = {nb.x, nb.y}, rh[2] = {rhs.x, rhs.y};
                        double gs[2], ds[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const double u = cs[e];
                            const double Au = fma(a.a_off, nbs[e], a.a_diag * u);
                            const double un = ipow(u, a.nu_exp);
                            if (a.variant == 0) {  // AllenCahn_2D_FD.py:170,183
                                const double react = __dmul_rn(__dmul_rn(a.inv_eps2, u), __dsub_rn(1.0, un));
                                gs[e] = __dsub_rn(__dsub_rn(u, __dmul_rn(factor, __dadd_rn(Au, react))), rh[e]);
                                const double jr = __dmul_rn(a.inv_eps2, __dsub_rn(1.0, __dmul_rn((double)(a.nu_exp + 1), un)));
                                ds[e] = __dsub_rn(1.0, __dmul_rn(factor, __dadd_rn(a.a_diag, jr)));
                            } else {  // allencahn_semiimplicit_v2, :447,453: implicit part A u - u^(nu+1)/eps^2
                                const double term = __dmul_rn(a.inv_eps2, __dmul_rn(un, u));
                                gs[e] = __dsub_rn(__dsub_rn(u, __dmul_rn(factor, __dsub_rn(Au, term))), rh[e]);
                                const double jr = __dmul_rn(a.inv_eps2, __dmul_rn((double)(a.nu_exp + 1), un));
                                ds[e] = __dsub_rn(1.0, __dmul_rn(factor, __dsub_rn(a.a_diag, jr)));
                            }
                        }
                        gv = make_double2(gs[0], gs[1]);
                        dv = make_double2(ds[0], ds[1]);
                        st2(gvec + idx, gv);
                        st2(dvec + idx, dv);
                        st2(S.x + idx, make_double2(0.0, 0.0));
                        gmax = fmax(gmax, fmax(fabs(gv.x), fabs(gv.y)));
                        if (gv.x != gv.x || gv.y != gv.y) gmax = INFINITY;  // NaN: never "converged"
                    });
                }
            }
            gmax = block_max(gmax, sh.scratch);
            put_partial(cg.partials, kSlotNewton, b, gmax);
        }
        grid_barrier(cg.bar);
        for (int b = 0; b < B; ++b) {
            if (!(act >> b & 1u)) continue;
            const double res = grid_max(cg.partials, kSlotNewton, b, sh.scratch);
            if (threadIdx.x == 0) {
                if (a.inexact_ratio > 0.0) sh.rtol[b] = res * a.inexact_ratio;
                if (res < a.newton_tol || s_newton[b] >= a.newton_maxiter) s_active &= ~(1u << b);
            }
        }
        __syncthreads();
        const unsigned go = s_active;
        if (go == 0) break;
        grid_barrier(cg.bar);  // everybody has read the residual norms before anybody can come back and rewrite them

        cg_collective_pipe<2, false, true, true>(g, B, go, cg.s, npa.maps, a.lin_maxiter, cg.partials, cg.bar, sh, sm,
                                                 nullptr, kstep, cg.timeline);
        if ((int)threadIdx.x < B && (go >> threadIdx.x & 1u)) {
            s_linear[threadIdx.x] += sh.iters[threadIdx.x];
            s_newton[threadIdx.x] += 1;
        }
        // u -= z
        for (int b = 0; b < B; ++b) {
            if (!(go >> b & 1u)) continue;
            double* u = a.ns[b].u;
            const double* z = cg.s[b].x;
            for (long long i = gtid; i < n2; i += gstride) {
                double2 uv = ld2(u + 2 * i);
                const double2 zv = ld2(z + 2 * i);
                uv.x = __dsub_rn(uv.x, zv.x);
                uv.y = __dsub_rn(uv.y, zv.y);
                st2(u + 2 * i, uv);
            }
        }
        grid_barrier(cg.bar);
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.counters_out != nullptr) {
        int nn = 0, nl = 0;
        for (int b = 0; b < B; ++b) {
            nn += s_newton[b];
            nl += s_linear[b];
        }
        a.counters_out[0] += nn;
        a.counters_out[1] += nl;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
template <class K>
int coresident_ctas(K kernel, int* out) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
    if (e != cudaSuccess) return fail_cuda("coresident_ctas", e);
    if (per_sm < 1) return fail("coresident_ctas", "solver kernel does not fit on an SM");
    if (per_sm > 4) per_sm = 4;  // 1024 threads/SM already saturate HBM; more CTAs only lengthen the barriers
    *out = per_sm * sm_count();
    return 0;
}

constexpr int kMaxGrid = 148 * 8;  // upper bound used to size the partials area
unsigned long long* g_timeline = nullptr;  // sdcb200_set_timeline

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct WorkLayout {
    size_t field;          // bytes of one guarded field
    size_t guard_bytes;
    size_t partials_off, bar_off, fields_off, total;
};
WorkLayout work_layout(int ndim, int n, int nfields) {
    WorkLayout w;
    const size_t guard = (size_t)sdcb200_guard(ndim, n), vol = (size_t)sdcb200_volume(ndim, n);
    w.guard_bytes = guard * sizeof(double);
    w.field = align_up((guard + vol) * sizeof(double), 256);
    w.partials_off = 0;
    w.bar_off = align_up(kPartialSlots * SDCB200_MAX_NODES * kMaxGrid * sizeof(double), 256);
    w.fields_off = w.bar_off + 256;
    w.total = w.fields_off + (size_t)nfields * w.field;
    return w;
}

// Slab workspace: identical offsets on every rank (sized for the thickest slab) so that peer addresses are
// "peer base + my offset":  partials | barrier word | mailbox (counter, error flag, flags, values) | 3*B work fields
struct SlabWorkLayout {
    size_t field, guard_bytes;
    size_t partials_off, bar_off, seq_off, err_off, flags_off, vals_off, fields_off, total;
};
SlabWorkLayout slab_work_layout(int n, int nz_max, int nfields) {
    SlabWorkLayout w;
    const size_t P = (size_t)(n + (n & 1)), sz = P * P;
    const size_t guard = (sz + 15) / 16 * 16;
    w.guard_bytes = guard * sizeof(double);
    w.field = align_up((guard + sz * (size_t)(nz_max + 1)) * sizeof(double), 256);
    w.partials_off = 0;
    w.bar_off = align_up(kPartialSlots * SDCB200_MAX_NODES * kMaxGrid * sizeof(double), 256);
    w.seq_off = w.bar_off + 256;
    w.err_off = w.seq_off + 128;
    w.flags_off = w.seq_off + 256;
    w.vals_off = w.flags_off + align_up(2 * kMaxRanks * sizeof(unsigned long long), 256);
    w.fields_off = w.vals_off + align_up(2 * kMaxRanks * kMailVals * sizeof(double), 256);
    w.total = w.fields_off + (size_t)nfields * w.field;
    return w;
}

// Degree-1 Chebyshev polynomial preconditioner for M = m_diag I + m_off S (S = sum of the 2*ndim neighbours, spectrum
// inside (-2 ndim, 2 ndim)): two steps of the Chebyshev semi-iteration for M z = r from z = 0 give
// z = pc_a r + pc_b M r.  It roughly halves the CG iteration count at the price of one more stencil pass per iteration.
inline void chebyshev1(int ndim, double m_diag, double m_off, double* pc_a, double* pc_b) {
    const double w = 2.0 * ndim * fabs(m_off);
    const double lmin = m_diag - w, lmax = m_diag + w;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin);
    if (!(delta > 0.0) || !(lmin > 0.0)) {  // M is (a multiple of) the identity, or not known to be definite
        *pc_a = 1.0 / theta;
        *pc_b = 0.0;
        return;
    }
    const double sigma = theta / delta, rho0 = 1.0 / sigma, rho1 = 1.0 / (2.0 * sigma - rho0);
    *pc_a = (1.0 + rho1 * rho0) / theta + 2.0 * rho1 / delta;
    *pc_b = -2.0 * rho1 / (delta * theta);
}

template <int NDIM, bool SLAB = false, bool PER = false>
int launch_cg_pipe(CgArgs& cg, cudaStream_t s, const SlabLink* link = nullptr) {
    int grid = 0;
    const size_t smem = sizeof(PipeSmemT<PER, false>);
    if (int rc = pipe_grid<cg_pipe_kernel<NDIM, SLAB, PER>>(smem, &grid)) return rc;
    if (grid > kMaxGrid) grid = kMaxGrid;
    static thread_local PipeArgs a;  // ~16 KB: kept off the stack
    a.cg = cg;
    a.cg.timeline = g_timeline;
    if (link != nullptr) a.link = *link;
    for (int b = 0; b < cg.B; ++b)
        if (int rc = encode_system_maps(a.maps.m[b], cg.g, cg.s[b])) return rc;
    void* params[] = {&a};
    SDC_CUDA_OK(cudaLaunchCooperativeKernel((void*)cg_pipe_kernel<NDIM, SLAB, PER>, dim3(grid), dim3(kPipeThreads), params,
                                            smem, s));
    return 0;
}

template <int NDIM, bool PER>
int launch_cg(CgArgs& a, cudaStream_t s) {
    int grid = 0;
    if (int rc = coresident_ctas(cg_kernel<NDIM, PER>, &grid)) return rc;
    if (grid > kMaxGrid) grid = kMaxGrid;
    void* params[] = {&a};
    SDC_CUDA_OK(cudaLaunchCooperativeKernel((void*)cg_kernel<NDIM, PER>, dim3(grid), dim3(kThreads), params, 0, s));
    return 0;
}

}  // namespace
}  // namespace sdcb200

using namespace sdcb200;

extern "C" {

int sdcb200_device_info(int* sm, int* cc_major, int* cc_minor, int* solver_ctas) {
    int dev = 0;
    SDC_CUDA_OK(cudaGetDevice(&dev));
    if (sm) SDC_CUDA_OK(cudaDeviceGetAttribute(sm, cudaDevAttrMultiProcessorCount, dev));
    if (cc_major) SDC_CUDA_OK(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_minor) SDC_CUDA_OK(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (solver_ctas) {
        if (int rc = pipe_grid<cg_pipe_kernel<3, false, false>>(sizeof(PipeSmemT<false, false>), solver_ctas)) return rc;
    }
    return 0;
}

int sdcb200_set_timeline(unsigned long long* dev_ns12) {
    g_timeline = dev_ns12;
    return 0;
}

size_t sdcb200_cg_workspace_bytes(int ndim, int n, int B) { return work_layout(ndim, n, 4 * B).total; }

int sdcb200_heat_cg_solve(int ndim, int n, int bc, int B, const double* m_diag_host, const double* m_off_host,
                          const double* const* rhs, double* const* x, double rtol, int maxiter, int precond,
                          void* work, size_t work_bytes, int* iters_dev, void* stream) {
    SDC_REQUIRE(ndim >= 1 && ndim <= 3, "ndim must be 1, 2 or 3");
    SDC_REQUIRE(B >= 1 && B <= SDCB200_MAX_NODES, "B out of range");
    SDC_REQUIRE(n >= 2, "grid too small");
    SDC_REQUIRE(bc == SDCB200_BC_PERIODIC || (n & 1), "dirichlet-zero grids need an odd number of points per dimension");
    SDC_REQUIRE(bc != SDCB200_BC_PERIODIC || !(n & 1), "periodic grids need an even number of points per dimension");
    SDC_REQUIRE(precond == SDCB200_PRECOND_NONE || precond == SDCB200_PRECOND_CHEBYSHEV1, "unknown preconditioner");
    SDC_REQUIRE(precond == SDCB200_PRECOND_NONE || (ndim >= 2 && bc == SDCB200_BC_DIRICHLET),
                "the polynomial preconditioner is implemented for 2-D / 3-D dirichlet-zero grids");
    const WorkLayout w = work_layout(ndim, n, 4 * B);
    SDC_REQUIRE(work != nullptr && work_bytes >= w.total, "workspace too small (see sdcb200_cg_workspace_bytes)");
    SDC_REQUIRE((reinterpret_cast<size_t>(work) & 255u) == 0, "workspace must be 256-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CgArgs a;
    memset(&a, 0, sizeof(a));
    a.g = make_geom(ndim, n, bc);
    a.B = B;
    a.rtol = rtol;
    a.maxiter = maxiter;
    char* base = static_cast<char*>(work);
    a.partials = reinterpret_cast<double*>(base + w.partials_off);
    a.bar = reinterpret_cast<unsigned*>(base + w.bar_off);
    a.iters_out = iters_dev;
    for (int b = 0; b < B; ++b) {
        SDC_REQUIRE(rhs[b] && x[b] && !(reinterpret_cast<size_t>(rhs[b]) & 15u) && !(reinterpret_cast<size_t>(x[b]) & 15u),
                    "rhs / x missing or misaligned");
        Sys& S = a.s[b];
        S.b = rhs[b];
        S.x = x[b];
        char* f = base + w.fields_off + (size_t)(4 * b) * w.field;
        S.r = reinterpret_cast<double*>(f + w.guard_bytes);
        S.p = reinterpret_cast<double*>(f + w.field + w.guard_bytes);
        S.q = reinterpret_cast<double*>(f + 2 * w.field + w.guard_bytes);
        S.dvec = nullptr;
        S.m_diag = m_diag_host[b];
        S.m_off = m_off_host[b];
        if (precond == SDCB200_PRECOND_CHEBYSHEV1) {
            S.z = reinterpret_cast<double*>(f + 3 * w.field + w.guard_bytes);
            chebyshev1(ndim, S.m_diag, S.m_off, &S.pc_a, &S.pc_b);
        }
    }
    SDC_CUDA_OK(cudaMemsetAsync(a.bar, 0, 256, s));
    int rc = 1;
    const bool per = a.g.periodic;
    // 2-D / 3-D grids: bulk-async pipelined passes (periodic grids with wrap sets); 1-D: register-marching passes
    if (ndim == 1) rc = per ? launch_cg<1, true>(a, s) : launch_cg<1, false>(a, s);
    if (ndim == 2) rc = per ? launch_cg_pipe<2, false, true>(a, s) : launch_cg_pipe<2>(a, s);
    if (ndim == 3) rc = per ? launch_cg_pipe<3, false, true>(a, s) : launch_cg_pipe<3>(a, s);
    return rc;
}

size_t sdcb200_slab_cg_workspace_bytes(int n, int nz_max, int B) { return slab_work_layout(n, nz_max, 4 * B).total; }

int sdcb200_heat_cg_solve_slab(int n, int nz, int nz_max, int bc, int B, const double* m_diag_host,
                               const double* m_off_host, const double* const* rhs, double* const* x, double rtol,
                               int maxiter, int precond, int rank, int nranks, const int* nz_of_rank,
                               void* const* work_of_rank, size_t work_bytes, int* iters_dev, void* stream) {
    SDC_REQUIRE(precond == SDCB200_PRECOND_NONE || precond == SDCB200_PRECOND_CHEBYSHEV1, "unknown preconditioner");
    SDC_REQUIRE(bc == SDCB200_BC_DIRICHLET, "slab-decomposed solves are implemented for dirichlet-zero grids");
    SDC_REQUIRE(n >= 3 && (n & 1), "dirichlet-zero grids need an odd number of points per dimension");
    SDC_REQUIRE(B >= 1 && B <= SDCB200_MAX_NODES, "B out of range");
    SDC_REQUIRE(nranks >= 1 && nranks <= kMaxRanks && rank >= 0 && rank < nranks, "rank / nranks out of range");
    SDC_REQUIRE(nz >= 1 && nz <= nz_max && nz_of_rank != nullptr && nz_of_rank[rank] == nz, "inconsistent slab sizes");
    const SlabWorkLayout w = slab_work_layout(n, nz_max, 4 * B);
    SDC_REQUIRE(work_of_rank != nullptr && work_bytes >= w.total, "workspace too small (see sdcb200_slab_cg_workspace_bytes)");
    for (int r = 0; r < nranks; ++r)
        SDC_REQUIRE(work_of_rank[r] != nullptr && (reinterpret_cast<size_t>(work_of_rank[r]) & 255u) == 0,
                    "peer workspace missing or misaligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CgArgs a;
    memset(&a, 0, sizeof(a));
    a.g = make_slab_geom(n, nz, bc);
    a.B = B;
    a.rtol = rtol;
    a.maxiter = maxiter;
    char* base = static_cast<char*>(work_of_rank[rank]);
    a.partials = reinterpret_cast<double*>(base + w.partials_off);
    a.bar = reinterpret_cast<unsigned*>(base + w.bar_off);
    a.iters_out = iters_dev;
    SlabLink L;
    memset(&L, 0, sizeof(L));
    L.rank = rank;
    L.nranks = nranks;
    L.has_lo = rank > 0;
    L.has_hi = rank + 1 < nranks;
    // Ranks launch their solver kernels without host synchronisation: a peer may legitimately be late (host stall, GC,
    // a debugger).  Default 300 s; the trap that follows a time-out kills this rank's context, which the host sees as a
    // CUDA error at its next call instead of a GPU that spins forever.
    {
        const char* env = getenv("SDCB200_PEER_TIMEOUT_S");
        double secs = env != nullptr ? atof(env) : 300.0;
        if (!(secs > 0.0)) secs = 300.0;
        L.timeout_ns = (unsigned long long)(secs * 1e9);
    }
    L.seq = reinterpret_cast<unsigned long long*>(base + w.seq_off);
    L.error = reinterpret_cast<int*>(base + w.err_off);
    for (int r = 0; r < nranks; ++r) {
        char* pb = static_cast<char*>(work_of_rank[r]);
        L.flags_of[r] = reinterpret_cast<unsigned long long*>(pb + w.flags_off);
        L.vals_of[r] = reinterpret_cast<double*>(pb + w.vals_off);
    }
    const long long sz = a.g.sz;
    for (int b = 0; b < B; ++b) {
        SDC_REQUIRE(rhs[b] && x[b] && !(reinterpret_cast<size_t>(rhs[b]) & 15u) && !(reinterpret_cast<size_t>(x[b]) & 15u),
                    "rhs / x missing or misaligned");
        Sys& S = a.s[b];
        S.b = rhs[b];
        S.x = x[b];
        const size_t off_r = w.fields_off + (size_t)(4 * b) * w.field + w.guard_bytes;
        const size_t off_z = off_r + 3 * w.field;
        S.r = reinterpret_cast<double*>(base + off_r);
        S.p = reinterpret_cast<double*>(base + off_r + w.field);
        S.q = reinterpret_cast<double*>(base + off_r + 2 * w.field);
        S.dvec = nullptr;
        S.m_diag = m_diag_host[b];
        S.m_off = m_off_host[b];
        if (precond == SDCB200_PRECOND_CHEBYSHEV1) {
            S.z = reinterpret_cast<double*>(base + off_z);
            chebyshev1(3, S.m_diag, S.m_off, &S.pc_a, &S.pc_b);
            if (L.has_lo)
                L.lo_z_halo[b] = reinterpret_cast<double*>(static_cast<char*>(work_of_rank[rank - 1]) + off_z) +
                                 sz * nz_of_rank[rank - 1];
            if (L.has_hi)
                L.hi_z_halo[b] = reinterpret_cast<double*>(static_cast<char*>(work_of_rank[rank + 1]) + off_z) - sz;
        }
        if (L.has_lo)  // the lower neighbour's upper halo plane: plane nz_lo of its r
            L.lo_r_halo[b] = reinterpret_cast<double*>(static_cast<char*>(work_of_rank[rank - 1]) + off_r) +
                             sz * nz_of_rank[rank - 1];
        if (L.has_hi)  // the upper neighbour's lower halo plane: plane -1 of its r
            L.hi_r_halo[b] = reinterpret_cast<double*>(static_cast<char*>(work_of_rank[rank + 1]) + off_r) - sz;
    }
    SDC_CUDA_OK(cudaMemsetAsync(a.bar, 0, 256, s));
    return launch_cg_pipe<3, true, false>(a, s, &L);
}

size_t sdcb200_newton_workspace_bytes(int n, int B) { return work_layout(2, n, 6 * B).total; }

int sdcb200_allencahn_newton_solve(int n, int B, int variant, const double* factor_host, double a_diag, double a_off,
                                   double inv_eps2, int nu_exp, const double* const* rhs, double* const* u, double newton_tol,
                                   int newton_maxiter, double lin_tol, int lin_maxiter, double inexact_ratio, void* work,
                                   size_t work_bytes, int* counters_dev, void* stream) {
    SDC_REQUIRE(n >= 2 && !(n & 1), "periodic grid needs an even number of points per dimension");
    SDC_REQUIRE(nu_exp >= 1, "nu must be a positive integer");
    SDC_REQUIRE(B >= 1 && B <= SDCB200_MAX_NODES, "B out of range");
    SDC_REQUIRE(variant == 0 || variant == 1, "variant must be 0 (fully implicit) or 1 (semi-implicit v2)");
    const WorkLayout w = work_layout(2, n, 6 * B);
    SDC_REQUIRE(work != nullptr && work_bytes >= w.total, "workspace too small (see sdcb200_newton_workspace_bytes)");
    SDC_REQUIRE((reinterpret_cast<size_t>(work) & 255u) == 0, "workspace must be 256-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    static thread_local NewtonPipeArgs npa;
    memset(&npa.nw, 0, sizeof(npa.nw));
    NewtonArgs& a = npa.nw;
    a.cg.g = make_geom(2, n, SDCB200_BC_PERIODIC);
    a.cg.B = B;
    a.cg.maxiter = lin_maxiter;
    a.a_diag = a_diag;
    a.a_off = a_off;
    a.inv_eps2 = inv_eps2;
    a.nu_exp = nu_exp;
    a.variant = variant;
    char* base = static_cast<char*>(work);
    for (int b = 0; b < B; ++b) {
        SDC_REQUIRE(rhs[b] && u[b] && !(reinterpret_cast<size_t>(rhs[b]) & 15u) && !(reinterpret_cast<size_t>(u[b]) & 15u),
                    "rhs / u missing or misaligned");
        auto fieldp = [&](int k) {
            return reinterpret_cast<double*>(base + w.fields_off + (size_t)(6 * b + k) * w.field + w.guard_bytes);
        };
        a.ns[b].rhs = rhs[b];
        a.ns[b].u = u[b];
        a.ns[b].factor = factor_host[b];
        Sys& S = a.cg.s[b];
        S.b = fieldp(0);     // Newton residual g
        S.x = fieldp(1);     // Newton update z
        S.dvec = fieldp(2);  // Jacobian diagonal
        S.r = fieldp(3);
        S.p = fieldp(4);
        S.q = fieldp(5);
        S.m_diag = 0.0;
        S.m_off = -(factor_host[b] * a_off);
        if (int rc = encode_system_maps(npa.maps.m[b], a.cg.g, S)) return rc;
    }
    a.newton_tol = newton_tol;
    a.lin_tol = lin_tol;
    a.inexact_ratio = inexact_ratio;
    a.newton_maxiter = newton_maxiter;
    a.lin_maxiter = lin_maxiter;
    a.cg.partials = reinterpret_cast<double*>(base + w.partials_off);
    a.cg.bar = reinterpret_cast<unsigned*>(base + w.bar_off);
    a.cg.timeline = g_timeline;
    a.counters_out = counters_dev;
    SDC_CUDA_OK(cudaMemsetAsync(a.cg.bar, 0, 256, s));
    int grid = 0;
    const size_t smem = sizeof(PipeSmemT<true, true>);
    if (int rc = pipe_grid<newton_pipe_kernel>(smem, &grid)) return rc;
    if (grid > kMaxGrid) grid = kMaxGrid;
    void* params[] = {&npa};
    SDC_CUDA_OK(cudaLaunchCooperativeKernel((void*)newton_pipe_kernel, dim3(grid), dim3(kPipeThreads), params, smem, s));
    return 0;
}

}  // extern "C"
