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
template <bool SLAB>
__device__ __forceinline__ void reduce_all(double* partials, unsigned* bar, CgShared& sh, const SlabLink* link,
                                           unsigned long long& seq, const int* slots, const int* systems, int nv) {
    if (SLAB) grid_barrier_sys(bar);
    else grid_barrier(bar);
    for (int i = 0; i < nv; ++i) {
        const double v = grid_sum(partials, slots[i], systems[i], sh.scratch);
        if (threadIdx.x == 0) (SLAB ? sh.loc : sh.glob)[i] = v;
    }
    if (SLAB) cross_rank_sum(*link, sh, nv, ++seq);
    else __syncthreads();
}

template <int NDIM, bool SLAB>
__device__ void cg_collective_pipe(const Geom& g, int B, const Sys* s, const PipeMaps& maps, double rtol, int maxiter,
                                   double* partials, unsigned* bar, CgShared& sh, PipeSmem& sm, const SlabLink* link) {
    const Units U = make_units(g, (int)gridDim.x);
    const PUnits PU = make_punits(g, 1, (int)gridDim.x);
    const long long n2 = g.owned / 2;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gstride = (long long)gridDim.x * blockDim.x;
    unsigned kstep = 0;
    unsigned long long seq = SLAB ? *link->seq : 0ull;
    __shared__ int r_slots[kMailVals], r_sys[kMailVals];

    // ---- r = b - M x0, ||b||^2, ||r||^2 ------------------------------------------------------------------------------
    for (int b = 0; b < B; ++b) {
        const Sys& S = s[b];
        double bb = 0.0, rr = 0.0;
        if (threadIdx.x < kThreads) {
            double* r_lo = SLAB && link->has_lo ? link->lo_r_halo[b] : nullptr;
            double* r_hi = SLAB && link->has_hi ? link->hi_r_halo[b] : nullptr;
            for (int unit = blockIdx.x; unit < U.per_field; unit += gridDim.x) {
                stencil_unit<NDIM, false>(g, U, S.x, unit, [&](long long idx, double2 c, double2 nb, bool v0, bool v1) {
                    const double2 rhs = ld2(S.b + idx);
                    double2 r;
                    r.x = v0 ? rhs.x - fma(S.m_off, nb.x, S.m_diag * c.x) : 0.0;
                    r.y = v1 ? rhs.y - fma(S.m_off, nb.y, S.m_diag * c.y) : 0.0;
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
        for (int b = 0; b < B; ++b) {
            r_slots[2 * b] = kSlotSetup0;
            r_sys[2 * b] = b;
            r_slots[2 * b + 1] = kSlotSetup1;
            r_sys[2 * b + 1] = b;
        }
    }
    fence_proxy_async_global();
    reduce_all<SLAB>(partials, bar, sh, link, seq, r_slots, r_sys, 2 * B);
    if (threadIdx.x == 0) {
        unsigned act = 0;
        for (int b = 0; b < B; ++b) {
            sh.bb[b] = sh.glob[2 * b];
            sh.rr[b] = sh.glob[2 * b + 1];
            sh.iters[b] = 0;
            sh.rho_prev[b] = 1.0;
            if (sh.bb[b] != 0.0) act |= 1u << b;  // scipy: ||b|| == 0 -> return b
        }
        sh.active = act;
    }
    __syncthreads();
    for (int b = 0; b < B; ++b) {
        if (sh.bb[b] == 0.0) {
            for (long long i = gtid; i < n2; i += gstride) st2(s[b].x + 2 * i, make_double2(0.0, 0.0));
        }
    }

    // preconditioned runs: z = C(M) r for the systems that will iterate, r.z
    const bool pc = s[0].z != nullptr;
    if (pc) {
        if (threadIdx.x == 0) {
            int na = 0;
            for (int b = 0; b < B; ++b)
                if (sh.active >> b & 1u) {
                    sm.act_list[na] = b;
                    r_sys[na] = b;
                    r_slots[na] = kSlotC;
                    ++na;
                }
            sm.nact = na;
        }
        __syncthreads();
        if (sm.nact > 0) {
            const int nc = sm.nact;
            fence_proxy_async_global();
            pipe_phase_c<NDIM>(g, PU, s, maps, sm, partials, kstep, link);
            fence_proxy_async_global();
            reduce_all<SLAB>(partials, bar, sh, link, seq, r_slots, r_sys, nc);
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
                const double atol = rtol * sqrt(sh.bb[b]);
                if (sqrt(sh.rr[b]) < atol || it >= maxiter) {
                    act &= ~(1u << b);
                } else {
                    sh.beta[b] = it > 0 ? sh.rz[b] / sh.rho_prev[b] : 0.0;
                    sm.act_list[na] = b;
                    r_sys[na] = b;
                    ++na;
                }
            }
            sh.active = act;
            sm.nact = na;
        }
        __syncthreads();
        const unsigned act = sh.active;
        if (act == 0) break;
        const int nact = sm.nact;
        const int cur = it & 1;  // p_new goes to (cur ? S.q : S.p), p_old is the other buffer

        // ---- phase A ---------------------------------------------------------------------------------------------------
        fence_proxy_async_global();
        pipe_phase_a<NDIM>(g, PU, s, maps, it == 0, cur, sh, sm, partials, kstep);
        if (threadIdx.x < nact) r_slots[threadIdx.x] = kSlotA;
        fence_proxy_async_global();
        reduce_all<SLAB>(partials, bar, sh, link, seq, r_slots, r_sys, nact);
        if ((int)threadIdx.x < nact) sh.alpha[r_sys[threadIdx.x]] = sh.rz[r_sys[threadIdx.x]] / sh.glob[threadIdx.x];
        __syncthreads();

        // ---- phase B ---------------------------------------------------------------------------------------------------
        fence_proxy_async_global();
        pipe_phase_b<NDIM>(g, PU, s, maps, cur, sh, sm, partials, kstep, link);
        if (threadIdx.x < nact) r_slots[threadIdx.x] = kSlotB;
        fence_proxy_async_global();
        reduce_all<SLAB>(partials, bar, sh, link, seq, r_slots, r_sys, nact);
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
                    if (sqrt(sh.rr[b]) < rtol * sqrt(sh.bb[b]) || it + 1 >= maxiter) continue;
                    sm.act_list[na] = b;
                    r_sys[na] = b;
                    r_slots[na] = kSlotC;
                    ++na;
                }
                sm.nact = na;
            }
            __syncthreads();
            const int nc = sm.nact;
            __syncthreads();  // everybody holds nc before thread 0 may rewrite the list at the top of the loop
            if (nc > 0) {
                fence_proxy_async_global();
                pipe_phase_c<NDIM>(g, PU, s, maps, sm, partials, kstep, link);
                fence_proxy_async_global();
                reduce_all<SLAB>(partials, bar, sh, link, seq, r_slots, r_sys, nc);
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

struct PipeArgs {
    CgArgs cg;
    PipeMaps maps;
    SlabLink link;  // used by the SLAB instantiation only
};

template <int NDIM, bool SLAB>
__global__ void __maxnreg__(112) cg_pipe_kernel(const __grid_constant__ PipeArgs pa) {
    extern __shared__ __align__(128) unsigned char pipe_smem_raw[];
    PipeSmem& sm = *reinterpret_cast<PipeSmem*>(pipe_smem_raw);
    __shared__ CgShared sh;
    const CgArgs& a = pa.cg;
    pipe_smem_init(sm);
    cg_collective_pipe<NDIM, SLAB>(a.g, a.B, a.s, pa.maps, a.rtol, a.maxiter, a.partials, a.bar, sh, sm,
                                   SLAB ? &pa.link : nullptr);
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

__global__ void __launch_bounds__(kThreads) newton_kernel(const __grid_constant__ NewtonArgs a) {
    __shared__ CgShared sh;
    __shared__ Sys sys;
    __shared__ int s_newton, s_linear;
    const Geom& g = a.g;
    const Units U = make_units(g);
    const long long n2 = g.vol / 2;
    const long long gtid = (long long)blockIdx.x * kThreads + threadIdx.x;
    const long long gstride = (long long)gridDim.x * kThreads;
    if (threadIdx.x == 0) {
        s_newton = 0;
        s_linear = 0;
        sys.b = a.gvec;
        sys.x = a.z;
        sys.r = a.r;
        sys.p = a.p;
        sys.q = a.q;
        sys.dvec = a.dvec;
        sys.m_diag = 0.0;
        sys.m_off = -(a.factor * a.a_off);
    }
    __syncthreads();
    double lin_tol = a.lin_tol;
    int n = 0;
    while (n < a.newton_maxiter) {
        // g = u - factor*(A u + 1/eps^2 u (1 - u^nu)) - rhs ;  Jacobian diagonal ;  z = 0  (AllenCahn_2D_FD.py:170,183)
        double gmax = 0.0;
        for (int unit = blockIdx.x; unit < U.per_field; unit += gridDim.x) {
            stencil_unit<2, true>(g, U, a.u, unit, [&](long long idx, double2 c, double2 nb, bool, bool) {
                const double2 rhs = ld2(a.rhs + idx);
                double2 gv, dv;
                {
                    const double Au = fma(a.a_off, nb.x, a.a_diag * c.x);
                    const double un = ipow(c.x, a.nu_exp);
                    const double react = __dmul_rn(__dmul_rn(a.inv_eps2, c.x), __dsub_rn(1.0, un));
                    gv.x = __dsub_rn(__dsub_rn(c.x, __dmul_rn(a.factor, __dadd_rn(Au, react))), rhs.x);
                    const double jr = __dmul_rn(a.inv_eps2, __dsub_rn(1.0, __dmul_rn((double)(a.nu_exp + 1), un)));
                    dv.x = __dsub_rn(1.0, __dmul_rn(a.factor, __dadd_rn(a.a_diag, jr)));
                }
                {
                    const double Au = fma(a.a_off, nb.y, a.a_diag * c.y);
                    const double un = ipow(c.y, a.nu_exp);
                    const double react = __dmul_rn(__dmul_rn(a.inv_eps2, c.y), __dsub_rn(1.0, un));
                    gv.y = __dsub_rn(__dsub_rn(c.y, __dmul_rn(a.factor, __dadd_rn(Au, react))), rhs.y);
                    const double jr = __dmul_rn(a.inv_eps2, __dsub_rn(1.0, __dmul_rn((double)(a.nu_exp + 1), un)));
                    dv.y = __dsub_rn(1.0, __dmul_rn(a.factor, __dadd_rn(a.a_diag, jr)));
                }
                st2(a.gvec + idx, gv);
                st2(a.dvec + idx, dv);
                st2(a.z + idx, make_double2(0.0, 0.0));
                gmax = fmax(gmax, fmax(fabs(gv.x), fabs(gv.y)));
                if (gv.x != gv.x || gv.y != gv.y) gmax = INFINITY;  // NaN: never "converged"
            });
        }
        gmax = block_max(gmax, sh.scratch);
        put_partial(a.partials, kSlotC, 0, gmax);
        grid_barrier(a.bar);
        const double res = grid_max(a.partials, kSlotC, 0, sh.scratch);
        if (a.inexact_ratio > 0.0) lin_tol = res * a.inexact_ratio;
        if (res < a.newton_tol) break;
        grid_barrier(a.bar);  // everybody has read the residual norm before anybody can come back here and rewrite it

        cg_collective<2, true>(g, 1, &sys, lin_tol, a.lin_maxiter, a.partials, a.bar, sh);
        if (threadIdx.x == 0) {
            s_linear += sh.iters[0];
            s_newton += 1;
        }
        // u -= z
        for (long long i = gtid; i < n2; i += gstride) {
            double2 u = ld2(a.u + 2 * i);
            const double2 z = ld2(a.z + 2 * i);
            u.x = __dsub_rn(u.x, z.x);
            u.y = __dsub_rn(u.y, z.y);
            st2(a.u + 2 * i, u);
        }
        grid_barrier(a.bar);
        ++n;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.counters_out != nullptr) {
        a.counters_out[0] += s_newton;
        a.counters_out[1] += s_linear;
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

template <int NDIM, bool SLAB>
int pipe_grid(int* out) {
    static int cached = 0;
    if (cached == 0) {
        SDC_CUDA_OK(cudaFuncSetAttribute(cg_pipe_kernel<NDIM, SLAB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(PipeSmem)));
        int per_sm = 0;
        SDC_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_pipe_kernel<NDIM, SLAB>, kPipeThreads,
                                                                  sizeof(PipeSmem)));
        if (per_sm < 1) return fail("pipe_grid", "pipelined solver kernel does not fit on an SM");
        if (per_sm > 2) per_sm = 2;
        cached = per_sm * sm_count();
    }
    *out = cached;
    return 0;
}

// ---- tensor maps (driver entry point resolved through the runtime: no link-time dependency on libcuda) ------------
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

// Tiled fp64 map over a walled field: 2-D  {P, P} from element (0,0);  3-D  {P, P, nz+2} starting ONE PLANE BELOW the
// field (guard = lower halo plane) up to and including plane nz (wall / upper halo plane).
int encode_field_map(CUtensorMap* map, const Geom& g, const double* field, int box_x, int box_y) {
    TensorMapEncodeFn enc = nullptr;
    if (int rc = tensor_map_encoder(&enc)) return rc;
    const cuuint32_t rank = (cuuint32_t)g.ndim;
    cuuint64_t dims[3] = {(cuuint64_t)g.P, (cuuint64_t)g.P, (cuuint64_t)(g.nz + 2)};
    cuuint64_t strides[2] = {(cuuint64_t)g.sy * 8u, (cuuint64_t)g.sz * 8u};
    cuuint32_t box[3] = {(cuuint32_t)box_x, (cuuint32_t)box_y, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    void* base = const_cast<double*>(g.ndim == 3 ? field - g.sz : field);
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank, base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("encode_field_map", "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return 0;
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

template <int NDIM, bool SLAB = false>
int launch_cg_pipe(CgArgs& cg, cudaStream_t s, const SlabLink* link = nullptr) {
    int grid = 0;
    if (int rc = pipe_grid<NDIM, SLAB>(&grid)) return rc;
    if (grid > kMaxGrid) grid = kMaxGrid;
    static thread_local PipeArgs a;  // 6 KB: kept off the stack
    a.cg = cg;
    if (link != nullptr) a.link = *link;
    for (int b = 0; b < cg.B; ++b) {
        const Sys& S = cg.s[b];
        if (int rc = encode_field_map(&a.maps.m[b][kMapRHalo], cg.g, S.r, kPHX, kPHY)) return rc;
        if (int rc = encode_field_map(&a.maps.m[b][kMapPHalo], cg.g, S.p, kPHX, kPHY)) return rc;
        if (int rc = encode_field_map(&a.maps.m[b][kMapQHalo], cg.g, S.q, kPHX, kPHY)) return rc;
        if (int rc = encode_field_map(&a.maps.m[b][kMapRCentre], cg.g, S.r, kPX, kPY)) return rc;
        if (int rc = encode_field_map(&a.maps.m[b][kMapXCentre], cg.g, S.x, kPX, kPY)) return rc;
        if (S.z != nullptr)
            if (int rc = encode_field_map(&a.maps.m[b][kMapZHalo], cg.g, S.z, kPHX, kPHY)) return rc;
    }
    void* params[] = {&a};
    SDC_CUDA_OK(cudaLaunchCooperativeKernel((void*)cg_pipe_kernel<NDIM, SLAB>, dim3(grid), dim3(kPipeThreads), params,
                                            sizeof(PipeSmem), s));
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
        if (int rc = pipe_grid<3, false>(solver_ctas)) return rc;
    }
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
    if (ndim == 1) rc = per ? launch_cg<1, true>(a, s) : launch_cg<1, false>(a, s);
    // heat solves on Dirichlet grids (the 2-D / 3-D benchmark configurations): bulk-async pipelined passes;
    // periodic and 1-D grids: register-marching passes
    if (ndim == 2) rc = per ? launch_cg<2, true>(a, s) : launch_cg_pipe<2>(a, s);
    if (ndim == 3) rc = per ? launch_cg<3, true>(a, s) : launch_cg_pipe<3>(a, s);
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
    return launch_cg_pipe<3, true>(a, s, &L);
}

size_t sdcb200_newton_workspace_bytes(int n) { return work_layout(2, n, 6).total; }

int sdcb200_allencahn_newton_solve(int n, double factor, double a_diag, double a_off, double inv_eps2, int nu_exp,
                                   const double* rhs, double* u, double newton_tol, int newton_maxiter,
                                   double lin_tol, int lin_maxiter, double inexact_ratio, void* work,
                                   size_t work_bytes, int* counters_dev, void* stream) {
    SDC_REQUIRE(n >= 2 && !(n & 1), "periodic grid needs an even number of points per dimension");
    SDC_REQUIRE(nu_exp >= 1, "nu must be a positive integer");
    const WorkLayout w = work_layout(2, n, 6);
    SDC_REQUIRE(work != nullptr && work_bytes >= w.total, "workspace too small (see sdcb200_newton_workspace_bytes)");
    SDC_REQUIRE((reinterpret_cast<size_t>(work) & 255u) == 0, "workspace must be 256-byte aligned");
    SDC_REQUIRE(rhs && u && !(reinterpret_cast<size_t>(rhs) & 15u) && !(reinterpret_cast<size_t>(u) & 15u),
                "rhs / u missing or misaligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    NewtonArgs a;
    memset(&a, 0, sizeof(a));
    a.g = make_geom(2, n, SDCB200_BC_PERIODIC);
    a.factor = factor;
    a.a_diag = a_diag;
    a.a_off = a_off;
    a.inv_eps2 = inv_eps2;
    a.nu_exp = nu_exp;
    a.rhs = rhs;
    a.u = u;
    char* base = static_cast<char*>(work);
    auto fieldp = [&](int k) { return reinterpret_cast<double*>(base + w.fields_off + (size_t)k * w.field + w.guard_bytes); };
    a.gvec = fieldp(0);
    a.z = fieldp(1);
    a.dvec = fieldp(2);
    a.r = fieldp(3);
    a.p = fieldp(4);
    a.q = fieldp(5);
    a.newton_tol = newton_tol;
    a.lin_tol = lin_tol;
    a.inexact_ratio = inexact_ratio;
    a.newton_maxiter = newton_maxiter;
    a.lin_maxiter = lin_maxiter;
    a.partials = reinterpret_cast<double*>(base + w.partials_off);
    a.bar = reinterpret_cast<unsigned*>(base + w.bar_off);
    a.counters_out = counters_dev;
    SDC_CUDA_OK(cudaMemsetAsync(a.bar, 0, 256, s));
    int grid = 0;
    if (int rc = coresident_ctas(newton_kernel, &grid)) return rc;
    if (grid > kMaxGrid) grid = kMaxGrid;
    void* params[] = {&a};
    SDC_CUDA_OK(cudaLaunchCooperativeKernel((void*)newton_kernel, dim3(grid), dim3(kThreads), params, 0, s));
    return 0;
}

}  // extern "C"
