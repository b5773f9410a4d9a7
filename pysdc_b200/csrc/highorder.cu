// Centred finite-difference Laplacians of order 4, 6 and 8: eval_f and the CG node solves for the
// stencils of pySDC/helpers/problem_helper.py:19-80 with the boundary treatment of :133-201 - periodic wrap, or, on
// dirichlet-zero grids, the one-sided closure rows the reference derives for the `order/2` points next to each boundary
// (which make the matrix slightly non-symmetric; the reference runs CG on it all the same, and so does this).
//
// The operator is a Kronecker sum: A = sum_d (1-D operator along axis d).  Every thread owns one grid point and gathers
// its 2*h*ndim neighbours with plain global loads (L1/L2 serve the re-reads); rows near a Dirichlet boundary take their
// order+1 coefficients from the closure table.  These wide stencils are NOT on the TMA pipeline of the order-2 path:
// correctness and coverage first ("next" row of the scope table), the halo boxes of cg_pipe.cuh would need to be
// 2h wide.  The CG is the same persistent, node-batched, device-resident iteration as cg.cu (scipy's recurrence and
// stopping test, fixed-order reductions), run as three passes per iteration.
#include "fdop.cuh"

namespace sdcb200 {
namespace {

struct HoEvalArgs {
    Geom g;
    HoOp op;
    int B;
    const double* u[SDCB200_MAX_NODES + 1];
    double* f[SDCB200_MAX_NODES + 1];
    double* fexpl[SDCB200_MAX_NODES + 1];
    double gt[SDCB200_MAX_NODES + 1];
    const double* profile;
};

template <int H>
__global__ void __launch_bounds__(kThreads) ho_eval_kernel(const __grid_constant__ HoEvalArgs a) {
    for (int b = 0; b < a.B; ++b) {
        const double* u = a.u[b];
        ho_points(a.g, [&](long long idx, int x, int y, int z) {
            a.f[b][idx] = ho_apply<H>(a.op, a.g, u, x, y, z);
            if (a.profile != nullptr) a.fexpl[b][idx] = __dmul_rn(a.profile[idx], a.gt[b]);
        });
    }
}

struct HoCgArgs {
    Geom g;
    HoOp op;
    int B;
    Sys s[SDCB200_MAX_NODES];
    double factor[SDCB200_MAX_NODES];
    double rtol;
    int maxiter;
    double* partials;
    unsigned* bar;
    int* iters_out;
};

// M v = v - factor * A v
template <int H>
__device__ __forceinline__ double ho_m(const HoCgArgs& a, int b, const double* v, long long idx, int x, int y, int z) {
    return fma(-a.factor[b], ho_apply<H>(a.op, a.g, v, x, y, z), v[idx]);
}

template <int H>
__global__ void __launch_bounds__(kThreads) ho_cg_kernel(const __grid_constant__ HoCgArgs a) {
    __shared__ CgShared sh;
    const Geom& g = a.g;
    const int B = a.B;
    // r = b - M x0, ||b||^2, ||r||^2
    for (int b = 0; b < B; ++b) {
        const Sys& S = a.s[b];
        double bb = 0.0, rr = 0.0;
        ho_points(g, [&](long long idx, int x, int y, int z) {
            const double rhs = S.b[idx];
            const double r = rhs - ho_m<H>(a, b, S.x, idx, x, y, z);
            S.r[idx] = r;
            bb = fma(rhs, rhs, bb);
            rr = fma(r, r, rr);
        });
        bb = block_sum(bb, sh.scratch);
        rr = block_sum(rr, sh.scratch);
        put_partial(a.partials, kSlotSetup0, b, bb);
        put_partial(a.partials, kSlotSetup1, b, rr);
    }
    grid_barrier(a.bar);
    for (int b = 0; b < B; ++b) {
        const double bb = grid_sum(a.partials, kSlotSetup0, b, sh.scratch);
        const double rr = grid_sum(a.partials, kSlotSetup1, b, sh.scratch);
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
    for (int b = 0; b < B; ++b)
        if (sh.bb[b] == 0.0) {
            Quad zero;
            zero.a = zero.b = make_double2(0.0, 0.0);
            flat_quads(g.vol, [&](long long i, bool full) { stq(a.s[b].x, i, zero, full); });
        }

    for (int it = 0;; ++it) {
        if (threadIdx.x == 0) {
            unsigned act = sh.active;
            for (int b = 0; b < B; ++b) {
                if (!(act >> b & 1u)) continue;
                if (sqrt(sh.rr[b]) < a.rtol * sqrt(sh.bb[b]) || it >= a.maxiter) act &= ~(1u << b);
                else sh.beta[b] = it > 0 ? sh.rr[b] / sh.rho_prev[b] : 0.0;
            }
            sh.active = act;
        }
        __syncthreads();
        const unsigned act = sh.active;
        if (act == 0) break;
        // pass 0: p = r + beta p   (scipy: p *= beta; p += r)
        for (int b = 0; b < B; ++b) {
            if (!(act >> b & 1u)) continue;
            const Sys& S = a.s[b];
            const double beta = sh.beta[b];
            flat_quads(g.vol, [&](long long i, bool full) {
                const Quad r = ldq(S.r, i, full);
                if (it == 0) {
                    stq(S.p, i, r, full);
                } else {
                    const Quad pq = ldq(S.p, i, full);
                    stq(S.p, i, qmap([&](int e) { return __dadd_rn(__dmul_rn(qe(pq, e), beta), qe(r, e)); }), full);
                }
            });
        }
        grid_barrier(a.bar);
        // pass 1: q = M p, p.q
        for (int b = 0; b < B; ++b) {
            if (!(act >> b & 1u)) continue;
            const Sys& S = a.s[b];
            double pq = 0.0;
            ho_points(g, [&](long long idx, int x, int y, int z) {
                const double q = ho_m<H>(a, b, S.p, idx, x, y, z);
                S.q[idx] = q;
                pq = fma(S.p[idx], q, pq);
            });
            pq = block_sum(pq, sh.scratch);
            put_partial(a.partials, kSlotA, b, pq);
        }
        grid_barrier(a.bar);
        for (int b = 0; b < B; ++b) {
            if (!(act >> b & 1u)) continue;
            const double pq = grid_sum(a.partials, kSlotA, b, sh.scratch);
            if (threadIdx.x == 0) sh.alpha[b] = sh.rr[b] / pq;
        }
        __syncthreads();
        // pass 2: x += alpha p, r -= alpha q, r.r
        for (int b = 0; b < B; ++b) {
            if (!(act >> b & 1u)) continue;
            const Sys& S = a.s[b];
            const double alpha = sh.alpha[b];
            double rr = 0.0;
            flat_quads(g.vol, [&](long long i, bool full) {
                const Quad xq = ldq(S.x, i, full), pq = ldq(S.p, i, full), rq = ldq(S.r, i, full), qq = ldq(S.q, i, full);
                stq(S.x, i, qmap([&](int e) { return __dadd_rn(qe(xq, e), __dmul_rn(alpha, qe(pq, e))); }), full);
                const Quad r = qmap([&](int e) { return __dsub_rn(qe(rq, e), __dmul_rn(alpha, qe(qq, e))); });
                stq(S.r, i, r, full);
                rr = qdot(r, r, rr);
            });
            rr = block_sum(rr, sh.scratch);
            put_partial(a.partials, kSlotB, b, rr);
        }
        grid_barrier(a.bar);
        for (int b = 0; b < B; ++b) {
            if (!(act >> b & 1u)) continue;
            const double rr = grid_sum(a.partials, kSlotB, b, sh.scratch);
            if (threadIdx.x == 0) {
                sh.rho_prev[b] = sh.rr[b];
                sh.rr[b] = rr;
                sh.iters[b] += 1;
            }
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.iters_out != nullptr)
        for (int b = 0; b < B; ++b) a.iters_out[b] += sh.iters[b];
}

int fill_op(HoOp& op, int order, int bc, const double* centre, const double* lo, const double* hi) {
    SDC_REQUIRE(order == 4 || order == 6 || order == 8, "order must be 4, 6 or 8 (order 2 has its own kernels)");
    memset(&op, 0, sizeof(op));
    op.h = order / 2;
    for (int k = -op.h; k <= op.h; ++k) op.cf[k + kHoMaxH] = centre[k < 0 ? -k : k];
    if (bc == SDCB200_BC_DIRICHLET) {
        SDC_REQUIRE(lo != nullptr && hi != nullptr, "dirichlet-zero grids need the closure rows");
        for (int i = 0; i < op.h; ++i)
            for (int j = 0; j <= order; ++j) {
                op.lo[i][j] = lo[i * (order + 1) + j];
                op.hi[i][j] = hi[i * (order + 1) + j];
            }
    }
    return 0;
}

}  // namespace
}  // namespace sdcb200

using namespace sdcb200;

extern "C" {

int sdcb200_heat_eval_f_ho(int ndim, int n, int bc, int order, const double* centre_host, const double* lo_host,
                           const double* hi_host, int B, const double* const* u, double* const* f_impl,
                           const double* profile, const double* gt_host, double* const* f_expl, void* stream) {
    SDC_REQUIRE(ndim >= 1 && ndim <= 3, "ndim must be 1, 2 or 3");
    SDC_REQUIRE(B >= 1 && B <= SDCB200_MAX_NODES + 1, "B out of range");
    SDC_REQUIRE(n > order, "grid too small for the stencil");
    static thread_local HoEvalArgs a;
    memset(&a, 0, sizeof(a));
    a.g = make_geom(ndim, n, bc);
    if (int rc = fill_op(a.op, order, bc, centre_host, lo_host, hi_host)) return rc;
    a.B = B;
    a.profile = profile;
    for (int b = 0; b < B; ++b) {
        SDC_REQUIRE(ok8(u[b]) && ok8(f_impl[b]), "u / f missing or misaligned");
        a.u[b] = u[b];
        a.f[b] = f_impl[b];
        if (profile != nullptr) {
            SDC_REQUIRE(f_expl && ok8(f_expl[b]) && gt_host, "forcing arguments missing");
            a.fexpl[b] = f_expl[b];
            a.gt[b] = gt_host[b];
        }
    }
    SDC_DISPATCH_H(a.op.h, (ho_eval_kernel<HW><<<sm_count() * 8, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a)));
    SDC_CUDA_OK(cudaGetLastError());
    return 0;
}

size_t sdcb200_cg_ho_workspace_bytes(int ndim, int n, int B) {
    const size_t guard = (size_t)sdcb200_guard(ndim, n), vol = (size_t)sdcb200_volume(ndim, n);
    const size_t field = align_up((guard + vol) * sizeof(double), 256);
    return align_up(kPartialSlots * SDCB200_MAX_NODES * kMaxGrid * sizeof(double), 256) + 256 + (size_t)(3 * B) * field;
}

int sdcb200_heat_cg_solve_ho(int ndim, int n, int bc, int order, const double* centre_host, const double* lo_host,
                             const double* hi_host, int B, const double* factor_host, const double* const* rhs,
                             double* const* x, double rtol, int maxiter, void* work, size_t work_bytes, int* iters_dev,
                             void* stream) {
    SDC_REQUIRE(ndim >= 1 && ndim <= 3, "ndim must be 1, 2 or 3");
    SDC_REQUIRE(B >= 1 && B <= SDCB200_MAX_NODES, "B out of range");
    SDC_REQUIRE(n > order, "grid too small for the stencil");
    SDC_REQUIRE(work != nullptr && work_bytes >= sdcb200_cg_ho_workspace_bytes(ndim, n, B), "workspace too small");
    SDC_REQUIRE((reinterpret_cast<size_t>(work) & 255u) == 0, "workspace must be 256-byte aligned");
    static thread_local HoCgArgs a;
    memset(&a, 0, sizeof(a));
    a.g = make_geom(ndim, n, bc);
    if (int rc = fill_op(a.op, order, bc, centre_host, lo_host, hi_host)) return rc;
    a.B = B;
    a.rtol = rtol;
    a.maxiter = maxiter;
    a.iters_out = iters_dev;
    const size_t guard = (size_t)sdcb200_guard(ndim, n), vol = (size_t)sdcb200_volume(ndim, n);
    const size_t field = align_up((guard + vol) * sizeof(double), 256);
    char* base = static_cast<char*>(work);
    a.partials = reinterpret_cast<double*>(base);
    const size_t bar_off = align_up(kPartialSlots * SDCB200_MAX_NODES * kMaxGrid * sizeof(double), 256);
    a.bar = reinterpret_cast<unsigned*>(base + bar_off);
    for (int b = 0; b < B; ++b) {
        SDC_REQUIRE(ok8(rhs[b]) && ok8(x[b]), "rhs / x missing or misaligned");
        Sys& S = a.s[b];
        S.b = rhs[b];
        S.x = x[b];
        char* f = base + bar_off + 256 + (size_t)(3 * b) * field + guard * sizeof(double);
        S.r = reinterpret_cast<double*>(f);
        S.p = reinterpret_cast<double*>(f + field);
        S.q = reinterpret_cast<double*>(f + 2 * field);
        a.factor[b] = factor_host[b];
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SDC_CUDA_OK(cudaMemsetAsync(a.bar, 0, 256, s));
    void* kernel = nullptr;
    SDC_DISPATCH_H(a.op.h, kernel = (void*)ho_cg_kernel<HW>);
    int per_sm = 0;
    SDC_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0));
    SDC_REQUIRE(per_sm >= 1, "solver kernel does not fit on an SM");
    if (per_sm > 4) per_sm = 4;
    int grid = per_sm * sm_count();
    if (grid > kMaxGrid) grid = kMaxGrid;
    void* params[] = {&a};
    SDC_CUDA_OK(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kThreads), params, 0, s));
    return 0;
}

}  // extern "C"
