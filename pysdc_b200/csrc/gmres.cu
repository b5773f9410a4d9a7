// Restarted GMRES for (I - factor A) x = rhs with A any of the finite-difference operators of fdop.cuh - the
// `solver_type='GMRES'` branch of GenericNDimFinDiff.solve_system (pySDC/implementations/problem_classes/
// generic_ND_FD.py:241-250), which is what the reference's advection problems run on (AdvectionEquation_ND_FD.py,
// tutorial/step_8/C_iteration_estimator.py:106-107) - and the general eval_f  f = A u  for those operators.
//
// The iteration is scipy's (scipy 1.18 sparse/linalg/_isolve/iterative.py::gmres, no preconditioner, atol = 0,
// callback_type='legacy'): restart length min(20, n); modified Gram-Schmidt with one dot product after the other; Givens
// rotations; inner stop on the rotated residual `presid <= ptol` with scipy's adaptive `ptol`; outer stop on the true
// residual `||b - M x|| <= rtol ||b||`; `maxiter` counts INNER iterations and every inner iteration is one call of the
// reference's work counter.  The whole solve is one persistent cooperative launch: the Krylov basis lives in the
// workspace, every dot product is a fixed-order grid reduction, so all CTAs hold bit-identical scalars, run the small
// Hessenberg / Givens / triangular-solve arithmetic redundantly (thread 0 of every CTA) and take identical branches; the
// host never synchronises inside a solve.  Passes are fused where the recurrence allows it: the Gram-Schmidt update with
// the previous coefficient rides on the pass that forms the next dot product, the last update on the pass that forms
// ||w||.
#include "fdop.cuh"

namespace sdcb200 {
namespace {

constexpr int kGmresMaxRestart = 20;  // scipy's default, the only value the reference uses

struct GmresArgs {
    Geom g;
    HoOp op;
    double factor;
    const double* b;
    double* x;
    double* V;            // (restart + 1) fields of `field` doubles each, V[k] = V + k * field (+ guard inside)
    long long field;
    int restart;
    double rtol;
    int maxiter;
    double* partials;     // [kPartialSlots][MAX_NODES][gridDim.x]; slots alternate between consecutive reductions
    unsigned* bar;
    int* iters_out;
};

struct GmresShared {
    double scratch[33];
    double h[kGmresMaxRestart][kGmresMaxRestart + 1];
    double giv[kGmresMaxRestart][2];
    double S[kGmresMaxRestart + 1];
    double y[kGmresMaxRestart];
    double val;      // the scalar of the current step, broadcast from thread 0
    int flag;
};

struct FdEvalArgs {
    Geom g;
    HoOp op;
    int B;
    const double* u[SDCB200_MAX_NODES + 1];
    double* f[SDCB200_MAX_NODES + 1];
};

template <int H>
__global__ void __launch_bounds__(kThreads) fd_eval_kernel(const __grid_constant__ FdEvalArgs a) {
    for (int b = 0; b < a.B; ++b) {
        const double* u = a.u[b];
        double* f = a.f[b];
        ho_points(a.g, [&](long long idx, int x, int y, int z) { f[idx] = ho_apply<H>(a.op, a.g, u, x, y, z); });
    }
}

// LAPACK dlartg (3.10+): plane rotation with c*f + s*g = r, -s*f + c*g = 0
__device__ inline void lartg(double f, double g, double& c, double& s, double& r) {
    const double safmin = 2.2250738585072014e-308, safmax = 1.0 / safmin;
    const double rtmin = sqrt(safmin), rtmax = sqrt(safmax / 2);
    const double f1 = fabs(f), g1 = fabs(g);
    if (g == 0.0) {
        c = 1.0;
        s = 0.0;
        r = f;
    } else if (f == 0.0) {
        c = 0.0;
        s = copysign(1.0, g);
        r = g1;
    } else if (f1 > rtmin && f1 < rtmax && g1 > rtmin && g1 < rtmax) {
        const double d = sqrt(f * f + g * g);
        c = f1 / d;
        r = copysign(d, f);
        s = g / r;
    } else {
        const double u = fmin(safmax, fmax(safmin, fmax(f1, g1)));
        const double fs = f / u, gs = g / u;
        const double d = sqrt(fs * fs + gs * gs);
        c = fabs(fs) / d;
        r = copysign(d, f);
        s = gs / r;
        r = r * u;
    }
}

// one grid-wide sum: per-CTA partial -> barrier -> fixed-order sum, identical in every thread of the grid.  Consecutive
// reductions use different slots (a slot may only be rewritten after a barrier every CTA entered after its last read).
__device__ __forceinline__ double gmres_reduce(const GmresArgs& a, GmresShared& sh, double v, unsigned& nred) {
    const int slot = (int)(nred++ % (unsigned)kPartialSlots);
    v = block_sum(v, sh.scratch);
    put_partial(a.partials, slot, 0, v);
    grid_barrier(a.bar);
    return grid_sum(a.partials, slot, 0, sh.scratch);
}

template <int H>
__global__ void __launch_bounds__(kThreads) fd_gmres_kernel(const __grid_constant__ GmresArgs a) {
    __shared__ GmresShared sh;
    const Geom& g = a.g;
    const int restart = a.restart;
    const double eps = 2.220446049250313e-16;
    unsigned nred = 0;
    auto Vk = [&](int k) { return a.V + (long long)k * a.field; };
    // M v at a point
    auto Mv = [&](const double* v, long long idx, int x, int y, int z) {
        return __dsub_rn(v[idx], __dmul_rn(a.factor, ho_apply<H>(a.op, g, v, x, y, z)));
    };

    // ||b||
    double acc = 0.0;
    flat_quads(g.vol, [&](long long i, bool full) {
        const Quad b = ldq(a.b, i, full);
        acc = qdot(b, b, acc);
    });
    const double bnrm2 = sqrt(gmres_reduce(a, sh, acc, nred));
    if (bnrm2 == 0.0) {  // scipy: return b
        flat_quads(g.vol, [&](long long i, bool full) { stq(a.x, i, ldq(a.b, i, full), full); });
        return;
    }
    const double atol = a.rtol * bnrm2;  // max(atol = 0, rtol * ||b||)
    double ptol_max_factor = 1.0;
    double ptol = bnrm2 * fmin(ptol_max_factor, atol / bnrm2);
    double presid = 0.0;
    int inner_iter = 0;

    // r = b - M x0 -> V[0]
    acc = 0.0;
    {
        double* v0 = Vk(0);
        ho_points(g, [&](long long idx, int x, int y, int z) {
            const double r = __dsub_rn(a.b[idx], Mv(a.x, idx, x, y, z));
            v0[idx] = r;
            acc = fma(r, r, acc);
        });
    }
    double rnorm = sqrt(gmres_reduce(a, sh, acc, nred));
    if (rnorm < atol) return;

    for (int iteration = 0; iteration < a.maxiter; ++iteration) {
        // v0 = r / ||r||
        {
            double* v0 = Vk(0);
            const double inv = 1.0 / rnorm;
            flat_quads(g.vol, [&](long long i, bool full) {
                const Quad v = ldq(v0, i, full);
                stq(v0, i, qmap([&](int e) { return __dmul_rn(qe(v, e), inv); }), full);
            });
        }
        if (threadIdx.x == 0) {
            for (int i = 0; i <= restart; ++i) sh.S[i] = 0.0;
            sh.S[0] = rnorm;
        }
        grid_barrier(a.bar);  // v0 complete before its neighbours are read; also orders the writes of S

        bool breakdown = false;
        int col = 0;
        for (col = 0; col < restart; ++col) {
            const double* vc = Vk(col);
            double* w = Vk(col + 1);
            // w = M v[col], ||w||
            acc = 0.0;
            ho_points(g, [&](long long idx, int x, int y, int z) {
                const double t = Mv(vc, idx, x, y, z);
                w[idx] = t;
                acc = fma(t, t, acc);
            });
            const double h0 = sqrt(gmres_reduce(a, sh, acc, nred));
            // modified Gram-Schmidt: h[col][k] = v[k].w ; w -= h[col][k] v[k]   (the update with coefficient k-1 rides on
            // the pass that forms dot product k)
            double hprev = 0.0;
            for (int k = 0; k <= col; ++k) {
                const double* vk = Vk(k);
                const double* vkm = k > 0 ? Vk(k - 1) : nullptr;
                acc = 0.0;
                flat_quads(g.vol, [&](long long i, bool full) {
                    Quad t = ldq(w, i, full);
                    const Quad kq = ldq(vk, i, full);
                    if (vkm != nullptr) {
                        const Quad m = ldq(vkm, i, full);
                        t = qmap([&](int e) { return __dsub_rn(qe(t, e), __dmul_rn(hprev, qe(m, e))); });
                        stq(w, i, t, full);
                    }
                    acc = qdot(kq, t, acc);
                });
                hprev = gmres_reduce(a, sh, acc, nred);
                if (threadIdx.x == 0) sh.h[col][k] = hprev;
            }
            acc = 0.0;
            flat_quads(g.vol, [&](long long i, bool full) {
                const Quad wq = ldq(w, i, full), c = ldq(vc, i, full);
                const Quad t = qmap([&](int e) { return __dsub_rn(qe(wq, e), __dmul_rn(hprev, qe(c, e))); });
                stq(w, i, t, full);
                acc = qdot(t, t, acc);
            });
            const double h1 = sqrt(gmres_reduce(a, sh, acc, nred));
            if (h1 <= eps * h0) {
                breakdown = true;  // exact solution indicator
            } else {
                const double inv = 1.0 / h1;
                flat_quads(g.vol, [&](long long i, bool full) {
                    const Quad wq = ldq(w, i, full);
                    stq(w, i, qmap([&](int e) { return __dmul_rn(qe(wq, e), inv); }), full);
                });
            }
            if (threadIdx.x == 0) {
                sh.h[col][col + 1] = breakdown ? 0.0 : h1;
                for (int k = 0; k < col; ++k) {  // past rotations on the new column
                    const double c = sh.giv[k][0], s = sh.giv[k][1];
                    const double n0 = sh.h[col][k], n1 = sh.h[col][k + 1];
                    sh.h[col][k] = c * n0 + s * n1;
                    sh.h[col][k + 1] = -s * n0 + c * n1;
                }
                double c, s, mag;
                lartg(sh.h[col][col], sh.h[col][col + 1], c, s, mag);
                sh.giv[col][0] = c;
                sh.giv[col][1] = s;
                sh.h[col][col] = mag;
                sh.h[col][col + 1] = 0.0;
                const double tmp = -s * sh.S[col];
                sh.S[col] = c * sh.S[col];
                sh.S[col + 1] = tmp;
                sh.val = fabs(tmp);
            }
            grid_barrier(a.bar);  // v[col+1] complete for the next matvec / the update of x; publishes sh.val
            presid = sh.val;
            ++inner_iter;
            if (inner_iter == a.maxiter) break;  // legacy: maxiter counts inner iterations
            if (presid <= ptol || breakdown) break;
        }
        if (col == restart) col = restart - 1;  // the loop ran to its end

        // y = triangular solve of h[:col+1, :col+1]^T y = S[:col+1], tolerating singular pivots like scipy
        if (threadIdx.x == 0) {
            if (sh.h[col][col] == 0.0) sh.S[col] = 0.0;
            for (int i = 0; i <= col; ++i) sh.y[i] = sh.S[i];
            for (int k = col; k > 0; --k) {
                if (sh.y[k] != 0.0) {
                    sh.y[k] /= sh.h[k][k];
                    const double t = sh.y[k];
                    for (int i = 0; i < k; ++i) sh.y[i] -= t * sh.h[k][i];
                }
            }
            if (sh.y[0] != 0.0) sh.y[0] /= sh.h[0][0];
        }
        __syncthreads();
        // x += y @ v[:col+1]
        flat_quads(g.vol, [&](long long i, bool full) {
            const Quad xq = ldq(a.x, i, full);
            Quad t;
            t.a = t.b = make_double2(0.0, 0.0);
            for (int k = 0; k <= col; ++k) {
                const Quad v = ldq(Vk(k), i, full);
                const double yk = sh.y[k];
                t = qmap([&](int e) { return fma(yk, qe(v, e), qe(t, e)); });
            }
            stq(a.x, i, qmap([&](int e) { return __dadd_rn(qe(xq, e), qe(t, e)); }), full);
        });
        grid_barrier(a.bar);
        // r = b - M x -> V[0]
        acc = 0.0;
        {
            double* v0 = Vk(0);
            ho_points(g, [&](long long idx, int x, int y, int z) {
                const double r = __dsub_rn(a.b[idx], Mv(a.x, idx, x, y, z));
                v0[idx] = r;
                acc = fma(r, r, acc);
            });
        }
        rnorm = sqrt(gmres_reduce(a, sh, acc, nred));
        if (inner_iter == a.maxiter) break;  // legacy exit
        if (rnorm <= atol) break;
        if (breakdown) break;
        if (presid <= ptol) ptol_max_factor = fmax(eps, 0.25 * ptol_max_factor);
        else ptol_max_factor = fmin(1.0, 1.5 * ptol_max_factor);
        ptol = presid * fmin(ptol_max_factor, atol / rnorm);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.iters_out != nullptr) a.iters_out[0] += inner_iter;
}

size_t gmres_layout(int ndim, int n, int restart, size_t* field_bytes, size_t* bar_off, size_t* v_off) {
    const size_t guard = (size_t)sdcb200_guard(ndim, n), vol = (size_t)sdcb200_volume(ndim, n);
    const size_t field = align_up((guard + vol) * sizeof(double), 256);
    const size_t part = align_up((size_t)kPartialSlots * SDCB200_MAX_NODES * kMaxGrid * sizeof(double), 256);
    if (field_bytes) *field_bytes = field;
    if (bar_off) *bar_off = part;
    if (v_off) *v_off = part + 256;
    return part + 256 + (size_t)(restart + 1) * field;
}

}  // namespace
}  // namespace sdcb200

using namespace sdcb200;

extern "C" {

int sdcb200_fd_eval_f(int ndim, int n, int bc, int h, const double* coef_host, const double* lo_host,
                      const double* hi_host, int B, const double* const* u, double* const* f, void* stream) {
    SDC_REQUIRE(ndim >= 1 && ndim <= 3, "ndim must be 1, 2 or 3");
    SDC_REQUIRE(B >= 1 && B <= SDCB200_MAX_NODES + 1, "B out of range");
    SDC_REQUIRE(n > 2 * h, "grid too small for the stencil");
    static thread_local FdEvalArgs a;
    memset(&a, 0, sizeof(a));
    a.g = make_geom(ndim, n, bc);
    if (int rc = fill_op_general(a.op, h, bc, coef_host, lo_host, hi_host)) return rc;
    a.B = B;
    for (int b = 0; b < B; ++b) {
        SDC_REQUIRE(ok8(u[b]) && ok8(f[b]), "u / f missing or misaligned");
        a.u[b] = u[b];
        a.f[b] = f[b];
    }
    SDC_DISPATCH_H(a.op.h, (fd_eval_kernel<HW><<<sm_count() * 8, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a)));
    SDC_CUDA_OK(cudaGetLastError());
    return 0;
}

size_t sdcb200_fd_gmres_workspace_bytes(int ndim, int n, int restart) {
    if (restart < 1 || restart > kGmresMaxRestart) return 0;
    return gmres_layout(ndim, n, restart, nullptr, nullptr, nullptr);
}

int sdcb200_fd_gmres_solve(int ndim, int n, int bc, int h, const double* coef_host, const double* lo_host,
                           const double* hi_host, double factor, const double* rhs, double* x, double rtol, int maxiter,
                           int restart, void* work, size_t work_bytes, int* iters_dev, void* stream) {
    SDC_REQUIRE(ndim >= 1 && ndim <= 3, "ndim must be 1, 2 or 3");
    SDC_REQUIRE(n > 2 * h, "grid too small for the stencil");
    SDC_REQUIRE(restart >= 1 && restart <= kGmresMaxRestart, "restart must be 1 .. 20");
    SDC_REQUIRE(maxiter >= 1, "maxiter must be positive");
    size_t field = 0, bar_off = 0, v_off = 0;
    const size_t total = gmres_layout(ndim, n, restart, &field, &bar_off, &v_off);
    SDC_REQUIRE(work != nullptr && work_bytes >= total, "workspace too small (see sdcb200_fd_gmres_workspace_bytes)");
    SDC_REQUIRE((reinterpret_cast<size_t>(work) & 255u) == 0, "workspace must be 256-byte aligned");
    SDC_REQUIRE(ok8(rhs) && ok8(x), "rhs / x missing or misaligned");
    static thread_local GmresArgs a;
    memset(&a, 0, sizeof(a));
    a.g = make_geom(ndim, n, bc);
    if (int rc = fill_op_general(a.op, h, bc, coef_host, lo_host, hi_host)) return rc;
    a.factor = factor;
    a.b = rhs;
    a.x = x;
    a.restart = restart;
    a.rtol = rtol;
    a.maxiter = maxiter;
    a.iters_out = iters_dev;
    char* base = static_cast<char*>(work);
    a.partials = reinterpret_cast<double*>(base);
    a.bar = reinterpret_cast<unsigned*>(base + bar_off);
    a.field = (long long)(field / sizeof(double));
    a.V = reinterpret_cast<double*>(base + v_off) + sdcb200_guard(ndim, n);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SDC_CUDA_OK(cudaMemsetAsync(a.bar, 0, 256, s));
    void* kernel = nullptr;
    SDC_DISPATCH_H(a.op.h, kernel = (void*)fd_gmres_kernel<HW>);
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
