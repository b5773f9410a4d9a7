// Point-wise Newton solve of the Allen-Cahn reaction part,  u - factor * (1/eps^2) u (1 - u^nu) = rhs  - the second
// implicit solve of the multi-implicit splitting (allencahn_multiimplicit.solve_system_2,
// pySDC/implementations/problem_classes/AllenCahn_2D_FD.py:594-651) - for B node systems in one persistent launch.
//
// The reference runs a GLOBAL Newton loop: all grid points are updated until the max-norm of g over the grid drops
// below newton_tol (or newton_maxiter updates were made), and it solves the DIAGONAL Jacobian system with scipy's CG to
// lin_tol.  Here the loop has the same global semantics (every point takes the same number of updates, decided by
// the grid-wide max-norm) and the diagonal system is solved exactly, z = g / dg: the limit the reference's CG converges to
// (its iterate differs from it by lin_tol * ||g||, which the next Newton step - or the stopping test on g itself -
// absorbs).  One pass per Newton update: the pass that applies update k also evaluates g at the new point and its
// max-norm, so an iteration reads u and rhs once, writes u once, and costs one grid barrier.
#include "fdop.cuh"

namespace sdcb200 {
namespace {

struct ReactSys {
    const double* rhs;
    double* u;
    double factor;
};
struct ReactArgs {
    long long count2;  // double2 elements per field
    int B;
    ReactSys s[SDCB200_MAX_NODES];
    double inv_eps2;
    int nu_exp;
    double tol;
    int maxiter;
    double* partials;  // [2][MAX_NODES][gridDim.x], the two slots alternate between consecutive reductions
    unsigned* bar;
    int* counters_out;  // [0] += Newton updates summed over the systems
};

__device__ __forceinline__ double upow(double u, int k) {  // numpy's u**2 is a multiplication; general k likewise
    double r = u;
    for (int i = 1; i < k; ++i) r = __dmul_rn(r, u);
    return r;
}
// g = u - factor * (((1/eps^2) * u) * (1 - u**nu)) - rhs, in the reference's order of operations (:620)
__device__ __forceinline__ double react_g(double u, double rhs, double factor, double inv_eps2, int nu) {
    const double react = __dmul_rn(__dmul_rn(inv_eps2, u), __dsub_rn(1.0, upow(u, nu)));
    return __dsub_rn(__dsub_rn(u, __dmul_rn(factor, react)), rhs);
}
// dg = 1 - factor * ((1/eps^2) * (1 - (nu + 1) * u**nu))  (:629)
__device__ __forceinline__ double react_dg(double u, double factor, double inv_eps2, int nu) {
    const double j = __dmul_rn(inv_eps2, __dsub_rn(1.0, __dmul_rn((double)(nu + 1), upow(u, nu))));
    return __dsub_rn(1.0, __dmul_rn(factor, j));
}
__device__ __forceinline__ double absmax_nan(double m, double g) {
    if (g != g) return INFINITY;  // NaN: never "converged"
    return fmax(m, fabs(g));
}

__global__ void __launch_bounds__(kThreads) reaction_newton_kernel(const __grid_constant__ ReactArgs a) {
    __shared__ double scratch[33];
    __shared__ unsigned s_active;
    __shared__ int s_iters[SDCB200_MAX_NODES];
    if (threadIdx.x == 0) {
        s_active = (1u << a.B) - 1u;
        for (int b = 0; b < a.B; ++b) s_iters[b] = 0;
    }
    __syncthreads();
    int slot = 0;
    bool first = true;
    while (true) {
        const unsigned act = s_active;
        for (int b = 0; b < a.B; ++b) {
            if (!(act >> b & 1u)) continue;
            const ReactSys S = a.s[b];
            double gmax = 0.0;
            flat_quads(2 * a.count2, [&](long long i, bool full) {  // four points per thread and step, loads first
                Quad u = ldq(S.u, i, full);
                const Quad rhs = ldq(S.rhs, i, full);
                if (!first) {  // Newton update k, then g at the new point
                    u = qmap([&](int e) {
                        const double v = qe(u, e), r = qe(rhs, e);
                        return __dsub_rn(v, __ddiv_rn(react_g(v, r, S.factor, a.inv_eps2, a.nu_exp),
                                                      react_dg(v, S.factor, a.inv_eps2, a.nu_exp)));
                    });
                    stq(S.u, i, u, full);
                }
                for (int e = 0; e < (full ? 4 : 2); ++e)
                    gmax = absmax_nan(gmax, react_g(qe(u, e), qe(rhs, e), S.factor, a.inv_eps2, a.nu_exp));
            });
            gmax = block_max(gmax, scratch);
            put_partial(a.partials, slot, b, gmax);
        }
        grid_barrier(a.bar);
        for (int b = 0; b < a.B; ++b) {
            if (!(act >> b & 1u)) continue;
            const double res = grid_max(a.partials, slot, b, scratch);
            if (threadIdx.x == 0) {
                if (!first) ++s_iters[b];
                if (res < a.tol || s_iters[b] >= a.maxiter) s_active &= ~(1u << b);
            }
        }
        __syncthreads();
        if (s_active == 0) break;
        first = false;
        slot ^= 1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.counters_out != nullptr) {
        int n = 0;
        for (int b = 0; b < a.B; ++b) n += s_iters[b];
        a.counters_out[0] += n;
    }
}

constexpr int kReactMaxGrid = kMaxGrid;

}  // namespace
}  // namespace sdcb200

using namespace sdcb200;

extern "C" {

size_t sdcb200_reaction_workspace_bytes(void) { return 256 + 2 * SDCB200_MAX_NODES * kReactMaxGrid * sizeof(double); }

int sdcb200_allencahn_reaction_newton(long long count, int B, const double* factor_host, double inv_eps2, int nu_exp,
                                      const double* const* rhs, double* const* u, double newton_tol, int newton_maxiter,
                                      void* work, size_t work_bytes, int* counters_dev, void* stream) {
    SDC_REQUIRE(count >= 0 && !(count & 1), "count must be even (fields are moved as double2)");
    SDC_REQUIRE(nu_exp >= 1, "nu must be a positive integer");
    SDC_REQUIRE(B >= 1 && B <= SDCB200_MAX_NODES, "B out of range");
    SDC_REQUIRE(work != nullptr && work_bytes >= sdcb200_reaction_workspace_bytes(), "workspace too small");
    SDC_REQUIRE((reinterpret_cast<size_t>(work) & 255u) == 0, "workspace must be 256-byte aligned");
    ReactArgs a;
    memset(&a, 0, sizeof(a));
    a.count2 = count / 2;
    a.B = B;
    for (int b = 0; b < B; ++b) {
        SDC_REQUIRE(rhs[b] && u[b] && !(reinterpret_cast<size_t>(rhs[b]) & 15u) && !(reinterpret_cast<size_t>(u[b]) & 15u),
                    "rhs / u missing or misaligned");
        a.s[b].rhs = rhs[b];
        a.s[b].u = u[b];
        a.s[b].factor = factor_host[b];
    }
    a.inv_eps2 = inv_eps2;
    a.nu_exp = nu_exp;
    a.tol = newton_tol;
    a.maxiter = newton_maxiter;
    a.bar = reinterpret_cast<unsigned*>(work);
    a.partials = reinterpret_cast<double*>(static_cast<char*>(work) + 256);
    a.counters_out = counters_dev;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SDC_CUDA_OK(cudaMemsetAsync(a.bar, 0, 256, s));
    if (count == 0) return 0;
    static int grid = 0;
    if (grid == 0) {
        int per_sm = 0;
        SDC_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reaction_newton_kernel, kThreads, 0));
        if (per_sm < 1) return fail("sdcb200_allencahn_reaction_newton", "kernel does not fit on an SM");
        if (per_sm > 4) per_sm = 4;
        grid = per_sm * sm_count();
        if (grid > kReactMaxGrid) grid = kReactMaxGrid;
    }
    void* params[] = {&a};
    SDC_CUDA_OK(cudaLaunchCooperativeKernel((void*)reaction_newton_kernel, dim3(grid), dim3(kThreads), params, 0, s));
    return 0;
}

}  // extern "C"
