// Direct solver for 1-D grids: (I - factor*A) is a constant-coefficient (cyclic) tridiagonal matrix [e, d, e].
// One CTA per system: the right-hand side is staged into shared memory with coalesced loads, one thread runs the
// Thomas recurrence there (n ~ 1e3: a few tens of microseconds, latency bound by construction), the CTA writes the
// solution back.  Periodic grids add the Sherman-Morrison correction for the two corner entries.
// Replaces scipy's spsolve in GenericNDimFinDiff.solve_system for solver_type='direct' (generic_ND_FD.py:239), ndim 1.
#include "common.cuh"

namespace sdcb200 {
namespace {

constexpr int kMaxDirectN = 8192;

struct DirectArgs {
    int n, periodic, B;
    const double* rhs[SDCB200_MAX_NODES];
    double* x[SDCB200_MAX_NODES];
    double d[SDCB200_MAX_NODES], e[SDCB200_MAX_NODES];
};

__global__ void __launch_bounds__(kThreads) thomas_kernel(const __grid_constant__ DirectArgs a) {
    extern __shared__ double sm[];
    const int n = a.n, b = blockIdx.x;
    double* c = sm;          // modified super-diagonal
    double* y = sm + n;      // rhs -> solution
    double* q = sm + 2 * n;  // periodic only: solution for the corner vector
    for (int i = threadIdx.x; i < n; i += kThreads) y[i] = a.rhs[b][i];
    __syncthreads();
    if (threadIdx.x == 0) {
        const double d = a.d[b], e = a.e[b];
        if (!a.periodic) {
            double denom = d;
            c[0] = e / denom;
            y[0] = y[0] / denom;
            for (int i = 1; i < n; ++i) {
                denom = d - e * c[i - 1];
                c[i] = e / denom;
                y[i] = (y[i] - e * y[i - 1]) / denom;
            }
            for (int i = n - 2; i >= 0; --i) y[i] -= c[i] * y[i + 1];
        } else {
            // M = T + u v^T,  u = (gamma, 0, .., 0, e)^T,  v = (1, 0, .., 0, e/gamma)^T,  gamma = -d
            const double gamma = -d;
            const double d0 = d - gamma, dn = d - e * e / gamma;
            for (int i = 0; i < n; ++i) q[i] = 0.0;
            q[0] = gamma;
            q[n - 1] = e;
            double denom = d0;
            c[0] = e / denom;
            y[0] /= denom;
            q[0] /= denom;
            for (int i = 1; i < n; ++i) {
                const double di = (i == n - 1) ? dn : d;
                denom = di - e * c[i - 1];
                c[i] = e / denom;
                y[i] = (y[i] - e * y[i - 1]) / denom;
                q[i] = (q[i] - e * q[i - 1]) / denom;
            }
            for (int i = n - 2; i >= 0; --i) {
                y[i] -= c[i] * y[i + 1];
                q[i] -= c[i] * q[i + 1];
            }
            const double vy = y[0] + e / gamma * y[n - 1], vq = q[0] + e / gamma * q[n - 1];
            const double fac = vy / (1.0 + vq);
            for (int i = 0; i < n; ++i) y[i] -= fac * q[i];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kThreads) a.x[b][i] = y[i];
}

}  // namespace
}  // namespace sdcb200

using namespace sdcb200;

extern "C" int sdcb200_heat_direct_solve_1d(int n, int bc, int B, const double* m_diag_host, const double* m_off_host,
                                            const double* const* rhs, double* const* x, void* stream) {
    SDC_REQUIRE(n >= 3 && n <= kMaxDirectN, "1-D direct solver supports 3 <= n <= 8192; use solver_type='CG'");
    SDC_REQUIRE(B >= 1 && B <= SDCB200_MAX_NODES, "B out of range");
    SDC_REQUIRE(bc == SDCB200_BC_PERIODIC ? !(n & 1) : (n & 1), "n parity does not match the boundary condition");
    DirectArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n;
    a.periodic = bc == SDCB200_BC_PERIODIC;
    a.B = B;
    for (int b = 0; b < B; ++b) {
        SDC_REQUIRE(rhs[b] && x[b], "rhs / x missing");
        a.rhs[b] = rhs[b];
        a.x[b] = x[b];
        a.d[b] = m_diag_host[b];
        a.e[b] = m_off_host[b];
    }
    const size_t smem = (size_t)3 * n * sizeof(double);
    SDC_CUDA_OK(cudaFuncSetAttribute(thomas_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    thomas_kernel<<<B, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(a);
    SDC_CUDA_OK(cudaGetLastError());
    return 0;
}
