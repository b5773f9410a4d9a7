// K1 (collocation integration / rhs assembly / residual max-norm) and K6 (datatype utilities).
// All kernels are single-pass, coalesced, double2-vectorised streaming kernels: HBM-bound, no tensor cores.
#include "common.cuh"

namespace sdcb200 {

namespace {

thread_local std::string g_last_error;
int g_sm_count = 0;

struct CollocArgs {
    const double* in[SDCB200_MAX_TERMS];
    double* out[SDCB200_MAX_NODES];
    const double* add[SDCB200_MAX_NODES];
    const double* u[SDCB200_MAX_NODES];
    const double* base;
    double W[SDCB200_MAX_NODES * SDCB200_MAX_TERMS];
    double* resnorm;
    long long count2;  // number of double2 elements
    int nin;
};

// Rounding: every product and every sum is rounded on its own (__dmul_rn / __dadd_rn are never contracted into an FMA)
// and the terms are accumulated in ascending k starting from 0, which is what the reference's loops
// `me[-1] += L.dt * self.coll.Qmat[m, j] * L.f[j]` (generic_implicit.py:46-47) do with numpy: scalar coefficient
// first, one rounded product, one rounded sum per term.
__device__ __forceinline__ void mul_add(double2& acc, double w, const double2 v) {
    acc.x = __dadd_rn(acc.x, __dmul_rn(w, v.x));
    acc.y = __dadd_rn(acc.y, __dmul_rn(w, v.y));
}
__device__ __forceinline__ void add2(double2& acc, const double2 v) {
    acc.x = __dadd_rn(acc.x, v.x);
    acc.y = __dadd_rn(acc.y, v.y);
}

// out[m] = sum_k W[m][k] in[k] + base + add[m]   (generic linear combination)
template <int NOUT>
__global__ void __launch_bounds__(kThreads) colloc_apply_kernel(const __grid_constant__ CollocArgs a) {
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < a.count2; i += stride) {
        double2 acc[NOUT];
#pragma unroll
        for (int m = 0; m < NOUT; ++m) acc[m] = make_double2(0.0, 0.0);
#pragma unroll 4
        for (int k = 0; k < a.nin; ++k) {
            const double2 v = ld2(a.in[k] + 2 * i);
#pragma unroll
            for (int m = 0; m < NOUT; ++m) mul_add(acc[m], a.W[m * a.nin + k], v);
        }
        if (a.base != nullptr) {
            const double2 b = ld2(a.base + 2 * i);
#pragma unroll
            for (int m = 0; m < NOUT; ++m) add2(acc[m], b);
        }
#pragma unroll
        for (int m = 0; m < NOUT; ++m) {
            if (a.add[m] != nullptr) add2(acc[m], ld2(a.add[m] + 2 * i));
            st2(a.out[m] + 2 * i, acc[m]);
        }
    }
}

// The sweep's node combinations in the reference's own order of operations (see sdc_b200.h, sdcb200_colloc_sweep).
struct SweepArgs {
    const double* in[SDCB200_MAX_TERMS];  // f[j] (NCOMP = 1) or f[j].impl, f[j].expl interleaved (NCOMP = 2)
    double* out[SDCB200_MAX_NODES];
    const double* add[SDCB200_MAX_NODES];
    const double* u[SDCB200_MAX_NODES];
    const double* base;
    double Wq[SDCB200_MAX_NODES * SDCB200_MAX_NODES];  // phase 1: quadrature coefficients dt*Q[m][j]
    double Wi[SDCB200_MAX_NODES * SDCB200_MAX_NODES];  // phase 2: implicit QDelta coefficients (sign folded in)
    double We[SDCB200_MAX_NODES * SDCB200_MAX_NODES];  // phase 2: explicit QDelta coefficients (NCOMP = 2 only)
    double dt2;                                        // phase 2, NCOMP = 2: outer factor dt
    double* resnorm;
    long long count2;
    int nj;     // input nodes
    int flags;  // SDCB200_SWEEP_*
};

template <int NCOMP>
__device__ __forceinline__ double2 node_value(const SweepArgs& a, int j, long long i) {
    if (NCOMP == 1) return ld2(a.in[j] + 2 * i);
    double2 v = ld2(a.in[2 * j] + 2 * i);
    add2(v, ld2(a.in[2 * j + 1] + 2 * i));  // f[j].impl + f[j].expl  (imex_1st_order.py:53)
    return v;
}

// SQUARE: as many input nodes as output nodes (the sweep's M x M combinations) - the trip counts are compile-time
// constants, the loops unroll completely and all M*NCOMP loads of a point are in flight together; otherwise (node
// additions with j <= m inputs, end point with one output) the generic loops run.
template <int NOUT, int NCOMP, bool SQUARE>
__device__ __forceinline__ void sweep_terms(const SweepArgs& a, long long i, double2 (&acc)[NOUT]) {
    const int nj = SQUARE ? NOUT : a.nj;
    if (SQUARE) {
        double2 v[NOUT * NCOMP];
#pragma unroll
        for (int k = 0; k < NOUT * NCOMP; ++k) v[k] = ld2(a.in[k] + 2 * i);
        if (a.flags & SDCB200_SWEEP_QUADRATURE) {
#pragma unroll
            for (int j = 0; j < NOUT; ++j) {
                double2 f = v[NCOMP * j];
                if (NCOMP == 2) add2(f, v[2 * j + 1]);  // f[j].impl + f[j].expl  (imex_1st_order.py:53)
#pragma unroll
                for (int m = 0; m < NOUT; ++m) mul_add(acc[m], a.Wq[m * NOUT + j], f);
            }
        }
        if (a.flags & SDCB200_SWEEP_QDELTA) {
#pragma unroll
            for (int j = 0; j < NOUT; ++j) {
#pragma unroll
                for (int m = 0; m < NOUT; ++m) {
                    if (NCOMP == 1) {
                        mul_add(acc[m], a.Wi[m * NOUT + j], v[j]);
                    } else {
                        double2 t;
                        t.x = __dadd_rn(__dmul_rn(a.Wi[m * NOUT + j], v[2 * j].x), __dmul_rn(a.We[m * NOUT + j], v[2 * j + 1].x));
                        t.y = __dadd_rn(__dmul_rn(a.Wi[m * NOUT + j], v[2 * j].y), __dmul_rn(a.We[m * NOUT + j], v[2 * j + 1].y));
                        mul_add(acc[m], a.dt2, t);
                    }
                }
            }
        }
        return;
    }
    if (a.flags & SDCB200_SWEEP_QUADRATURE) {
#pragma unroll 4
        for (int j = 0; j < nj; ++j) {
            const double2 v = node_value<NCOMP>(a, j, i);
#pragma unroll
            for (int m = 0; m < NOUT; ++m) mul_add(acc[m], a.Wq[m * a.nj + j], v);
        }
    }
    if (a.flags & SDCB200_SWEEP_QDELTA) {
#pragma unroll 2
        for (int j = 0; j < a.nj; ++j) {  // second read of f[j]: served by L1/L2, no extra DRAM traffic
            if (NCOMP == 1) {
                const double2 v = ld2(a.in[j] + 2 * i);
#pragma unroll
                for (int m = 0; m < NOUT; ++m) mul_add(acc[m], a.Wi[m * a.nj + j], v);  // (dt*QI[m][j]) * f[j]
            } else {
                const double2 vi = ld2(a.in[2 * j] + 2 * i), ve = ld2(a.in[2 * j + 1] + 2 * i);
#pragma unroll
                for (int m = 0; m < NOUT; ++m) {  // dt * (QI[m][j]*f[j].impl + QE[m][j]*f[j].expl), imex_1st_order.py:83,95
                    double2 t;
                    t.x = __dadd_rn(__dmul_rn(a.Wi[m * a.nj + j], vi.x), __dmul_rn(a.We[m * a.nj + j], ve.x));
                    t.y = __dadd_rn(__dmul_rn(a.Wi[m * a.nj + j], vi.y), __dmul_rn(a.We[m * a.nj + j], ve.y));
                    mul_add(acc[m], a.dt2, t);
                }
            }
        }
    }
}

template <int NOUT, int NCOMP, bool SQUARE>
__global__ void __launch_bounds__(kThreads) colloc_sweep_kernel(const __grid_constant__ SweepArgs a) {
    const long long stride = (long long)gridDim.x * kThreads;
    const bool base_first = a.flags & SDCB200_SWEEP_BASE_FIRST;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < a.count2; i += stride) {
        double2 acc[NOUT];
        double2 b = make_double2(0.0, 0.0);
        if (a.base != nullptr) b = ld2(a.base + 2 * i);
#pragma unroll
        for (int m = 0; m < NOUT; ++m) acc[m] = base_first ? b : make_double2(0.0, 0.0);
        sweep_terms<NOUT, NCOMP, SQUARE>(a, i, acc);
#pragma unroll
        for (int m = 0; m < NOUT; ++m) {
            if (!base_first && a.base != nullptr) add2(acc[m], b);
            if (a.add[m] != nullptr) add2(acc[m], ld2(a.add[m] + 2 * i));
            st2(a.out[m] + 2 * i, acc[m]);
        }
    }
}

// res[m] = sum_j (dt*Q[m][j]) f[j] + (u0 - u[m]) + tau[m];  resnorm[m] = max |res[m]|   (core/sweeper.py:186-195)
template <int NOUT, int NCOMP, bool SQUARE>
__global__ void __launch_bounds__(kThreads) colloc_residual_kernel(const __grid_constant__ SweepArgs a) {
    __shared__ double scratch[33];
    double vmax[NOUT];
    unsigned bad = 0;  // bit m: a NaN was seen in res[m] (fmax would silently drop it; numpy's max would not)
#pragma unroll
    for (int m = 0; m < NOUT; ++m) vmax[m] = 0.0;
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < a.count2; i += stride) {
        // all loads of this point first (node values, u0, then the right-hand sides inside sweep_terms): 2M+1 (3M+1 for
        // IMEX) independent 16-byte loads in flight per thread
        double2 um[NOUT];
#pragma unroll
        for (int m = 0; m < NOUT; ++m) um[m] = ld2(a.u[m] + 2 * i);
        const double2 u0 = ld2(a.base + 2 * i);
        double2 acc[NOUT];
#pragma unroll
        for (int m = 0; m < NOUT; ++m) acc[m] = make_double2(0.0, 0.0);
        sweep_terms<NOUT, NCOMP, SQUARE>(a, i, acc);
#pragma unroll
        for (int m = 0; m < NOUT; ++m) {
            acc[m].x = __dadd_rn(acc[m].x, __dadd_rn(u0.x, -um[m].x));  // res += u[0] - u[m+1]  (sweeper.py:188)
            acc[m].y = __dadd_rn(acc[m].y, __dadd_rn(u0.y, -um[m].y));
            if (a.add[m] != nullptr) add2(acc[m], ld2(a.add[m] + 2 * i));
            if (a.out[m] != nullptr) st2(a.out[m] + 2 * i, acc[m]);
            vmax[m] = fmax(vmax[m], fmax(fabs(acc[m].x), fabs(acc[m].y)));
            if (acc[m].x != acc[m].x || acc[m].y != acc[m].y) bad |= 1u << m;
        }
    }
#pragma unroll
    for (int m = 0; m < NOUT; ++m) {
        const double v = block_max(vmax[m], scratch);
        const double nanflag = block_max((bad >> m & 1u) ? 1.0 : 0.0, scratch);
        if (threadIdx.x == 0) atomic_max_nonneg(a.resnorm + m, nanflag > 0.0 ? fabs(nan("")) : v);
    }
}

__global__ void __launch_bounds__(kThreads) maxabs_kernel(const double* __restrict__ x, long long count2, double* out) {
    __shared__ double scratch[33];
    double vmax = 0.0;
    bool bad = false;
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < count2; i += stride) {
        const double2 v = ld2(x + 2 * i);
        vmax = fmax(vmax, fmax(fabs(v.x), fabs(v.y)));
        bad |= (v.x != v.x) || (v.y != v.y);
    }
    const double nanflag = block_max(bad ? 1.0 : 0.0, scratch);
    const double v = block_max(vmax, scratch);
    if (threadIdx.x == 0) atomic_max_nonneg(out, nanflag > 0.0 ? fabs(nan("")) : v);
}

__global__ void __launch_bounds__(kThreads) axpby_kernel(long long count2, double a, const double* __restrict__ x,
                                                          double b, const double* __restrict__ y, double* out) {
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < count2; i += stride) {
        double2 v = ld2(x + 2 * i);
        v.x *= a;
        v.y *= a;
        if (y != nullptr) {
            const double2 w = ld2(y + 2 * i);
            v.x = __dadd_rn(v.x, __dmul_rn(b, w.x));  // numpy rounds a*x, b*y and the sum separately
            v.y = __dadd_rn(v.y, __dmul_rn(b, w.y));
        }
        st2(out + 2 * i, v);
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<size_t>(p) & 15u) == 0; }

inline int stream_grid(long long count2) {
    const long long want = (count2 + kThreads - 1) / kThreads;
    const long long cap = (long long)sm_count() * 8;  // a whole number of waves: 8 resident CTAs per SM
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

// Grid of a grid-stride streaming kernel: exactly ONE wave of co-resident CTAs (what the kernel's registers allow per SM
// x SM count), so that no partially filled last wave trails behind (the collocation kernels hold 60-150 registers: 2-4
// CTAs per SM, not the 8 that stream_grid assumes).
template <auto kernel>
int stream_grid_of(long long count2) {
    static int per_sm = 0;
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
    }
    const long long want = (count2 + kThreads - 1) / kThreads;
    const long long cap = (long long)sm_count() * per_sm;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}
#define LAUNCH(...) __VA_ARGS__<<<stream_grid_of<__VA_ARGS__>(a.count2), kThreads, 0, s>>>(a)

}  // namespace

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(const char* where, const std::string& msg) {
    g_last_error = std::string(where) + ": " + msg;
    return 1;
}
int fail_cuda(const char* where, cudaError_t e) {
    g_last_error = std::string(where) + ": CUDA error " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return 2;
}
int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) g_sm_count = 148;
    }
    return g_sm_count;
}

}  // namespace sdcb200

using namespace sdcb200;

extern "C" {

int sdcb200_version(void) { return 100; }
const char* sdcb200_last_error(void) { return g_last_error.c_str(); }

long long sdcb200_pitch(int n) { return n + (n & 1); }
long long sdcb200_volume(int ndim, int n) {
    long long P = sdcb200_pitch(n), v = 1;
    for (int d = 0; d < ndim; ++d) v *= P;
    return v;
}
long long sdcb200_guard(int ndim, int n) {
    long long P = sdcb200_pitch(n), g = 1;
    for (int d = 1; d < ndim; ++d) g *= P;
    return (g + 15) / 16 * 16;
}

int sdcb200_maxabs(const double* x, long long count, double* out_dev, void* stream) {
    SDC_REQUIRE(count >= 0 && count % 2 == 0 && aligned16(x), "count must be even and x 16-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SDC_CUDA_OK(cudaMemsetAsync(out_dev, 0, sizeof(double), s));
    if (count == 0) return 0;
    maxabs_kernel<<<stream_grid(count / 2), kThreads, 0, s>>>(x, count / 2, out_dev);
    SDC_CUDA_OK(cudaGetLastError());
    return 0;
}

int sdcb200_axpby(long long count, double a, const double* x, double b, const double* y, double* out, void* stream) {
    SDC_REQUIRE(count >= 0 && count % 2 == 0 && aligned16(x) && aligned16(y) && aligned16(out),
                "count must be even and pointers 16-byte aligned");
    if (count == 0) return 0;
    axpby_kernel<<<stream_grid(count / 2), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(count / 2, a, x, b, y, out);
    SDC_CUDA_OK(cudaGetLastError());
    return 0;
}

int sdcb200_colloc_apply(long long count, int nout, int nin, const double* W_host, const double* const* in,
                         const double* base, const double* const* add, double* const* out, void* stream) {
    SDC_REQUIRE(nout >= 1 && nout <= SDCB200_MAX_NODES, "nout out of range");
    SDC_REQUIRE(nin >= 0 && nin <= SDCB200_MAX_TERMS, "nin out of range");
    SDC_REQUIRE(count >= 0 && count % 2 == 0, "count must be even");
    CollocArgs a;
    memset(&a, 0, sizeof(a));
    for (int k = 0; k < nin; ++k) {
        SDC_REQUIRE(in[k] != nullptr && aligned16(in[k]), "input field missing or misaligned");
        a.in[k] = in[k];
    }
    for (int m = 0; m < nout; ++m) {
        SDC_REQUIRE(out[m] != nullptr && aligned16(out[m]), "output field missing or misaligned");
        a.out[m] = out[m];
        a.add[m] = add ? add[m] : nullptr;
        SDC_REQUIRE(aligned16(a.add[m]), "add field misaligned");
        for (int k = 0; k < nin; ++k) a.W[m * nin + k] = W_host[m * nin + k];
    }
    SDC_REQUIRE(aligned16(base), "base field misaligned");
    a.base = base;
    a.nin = nin;
    a.count2 = count / 2;
    if (count == 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (nout) {
#define CASE(N) case N: LAUNCH(colloc_apply_kernel<N>); break;
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    }
    SDC_CUDA_OK(cudaGetLastError());
    return 0;
}

namespace {
int fill_sweep_args(SweepArgs& a, long long count, int nout, int nj, int ncomp, int flags, const double* Wq,
                    const double* Wi, const double* We, double dt2, const double* const* in, const double* base,
                    const double* const* add) {
    SDC_REQUIRE(nout >= 1 && nout <= SDCB200_MAX_NODES, "number of output nodes out of range");
    SDC_REQUIRE(nj >= 0 && nj <= SDCB200_MAX_NODES, "number of input nodes out of range");
    SDC_REQUIRE(ncomp == 1 || ncomp == 2, "ncomp must be 1 (mesh) or 2 (imex_mesh)");
    SDC_REQUIRE(count >= 0 && count % 2 == 0, "count must be even");
    SDC_REQUIRE(!(flags & SDCB200_SWEEP_QUADRATURE) || Wq != nullptr, "quadrature coefficients missing");
    SDC_REQUIRE(!(flags & SDCB200_SWEEP_QDELTA) || (Wi != nullptr && (ncomp == 1 || We != nullptr)),
                "QDelta coefficients missing");
    memset(&a, 0, sizeof(a));
    for (int k = 0; k < nj * ncomp; ++k) {
        SDC_REQUIRE(in[k] != nullptr && aligned16(in[k]), "input field missing or misaligned");
        a.in[k] = in[k];
    }
    for (int m = 0; m < nout; ++m) {
        a.add[m] = add ? add[m] : nullptr;
        SDC_REQUIRE(aligned16(a.add[m]), "tau field misaligned");
        for (int j = 0; j < nj; ++j) {
            a.Wq[m * nj + j] = (flags & SDCB200_SWEEP_QUADRATURE) ? Wq[m * nj + j] : 0.0;
            a.Wi[m * nj + j] = (flags & SDCB200_SWEEP_QDELTA) ? Wi[m * nj + j] : 0.0;
            a.We[m * nj + j] = ((flags & SDCB200_SWEEP_QDELTA) && ncomp == 2) ? We[m * nj + j] : 0.0;
        }
    }
    SDC_REQUIRE(aligned16(base), "base field misaligned");
    a.base = base;
    a.dt2 = dt2;
    a.nj = nj;
    a.flags = flags;
    a.count2 = count / 2;
    return 0;
}
}  // namespace

int sdcb200_colloc_sweep(long long count, int nout, int nj, int ncomp, int flags, const double* Wq_host,
                         const double* Wi_host, const double* We_host, double dt2, const double* const* in,
                         const double* base, const double* const* add, double* const* out, void* stream) {
    SweepArgs a;
    if (int rc = fill_sweep_args(a, count, nout, nj, ncomp, flags, Wq_host, Wi_host, We_host, dt2, in, base, add)) return rc;
    SDC_REQUIRE(!(flags & SDCB200_SWEEP_BASE_FIRST) || base != nullptr, "BASE_FIRST needs a base field");
    for (int m = 0; m < nout; ++m) {
        SDC_REQUIRE(out[m] != nullptr && aligned16(out[m]), "output field missing or misaligned");
        a.out[m] = out[m];
    }
    if (count == 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool square = nj == nout && nout <= 4;  // (larger M: the register-resident variant would spill)
    switch (nout * 4 + (ncomp - 1) * 2 + (square ? 1 : 0)) {
#define CASE(N) \
    case 4 * N: LAUNCH(colloc_sweep_kernel<N, 1, false>); break; \
    case 4 * N + 1: LAUNCH(colloc_sweep_kernel<N, 1, (N <= 4)>); break; \
    case 4 * N + 2: LAUNCH(colloc_sweep_kernel<N, 2, false>); break; \
    case 4 * N + 3: LAUNCH(colloc_sweep_kernel<N, 2, (N <= 4)>); break;
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    }
    SDC_CUDA_OK(cudaGetLastError());
    return 0;
}

int sdcb200_colloc_residual(long long count, int M, int nj, int ncomp, const double* Wq_host, const double* const* in,
                            const double* u0, const double* const* u, const double* const* tau,
                            double* const* res_out, double* resnorm_dev, void* stream) {
    SDC_REQUIRE(u0 != nullptr && aligned16(u0), "u0 missing or misaligned");
    SweepArgs a;
    if (int rc = fill_sweep_args(a, count, M, nj, ncomp, SDCB200_SWEEP_QUADRATURE, Wq_host, nullptr, nullptr, 0.0, in, u0,
                                 tau)) return rc;
    for (int m = 0; m < M; ++m) {
        SDC_REQUIRE(u[m] != nullptr && aligned16(u[m]), "node value missing or misaligned");
        a.u[m] = u[m];
        a.out[m] = res_out ? res_out[m] : nullptr;
        SDC_REQUIRE(aligned16(a.out[m]), "residual field misaligned");
    }
    a.resnorm = resnorm_dev;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SDC_CUDA_OK(cudaMemsetAsync(resnorm_dev, 0, sizeof(double) * M, s));
    if (count == 0) return 0;
    // (the register-resident variant pays off for the sweep kernel only: with the node values and the norms on top it
    // needs 120 registers and was measured at half the bandwidth of the generic loops)
    switch (M * 2 + (ncomp - 1)) {
#define CASE(N) \
    case 2 * N: LAUNCH(colloc_residual_kernel<N, 1, false>); break; \
    case 2 * N + 1: LAUNCH(colloc_residual_kernel<N, 2, false>); break;
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    }
    SDC_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
