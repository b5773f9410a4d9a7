// K2: right-hand side evaluation (eval_f) for the FD heat equation and the fully implicit Allen-Cahn problem.
#include "pipe_host.cuh"

namespace sdcb200 {
namespace {

struct EvalArgs {
    Geom g;
    int B;
    const double* u[SDCB200_MAX_NODES + 1];
    double* f[SDCB200_MAX_NODES + 1];
    double* fexpl[SDCB200_MAX_NODES + 1];
    double gt[SDCB200_MAX_NODES + 1];
    const double* profile;
    double a_diag, a_off, inv_eps2;
    int nu_exp;
    int split;  // Allen-Cahn semi-implicit: reaction term into fexpl
};

__device__ __forceinline__ double ipow(double u, int k) {
    double r = u;
    for (int i = 1; i < k; ++i) r = __dmul_rn(r, u);
    return r;
}

// MODE 0: f = A u;  MODE 1: f = A u, fexpl = profile * g(t);  MODE 2: f = A u + inv_eps2 * u * (1 - u^nu)
template <int NDIM, bool PER, int MODE>
__global__ void __launch_bounds__(kThreads) eval_f_kernel(const __grid_constant__ EvalArgs a) {
    const Units U = make_units(a.g);
    const int total = a.B * U.per_field;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int b = w / U.per_field, unit = w - b * U.per_field;
        const double* u = a.u[b];
        double* f = a.f[b];
        stencil_unit<NDIM, PER>(a.g, U, u, unit, [&](long long idx, double2 c, double2 nb, bool v0, bool v1) {
            double2 out;
            out.x = fma(a.a_off, nb.x, a.a_diag * c.x);
            out.y = fma(a.a_off, nb.y, a.a_diag * c.y);
            if constexpr (MODE == 2) {
                // (A u) + ((1/eps^2) * u) * (1 - u**nu), evaluated left to right like AllenCahn_2D_FD.py:225
                out.x = __dadd_rn(out.x, __dmul_rn(__dmul_rn(a.inv_eps2, c.x), __dsub_rn(1.0, ipow(c.x, a.nu_exp))));
                out.y = __dadd_rn(out.y, __dmul_rn(__dmul_rn(a.inv_eps2, c.y), __dsub_rn(1.0, ipow(c.y, a.nu_exp))));
            }
            if (!v0) out.x = 0.0;
            if (!v1) out.y = 0.0;
            st2(f + idx, out);
            if constexpr (MODE == 1) {
                const double2 s = ld2(a.profile + idx);  // zero on the walls
                st2(a.fexpl[b] + idx, make_double2(__dmul_rn(s.x, a.gt[b]), __dmul_rn(s.y, a.gt[b])));
            }
        });
    }
}

template <int NDIM, bool PER, int MODE>
int launch_eval(const EvalArgs& a, cudaStream_t s) {
    const Units U = make_units(a.g);
    long long total = (long long)a.B * U.per_field;
    const long long cap = (long long)sm_count() * 8;
    const int grid = (int)(total < cap ? total : cap);
    eval_f_kernel<NDIM, PER, MODE><<<grid, kThreads, 0, s>>>(a);
    SDC_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- 2-D / 3-D grids: the same evaluation as a TMA-pipelined pass (phase F of cg_pipe.cuh): a producer warp streams
// the planes of u (with halo; periodic grids: wrap boxes) and of the forcing profile through shared memory, 8 consumer
// warps evaluate the stencil from there - enough bytes in flight to keep HBM busy where the register-marching loader
// tops out at ~70 % of the copy bandwidth.
struct EvalPipeArgs {
    Geom g;
    int B;
    EvalPhase ev;
    PipeMaps maps;
};

// SLIM: no tile-only box in the stages (no forcing profile) -> 10 (periodic: 8) stages instead of 5
template <int NDIM, bool PER, bool SLIM>
__global__ void __maxnreg__(96) eval_pipe_kernel(const __grid_constant__ EvalPipeArgs a) {
    using Smem = PipeSmemT<PER, false, SLIM ? 2 : 1>;
    extern __shared__ __align__(128) unsigned char pipe_smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(pipe_smem_raw);
    __shared__ CgShared sh;  // not used by this phase (no reductions)
    pipe_ctl_init(sm.ctl, Smem::Cfg::kStages);
    if (threadIdx.x == 0) {
        sm.ctl.nact = a.B;
        for (int b = 0; b < a.B; ++b) sm.ctl.act_list[b] = b;
    }
    __syncthreads();
    const PUnits PU = make_punits(a.g, a.B, (int)gridDim.x);
    unsigned kstep = 0;
    PassArgs pa;
    pa.ev = &a.ev;
    pipe_pass<NDIM, PER, false, kPhaseF>(a.g, PU, nullptr, a.maps, sh, sm, nullptr, kstep, pa);
}

template <int NDIM, bool PER, bool SLIM>
int launch_eval_pipe_t(const EvalArgs& e, cudaStream_t s) {
    static thread_local EvalPipeArgs a;
    const size_t smem = sizeof(PipeSmemT<PER, false, SLIM ? 2 : 1>);
    int ctas = 0;
    if (int rc = pipe_grid<eval_pipe_kernel<NDIM, PER, SLIM>>(smem, &ctas)) return rc;
    for (int b0 = 0; b0 < e.B; b0 += SDCB200_MAX_NODES) {  // at most MAX_NODES fields per launch (M + 1 fields in predict)
        const int nb = e.B - b0 < SDCB200_MAX_NODES ? e.B - b0 : SDCB200_MAX_NODES;
        a.g = e.g;
        a.B = nb;
        a.ev.a_diag = e.a_diag;
        a.ev.a_off = e.a_off;
        a.ev.inv_eps2 = e.inv_eps2;
        a.ev.nu_exp = e.nu_exp;
        a.ev.split = e.split;
        for (int b = 0; b < nb; ++b) {
            a.ev.e[b].f = e.f[b0 + b];
            a.ev.e[b].f_expl = (e.profile != nullptr || e.split) ? e.fexpl[b0 + b] : nullptr;
            a.ev.e[b].gt = e.gt[b0 + b];
            if (int rc = encode_halo_maps(a.maps.m[b], kMapPHalo, e.g, e.u[b0 + b])) return rc;
            if (e.profile != nullptr)
                if (int rc = encode_field_map(a.maps.m[b] + kMapDCentre, e.g, e.profile, kPX, kPY)) return rc;
        }
        const PUnits PU = make_punits(e.g, nb, ctas);
        const long long units = (long long)nb * PU.per_field;
        const int grid = (int)(units < ctas ? units : ctas);
        eval_pipe_kernel<NDIM, PER, SLIM><<<grid, kPipeThreads, smem, s>>>(a);
        SDC_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

template <int NDIM, bool PER>
int launch_eval_pipe(const EvalArgs& e, cudaStream_t s) {
    // The SLIM variant (no tile-only slot, 10 stages instead of 5) was measured SLOWER at 511^3 (0.74 vs 0.83-0.86 of the
    // measured HBM peak, profiles/r02/bench_driverlike_r2l.json vs bench_c3_r2k.json): the 5-stage shape is used always.
#ifdef SDCB200_EVAL_SLIM
    if (e.profile == nullptr) return launch_eval_pipe_t<NDIM, PER, true>(e, s);
#endif
    return launch_eval_pipe_t<NDIM, PER, false>(e, s);
}

template <int MODE>
int dispatch_eval(const EvalArgs& a, cudaStream_t s) {
    const bool per = a.g.periodic;
    switch (a.g.ndim) {
        case 1: return per ? launch_eval<1, true, MODE>(a, s) : launch_eval<1, false, MODE>(a, s);
        case 2: return per ? launch_eval_pipe<2, true>(a, s) : launch_eval_pipe<2, false>(a, s);
        case 3: return per ? launch_eval_pipe<3, true>(a, s) : launch_eval_pipe<3, false>(a, s);
    }
    return fail("dispatch_eval", "ndim must be 1, 2 or 3");
}

inline bool ok16(const void* p) { return p != nullptr && (reinterpret_cast<size_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace sdcb200

using namespace sdcb200;

extern "C" {

int sdcb200_heat_eval_f(int ndim, int n, int bc, double a_diag, double a_off, int B, const double* const* u,
                        double* const* f_impl, const double* profile, const double* gt_host, double* const* f_expl,
                        void* stream) {
    SDC_REQUIRE(ndim >= 1 && ndim <= 3, "ndim must be 1, 2 or 3");
    SDC_REQUIRE(B >= 1 && B <= SDCB200_MAX_NODES + 1, "B out of range");
    SDC_REQUIRE(bc == SDCB200_BC_PERIODIC ? !(n & 1) : (n & 1), "n parity does not match the boundary condition");
    EvalArgs a;
    memset(&a, 0, sizeof(a));
    a.g = make_geom(ndim, n, bc);
    a.B = B;
    a.a_diag = a_diag;
    a.a_off = a_off;
    a.profile = profile;
    for (int b = 0; b < B; ++b) {
        SDC_REQUIRE(ok16(u[b]) && ok16(f_impl[b]), "u / f missing or misaligned");
        a.u[b] = u[b];
        a.f[b] = f_impl[b];
        if (profile != nullptr) {
            SDC_REQUIRE(ok16(profile) && f_expl && ok16(f_expl[b]) && gt_host, "forcing arguments missing or misaligned");
            a.fexpl[b] = f_expl[b];
            a.gt[b] = gt_host[b];
        }
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return profile ? dispatch_eval<1>(a, s) : dispatch_eval<0>(a, s);
}

int sdcb200_heat_eval_f_slab(int n, int nz, int bc, double a_diag, double a_off, int B, const double* const* u,
                             double* const* f_impl, const double* profile, const double* gt_host,
                             double* const* f_expl, void* stream) {
    SDC_REQUIRE(B >= 1 && B <= SDCB200_MAX_NODES + 1, "B out of range");
    SDC_REQUIRE(nz >= 1, "empty slab");
    SDC_REQUIRE(bc == SDCB200_BC_PERIODIC ? !(n & 1) : (n & 1), "n parity does not match the boundary condition");
    EvalArgs a;
    memset(&a, 0, sizeof(a));
    a.g = make_slab_geom(n, nz, bc);
    a.B = B;
    a.a_diag = a_diag;
    a.a_off = a_off;
    a.profile = profile;
    for (int b = 0; b < B; ++b) {
        SDC_REQUIRE(ok16(u[b]) && ok16(f_impl[b]), "u / f missing or misaligned");
        a.u[b] = u[b];
        a.f[b] = f_impl[b];
        if (profile != nullptr) {
            SDC_REQUIRE(ok16(profile) && f_expl && ok16(f_expl[b]) && gt_host, "forcing arguments missing or misaligned");
            a.fexpl[b] = f_expl[b];
            a.gt[b] = gt_host[b];
        }
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return profile ? dispatch_eval<1>(a, s) : dispatch_eval<0>(a, s);
}

int sdcb200_allencahn_eval_f(int n, double a_diag, double a_off, double inv_eps2, int nu_exp, int split, int B,
                             const double* const* u, double* const* f, double* const* f_expl, void* stream) {
    SDC_REQUIRE(B >= 1 && B <= SDCB200_MAX_NODES + 1, "B out of range");
    SDC_REQUIRE(n >= 2 && !(n & 1), "periodic grid needs an even number of points per dimension");
    SDC_REQUIRE(nu_exp >= 1, "nu must be a positive integer");
    EvalArgs a;
    memset(&a, 0, sizeof(a));
    a.g = make_geom(2, n, SDCB200_BC_PERIODIC);
    a.B = B;
    a.a_diag = a_diag;
    a.a_off = a_off;
    a.inv_eps2 = inv_eps2;
    a.nu_exp = nu_exp;
    SDC_REQUIRE(split >= 0 && split <= 2 && (split == 0) == (f_expl == nullptr), "split / f_expl mismatch");
    a.split = split;
    for (int b = 0; b < B; ++b) {
        SDC_REQUIRE(ok16(u[b]) && ok16(f[b]), "u / f missing or misaligned");
        a.u[b] = u[b];
        a.f[b] = f[b];
        if (f_expl != nullptr) {
            SDC_REQUIRE(ok16(f_expl[b]), "f_expl missing or misaligned");
            a.fexpl[b] = f_expl[b];
        }
    }
    return launch_eval_pipe<2, true>(a, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
