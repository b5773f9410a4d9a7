// Pieces shared by the CG / Newton kernels: argument structs, the software grid barrier, fixed-order grid reductions.
#pragma once
#include "stencil.cuh"

namespace sdcb200 {

struct Sys {
    const double* b;     // right-hand side
    double* x;           // in: initial guess, out: solution
    double* r;
    double* p;
    double* q;
    const double* dvec;  // optional full diagonal of the operator (Allen-Cahn Jacobian); NULL -> m_diag
    double m_diag;       // 1 - factor*a_diag
    double m_off;        // -factor*a_off
};

struct CgArgs {
    Geom g;
    int B;
    Sys s[SDCB200_MAX_NODES];
    double rtol;
    int maxiter;
    double* partials;  // [2][MAX_NODES][gridDim.x]
    unsigned* bar;     // grid barrier word (zero before the launch)
    int* iters_out;    // [B], += iterations
};

struct NewtonArgs {
    Geom g;
    double factor, a_diag, a_off, inv_eps2;
    int nu_exp;
    const double* rhs;
    double* u;
    double* gvec;  // Newton residual
    double* z;     // Newton update (CG solution)
    double* dvec;  // Jacobian diagonal
    double* r;
    double* p;
    double* q;
    double newton_tol, lin_tol, inexact_ratio;
    int newton_maxiter, lin_maxiter;
    double* partials;
    unsigned* bar;
    int* counters_out;  // [0] += newton iterations, [1] += CG iterations
};

// ---------------------------------------------------------------------------------------------------------------------
// grid-wide barrier (all CTAs co-resident: cooperative launch).  Same protocol as cooperative groups' grid sync:
// CTA barrier, one thread arrives with a gpu-scope release and spins with gpu-scope acquire, CTA barrier.  CTA 0 adds
// the complement so that the top bit flips once per generation and the word never needs resetting.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned* bar) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned add = (blockIdx.x == 0) ? (0x80000000u - (gridDim.x - 1)) : 1u;
        unsigned old;
        asm volatile("atom.add.release.gpu.u32 %0,[%1],%2;" : "=r"(old) : "l"(bar), "r"(add) : "memory");
        unsigned cur;
        do {
            asm volatile("ld.acquire.gpu.u32 %0,[%1];" : "=r"(cur) : "l"(bar) : "memory");
        } while (((old ^ cur) & 0x80000000u) == 0);
    }
    __syncthreads();
}

// Sum the per-CTA partials of one quantity in a fixed order; identical bits in every thread of every CTA.
__device__ __forceinline__ double grid_sum(const double* partials, int slot, int b, double* scratch) {
    const double* src = partials + (size_t)(slot * SDCB200_MAX_NODES + b) * gridDim.x;
    double v = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) v += __ldcg(src + i);
    return block_sum(v, scratch);
}
__device__ __forceinline__ double grid_max(const double* partials, int slot, int b, double* scratch) {
    const double* src = partials + (size_t)(slot * SDCB200_MAX_NODES + b) * gridDim.x;
    double v = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) v = fmax(v, __ldcg(src + i));
    return block_max(v, scratch);
}
__device__ __forceinline__ void put_partial(double* partials, int slot, int b, double v) {
    if (threadIdx.x == 0) partials[(size_t)(slot * SDCB200_MAX_NODES + b) * gridDim.x + blockIdx.x] = v;
}

// scalar state of the solver, one copy per CTA in shared memory, written by thread 0 only
struct CgShared {
    double scratch[33];
    double bb[SDCB200_MAX_NODES], rr[SDCB200_MAX_NODES], rho_prev[SDCB200_MAX_NODES];
    double alpha[SDCB200_MAX_NODES], beta[SDCB200_MAX_NODES];
    int iters[SDCB200_MAX_NODES];
    unsigned active;  // bit b set: system b still iterating
};

}  // namespace sdcb200
