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
    // polynomial preconditioner  z = pc_a * r + pc_b * (M r)  (degree-1 Chebyshev polynomial in M); z == NULL: none
    double* z;
    double pc_a, pc_b;
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
    unsigned long long* timeline;  // optional [2][4] ns accumulators (sdcb200_set_timeline): where the time of a solve goes
};

// Batched Allen-Cahn Newton solves: system b solves  u - factor_b (A u + 1/eps^2 u (1 - u^nu)) = rhs_b  for u_b.  The
// inner CG systems (right-hand side g_b, solution z_b, Jacobian diagonal d_b, work fields) are described by CgArgs.s[b].
struct NewtonSys {
    const double* rhs;
    double* u;
    double factor;
};
struct NewtonArgs {
    CgArgs cg;     // g, B, s[b] = {b: g_b, x: z_b, r, p, q, dvec: d_b (written by the kernel), m_off}, partials, bar
    NewtonSys ns[SDCB200_MAX_NODES];
    double a_diag, a_off, inv_eps2;
    int nu_exp;
    int variant;   // 0: allencahn_fullyimplicit; 1: allencahn_semiimplicit_v2 (implicit part A u - u^(nu+1)/eps^2)
    double newton_tol, lin_tol, inexact_ratio;
    int newton_maxiter, lin_maxiter;
    int* counters_out;  // [0] += newton iterations, [1] += CG iterations (summed over the systems)
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

// Per-CTA partial sums live in kPartialSlots areas of [MAX_NODES][gridDim.x] doubles.  A slot may only be rewritten
// after a grid barrier that every CTA enters AFTER its last read of the slot, so consecutive reductions never share a
// slot: pass A uses 0, pass B 1, pass C (preconditioner) 2, the set-up pass 3 and 4 (written once per solve), the Newton
// residual norm 5.
constexpr int kPartialSlots = 6;
enum { kSlotA = 0, kSlotB = 1, kSlotC = 2, kSlotSetup0 = 3, kSlotSetup1 = 4, kSlotNewton = 5 };

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

// ---------------------------------------------------------------------------------------------------------------------
// Slab decomposition over the GPUs of one NVLink/NVSwitch node: everything the persistent solver needs to talk to its
// neighbours through PEER-MAPPED memory (cudaIpc), no host round trip and no NCCL call inside a solve.
//   - halo exchange: the pass that updates r stores its two boundary planes straight into the neighbours' halo planes
//   - all-reduce of the dot products: every rank writes its local sums into a mailbox slot on EVERY rank, raises a
//     sequence-numbered flag (release at system scope) and sums the slots of all ranks in rank order, so all ranks hold
//     bit-identical scalars and take identical branches.  Mailboxes are double-buffered by the parity of the sequence
//     number; a rank cannot run more than one synchronisation ahead of any other.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kMaxRanks = 8;
constexpr int kMailVals = 2 * SDCB200_MAX_NODES;  // doubles one rank publishes per synchronisation

struct SlabLink {
    int rank, nranks;
    int has_lo, has_hi;                          // a neighbouring slab below / above
    double* lo_r_halo[SDCB200_MAX_NODES];        // plane nz of the lower neighbour's r (its upper halo plane), per system
    double* hi_r_halo[SDCB200_MAX_NODES];        // plane -1 of the upper neighbour's r
    double* lo_z_halo[SDCB200_MAX_NODES];        // the same for the preconditioned residual z (preconditioned runs)
    double* hi_z_halo[SDCB200_MAX_NODES];
    unsigned long long* flags_of[kMaxRanks];     // [2][kMaxRanks] flags in every rank's mailbox (peer-mapped; own included)
    double* vals_of[kMaxRanks];                  // [2][kMaxRanks][kMailVals]
    unsigned long long* seq;                     // own persistent synchronisation counter
    int* error;                                  // own: set to 1 when a peer did not show up in time
    unsigned long long timeout_ns;               // how long to wait for a peer before giving up (SDCB200_PEER_TIMEOUT_S)
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// grid barrier whose arrival also publishes this CTA's peer stores system-wide
__device__ __forceinline__ void grid_barrier_sys(unsigned* bar) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
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

// scalar state of the solver, one copy per CTA in shared memory, written by thread 0 only
struct CgShared {
    double scratch[33];
    double loc[kMailVals];   // slab runs: this rank's sums / the global sums of one synchronisation
    double glob[kMailVals];
    double bb[SDCB200_MAX_NODES], rr[SDCB200_MAX_NODES], rho_prev[SDCB200_MAX_NODES];
    double rz[SDCB200_MAX_NODES];  // r.z of the preconditioned solver (== rr without a preconditioner)
    double alpha[SDCB200_MAX_NODES], beta[SDCB200_MAX_NODES];
    double rtol[SDCB200_MAX_NODES];  // relative tolerance of each system (set by the caller of the collective solver)
    int iters[SDCB200_MAX_NODES];
    unsigned active;  // bit b set: system b still iterating
    unsigned long long t_last;  // timeline: end of the previous synchronisation
};

// Cross-rank sum of `nv` values per rank (sh.loc[0..nv) on entry, identical in every CTA of the rank).  On return
// sh.glob[0..nv) holds the sums over all ranks, bit-identical on every thread of every rank.  Must be called by all
// threads of all CTAs of all ranks with the same `seq` (monotonically increasing across calls and launches).
__device__ __forceinline__ void cross_rank_sum(const SlabLink& L, CgShared& sh, int nv, unsigned long long seq) {
    const unsigned par = (unsigned)(seq & 1ull);
    const int t = threadIdx.x;
    __syncthreads();
    if (blockIdx.x == 0 && t < L.nranks) {
        double* dst = L.vals_of[t] + ((size_t)par * kMaxRanks + L.rank) * kMailVals;
        for (int i = 0; i < nv; ++i) asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(dst + i), "d"(sh.loc[i]) : "memory");
        __threadfence_system();
        unsigned long long* f = L.flags_of[t] + par * kMaxRanks + L.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(seq) : "memory");
    }
    if (t < L.nranks) {
        const unsigned long long* f = L.flags_of[L.rank] + par * kMaxRanks + t;
        unsigned long long v, t0 = 0;
        unsigned spins = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= seq) break;
            if ((++spins & 0x3ffu) == 0) {
                const unsigned long long now = global_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > L.timeout_ns) {  // the peer never showed up: give up loudly instead of hanging the GPU
                    *L.error = 1;
                    __threadfence_system();
                    __trap();
                }
            }
        }
    }
    __syncthreads();
    if (t < nv) {
        const double* src = L.vals_of[L.rank] + (size_t)par * kMaxRanks * kMailVals + t;
        double acc = 0.0;
        for (int r = 0; r < L.nranks; ++r) acc += __ldcg(src + (size_t)r * kMailVals);
        sh.glob[t] = acc;
    }
    __syncthreads();
}

}  // namespace sdcb200
