/*
 * sdc_b200.h — C ABI of libsdcb200.so: the sm_100a kernels behind pySDC's SDC sweep hot path.
 *
 * Boundary.  The reference (pySDC, pure Python) has no FFI: its "plugin API" is a set of Python classes chosen
 * through the `description` dict (pySDC/core/step.py:109-160, pySDC/core/level.py:61-88).  The Python classes in
 * `pysdc_b200/` mirror those classes one-to-one and call THIS library through ctypes; each entry point below names
 * the reference code it replaces (paths relative to /root/reference/pySDC).  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every pointer named *_dev / every `double*` field argument is a DEVICE pointer to fp64 data in the padded
 *     layout described below; `const double* const*` arguments are HOST arrays of such device pointers;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is asynchronous on it;
 *   - every function returns 0 on success, non-zero on failure; sdcb200_last_error() describes the last failure
 *     of the calling thread.  Nothing falls back to the CPU.
 *
 * Field layout ("walled" layout).  A field on an n^ndim grid is stored with pitch P = n + (n & 1) in every
 * dimension, i.e. volume V = P^ndim doubles, row-major (x fastest), preceded by a guard of G zero doubles
 * (G = P^(ndim-1) rounded up to a multiple of 16).  Field pointers passed to this library point at element
 * (0,..,0), i.e. just after the guard.
 *   - Dirichlet-zero grids have odd n (the reference enforces it: generic_ND_FD.py:127-131), so P = n+1 and the
 *     extra index in every dimension is a zero "wall": it IS the homogeneous boundary value, and the wall/guard
 *     make every 3/5/7-point neighbour access in-bounds and branch-free.  Kernels never write non-zero walls.
 *   - periodic grids have even n, so P = n (dense); neighbours wrap explicitly.
 *   Several fields of one batch may live anywhere; kernels take pointer arrays, not a stacked tensor.
 */
#ifndef SDC_B200_H
#define SDC_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDCB200_MAX_NODES 8   /* max collocation nodes / batched systems per launch */
#define SDCB200_MAX_TERMS 24  /* max input fields of one collocation launch ((M+1) * components) */

#define SDCB200_BC_DIRICHLET 0
#define SDCB200_BC_PERIODIC 1

#define SDCB200_PRECOND_NONE 0        /* plain CG: the reference's algorithm, iteration for iteration */
#define SDCB200_PRECOND_CHEBYSHEV1 1  /* CG preconditioned with a degree-1 Chebyshev polynomial in the operator */

int sdcb200_version(void);
const char* sdcb200_last_error(void);
/* SM count, compute capability and co-resident CTA capacity of the persistent solver kernel on the current device */
int sdcb200_device_info(int* sm_count, int* cc_major, int* cc_minor, int* solver_ctas);

/* ---- layout helpers (host only) -------------------------------------------------------------------------------- */
long long sdcb200_pitch(int n);
long long sdcb200_volume(int ndim, int n);
long long sdcb200_guard(int ndim, int n);

/* ---- K6: datatype utilities (datatype_classes/mesh.py) --------------------------------------------------------- */
/* max |x_i| over `count` doubles -> *out_dev (device double).  Replaces mesh.__abs__ (mesh.py:65-83). */
int sdcb200_maxabs(const double* x, long long count, double* out_dev, void* stream);
/* out = a*x + b*y (y may be NULL -> out = a*x).  Plain streaming helper for the datatype operators (mesh.py:51-63) */
int sdcb200_axpby(long long count, double a, const double* x, double b, const double* y, double* out, void* stream);

/* ---- K1: collocation --------------------------------------------------------------------------------------------
 * out[m] = sum_k W[m*nin + k] * in[k] (k ascending, products and sums rounded separately) + (base ? base : 0)
 *          + (add && add[m] ? add[m] : 0),   m < nout, k < nin
 * One coalesced, double2-vectorised pass: every input is read once, every output written once.  Generic linear
 * combination of fields: the node-to-node combinations of the FAS transfer (core/base_transfer.py:93-251).          */
int sdcb200_colloc_apply(long long count, int nout, int nin, const double* W_host,
                         const double* const* in, const double* base, const double* const* add,
                         double* const* out, void* stream);

/* The node combinations of one SDC sweep in the reference's OWN order of floating-point operations (every product and
 * sum rounded separately, no FMA contraction), so that they reproduce numpy bit for bit:
 *
 *   acc[m] = BASE_FIRST ? base : 0
 *   QUADRATURE:  for j < nj:  acc[m] += Wq[m*nj+j] * F_j        F_j = in[j]               (ncomp == 1, mesh)
 *                                                                F_j = in[2j] + in[2j+1]   (ncomp == 2, imex_mesh:
 *                                                                                           f[j].impl + f[j].expl)
 *   QDELTA:      for j < nj:  acc[m] += Wi[m*nj+j] * in[j]                                  (ncomp == 1)
 *                             acc[m] += dt2 * (Wi[m*nj+j]*in[2j] + We[m*nj+j]*in[2j+1])     (ncomp == 2)
 *   if !BASE_FIRST and base:  acc[m] += base
 *   if add && add[m]:         acc[m] += add[m]                                              (tau)
 *   out[m] = acc[m]                                                    m < nout; out[m] may alias base
 *
 * - integrate (generic_implicit.py:29-49, imex_1st_order.py:37-55): QUADRATURE with Wq = dt*Q.
 * - known terms of update_nodes (generic_implicit.py:70-82): QUADRATURE | QDELTA with Wq = dt*Q, Wi = -(dt*QI), base =
 *   u[0], add = tau; IMEX (imex_1st_order.py:77-88): Wi = -QI, We = -QE, dt2 = dt (negation is exact, so
 *   `integral -= dt*(QI f.impl + QE f.expl)` is reproduced).
 * - new-node additions (generic_implicit.py:87-89, imex_1st_order.py:92-95): BASE_FIRST | QDELTA on rhs[m] in place.
 * - compute_end_point (generic_implicit.py:123-129, imex_1st_order.py:128-135): BASE_FIRST | QUADRATURE with
 *   Wq = dt*weights, base = u[0], add = tau[-1].
 * One coalesced double2 pass; the second read of in[] in the QDELTA phase is served by L1/L2.                        */
#define SDCB200_SWEEP_QUADRATURE 1
#define SDCB200_SWEEP_QDELTA 2
#define SDCB200_SWEEP_BASE_FIRST 4
int sdcb200_colloc_sweep(long long count, int nout, int nj, int ncomp, int flags, const double* Wq_host,
                         const double* Wi_host, const double* We_host, double dt2, const double* const* in,
                         const double* base, const double* const* add, double* const* out, void* stream);

/* res[m] = sum_j Wq[m*nj+j]*F_j + (u0 - u[m]) + tau[m];  resnorm_dev[m] = max|res[m]|  (device, M doubles); F_j as in
 * sdcb200_colloc_sweep.  res_out may be NULL (norm only, nothing written) or an array of M device pointers.
 * Replaces Sweeper.compute_residual (core/sweeper.py:164-215) incl. the max-norm of mesh.__abs__.                 */
int sdcb200_colloc_residual(long long count, int M, int nj, int ncomp, const double* Wq_host,
                            const double* const* in, const double* u0, const double* const* u,
                            const double* const* tau, double* const* res_out, double* resnorm_dev, void* stream);

/* ---- K2: right-hand sides ------------------------------------------------------------------------------------------
 * f = A u with A = a_off * (sum of 2*ndim neighbours) + a_diag * u  (order-2 centred FD Laplacian times nu;
 * GenericNDimFinDiff.eval_f, generic_ND_FD.py:188-206; matrix entries as built by problem_helper.py:224-241).
 * B independent fields per launch.  If profile != NULL also writes f_expl[b] = profile * gt[b]
 * (heatNd_forced.eval_f, HeatEquation_ND_FD.py:162-204: forcing = spatial profile x scalar g(t)).                    */
int sdcb200_heat_eval_f(int ndim, int n, int bc, double a_diag, double a_off, int B,
                        const double* const* u, double* const* f_impl,
                        const double* profile, const double* gt_host, double* const* f_expl, void* stream);

/* split == 0: f = A u + inv_eps2 * u * (1 - u^nu_exp)   (allencahn_fullyimplicit.eval_f, AllenCahn_2D_FD.py:207-228);
 * split == 1: f = A u, f_expl = inv_eps2 * u * (1 - u^nu_exp)             (allencahn_semiimplicit.eval_f, :278-303);
 * split == 2: f = A u - inv_eps2 * u^(nu_exp+1), f_expl = inv_eps2 * u    (allencahn_semiimplicit_v2.eval_f, :402-424).
 * f_expl is NULL exactly when split == 0.                                                                            */
int sdcb200_allencahn_eval_f(int n, double a_diag, double a_off, double inv_eps2, int nu_exp, int split, int B,
                             const double* const* u, double* const* f, double* const* f_expl, void* stream);

/* ---- K3: node solves ------------------------------------------------------------------------------------------------
 * Solve (I - factor_b * A) x_b = rhs_b for b < B with unpreconditioned CG following scipy.sparse.linalg.cg
 * (scipy 1.18: atol' = rtol*||b||, test ||r|| < atol' before each iteration, x0 = incoming x_b), all iterations of
 * all B systems inside ONE persistent cooperative launch (device-side reductions, grid barriers, per-system
 * convergence).  Replaces GenericNDimFinDiff.solve_system with solver_type='CG' (generic_ND_FD.py:252-260).
 * m_diag[b] = 1 - factor_b*a_diag, m_off[b] = -factor_b*a_off (host arrays).  iters_dev[b] += iterations used.
 * precond = SDCB200_PRECOND_CHEBYSHEV1 (2-D / 3-D dirichlet-zero grids) preconditions the iteration with
 * z = a r + b M r, the degree-1 Chebyshev polynomial for the known spectrum (1, 1 + 4 ndim factor a_off) of M: the same
 * stopping test on ||r||, about half the iterations, one more stencil pass per iteration.
 * work must hold sdcb200_cg_workspace_bytes() bytes of device memory, 256-byte aligned, ZERO-FILLED by the caller
 * before its first use and otherwise left alone between calls (the solver keeps the walls of its work fields zero).  */
size_t sdcb200_cg_workspace_bytes(int ndim, int n, int B);
/* Measurement aid (bench.py --timeline): dev_ns8 != NULL makes the pipelined CG kernels of the calling thread's later
 * launches add, in nanoseconds of %globaltimer, for the first CTA ([0..3]) and the last CTA ([4..7]) of the grid: [0] time
 * spent working between synchronisations (passes incl. pipeline fill / drain and tail), [1] waiting in the grid barrier,
 * [2] summing the partials, [3] in the cross-rank exchange of the sums (slab runs); [8..10] = grid size, units per system, planes per unit of
 * the last launch; [16 + 4*c + k]: the same four accumulators for every CTA c.  The buffer holds 16 + 4*1184 values.
 * NULL switches it off.                                                                                             */
int sdcb200_set_timeline(unsigned long long* dev_ns12);
int sdcb200_heat_cg_solve(int ndim, int n, int bc, int B, const double* m_diag_host, const double* m_off_host,
                          const double* const* rhs, double* const* x, double rtol, int maxiter, int precond,
                          void* work, size_t work_bytes, int* iters_dev, void* stream);

/* ---- multi-GPU: slab decomposition of 3-D grids along the slowest axis ----------------------------------------------
 * No reference counterpart: the reference's FD problems are not space-parallel.  One process per GPU; a
 * rank owns nz consecutive planes of the n x n cross-section.  Slab fields use the walled layout with the guard plane
 * as LOWER halo plane and plane nz as UPPER halo plane (volume P*P*(nz+1) doubles after the guard).
 *
 * Peer memory: each rank allocates its solver workspace with sdcb200_peer_alloc (zero-filled device memory + a 64-byte
 * cudaIpc handle), ships the handle to the other ranks of the node through any host channel and maps theirs with
 * sdcb200_peer_open.  The persistent CG kernel then exchanges the halo planes of its residual and all-reduces its dot
 * products through that memory (NVLink loads/stores, sequence-numbered flags), with no host involvement per iteration.
 * The caller must have put the neighbours' boundary planes into the halo planes of x (initial guess) before the call;
 * on return the halo planes of x are stale.  All ranks must call with the same B, rtol, maxiter.                      */
#define SDCB200_IPC_HANDLE_BYTES 64
int sdcb200_peer_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64);
int sdcb200_peer_open(const unsigned char* handle64, void** peer_ptr);
int sdcb200_peer_close(void* peer_ptr);
int sdcb200_peer_free(void* dev_ptr);
size_t sdcb200_slab_cg_workspace_bytes(int n, int nz_max, int B);
int sdcb200_heat_cg_solve_slab(int n, int nz, int nz_max, int bc, int B, const double* m_diag_host,
                               const double* m_off_host, const double* const* rhs, double* const* x, double rtol,
                               int maxiter, int precond, int rank, int nranks, const int* nz_of_rank,
                               void* const* work_of_rank, size_t work_bytes, int* iters_dev, void* stream);
/* eval_f on a slab (halo planes of u filled by the caller) */
int sdcb200_heat_eval_f_slab(int n, int nz, int bc, double a_diag, double a_off, int B,
                             const double* const* u, double* const* f_impl,
                             const double* profile, const double* gt_host, double* const* f_expl, void* stream);

/* ---- K5: space transfer (FAS restriction / prolongation between nested grids) ---------------------------------------
 * out[o, i, c] = sum_t W[i*width + t] * in[o, col[i*width + t], c]   (col < 0: unused slot), inner stride 1.
 * One 1-D interpolation / restriction operator in ELL form applied along one axis of a field; N-D transfers are one
 * call per axis.  Replaces the sparse Kronecker-product matvecs of mesh_to_mesh.restrict / prolong
 * (transfer_classes/TransferMesh.py:149-218; operators as built by helpers/transfer_helper.py:139-247).             */
int sdcb200_axis_apply(long long n_outer, int n_out, long long n_inner, int width, const double* W_dev,
                       const int* col_dev, const double* in, long long in_stride_outer, long long in_stride_axis,
                       double* out, long long out_stride_outer, long long out_stride_axis, void* stream);

/* Direct solve on 1-D grids: (I - factor*A) is a constant-coefficient (cyclic) tridiagonal matrix; Thomas algorithm in
 * shared memory, one CTA per system (3 <= n <= 8192).  Replaces solver_type='direct' (scipy spsolve,
 * generic_ND_FD.py:239) for ndim == 1, the reference's CPU tutorial configuration.                                  */
int sdcb200_heat_direct_solve_1d(int n, int bc, int B, const double* m_diag_host, const double* m_off_host,
                                 const double* const* rhs, double* const* x, void* stream);

/* ---- higher-order stencils (order 4 / 6 / 8) ---------------------------------------------------------------------------
 * A = nu * (Kronecker sum of the 1-D centred second-derivative stencil of the given order) / dx^2 as built by
 * helpers/problem_helper.py:42-80,83-242: centre_host[k] (k = 0 .. order/2) are the centred coefficients (already
 * scaled by nu/dx^2); on dirichlet-zero grids the order/2 rows next to each boundary use the reference's one-sided
 * closure stencils, lo_host / hi_host = (order/2) x (order+1) coefficient rows acting on the first / last order+1
 * points of a line (NULL on periodic grids).  eval_f and the CG node solves (I - factor_b A) x_b = rhs_b with the same
 * persistent, node-batched, device-resident iteration as sdcb200_heat_cg_solve.  Replaces GenericNDimFinDiff.eval_f /
 * solve_system for order > 2 (generic_ND_FD.py:188-264; tests/test_2d_fd_accuracy.py).                              */
int sdcb200_heat_eval_f_ho(int ndim, int n, int bc, int order, const double* centre_host, const double* lo_host,
                           const double* hi_host, int B, const double* const* u, double* const* f_impl,
                           const double* profile, const double* gt_host, double* const* f_expl, void* stream);
size_t sdcb200_cg_ho_workspace_bytes(int ndim, int n, int B);
int sdcb200_heat_cg_solve_ho(int ndim, int n, int bc, int order, const double* centre_host, const double* lo_host,
                             const double* hi_host, int B, const double* factor_host, const double* const* rhs,
                             double* const* x, double rtol, int maxiter, void* work, size_t work_bytes, int* iters_dev,
                             void* stream);

/* ---- general finite-difference operators + GMRES ---------------------------------------------------------------------
 * A = (Kronecker sum of ONE 1-D stencil) with offsets -h .. h (h <= 4): coef_host[k + h] is the coefficient of offset k,
 * already scaled by coeff / dx^derivative - any stencil helpers/problem_helper.py:4-80 produces within that width
 * (centred, upwind, forward, backward; first or second derivative).  Periodic wrap, or - dirichlet-zero - closure rows
 * lo_host / hi_host = h x (2h+1) as above (NULL on periodic grids).
 *   sdcb200_fd_eval_f:     f_b = A u_b                      (GenericNDimFinDiff.eval_f, generic_ND_FD.py:188-206; the
 *                                                            advection problems of AdvectionEquation_ND_FD.py)
 *   sdcb200_fd_gmres_solve: (I - factor A) x = rhs by restarted GMRES exactly as the reference calls scipy's
 *                           (generic_ND_FD.py:241-250: x0, rtol = lintol, maxiter = liniter counting INNER iterations
 *                           - callback_type='legacy' -, atol = 0, restart = 20, no preconditioner); x: in = initial
 *                           guess, out = solution; iters_dev[0] += inner iterations (= calls of the reference's work
 *                           counter).  One persistent cooperative launch per solve.  work:
 *                           sdcb200_fd_gmres_workspace_bytes(ndim, n, restart) bytes, 256-byte aligned.              */
int sdcb200_fd_eval_f(int ndim, int n, int bc, int h, const double* coef_host, const double* lo_host,
                      const double* hi_host, int B, const double* const* u, double* const* f, void* stream);
size_t sdcb200_fd_gmres_workspace_bytes(int ndim, int n, int restart);
int sdcb200_fd_gmres_solve(int ndim, int n, int bc, int h, const double* coef_host, const double* lo_host,
                           const double* hi_host, double factor, const double* rhs, double* x, double rtol, int maxiter,
                           int restart, void* work, size_t work_bytes, int* iters_dev, void* stream);

/* ---- K4: Allen-Cahn Newton ------------------------------------------------------------------------------------------
 * Newton iteration with inner CG on the Jacobian for B node systems  u_b - factor_b (A u_b + 1/eps^2 u_b (1 - u_b^nu)) =
 * rhs_b, whole solve in one persistent launch (allencahn_fullyimplicit.solve_system, AllenCahn_2D_FD.py:137-205; B > 1:
 * the independent node systems of a diagonal QDelta).  u[b]: in = initial guess, out = solution.  Systems leave the
 * Newton loop and the inner CG individually.  counters_dev[0] += Newton iterations, counters_dev[1] += CG iterations
 * (summed over the systems).  inexact_ratio <= 0 disables :176-177.  variant == 1: the Newton system of
 * allencahn_semiimplicit_v2 (:426-466), u - factor (A u - u^(nu+1)/eps^2) = rhs.  The inner CG runs the TMA-pipelined passes with
 * periodic wrap boxes and the Jacobian diagonal as a tile-only box.  work: sdcb200_newton_workspace_bytes(n, B) bytes,
 * 256-byte aligned, zero-filled before its first use.                                                               */
size_t sdcb200_newton_workspace_bytes(int n, int B);
int sdcb200_allencahn_newton_solve(int n, int B, int variant, const double* factor_host, double a_diag, double a_off,
                                   double inv_eps2, int nu_exp, const double* const* rhs, double* const* u, double newton_tol,
                                   int newton_maxiter, double lin_tol, int lin_maxiter, double inexact_ratio,
                                   void* work, size_t work_bytes, int* counters_dev, void* stream);

/* Point-wise Newton solve of the reaction part,  u_b - factor_b (1/eps^2) u_b (1 - u_b^nu) = rhs_b  on `count` grid
 * points for B systems in one persistent launch: allencahn_multiimplicit.solve_system_2 (AllenCahn_2D_FD.py:594-651).
 * Global loop as in the reference (all points take the same number of updates, decided by max|g| < newton_tol over the
 * grid, at most newton_maxiter); its diagonal Jacobian system is solved exactly (the limit of the reference's CG).
 * u[b]: in = initial guess, out = solution.  counters_dev[0] += Newton updates (summed over the systems).
 * work: sdcb200_reaction_workspace_bytes() bytes, 256-byte aligned.                                                  */
size_t sdcb200_reaction_workspace_bytes(void);
int sdcb200_allencahn_reaction_newton(long long count, int B, const double* factor_host, double inv_eps2, int nu_exp,
                                      const double* const* rhs, double* const* u, double newton_tol, int newton_maxiter,
                                      void* work, size_t work_bytes, int* counters_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDC_B200_H */
