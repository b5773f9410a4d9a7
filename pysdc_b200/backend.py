"""ctypes binding of ``libsdcb200.so`` (C ABI in ``include/sdc_b200.h``) and the ``Backend`` object the host classes use.

PyTorch is plumbing here: device memory (``torch.Tensor`` storage), the current CUDA stream and, for multi-GPU runs,
``torch.distributed``.  Every numerical operation of the sweep path is a call into the CUDA library; if the library is
missing or no CUDA device is present the backend raises ``BackendError`` — there is no CPU fallback.

Field arguments are 1-D fp64 tensors viewing the *volume* of a walled field (``layout.Layout``); their ``data_ptr()``
is element (0,..,0) and the zero guard sits in front of it in the same storage.
"""
import ctypes
import os

import numpy as np
import torch

from .errors import BackendError

_LIB_NAME = "libsdcb200.so"
_c_dp = ctypes.c_void_p


def lib_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", _LIB_NAME)


def _declare(lib):
    c_int, c_ll, c_d, c_sz = ctypes.c_int, ctypes.c_longlong, ctypes.c_double, ctypes.c_size_t
    PP = ctypes.POINTER(ctypes.c_void_p)
    PD = ctypes.POINTER(ctypes.c_double)
    sig = {
        "sdcb200_version": (c_int, []),
        "sdcb200_last_error": (ctypes.c_char_p, []),
        "sdcb200_device_info": (c_int, [ctypes.POINTER(c_int)] * 4),
        "sdcb200_pitch": (c_ll, [c_int]),
        "sdcb200_volume": (c_ll, [c_int, c_int]),
        "sdcb200_guard": (c_ll, [c_int, c_int]),
        "sdcb200_maxabs": (c_int, [_c_dp, c_ll, _c_dp, _c_dp]),
        "sdcb200_axpby": (c_int, [c_ll, c_d, _c_dp, c_d, _c_dp, _c_dp, _c_dp]),
        "sdcb200_colloc_apply": (c_int, [c_ll, c_int, c_int, PD, PP, _c_dp, PP, PP, _c_dp]),
        "sdcb200_colloc_sweep": (c_int, [c_ll, c_int, c_int, c_int, c_int, PD, PD, PD, c_d, PP, _c_dp, PP, PP, _c_dp]),
        "sdcb200_colloc_residual": (c_int, [c_ll, c_int, c_int, c_int, PD, PP, _c_dp, PP, PP, PP, _c_dp, _c_dp]),
        "sdcb200_heat_eval_f": (c_int, [c_int, c_int, c_int, c_d, c_d, c_int, PP, PP, _c_dp, PD, PP, _c_dp]),
        "sdcb200_allencahn_eval_f": (c_int, [c_int, c_d, c_d, c_d, c_int, c_int, c_int, PP, PP, PP, _c_dp]),
        "sdcb200_cg_workspace_bytes": (c_sz, [c_int, c_int, c_int]),
        "sdcb200_set_timeline": (c_int, [_c_dp]),
        "sdcb200_heat_cg_solve": (c_int, [c_int, c_int, c_int, c_int, PD, PD, PP, PP, c_d, c_int, c_int, _c_dp, c_sz, _c_dp,
                                          _c_dp]),
        "sdcb200_peer_alloc": (c_int, [c_sz, PP, ctypes.c_char_p]),
        "sdcb200_peer_open": (c_int, [ctypes.c_char_p, PP]),
        "sdcb200_peer_close": (c_int, [_c_dp]),
        "sdcb200_peer_free": (c_int, [_c_dp]),
        "sdcb200_slab_cg_workspace_bytes": (c_sz, [c_int, c_int, c_int]),
        "sdcb200_heat_cg_solve_slab": (c_int, [c_int, c_int, c_int, c_int, c_int, PD, PD, PP, PP, c_d, c_int, c_int, c_int,
                                               c_int, ctypes.POINTER(c_int), PP, c_sz, _c_dp, _c_dp]),
        "sdcb200_heat_eval_f_slab": (c_int, [c_int, c_int, c_int, c_d, c_d, c_int, PP, PP, _c_dp, PD, PP, _c_dp]),
        "sdcb200_axis_apply": (c_int, [c_ll, c_int, c_ll, c_int, _c_dp, _c_dp, _c_dp, c_ll, c_ll, _c_dp, c_ll, c_ll, _c_dp]),
        "sdcb200_heat_direct_solve_1d": (c_int, [c_int, c_int, c_int, PD, PD, PP, PP, _c_dp]),
        "sdcb200_heat_eval_f_ho": (c_int, [c_int, c_int, c_int, c_int, PD, PD, PD, c_int, PP, PP, _c_dp, PD, PP, _c_dp]),
        "sdcb200_cg_ho_workspace_bytes": (c_sz, [c_int, c_int, c_int]),
        "sdcb200_heat_cg_solve_ho": (c_int, [c_int, c_int, c_int, c_int, PD, PD, PD, c_int, PD, PP, PP, c_d, c_int, _c_dp,
                                             c_sz, _c_dp, _c_dp]),
        "sdcb200_fd_eval_f": (c_int, [c_int, c_int, c_int, c_int, PD, PD, PD, c_int, PP, PP, _c_dp]),
        "sdcb200_fd_gmres_workspace_bytes": (c_sz, [c_int, c_int, c_int]),
        "sdcb200_fd_gmres_solve": (c_int, [c_int, c_int, c_int, c_int, PD, PD, PD, c_d, _c_dp, _c_dp, c_d, c_int, c_int,
                                           _c_dp, c_sz, _c_dp, _c_dp]),
        "sdcb200_newton_workspace_bytes": (c_sz, [c_int, c_int]),
        "sdcb200_allencahn_newton_solve": (c_int, [c_int, c_int, c_int, PD, c_d, c_d, c_d, c_int, PP, PP, c_d, c_int, c_d,
                                                   c_int, c_d, _c_dp, c_sz, _c_dp, _c_dp]),
        "sdcb200_reaction_workspace_bytes": (c_sz, []),
        "sdcb200_allencahn_reaction_newton": (c_int, [c_ll, c_int, PD, c_d, c_int, PP, PP, c_d, c_int, _c_dp, c_sz, _c_dp,
                                                      _c_dp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here = the .so does not export what include/sdc_b200.h declares
        fn.restype, fn.argtypes = res, args
    return sig


def load_library(path=None):
    """Load the shared library and declare every prototype of ``include/sdc_b200.h`` (no GPU needed)."""
    path = path or lib_path()
    if not os.path.exists(path):
        raise BackendError(f"{path} not found: build it with `python -m pysdc_b200.build` "
                           "(or __graft_entry__.build()); there is no CPU fallback")
    try:
        lib = ctypes.CDLL(path)
    except OSError as e:
        raise BackendError(f"cannot load {path}: {e}") from e
    lib._sdc_signatures = _declare(lib)
    return lib


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def _dbl_array(values):
    values = np.ascontiguousarray(values, dtype=np.float64).ravel()
    return (ctypes.c_double * values.size)(*values.tolist())


class CudaBackend:
    """Thin, stateless wrapper: torch tensors in, kernel launches on the current torch CUDA stream out."""

    name = "cuda"

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise BackendError("no CUDA device visible: pysdc_b200 runs its sweep path on the GPU only")
        self.lib = load_library()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.launches = 0  # kernels launched through this backend (bench.py reports it)

    # -- helpers ------------------------------------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc):
        if rc != 0:
            raise BackendError(self.lib.sdcb200_last_error().decode())

    def zeros(self, count, dtype=torch.float64):
        return torch.zeros(count, dtype=dtype, device=self.device)

    def synchronize(self):
        torch.cuda.current_stream(self.device).synchronize()

    def device_info(self):
        v = [ctypes.c_int(0) for _ in range(4)]
        self._check(self.lib.sdcb200_device_info(*[ctypes.byref(x) for x in v]))
        return dict(sm_count=v[0].value, cc=(v[1].value, v[2].value), solver_ctas=v[3].value)

    # -- K6 -----------------------------------------------------------------------------------------------------------
    def maxabs_async(self, x, out):
        self.launches += 1
        self._check(self.lib.sdcb200_maxabs(x.data_ptr(), x.numel(), out.data_ptr(), self._stream()))

    def maxabs(self, x):
        out = torch.empty(1, dtype=torch.float64, device=self.device)
        self.maxabs_async(x, out)
        return float(out.item())

    def axpby(self, a, x, b, y, out):
        self.launches += 1
        self._check(self.lib.sdcb200_axpby(x.numel(), float(a), x.data_ptr(), float(b),
                                           None if y is None else y.data_ptr(), out.data_ptr(), self._stream()))

    # -- K1 -----------------------------------------------------------------------------------------------------------
    def colloc_apply(self, W, ins, base, adds, outs):
        W = np.asarray(W, dtype=np.float64).reshape(len(outs), len(ins))
        self.launches += 1
        self._check(self.lib.sdcb200_colloc_apply(
            outs[0].numel(), len(outs), len(ins), _dbl_array(W), _ptr_array(ins),
            None if base is None else base.data_ptr(), None if adds is None else _ptr_array(adds),
            _ptr_array(outs), self._stream()))

    def colloc_sweep(self, ins, ncomp, outs, Wq=None, Wi=None, We=None, dt2=0.0, base=None, adds=None,
                     base_first=False):
        """The node combinations of a sweep in the reference's order of operations (sdc_b200.h: sdcb200_colloc_sweep).
        ``ins`` = f[j] per node (ncomp = 1) or f[j].impl, f[j].expl interleaved (ncomp = 2); ``Wq`` (nout x nj) the
        quadrature phase, ``Wi`` / ``We`` the QDelta phase."""
        nj = len(ins) // ncomp
        flags = (1 if Wq is not None else 0) | (2 if Wi is not None else 0) | (4 if base_first else 0)
        arr = lambda W: None if W is None else _dbl_array(np.asarray(W, dtype=np.float64).reshape(len(outs), nj))  # noqa: E731
        self.launches += 1
        self._check(self.lib.sdcb200_colloc_sweep(
            outs[0].numel(), len(outs), nj, ncomp, flags, arr(Wq), arr(Wi), arr(We), float(dt2), _ptr_array(ins),
            None if base is None else base.data_ptr(), None if adds is None else _ptr_array(adds), _ptr_array(outs),
            self._stream()))

    def colloc_residual(self, Wq, ins, ncomp, u0, us, taus, res_outs, resnorm):
        nj = len(ins) // ncomp
        Wq = np.asarray(Wq, dtype=np.float64).reshape(len(us), nj)
        self.launches += 1
        self._check(self.lib.sdcb200_colloc_residual(
            u0.numel(), len(us), nj, ncomp, _dbl_array(Wq), _ptr_array(ins), u0.data_ptr(), _ptr_array(us),
            None if taus is None else _ptr_array(taus), None if res_outs is None else _ptr_array(res_outs),
            resnorm.data_ptr(), self._stream()))

    # -- K2 -----------------------------------------------------------------------------------------------------------
    def heat_eval_f(self, lay, bc, a_diag, a_off, us, fs, profile=None, gts=None, fexpls=None):
        self.launches += 1
        tail = (a_diag, a_off, len(us), _ptr_array(us), _ptr_array(fs),
                None if profile is None else profile.data_ptr(), None if gts is None else _dbl_array(gts),
                None if fexpls is None else _ptr_array(fexpls), self._stream())
        if lay.is_slab:  # halo planes of us filled by the caller (SlabComm.exchange_halos)
            self._check(self.lib.sdcb200_heat_eval_f_slab(lay.n, lay.nz, bc, *tail))
        else:
            self._check(self.lib.sdcb200_heat_eval_f(lay.ndim, lay.n, bc, *tail))

    def allencahn_eval_f(self, lay, a_diag, a_off, inv_eps2, nu_exp, us, fs, fexpls=None, split=None):
        """fexpls given: fs = A u, fexpls = reaction term (semi-implicit splitting, split = 1) or fs = A u - u^(nu+1)/eps^2,
        fexpls = u/eps^2 (split = 2); else fs = A u + reaction term."""
        self.launches += 1
        split = (0 if fexpls is None else 1) if split is None else int(split)
        self._check(self.lib.sdcb200_allencahn_eval_f(lay.n, a_diag, a_off, inv_eps2, int(nu_exp), split, len(us),
                                                      _ptr_array(us), _ptr_array(fs),
                                                      None if fexpls is None else _ptr_array(fexpls), self._stream()))

    # -- K3 / K4 ------------------------------------------------------------------------------------------------------
    def set_timeline(self, buf):
        """buf: 8-element int64 device tensor the pipelined CG kernels add their phase timings to (None: off)."""
        self._check(self.lib.sdcb200_set_timeline(None if buf is None else buf.data_ptr()))

    def cg_workspace(self, lay, B):
        nbytes = self.lib.sdcb200_cg_workspace_bytes(lay.ndim, lay.n, B)
        return torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=self.device)

    def heat_cg_solve(self, lay, bc, m_diag, m_off, rhs, xs, rtol, maxiter, work, iters_dev, precond=0):
        self.launches += 1
        self._check(self.lib.sdcb200_heat_cg_solve(
            lay.ndim, lay.n, bc, len(xs), _dbl_array(m_diag), _dbl_array(m_off), _ptr_array(rhs), _ptr_array(xs),
            float(rtol), int(maxiter), int(precond), work.data_ptr(), work.numel() * 8, iters_dev.data_ptr(),
            self._stream()))

    # -- higher-order stencils (order 4 / 6 / 8) ------------------------------------------------------------------------
    @staticmethod
    def _ho_tables(op):
        lo = None if op["lo"] is None else _dbl_array(op["lo"])
        hi = None if op["hi"] is None else _dbl_array(op["hi"])
        return _dbl_array(op["centre"]), lo, hi

    def heat_eval_f_ho(self, lay, bc, op, us, fs, profile=None, gts=None, fexpls=None):
        self.launches += 1
        c, lo, hi = self._ho_tables(op)
        self._check(self.lib.sdcb200_heat_eval_f_ho(
            lay.ndim, lay.n, bc, op["order"], c, lo, hi, len(us), _ptr_array(us), _ptr_array(fs),
            None if profile is None else profile.data_ptr(), None if gts is None else _dbl_array(gts),
            None if fexpls is None else _ptr_array(fexpls), self._stream()))

    def cg_ho_workspace(self, lay, B):
        nbytes = self.lib.sdcb200_cg_ho_workspace_bytes(lay.ndim, lay.n, B)
        return torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=self.device)

    def heat_cg_solve_ho(self, lay, bc, op, factors, rhs, xs, rtol, maxiter, work, iters_dev):
        self.launches += 1
        c, lo, hi = self._ho_tables(op)
        self._check(self.lib.sdcb200_heat_cg_solve_ho(
            lay.ndim, lay.n, bc, op["order"], c, lo, hi, len(xs), _dbl_array(factors), _ptr_array(rhs), _ptr_array(xs),
            float(rtol), int(maxiter), work.data_ptr(), work.numel() * 8, iters_dev.data_ptr(), self._stream()))

    # -- general finite-difference operators, GMRES (gmres.cu) ---------------------------------------------------------
    @staticmethod
    def _fd_tables(op):
        lo = None if op["lo"] is None else _dbl_array(np.asarray(op["lo"], dtype=np.float64).ravel())
        hi = None if op["hi"] is None else _dbl_array(np.asarray(op["hi"], dtype=np.float64).ravel())
        return _dbl_array(op["coef"]), lo, hi

    def fd_eval_f(self, lay, bc, op, us, fs):
        """fs[i] = A us[i] for the operator tables of problems.fd_operator_tables."""
        self.launches += 1
        c, lo, hi = self._fd_tables(op)
        self._check(self.lib.sdcb200_fd_eval_f(lay.ndim, lay.n, bc, op["h"], c, lo, hi, len(us), _ptr_array(us),
                                               _ptr_array(fs), self._stream()))

    def fd_gmres_workspace(self, lay, restart):
        nbytes = self.lib.sdcb200_fd_gmres_workspace_bytes(lay.ndim, lay.n, restart)
        return torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=self.device)

    def fd_gmres_solve(self, lay, bc, op, factor, rhs, x, rtol, maxiter, restart, work, iters_dev):
        """(I - factor A) x = rhs by restarted GMRES in one persistent launch, in place on x (initial guess in)."""
        self.launches += 1
        c, lo, hi = self._fd_tables(op)
        self._check(self.lib.sdcb200_fd_gmres_solve(
            lay.ndim, lay.n, bc, op["h"], c, lo, hi, float(factor), rhs.data_ptr(), x.data_ptr(), float(rtol),
            int(maxiter), int(restart), work.data_ptr(), work.numel() * 8, iters_dev.data_ptr(), self._stream()))

    # -- slab-decomposed solves over peer-mapped memory -----------------------------------------------------------------
    def slab_cg_workspace(self, lay, comm, B):
        """This rank's solver workspace in IPC-exportable device memory plus the mapped workspaces of all other ranks of
        ``comm`` (collective call).  Returns a ``SlabWork`` to pass to ``heat_cg_solve_slab``."""
        planes = comm.planes(lay.n)
        nbytes = self.lib.sdcb200_slab_cg_workspace_bytes(lay.n, max(planes), B)
        ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        self._check(self.lib.sdcb200_peer_alloc(nbytes, ctypes.byref(ptr), handle))
        handles = comm.allgather(handle.raw)
        ptrs = []
        for r, h in enumerate(handles):
            if r == comm.rank:
                ptrs.append(ptr.value)
            else:
                q = ctypes.c_void_p()
                self._check(self.lib.sdcb200_peer_open(h, ctypes.byref(q)))
                ptrs.append(q.value)
        torch.cuda.synchronize(self.device)  # the zero-fill of every segment is complete before anybody writes into it
        comm.barrier()
        return SlabWork(self, comm.rank, ptrs, nbytes, planes)

    def heat_cg_solve_slab(self, lay, comm, bc, m_diag, m_off, rhs, xs, rtol, maxiter, work, iters_dev, precond=0):
        self.launches += 1
        planes = (ctypes.c_int * len(work.planes))(*work.planes)
        peers = (ctypes.c_void_p * len(work.ptrs))(*work.ptrs)
        self._check(self.lib.sdcb200_heat_cg_solve_slab(
            lay.n, lay.nz, max(work.planes), bc, len(xs), _dbl_array(m_diag), _dbl_array(m_off), _ptr_array(rhs),
            _ptr_array(xs), float(rtol), int(maxiter), int(precond), comm.rank, comm.size, planes, peers, work.nbytes,
            iters_dev.data_ptr(), self._stream()))

    # -- K5 -----------------------------------------------------------------------------------------------------------
    def upload_operator(self, W, col):
        """ELL operator (weights, column indices) -> device tensors."""
        return (torch.from_numpy(np.ascontiguousarray(W, dtype=np.float64)).to(self.device),
                torch.from_numpy(np.ascontiguousarray(col, dtype=np.int32)).to(self.device))

    def axis_apply(self, op, n_outer, n_out, n_inner, src, src_so, src_sa, dst, dst_so, dst_sa):
        W, col = op
        self.launches += 1
        self._check(self.lib.sdcb200_axis_apply(n_outer, n_out, n_inner, W.shape[1], W.data_ptr(), col.data_ptr(),
                                                src.data_ptr(), src_so, src_sa, dst.data_ptr(), dst_so, dst_sa,
                                                self._stream()))

    def heat_direct_solve_1d(self, lay, bc, m_diag, m_off, rhs, xs):
        self.launches += 1
        self._check(self.lib.sdcb200_heat_direct_solve_1d(lay.n, bc, len(xs), _dbl_array(m_diag), _dbl_array(m_off),
                                                          _ptr_array(rhs), _ptr_array(xs), self._stream()))

    def newton_workspace(self, lay, B):
        nbytes = self.lib.sdcb200_newton_workspace_bytes(lay.n, B)
        return torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=self.device)

    def allencahn_newton_solve(self, lay, factors, a_diag, a_off, inv_eps2, nu_exp, rhs, us, newton_tol, newton_maxiter,
                               lin_tol, lin_maxiter, inexact_ratio, work, counters_dev, variant=0):
        """Newton + inner CG for the len(us) systems (rhs[b], factors[b]) in ONE persistent launch, in place on us[b].
        variant 1: the implicit part of allencahn_semiimplicit_v2."""
        self.launches += 1
        self._check(self.lib.sdcb200_allencahn_newton_solve(
            lay.n, len(us), int(variant), _dbl_array(factors), a_diag, a_off, inv_eps2, int(nu_exp), _ptr_array(rhs), _ptr_array(us),
            float(newton_tol), int(newton_maxiter), float(lin_tol), int(lin_maxiter),
            float(inexact_ratio or 0.0), work.data_ptr(), work.numel() * 8, counters_dev.data_ptr(), self._stream()))

    def reaction_workspace(self):
        nbytes = self.lib.sdcb200_reaction_workspace_bytes()
        return torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=self.device)

    def allencahn_reaction_newton(self, factors, inv_eps2, nu_exp, rhs, us, newton_tol, newton_maxiter, work, counters_dev):
        """Point-wise Newton on the reaction part for the len(us) systems in ONE persistent launch, in place on us[b]
        (``rhs`` / ``us``: flat views of whole fields)."""
        self.launches += 1
        self._check(self.lib.sdcb200_allencahn_reaction_newton(
            us[0].numel(), len(us), _dbl_array(factors), float(inv_eps2), int(nu_exp), _ptr_array(rhs), _ptr_array(us),
            float(newton_tol), int(newton_maxiter), work.data_ptr(), work.numel() * 8, counters_dev.data_ptr(),
            self._stream()))


class SlabWork:
    """Peer-mapped solver workspaces of all ranks (index = rank; own segment at ``rank``)."""

    def __init__(self, backend, rank, ptrs, nbytes, planes):
        self._be, self.rank, self.ptrs, self.nbytes, self.planes = backend, rank, ptrs, nbytes, list(planes)

    def close(self):
        if self.ptrs is None:
            return
        for r, p in enumerate(self.ptrs):
            if r == self.rank:
                self._be.lib.sdcb200_peer_free(ctypes.c_void_p(p))
            else:
                self._be.lib.sdcb200_peer_close(ctypes.c_void_p(p))
        self.ptrs = None


_backend = None


def get_backend():
    """The process-wide backend (created on first use; raises ``BackendError`` without library or GPU)."""
    global _backend
    if _backend is None:
        _backend = CudaBackend()
    return _backend


def set_backend(backend):
    """Install a backend object (used by the CPU tests to drive the host logic with a numpy stand-in)."""
    global _backend
    _backend = backend
    return backend
