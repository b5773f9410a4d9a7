"""Device datatypes with the surface of the reference's ``mesh`` / ``imex_mesh``
(``pySDC/implementations/datatype_classes/mesh.py:12-190``).

A ``mesh`` owns (or views) fp64 device storage in the walled layout (``layout.py``); ``abs()`` is the global max-norm
computed by the ``sdcb200_maxabs`` kernel (+ a MAX all-reduce when a communicator is attached, mesh.py:65-83);
``+ - *`` with meshes and scalars return new meshes of the same type (mesh.py:51-63) through the ``sdcb200_axpby``
kernel; ``u[:] = v`` writes in place; ``isend / irecv / bcast`` move the field between ranks with
``torch.distributed`` (NCCL on device memory) behind the mpi4py-style signatures the controllers use (mesh.py:85-125).

The SDC sweep itself never goes through these generic operators — the sweepers call fused kernels on the underlying
buffers — they exist so that code written against the reference's datatype (controllers, transfer classes, user
scripts such as ``abs(uex - uend)``) keeps working.
"""
import numpy as np
import torch

from .backend import get_backend
from .layout import get_layout


def _as_shape(s):
    return (int(s),) if isinstance(s, (int, np.integer)) else tuple(int(v) for v in s)


class mesh:
    """Single-component field.  ``mesh(init, val=0.0)`` with ``init`` another mesh (deep copy) or
    ``(shape, comm, numpy dtype)`` (mesh.py:24-49)."""

    components = ()
    comm = None
    __array_priority__ = 1000
    __array_ufunc__ = None  # numpy scalars / arrays on the left defer to our reflected operators (np.float64 * mesh)

    def __init__(self, init, val=0.0, *, _buf=None, _lay=None, _ncomp=None):
        if _buf is not None:  # internal: wrap existing storage (component views, arena slices)
            self._lay, self._ncomp, self._buf = _lay, _ncomp, _buf
            return
        if isinstance(init, mesh):
            self._lay, self._ncomp = init._lay, init._ncomp
            self._buf = init._buf.clone()
            self.comm = init.comm
            return
        if isinstance(init, tuple) and len(init) == 3 and isinstance(init[2], np.dtype):
            if init[2] != np.dtype("float64"):
                raise NotImplementedError(f"only float64 fields are implemented on the device, got {init[2]}")
            shape = _as_shape(init[0])
            ncomp = max(len(type(self).components), 1)
            if len(type(self).components) and len(shape) > 1 and shape[0] == ncomp and not self._shape_is_grid(shape):
                shape = shape[1:]
            comm = init[1]
            # a slab communicator (parallel.SlabComm) decomposes 3-D grids along axis 0: `shape` is the GLOBAL shape
            self._lay = comm.slab_layout(shape) if hasattr(comm, "slab_layout") else get_layout(shape)
            self._ncomp = ncomp
            be = get_backend()
            self._buf = be.zeros(self._lay.alloc(ncomp))
            if comm is not None:
                type(self).comm = comm  # class-level like the reference (mesh.py:46)
            if val != 0.0:
                for c in range(ncomp):
                    self._lay.interior(self._comp_vol(c)).fill_(float(val))
            return
        raise NotImplementedError(type(init))

    @staticmethod
    def _shape_is_grid(shape):
        return len(set(shape)) == 1

    # ---- storage views ----------------------------------------------------------------------------------------------
    @property
    def layout(self):
        return self._lay

    @property
    def vol(self):
        """1-D view of all components' volumes (what the streaming kernels see)."""
        return self._buf[self._lay.guard:]

    def _comp_vol(self, c):
        g, v, st = self._lay.guard, self._lay.vol, self._lay.stride
        return self._buf[g + c * st: g + c * st + v]

    @property
    def flat(self):
        """Kernel argument for a single-component field: 1-D view whose data_ptr() is grid point (0,..,0)."""
        assert self._ncomp == 1
        return self._comp_vol(0)

    @property
    def data(self):
        """Strided tensor view of the grid values, shape ``self.shape``."""
        if self._ncomp == 1 and not type(self).components:
            return self._lay.interior(self._comp_vol(0))
        return torch.stack([self._lay.interior(self._comp_vol(c)) for c in range(self._ncomp)])

    @property
    def shape(self):
        return self._lay.shape if not type(self).components else (self._ncomp,) + self._lay.shape

    @property
    def dtype(self):
        return np.dtype("float64")

    @property
    def size(self):
        return int(np.prod(self.shape))

    @property
    def ndim(self):
        return len(self.shape)

    # ---- element access ---------------------------------------------------------------------------------------------
    def _views(self):
        return [self._lay.interior(self._comp_vol(c)) for c in range(self._ncomp)]

    def __setitem__(self, key, value):
        if isinstance(value, mesh):
            if key == slice(None) or key is Ellipsis:
                if value._lay != self._lay or value._ncomp != self._ncomp:
                    raise ValueError(f"shape mismatch: {value.shape} into {self.shape}")
                self._buf.copy_(value._buf)
                return
            value = value.data
        elif isinstance(value, np.ndarray):
            value = self._localize(value)
            value = torch.from_numpy(np.array(value, dtype=np.float64, order="C")).to(self._buf.device)
        if type(self).components:
            views = self._views()
            if key == slice(None) or key is Ellipsis:
                if torch.is_tensor(value) and value.dim() == len(self.shape):
                    for c, v in enumerate(views):
                        v.copy_(value[c])
                else:
                    for v in views:
                        v[...] = value
                return
            raise IndexError("multi-component meshes support whole-array assignment and component access only")
        self.data[key] = value

    def __getitem__(self, key):
        return self.data[key]

    def _localize(self, arr):
        """Slab fields: an array of the GLOBAL grid shape is cut down to the planes this rank owns."""
        lay = self._lay
        if lay.is_slab and arr.shape[-3:] == lay.global_shape and lay.global_shape != lay.shape:
            return arr[..., lay.z0: lay.z0 + lay.nz, :, :]
        return arr

    def gather(self):
        """Global numpy array on every rank (slab fields: all-gather of the slabs; otherwise the same as ``get``)."""
        local = self.get()
        comm = self.comm
        if not self._lay.is_slab or comm is None or comm.size == 1:
            return local
        parts = comm.allgather(local)
        return np.concatenate(parts, axis=-3)

    def flatten(self):
        return self.data.reshape(-1)

    def get(self):
        """Host copy as a numpy array (synchronises)."""
        return self.data.detach().cpu().numpy()

    numpy = get

    def view(self, cls=None):
        """``u.view(numpy.ndarray)`` as used by the reference's output hooks (hooks/log_solution.py:109-110): a host
        copy of the grid values (the device field itself cannot be re-typed)."""
        return self.get()

    def __array__(self, dtype=None, copy=None):
        a = self.get()
        return a if dtype is None else a.astype(dtype)

    def copy(self):
        return type(self)(self)

    # ---- arithmetic -------------------------------------------------------------------------------------------------
    def _new_like(self):
        out = type(self).__new__(type(self))
        mesh.__init__(out, None, _buf=torch.empty_like(self._buf), _lay=self._lay, _ncomp=self._ncomp)
        out._buf[: self._lay.guard].zero_()
        return out

    def _same(self, other):
        return isinstance(other, mesh) and other._lay == self._lay and other._ncomp == self._ncomp

    def _touch(self):
        """Count a write that went through a kernel or the transport (torch's own version counter only sees torch
        operations); the sweepers use both counters to know whether a cached residual is still valid."""
        self._kver = getattr(self, "_kver", 0) + 1

    def _lin(self, a, b, other, out=None):
        """out = a*self + b*other through the axpby kernel (walls stay zero)."""
        if out is not None:
            out._touch()
        out = self._new_like() if out is None else out
        get_backend().axpby(a, self.vol, b, None if other is None else other.vol, out.vol)
        return out

    def _generic(self, fn, other, out=None):
        """Element-wise fallback on the grid points only (keeps walls zero for ops with f(0,0) != 0)."""
        out = self._new_like() if out is None else out
        if out is not self:
            out._buf[self._lay.guard:].zero_()
        o = other._views() if isinstance(other, mesh) else None
        for c, (dst, src) in enumerate(zip(out._views(), self._views())):
            dst.copy_(fn(src, o[c] if o is not None else other))
        return out

    @staticmethod
    def _scalar(x):
        return isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, bool)

    def __add__(self, other):
        if self._same(other):
            return self._lin(1.0, 1.0, other)
        if self._scalar(other) or torch.is_tensor(other) or isinstance(other, np.ndarray):
            return self._generic(lambda a, b: a + b, self._coerce(other))
        return NotImplemented

    __radd__ = __add__

    def __sub__(self, other):
        if self._same(other):
            return self._lin(1.0, -1.0, other)
        if self._scalar(other) or torch.is_tensor(other) or isinstance(other, np.ndarray):
            return self._generic(lambda a, b: a - b, self._coerce(other))
        return NotImplemented

    def __rsub__(self, other):
        return self._generic(lambda a, b: b - a, self._coerce(other))

    def __mul__(self, other):
        if self._scalar(other):
            return self._lin(float(other), 0.0, None)
        if self._same(other) or torch.is_tensor(other) or isinstance(other, np.ndarray):
            return self._generic(lambda a, b: a * b, self._coerce(other))
        return NotImplemented

    __rmul__ = __mul__

    def __truediv__(self, other):
        if self._scalar(other):
            return self._generic(lambda a, b: a / b, float(other))
        return self._generic(lambda a, b: a / b, self._coerce(other))

    def __neg__(self):
        return self._lin(-1.0, 0.0, None)

    def __pos__(self):
        return self.copy()

    def __pow__(self, k):
        return self._generic(lambda a, b: a**b, k)

    def __iadd__(self, other):
        if self._same(other):
            return self._lin(1.0, 1.0, other, out=self)
        return self._generic(lambda a, b: a + b, self._coerce(other), out=self)

    def __isub__(self, other):
        if self._same(other):
            return self._lin(1.0, -1.0, other, out=self)
        return self._generic(lambda a, b: a - b, self._coerce(other), out=self)

    def __imul__(self, other):
        if self._scalar(other):
            return self._lin(float(other), 0.0, None, out=self)
        return self._generic(lambda a, b: a * b, self._coerce(other), out=self)

    def _coerce(self, other):
        if isinstance(other, np.ndarray):
            return torch.from_numpy(np.array(other, dtype=np.float64, order="C")).to(self._buf.device)
        return other

    def __abs__(self):
        """Global max-norm as a Python float (mesh.py:65-83)."""
        if self._lay.halo:  # slab fields: the halo planes between components are not part of the field
            local = max(get_backend().maxabs(self._comp_vol(c)) for c in range(self._ncomp))
        else:
            local = get_backend().maxabs(self.vol)
        comm = self.comm
        if comm is not None and getattr(comm, "size", 1) > 1:
            from .comm import MAX
            return float(comm.allreduce(local, op=MAX))
        return float(local)

    # ---- communication (mesh.py:85-125) -----------------------------------------------------------------------------
    def isend(self, dest=None, tag=None, comm=None):
        return comm.Issend(self, dest=dest, tag=tag)

    def irecv(self, source=None, tag=None, comm=None):
        self._touch()
        return comm.Irecv(self, source=source, tag=tag)

    def bcast(self, root=None, comm=None):
        self._touch()
        comm.Bcast(self, root=root)
        return self

    def __repr__(self):
        return f"{type(self).__name__}(shape={self.shape}, device={self._buf.device})"


class MultiComponentMesh(mesh):
    """Leading axis of ``len(components)`` fields with attribute access to writable component views (mesh.py:128-186)."""

    components = ()

    def __getattr__(self, name):
        comps = type(self).components
        if name in comps:
            c = comps.index(name)
            lay, g, st = self._lay, self._lay.guard, self._lay.stride
            # the component's storage view starts one guard before its volume: for c > 0 that region is the previous
            # component's (zero) wall region, exactly what the stencil kernels expect in front of a field
            view = mesh.__new__(mesh)
            mesh.__init__(view, None, _buf=self._buf[c * st: g + (c + 1) * st], _lay=lay, _ncomp=1)
            return view
        raise AttributeError(f"{type(self)!r} does not have attribute {name!r}!")


class imex_mesh(MultiComponentMesh):
    components = ("impl", "expl")


class comp2_mesh(MultiComponentMesh):
    components = ("comp1", "comp2")


# names used by the reference's CuPy twin (datatype_classes/cupy_mesh.py), kept as aliases
cuda_mesh = mesh
imex_cuda_mesh = imex_mesh
