"""Checkpoint / output of device fields.

The reference's output hooks (``pySDC/implementations/hooks/log_solution.py:207-282`` ``LogToFile``) ask the problem for
``getOutputFile(fileName)`` and pass every solution through ``processSolutionForOutput(u)`` before
``outfile.addField(time, field)`` (``pySDC/core/problem.py:84-94``).  The file is a ``FieldsIO`` ``Rectilinear`` binary
(``pySDC/helpers/fieldsIO.py:388-463``):

    int8  sID = 1, int8 dtype id = 0 (float64)
    int32 nVar, dim, gridSizes[dim];  float64 coords of every axis
    then per stored solution:  float64 time, float64 field[nVar * prod(gridSizes)]   (C order)

``RectilinearFile`` writes and reads exactly that layout, so files written here open with the reference's
``FieldsIO.fromFile`` (and the hook's resume path, which re-opens an existing file and appends) and vice versa.  What is
device-specific:

* ``stage(u)`` copies the field device -> pinned host memory on a side stream and returns immediately, so the copy
  overlaps the next time step; ``addField`` waits for the staged copy (double-buffered) only when it writes it;
* slab-decomposed fields (``parallel.SlabComm``): every rank stages its own planes and writes them at its own offset of
  the record (``os.pwrite``), rank 0 writes the header and the time stamps - the on-disk layout is that of the global
  field, no gather through one rank.
"""
import os

import numpy as np
import torch

SID_RECTILINEAR, DTYPE_F64 = 1, 0


class Staged:
    """A device field on its way to pinned host memory (side stream); ``array()`` waits and returns the host view."""

    def __init__(self, host, event, shape):
        self._host, self._event, self.shape = host, event, shape

    def array(self):
        if self._event is not None:
            self._event.synchronize()
            self._event = None
        return self._host.numpy().reshape(self.shape)

    # numpy protocol: the reference's FieldsIO calls np.asarray(field)
    def __array__(self, dtype=None, copy=None):
        a = self.array()
        return a if dtype is None else a.astype(dtype)

    @property
    def dtype(self):
        return np.dtype("float64")

    @property
    def size(self):
        return int(np.prod(self.shape))


class Stager:
    """Double-buffered pinned staging area + copy stream of one problem instance."""

    def __init__(self, nbuf=2):
        self._bufs, self._next, self._nbuf = [], 0, nbuf
        self._stream = None

    def stage(self, u):
        t = u.data  # strided device view of the grid values (interior of the walled layout)
        if not t.is_cuda:
            return Staged(t.detach().clone().contiguous(), None, tuple(t.shape))
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=t.device)
        if len(self._bufs) < self._nbuf:
            self._bufs.append(torch.empty(t.numel(), dtype=torch.float64).pin_memory())
        host = self._bufs[self._next % len(self._bufs)]
        self._next += 1
        if host.numel() != t.numel():
            host = self._bufs[(self._next - 1) % len(self._bufs)] = torch.empty(t.numel(), dtype=torch.float64).pin_memory()
        self._stream.wait_stream(torch.cuda.current_stream(t.device))
        with torch.cuda.stream(self._stream):
            host.view(t.shape).copy_(t, non_blocking=True)
            t.record_stream(self._stream)
            ev = torch.cuda.Event()
            ev.record(self._stream)
        return Staged(host, ev, tuple(t.shape))


class RectilinearFile:
    """``FieldsIO`` ``Rectilinear`` file (float64).  ``comm`` / ``slab`` = (z0, nz): this rank owns planes z0 .. z0+nz-1
    along axis 0 of every variable."""

    ALLOW_OVERWRITE = False

    def __init__(self, fileName, nVar, coords, comm=None, slab=None):
        self.fileName, self.nVar = fileName, int(nVar)
        self.coords = [np.asarray(c, dtype=np.float64) for c in (coords if isinstance(coords, (list, tuple)) else [coords])]
        self.gridSizes = [c.size for c in self.coords]
        self.nItems = self.nVar * int(np.prod(self.gridSizes))
        self.comm, self.slab = comm, slab
        self.rank = 0 if comm is None else comm.rank
        self._pending = None  # (time, Staged): a device field whose copy to the host is still in flight

    # ---- layout ---------------------------------------------------------------------------------------------------------
    @property
    def hSize(self):
        return 2 + 4 * (2 + len(self.gridSizes)) + 8 * sum(self.gridSizes)

    @property
    def recordSize(self):
        return 8 + 8 * self.nItems

    def initialize(self):
        if self.rank == 0:
            import sys

            ref_io = sys.modules.get("pySDC.helpers.fieldsIO")  # LogToFile sets the flag on the reference's class
            allow = self.ALLOW_OVERWRITE or (ref_io is not None and ref_io.FieldsIO.ALLOW_OVERWRITE)
            if os.path.isfile(self.fileName) and not allow:
                raise FileExistsError(f"file {self.fileName!r} already exists, use RectilinearFile.ALLOW_OVERWRITE = True "
                                      "to allow overwriting")
            with open(self.fileName, "w+b") as f:
                np.array([SID_RECTILINEAR, DTYPE_F64], dtype=np.int8).tofile(f)
                np.array([self.nVar, len(self.gridSizes), *self.gridSizes], dtype=np.int32).tofile(f)
                for c in self.coords:
                    c.tofile(f)
        if self.comm is not None:
            self.comm.barrier()
        return self

    @classmethod
    def fromFile(cls, fileName, comm=None, slab=None):
        with open(fileName, "rb") as f:
            sid, dt = np.fromfile(f, dtype=np.int8, count=2)
            if sid != SID_RECTILINEAR or dt != DTYPE_F64:
                raise ValueError(f"{fileName!r} is not a float64 Rectilinear FieldsIO file (sID {sid}, dtype id {dt})")
            nVar, dim = np.fromfile(f, dtype=np.int32, count=2)
            sizes = np.fromfile(f, dtype=np.int32, count=dim)
            coords = [np.fromfile(f, dtype=np.float64, count=n) for n in sizes]
        return cls(fileName, nVar, coords, comm=comm, slab=slab)

    @property
    def nFields(self):
        self.flush()
        return int((os.path.getsize(self.fileName) - self.hSize) // self.recordSize)

    @property
    def times(self):
        out = []
        with open(self.fileName, "rb") as f:
            for i in range(self.nFields):
                f.seek(self.hSize + i * self.recordSize)
                out.append(float(np.fromfile(f, dtype=np.float64, count=1)[0]))
        return out

    # ---- records --------------------------------------------------------------------------------------------------------
    def addField(self, time, field):
        """Append one solution.  ``field``: array-like of shape (nVar, *grid); on slab runs every rank passes its own
        planes and all ranks must call.  A ``Staged`` device field is not waited for here: it is written when the next
        solution arrives (or the file is read / flushed / closed), so its copy overlaps the following time step."""
        self.flush()
        if isinstance(field, Staged):
            self._pending = (float(time), field)
            return
        self._write(time, field)

    def flush(self):
        if self._pending is not None:
            (time, field), self._pending = self._pending, None
            self._write(time, field)

    close = flush

    def __del__(self):
        try:
            self.flush()
        except Exception:
            pass

    def _write(self, time, field):
        data = np.ascontiguousarray(np.asarray(field), dtype=np.float64)
        nfields = int((os.path.getsize(self.fileName) - self.hSize) // self.recordSize)
        idx = nfields if self.comm is None else self.comm.bcast(nfields if self.rank == 0 else None, root=0)
        base = self.hSize + idx * self.recordSize
        if self.slab is None:
            if data.size != self.nItems:
                raise ValueError(f"expected {self.nItems} values, got {data.size}")
            with open(self.fileName, "r+b") as f:
                f.seek(base)
                np.array(time, dtype=np.float64).tofile(f)
                data.tofile(f)
            return
        z0, nz = self.slab
        plane = int(np.prod(self.gridSizes[1:]))
        data = data.reshape(self.nVar, nz, plane)
        fd = os.open(self.fileName, os.O_RDWR)
        try:
            if self.rank == 0:
                os.pwrite(fd, np.array(time, dtype=np.float64).tobytes(), base)
            for v in range(self.nVar):
                os.pwrite(fd, data[v].tobytes(), base + 8 + 8 * ((v * self.gridSizes[0] + z0) * plane))
        finally:
            os.close(fd)
        self.comm.barrier()

    def readField(self, idx):
        n = self.nFields  # (flushes a pending device field first)
        idx = idx + n if idx < 0 else idx
        if not 0 <= idx < n:
            raise IndexError(f"cannot read index {idx} from {n} fields")
        with open(self.fileName, "rb") as f:
            f.seek(self.hSize + idx * self.recordSize)
            t = float(np.fromfile(f, dtype=np.float64, count=1)[0])
            field = np.fromfile(f, dtype=np.float64, count=self.nItems)
        return t, field.reshape(self.nVar, *self.gridSizes)


class OutputMixin:
    """``getOutputFile`` / ``processSolutionForOutput`` (core/problem.py:84-94) for the device problem classes."""

    def _output_coords(self):
        return [np.asarray(self.xvalues, dtype=np.float64)] * len(self.nvars)

    def getOutputFile(self, fileName):
        comm = getattr(self, "_comm", None)
        slab = (self._lay.z0, self._lay.nz) if comm is not None else None
        return RectilinearFile(fileName, nVar=1, coords=self._output_coords(), comm=comm, slab=slab).initialize()

    def setUpFieldsIO(self):
        """Nothing to set up: the slab decomposition travels with the file object (core/problem.py:78-82)."""

    def openOutputFile(self, fileName):
        """Re-open an existing file to continue a run (the resume path of LogToFile.pre_run)."""
        comm = getattr(self, "_comm", None)
        slab = (self._lay.z0, self._lay.nz) if comm is not None else None
        return RectilinearFile.fromFile(fileName, comm=comm, slab=slab)

    def processSolutionForOutput(self, u):
        """Start the device -> pinned-host copy of ``u`` on a side stream and return at once; the returned object turns
        into the (1, *grid) array when the writer touches it."""
        if not hasattr(self, "_stager"):
            self._stager = Stager()
        st = self._stager.stage(u)
        st.shape = (1,) + tuple(st.shape)
        return st
