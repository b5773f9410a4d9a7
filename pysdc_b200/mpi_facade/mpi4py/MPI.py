"""``mpi4py.MPI`` facade: constants, ``Intracomm`` and request objects backed by ``pysdc_b200.parallel.TorchComm``.

Messages between a pair of ranks are matched in order (NCCL / gloo point-to-point semantics), not by tag, which is
sufficient because both ends of pySDC's time-parallel controller issue their sends and receives in the same
deterministic stage order (controller_MPI.py:218-305); with SDCB200_CHECK_TAGS=1 the tags of the field messages travel
in a header and a mismatch raises (the multi-process tests run that way).  Numpy buffers (``[array, MPI.DOUBLE]``) travel as
small tensors, Python objects through the object collectives."""
import numpy as np
import torch
import torch.distributed as dist

from pysdc_b200 import comm as _ops
from pysdc_b200.parallel import Request, TorchComm

# reduction operations / datatypes (opaque tokens)
LAND, LOR, MAX, MIN, SUM = _ops.LAND, _ops.LOR, _ops.MAX, _ops.MIN, _ops.SUM
DOUBLE, INT, BOOL = "double", "int", "bool"
_RED = {MAX: dist.ReduceOp.MAX, MIN: dist.ReduceOp.MIN, SUM: dist.ReduceOp.SUM}
REQUEST_NULL = None
UNDEFINED = -32766


def _array(buf):
    return buf[0] if isinstance(buf, (list, tuple)) else buf


class Intracomm:
    def __init__(self, tc=None):
        self._tc = tc
        self._inflight = []  # requests the caller dropped without waiting: kept alive until they complete

    def _track(self, req):
        self._inflight = [r for r in self._inflight if not r.Test()]
        self._inflight.append(req)
        return req

    # the communicator is created on first use so that `MPI.COMM_WORLD` can exist before init_process_group
    @property
    def tc(self):
        if self._tc is None:
            self._tc = TorchComm()
        return self._tc

    @property
    def rank(self):
        return self.tc.rank

    @property
    def size(self):
        return self.tc.size

    def Get_rank(self):
        return self.tc.rank

    def Get_size(self):
        return self.tc.size

    def Barrier(self):
        self.tc.barrier()

    barrier = Barrier

    def Free(self):
        pass

    def Split(self, color=0, key=0):
        """Sub-communicator of the ranks that passed the same ``color`` (bool or int), ordered by rank."""
        if self.tc.size != dist.get_world_size():
            # torch.distributed.new_group is collective over the DEFAULT group: every rank of the job has to enter it, so
            # only a communicator that spans all of them can split (see parallel.TorchComm.first)
            raise NotImplementedError("Split needs a communicator that spans all ranks of the job")
        colors = self.tc.allgather(int(color))
        mine = None
        for c in sorted(set(colors)):  # every rank creates every group, in the same order
            members = [r for r, v in enumerate(colors) if v == c]
            g = dist.new_group(ranks=[self.tc._global(r) for r in members])
            if int(color) == c:
                mine = Intracomm(TorchComm(g, self.tc.device))
        return mine

    # ---- Python objects ---------------------------------------------------------------------------------------------
    def allgather(self, sendobj):
        return self.tc.allgather(sendobj)

    def bcast(self, obj=None, root=0):
        return self.tc.bcast(obj, root=root)

    def allreduce(self, sendobj, op=SUM):
        return self.tc.allreduce(sendobj, op=op)

    def send(self, obj, dest, tag=0):
        dist.send_object_list([obj], dst=self.tc._global(dest), group=self.tc.group)

    def isend(self, obj, dest, tag=0):
        self.send(obj, dest, tag)
        return Request()

    def recv(self, buf=None, source=0, tag=0):
        box = [None]
        dist.recv_object_list(box, src=self.tc._global(source), group=self.tc.group)
        return box[0]

    # ---- buffers: device fields (mesh) or [numpy array, datatype] -----------------------------------------------
    def _send_buffer(self, buf, dest):
        if hasattr(buf, "_buf") or torch.is_tensor(buf):
            return self._track(self.tc.Issend(buf, dest=dest))
        a = np.ascontiguousarray(_array(buf)).astype(np.float64).ravel()
        t = torch.from_numpy(a).to(self.tc.device)
        return self._track(Request([dist.isend(t, self.tc._global(dest), group=self.tc.group)], keep=t))

    def _recv_buffer(self, buf, source):
        if hasattr(buf, "_buf") or torch.is_tensor(buf):
            return self.tc.Irecv(buf, source=source)
        a = _array(buf)
        t = torch.zeros(a.size, dtype=torch.float64, device=self.tc.device)

        def fill():
            a[...] = t.cpu().numpy().reshape(a.shape).astype(a.dtype)

        return Request([dist.irecv(t, self.tc._global(source), group=self.tc.group)], after=fill, keep=t)

    def Issend(self, buf, dest=0, tag=0):
        if hasattr(buf, "_buf") or torch.is_tensor(buf):
            return self._track(self.tc.Issend(buf, dest=dest, tag=tag))
        return self._send_buffer(buf, dest)

    Isend = Issend

    def Send(self, buf, dest=0, tag=0):
        self._send_buffer(buf, dest).Wait()

    def Irecv(self, buf, source=0, tag=0):
        if hasattr(buf, "_buf") or torch.is_tensor(buf):
            return self.tc.Irecv(buf, source=source, tag=tag)
        return self._recv_buffer(buf, source)

    def Recv(self, buf, source=0, tag=0):
        self._recv_buffer(buf, source).Wait()

    def Bcast(self, buf, root=0):
        if hasattr(buf, "_buf") or torch.is_tensor(buf):
            self.tc.Bcast(buf, root=root)
            return
        a = _array(buf)
        a[...] = np.asarray(self.tc.bcast(np.array(a), root=root)).reshape(a.shape)

    def Ibcast(self, buf, root=0):
        self.Bcast(buf, root=root)
        return Request()

    # ---- buffer reductions (the node-parallel sweepers: generic_implicit_MPI.py:176-196,241-267) -------------------
    def _reduce_buffer(self, sendbuf, recvbuf, op, root):
        """root=None: all ranks receive.  Device fields are reduced where they live (NCCL); numpy buffers as tensors."""
        tc = self.tc
        if hasattr(sendbuf, "_buf") or torch.is_tensor(sendbuf):
            src = tc._storage(sendbuf)
            t = src.cpu() if (tc._host_staged and src.is_cuda) else src.clone()
        else:
            a = np.ascontiguousarray(_array(sendbuf), dtype=np.float64)
            t = torch.from_numpy(a.copy().ravel()).to(tc.device)
        if root is None:
            dist.all_reduce(t, op=_RED[op], group=tc.group)
        else:
            dist.reduce(t, dst=tc._global(root), op=_RED[op], group=tc.group)
        if recvbuf is None or (root is not None and tc.rank != root):
            return
        if hasattr(recvbuf, "_buf") or torch.is_tensor(recvbuf):
            tc._storage(recvbuf).copy_(t)
            if hasattr(recvbuf, "_touch"):
                recvbuf._touch()
        else:
            out = _array(recvbuf)
            out[...] = t.cpu().numpy().reshape(out.shape)

    def Reduce(self, sendbuf, recvbuf, op=SUM, root=0):
        self._reduce_buffer(sendbuf, recvbuf, op, root)

    def Allreduce(self, sendbuf, recvbuf, op=SUM):
        self._reduce_buffer(sendbuf, recvbuf, op, None)


Comm = Intracomm
COMM_WORLD = Intracomm()
COMM_SELF = None
