"""Communicators: the mpi4py-style surface the reference's datatypes / controllers use, on ``torch.distributed``.

The reference talks to ``mpi4py`` (``datatype_classes/mesh.py:65-125``: ``comm.allreduce(..., op=MPI.MAX)``,
``comm.Issend / Irecv / Bcast``; ``controller_classes/controller_MPI.py:218-305``) or, for CuPy fields, to
``helpers/NCCL_communicator.py`` which only wraps the collectives.  Here one process drives one GPU and the transport
is NCCL over NVLink (``torch.distributed``, backend ``nccl``; ``gloo`` with CPU tensors in the CPU test-suite):

* ``TorchComm``  — rank / size, ``allreduce`` of Python scalars or lists, ``allgather`` / ``bcast`` of Python objects,
  ``Issend / Irecv / Bcast`` of device fields with request objects (``Wait`` / ``Test``), ``barrier``.  This is what the
  PFASST mode uses to hand ``uend`` from time slice to time slice.
* ``SlabComm``   — a ``TorchComm`` that also describes a slab decomposition of 3-D grids along the slowest axis
  (no reference counterpart): partition, neighbours, halo-plane exchange with grouped NCCL send/recv, and
  the registry of peer-mapped solver workspaces (``backend.slab_cg_workspace``).  Passing it as ``comm=`` to the heat
  problem classes turns their fields into slabs.
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from .comm import LAND, LOR, MAX, MIN, SUM  # noqa: F401
from .layout import get_slab_layout

_OPS = {MAX: dist.ReduceOp.MAX, MIN: dist.ReduceOp.MIN, SUM: dist.ReduceOp.SUM}


class Request:
    """mpi4py-like request around torch.distributed work handles (``after`` runs once on completion: used to copy a
    host-staged receive back to the device when the transport cannot move device memory itself)."""

    def __init__(self, works=(), after=None, keep=None):
        self._works = [w for w in works if w is not None]
        self._after, self._keep = after, keep

    def _finish(self):
        self._works = []
        if self._after is not None:
            self._after()
            self._after = None
        self._keep = None

    def Wait(self):
        for w in self._works:
            w.wait()
        self._finish()
        return True

    wait = Wait

    def Test(self):
        if not self._works:
            return True
        if all(w.is_completed() for w in self._works):
            self._finish()
            return True
        return False

    def Cancel(self):
        self._works, self._after, self._keep = [], None, None


class LocalComm:
    """Communicator of one process (serial runs of the time-parallel controller: SDC / MLSDC)."""

    rank, size = 0, 1

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def first(self, count):
        return self

    def barrier(self):
        pass

    Barrier = barrier

    def allreduce(self, value, op=SUM):
        return value

    def allgather(self, obj):
        return [obj]

    def bcast(self, obj, root=0):
        return obj

    def Bcast(self, field, root=0):
        pass


class TorchComm:
    def __init__(self, group=None, device=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun / init_process_group)")
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        backend = dist.get_backend(group)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        self.device = torch.device(device)
        self._host_staged = backend != "nccl"  # gloo moves host memory only: device fields are staged through the host
        self._subgroups = {}

    def first(self, count):
        """Communicator of the first ``count`` ranks (``None`` on the others).  Collective over THIS communicator the
        first time a given count is requested (torch.distributed.new_group must be entered by every rank)."""
        if count == self.size:
            return self
        if self.size != dist.get_world_size():
            # torch.distributed.new_group is collective over the DEFAULT group: a communicator that covers only part of
            # the job (time x space splittings) cannot create further groups without the ranks outside it
            raise NotImplementedError("sub-communicators can only be created from a communicator that spans the job")
        if count not in self._subgroups:
            ranks = [self._global(r) for r in range(count)]
            g = dist.new_group(ranks=ranks)
            self._subgroups[count] = TorchComm(g, self.device) if self.rank < count else None
        return self._subgroups[count]

    # mpi4py spellings
    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def _global(self, r):
        return r if self.group is None else dist.get_global_rank(self.group, r)

    def barrier(self):
        dist.barrier(group=self.group)

    Barrier = barrier

    # ---- small host values --------------------------------------------------------------------------------------------
    def allreduce(self, value, op=SUM):
        """Python float / int / bool or a list of them -> the same, reduced over the communicator."""
        is_list = isinstance(value, (list, tuple, np.ndarray))
        vals = list(value) if is_list else [value]
        if op in (LAND, LOR):
            t = torch.tensor([1.0 if v else 0.0 for v in vals], dtype=torch.float64, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN if op == LAND else dist.ReduceOp.MAX, group=self.group)
            out = [bool(v) for v in t.tolist()]
        else:
            t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device=self.device)
            dist.all_reduce(t, op=_OPS[op], group=self.group)
            out = t.tolist()
        return out if is_list else out[0]

    def allreduce_device(self, t, op=MAX):
        """In-place all-reduce of a small device vector (residual norms): NCCL reduces it where it lives; a host-staged
        transport (gloo) takes the round trip through the host."""
        if self._host_staged and t.is_cuda:
            h = t.cpu()
            dist.all_reduce(h, op=_OPS[op], group=self.group)
            t.copy_(h)
        else:
            dist.all_reduce(t, op=_OPS[op], group=self.group)
        return t

    def allgather(self, obj):
        out = [None] * self.size
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def bcast(self, obj, root=0):
        box = [obj]
        dist.broadcast_object_list(box, src=self._global(root), group=self.group)
        return box[0]

    # ---- device fields (mesh.isend / irecv / bcast, mesh.py:85-125) ------------------------------------------------
    @staticmethod
    def _storage(field):
        return field._buf if hasattr(field, "_buf") else field

    # Messages between two ranks are matched in ORDER (NCCL / gloo semantics), not by tag.  That is sufficient while both
    # ends post their sends and receives in the same order (pySDC's controllers do: controller_MPI.py:218-305).  With
    # SDCB200_CHECK_TAGS=1 every field message is preceded by a one-element header carrying the tag and the receiver
    # raises on a mismatch - the multi-process tests run that way.
    _check_tags = os.environ.get("SDCB200_CHECK_TAGS", "0") == "1"

    def _tag_header(self, tag):
        dev = torch.device("cpu") if self._host_staged else self.device
        return torch.tensor([-1 if tag is None else int(tag)], dtype=torch.int64, device=dev)

    def Issend(self, field, dest=None, tag=None):
        t = self._storage(field)
        head, works = None, []
        if self._check_tags:
            head = self._tag_header(tag)
            works.append(dist.isend(head, self._global(dest), group=self.group))
        if self._host_staged and t.is_cuda:
            h = t.cpu()
            return Request(works + [dist.isend(h, self._global(dest), group=self.group)], keep=(h, head))
        return Request(works + [dist.isend(t, self._global(dest), group=self.group)], keep=head)

    Isend = Issend

    def Irecv(self, field, source=None, tag=None):
        t = self._storage(field)
        head, works, check = None, [], None
        if self._check_tags:
            head = self._tag_header(None)
            works.append(dist.irecv(head, self._global(source), group=self.group))

            def check():
                got = int(head.item())
                if tag is not None and got != -1 and got != int(tag):
                    raise RuntimeError(f"message order mismatch: expected tag {tag} from rank {source}, got {got}")
        if self._host_staged and t.is_cuda:
            h = torch.empty(t.shape, dtype=t.dtype)

            def after():
                if check is not None:
                    check()
                t.copy_(h)
            return Request(works + [dist.irecv(h, self._global(source), group=self.group)], after=after, keep=(h, head))
        return Request(works + [dist.irecv(t, self._global(source), group=self.group)], after=check, keep=head)

    def Send(self, field, dest=None, tag=None):
        self.Issend(field, dest, tag).Wait()

    def Recv(self, field, source=None, tag=None):
        self.Irecv(field, source, tag).Wait()

    def Bcast(self, field, root=0):
        t = self._storage(field)
        if self._host_staged and t.is_cuda:
            h = t.cpu()
            dist.broadcast(h, src=self._global(root), group=self.group)
            t.copy_(h)
        else:
            dist.broadcast(t, src=self._global(root), group=self.group)

    def Allgather_fields(self, field, fields):
        """fields[r] <- the ``field`` of rank r, for every rank: ONE collective moving each field once over NVLink (the
        node-parallel sweepers gather the right-hand sides of all collocation nodes with it)."""
        send = self._storage(field)
        recv = [self._storage(f) for f in fields]
        if self._host_staged and send.is_cuda:
            hs = [torch.empty(t.shape, dtype=t.dtype) for t in recv]
            dist.all_gather(hs, send.cpu(), group=self.group)
            for t, h in zip(recv, hs):
                t.copy_(h)
        else:
            dist.all_gather(recv, send.contiguous(), group=self.group)

    # ---- one Python scalar between neighbours (convergence status ring, check_convergence.py:146-157) ----------
    def send_scalar(self, value, dest):
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device)
        dist.send(t, self._global(dest), group=self.group)

    def recv_scalar(self, source):
        t = torch.zeros(1, dtype=torch.float64, device=self.device)
        dist.recv(t, self._global(source), group=self.group)
        return float(t.item())


def split_planes(n, size):
    """Planes per rank: as even as possible, thicker slabs first (511 over 8 -> 7 x 64 + 63)."""
    base, extra = divmod(int(n), int(size))
    return [base + (1 if r < extra else 0) for r in range(size)]


class SlabComm(TorchComm):
    """Slab decomposition of 3-D grids along axis 0 over the ranks of the communicator."""

    def __init__(self, group=None, device=None):
        super().__init__(group, device)
        self._work = {}

    def planes(self, n):
        counts = split_planes(n, self.size)
        if min(counts) < 1:
            raise ValueError(f"cannot decompose {n} planes over {self.size} ranks")
        return counts

    def slab_layout(self, shape):
        shape = (shape,) * 3 if isinstance(shape, int) else tuple(int(s) for s in shape)
        if len(shape) != 3 or len(set(shape)) != 1:
            raise ValueError(f"slab decomposition needs a cubic 3-D grid, got {shape}")
        counts = self.planes(shape[0])
        return get_slab_layout(shape[0], counts[self.rank], sum(counts[: self.rank]))

    def exchange_planes(self, quads, periodic=False):
        """``quads`` = [(top_owned, lower_halo, bottom_owned, upper_halo), ...] contiguous plane tensors: send the top
        owned plane up and the bottom owned plane down, receive the neighbours' planes into the halo planes."""
        if self.size == 1:
            if periodic:
                for top, lower, bottom, upper in quads:
                    lower.copy_(top)
                    upper.copy_(bottom)
            return
        lo = self.rank - 1 if self.rank > 0 else (self.size - 1 if periodic else None)
        hi = self.rank + 1 if self.rank + 1 < self.size else (0 if periodic else None)
        ops = []
        for top, lower, bottom, upper in quads:
            # order matters when lo == hi (two ranks, periodic): sends / receives to one peer are matched in order
            if hi is not None:
                ops.append(dist.P2POp(dist.isend, top, self._global(hi), group=self.group))
            if lo is not None:
                ops.append(dist.P2POp(dist.irecv, lower, self._global(lo), group=self.group))
                ops.append(dist.P2POp(dist.isend, bottom, self._global(lo), group=self.group))
            if hi is not None:
                ops.append(dist.P2POp(dist.irecv, upper, self._global(hi), group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def exchange_halos(self, fields, periodic=False):
        """Fill the halo planes of single-component slab fields with the neighbours' boundary planes (grouped NCCL
        send/recv).  At the ends of a non-periodic domain the halo planes keep their zeros (= the Dirichlet value)."""
        self.exchange_planes([(f._lay.plane(f._buf, f._lay.nz - 1), f._lay.plane(f._buf, -1), f._lay.plane(f._buf, 0),
                               f._lay.plane(f._buf, f._lay.nz)) for f in fields], periodic)


def cartesian_comms(n_outer, n_space, device=None):
    """Two-dimensional process grid over ALL ranks of the job (world = n_outer * n_space, rank = outer * n_space + space):
    returns ``(outer_comm, space_comm)`` of this rank - ``outer_comm`` (a ``TorchComm``) joins the ranks with the same
    space index (time slices of PFASST, or the collocation nodes of the node-parallel sweepers), ``space_comm`` (a
    ``SlabComm``) the ranks of one outer index, i.e. the slabs of one field.  torch.distributed creates groups
    collectively over the default group, so every rank builds every group here, in the same order (this is how a
    communicator that does not span the job gets its sub-communicators: up front, not by splitting it later)."""
    world = dist.get_world_size()
    if n_outer * n_space != world:
        raise ValueError(f"{n_outer} x {n_space} ranks requested, the job has {world}")
    rank = dist.get_rank()
    outer_comm = space_comm = None
    for o in range(n_outer):
        g = dist.new_group(ranks=[o * n_space + s for s in range(n_space)])
        if rank // n_space == o:
            space_comm = SlabComm(g, device)
    for s_ in range(n_space):
        g = dist.new_group(ranks=[o * n_space + s_ for o in range(n_outer)])
        if rank % n_space == s_:
            outer_comm = TorchComm(g, device)
    return outer_comm, space_comm
