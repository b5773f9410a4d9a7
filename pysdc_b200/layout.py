"""The "walled" field layout shared by the datatypes and the kernels (see include/sdc_b200.h).

pitch P = n + (n & 1) in every dimension, volume P**ndim, preceded by a zero guard of P**(ndim-1) doubles (rounded up
to 16).  Odd n (Dirichlet-zero grids, generic_ND_FD.py:127-131) therefore carry one zero wall per dimension which *is*
the boundary value; even n (periodic grids) are dense.

``SlabLayout`` is the same layout for one slab of a 3-D grid decomposed along its slowest axis: the rank owns ``nz``
consecutive planes starting at global plane ``z0``; the guard plane in front doubles as the LOWER halo plane and one
extra plane behind the owned ones is the UPPER halo plane.
"""
import functools


class Layout:
    __slots__ = ("shape", "ndim", "n", "P", "vol", "guard", "padded_shape", "halo", "global_shape", "z0")

    def __init__(self, shape):
        if isinstance(shape, int):
            shape = (shape,)
        shape = tuple(int(s) for s in shape)
        if not 1 <= len(shape) <= 3:
            raise ValueError(f"fields must have 1 to 3 dimensions, got shape {shape}")
        if len(set(shape)) != 1:
            raise ValueError(f"need the same number of points in every dimension, got {shape}")
        self.shape = shape
        self.global_shape = shape
        self.ndim = len(shape)
        self.n = shape[0]
        self.P = self.n + (self.n & 1)
        self.vol = self.P**self.ndim          # doubles the streaming kernels see per component
        self.guard = (self.P ** (self.ndim - 1) + 15) // 16 * 16
        self.halo = 0                         # extra doubles behind the volume (upper halo plane of a slab)
        self.padded_shape = (self.P,) * self.ndim
        self.z0 = 0

    is_slab = False

    @property
    def stride(self):
        """Distance between the volumes of consecutive components of one buffer."""
        return self.vol + self.halo

    def alloc(self, ncomp=1):
        """Doubles to allocate for a field of ``ncomp`` components."""
        return self.guard + ncomp * self.stride

    def interior(self, vol_view):
        """Strided view of the grid points inside a flat volume view."""
        v = vol_view.view(self.padded_shape)
        return v[tuple(slice(0, s) for s in self.shape)]

    def __eq__(self, other):
        return type(other) is type(self) and other._key() == self._key()

    def _key(self):
        return self.shape

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        return f"Layout(shape={self.shape}, pitch={self.P}, guard={self.guard})"


class SlabLayout(Layout):
    """``nz`` planes (global planes ``z0 .. z0+nz-1``) of an ``n^3`` grid."""

    __slots__ = ("nz",)
    is_slab = True

    def __init__(self, n, nz, z0):
        n, nz, z0 = int(n), int(nz), int(z0)
        if nz < 1 or z0 < 0 or z0 + nz > n:
            raise ValueError(f"slab [{z0}, {z0 + nz}) does not fit an n={n} grid")
        self.shape = (nz, n, n)
        self.global_shape = (n, n, n)
        self.ndim = 3
        self.n = n
        self.nz = nz
        self.z0 = z0
        self.P = n + (n & 1)
        sz = self.P * self.P
        self.vol = sz * nz                    # owned planes only: what reductions and streaming kernels cover
        self.guard = (sz + 15) // 16 * 16     # >= one plane: the lower halo plane
        self.halo = sz                        # the upper halo plane
        self.padded_shape = (nz, self.P, self.P)

    def plane(self, buf_1comp, z):
        """1-D view of plane ``z`` (-1 = lower halo, nz = upper halo) of a single-component buffer (guard included)."""
        sz = self.P * self.P
        start = self.guard + z * sz
        return buf_1comp[start: start + sz]

    def _key(self):
        return (self.n, self.nz, self.z0)

    def __repr__(self):
        return f"SlabLayout(n={self.n}, planes=[{self.z0}, {self.z0 + self.nz}), pitch={self.P})"


@functools.lru_cache(maxsize=None)
def get_layout(shape):
    return Layout(shape)


@functools.lru_cache(maxsize=None)
def get_slab_layout(n, nz, z0):
    return SlabLayout(n, nz, z0)
