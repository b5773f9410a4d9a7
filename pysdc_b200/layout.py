"""The "walled" field layout shared by the datatypes and the kernels (see include/sdc_b200.h).

pitch P = n + (n & 1) in every dimension, volume P**ndim, preceded by a zero guard of P**(ndim-1) doubles (rounded up
to 16).  Odd n (Dirichlet-zero grids, generic_ND_FD.py:127-131) therefore carry one zero wall per dimension which *is*
the boundary value; even n (periodic grids) are dense.
"""
import functools


class Layout:
    __slots__ = ("shape", "ndim", "n", "P", "vol", "guard", "padded_shape")

    def __init__(self, shape):
        if isinstance(shape, int):
            shape = (shape,)
        shape = tuple(int(s) for s in shape)
        if not 1 <= len(shape) <= 3:
            raise ValueError(f"fields must have 1 to 3 dimensions, got shape {shape}")
        if len(set(shape)) != 1:
            raise ValueError(f"need the same number of points in every dimension, got {shape}")
        self.shape = shape
        self.ndim = len(shape)
        self.n = shape[0]
        self.P = self.n + (self.n & 1)
        self.vol = self.P**self.ndim
        self.guard = (self.P ** (self.ndim - 1) + 15) // 16 * 16
        self.padded_shape = (self.P,) * self.ndim

    def interior(self, vol_view):
        """Strided view of the n**ndim grid points inside a flat volume view."""
        v = vol_view.view(self.padded_shape)
        return v[tuple(slice(0, self.n) for _ in range(self.ndim))]

    def __eq__(self, other):
        return isinstance(other, Layout) and other.shape == self.shape

    def __hash__(self):
        return hash(self.shape)

    def __repr__(self):
        return f"Layout(shape={self.shape}, pitch={self.P}, guard={self.guard})"


@functools.lru_cache(maxsize=None)
def get_layout(shape):
    return Layout(shape)
