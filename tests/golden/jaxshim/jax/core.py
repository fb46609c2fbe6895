"""jax.core stand-in.  Test infrastructure only."""
from ._core import to_dtype


class ShapedArray:
    def __init__(self, shape, dtype, *a, **k):
        self.shape = tuple(shape)
        self.dtype = to_dtype(dtype)
