"""Periodic space, host mirror of jax_sph/jax_md/space.py:232-286.

``periodic(side)`` returns ``(displacement_fn, shift_fn)`` like the reference.
The engine never calls them -- the periodic fold is done inside the CUDA
kernels with the same float32 arithmetic -- but ``WCSPH`` and ``neighbor_list``
read the box from ``displacement_fn.side`` exactly as the reference closes over
``side`` (case_setup.py:133).  The callables themselves work on torch tensors
for API compatibility (e.g. user post-processing).
"""

import numpy as np


def periodic(side):
    side_np = np.asarray(side, dtype=np.float64).reshape(-1)

    def _side(t):
        import torch

        return torch.as_tensor(side_np, dtype=t.dtype, device=t.device)

    def displacement_fn(Ra, Rb, **kwargs):
        import torch

        s = _side(Ra)
        d = Ra - Rb
        return torch.remainder(d + s * 0.5, s) - 0.5 * s  # space.py:170-181

    def shift_fn(R, dR, **kwargs):
        import torch

        return torch.remainder(R + dR, _side(R))  # space.py:207-209

    displacement_fn.side = side_np
    shift_fn.side = side_np
    return displacement_fn, shift_fn
