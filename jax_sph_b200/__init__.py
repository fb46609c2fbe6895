"""jax_sph_b200 -- B200-native engine for the JAX-SPH per-step particle hot path.

Host-side mirror of the reference interface for that path:

* ``partition.neighbor_list``  <- jax_sph/partition.py:492-571
* ``solver.WCSPH``             <- jax_sph/solver.py:613-951
* ``integrator.si_euler``      <- jax_sph/integrator.py:8-58
* ``engine.Engine``            <- the step loop of jax_sph/simulate.py:110-134
* ``slab.SlabEngine``          <- the same loop, slab-decomposed over the GPUs of one node
* ``case_setup``               <- lattice starts of jax_sph/utils.py:35-54 + case_setup.py:127-181, on device

All compute goes through the C ABI of ``libsphb200.so`` (include/sphb200.h,
hand-written sm_100a CUDA); there is no CPU or PyTorch fallback.
"""

from . import _lib, case_setup, eos, integrator, partition, solver, space  # noqa: F401
from .engine import Engine, config_from_setup, make_config  # noqa: F401
from .slab import SlabEngine  # noqa: F401

__all__ = ["Engine", "SlabEngine", "make_config", "config_from_setup"]
