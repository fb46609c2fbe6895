"""jax.ffi registration of the sphb200 handlers (csrc/ffi_shim.cc).

Importable only where jax (>= 0.4.38) and the compiled shim exist; the build
image has neither, so this module is exercised on a user's machine, not in the
test suite (INTEGRATION.md shows the reference-side change).
"""

import ctypes
import os

try:  # pragma: no cover - jax is absent in the build image
    import jax
    import jax.ffi
    import jax.numpy as jnp
    import numpy as np
except ImportError as exc:  # pragma: no cover
    raise ImportError("jax_sph_b200.jax_ffi needs jax >= 0.4.38 (not present here)") from exc

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "libsphb200_ffi.so")
STATE_ORDER = ("r", "tag", "u", "v", "dudt", "dvdt", "drhodt", "rho", "p", "mass", "eta", "dTdt",
               "T", "kappa", "Cp", "nw")  # solver.py:930-947


def register():  # pragma: no cover
    lib = ctypes.CDLL(_SHIM)
    for name in ("sphb200_ffi_advance", "sphb200_ffi_neighbors"):
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, name)), platform="CUDA")


def advance_fn(cfg):  # pragma: no cover
    """`advance(dt, state, neighbors)` as one XLA custom call (drop-in for the jitted
    closure of integrator.py:22-56)."""
    blob = bytes(cfg).hex()  # a str attribute: what the shim's Attr<std::string_view> decodes

    def advance(dt, state, neighbors):
        args = [state[k] for k in STATE_ORDER]
        outs = [jax.ShapeDtypeStruct(a.shape, a.dtype) for a in args]
        outs.append(jax.ShapeDtypeStruct((1,), jnp.uint32))
        res = jax.ffi.ffi_call("sphb200_ffi_advance", outs,
                               input_output_aliases={i: i for i in range(len(args))})(
            *args, config=blob, dt=float(dt))
        new_state = dict(zip(STATE_ORDER, res[:-1]))
        return new_state, neighbors

    return advance
