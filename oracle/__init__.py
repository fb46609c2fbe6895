"""CPU oracle: a NumPy restatement of the JAX-SPH per-step particle hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product (``jax_sph_b200``) never does.

What it restates (all citations relative to the upstream reference tree):

* ``space``      -> jax_sph/jax_md/space.py:170-209,232-286
* ``partition``  -> jax_sph/jax_md/partition.py:114-199,243-428,818-983
* ``kernel``     -> jax_sph/kernel.py:51-103 (+ Cubic/WC4/WC6/Gaussian for the pins)
* ``eos``        -> jax_sph/eos.py:20-57
* ``solver``     -> jax_sph/solver.py:21-30,108-178,199-256,316-428,431-610,705-949
* ``integrator`` -> jax_sph/integrator.py:22-56
* ``cases``      -> jax_sph/case_setup.py:83-181, jax_sph/utils.py:22-118,169-194,
                    cases/{tgv,db,pf,cf,ht}.py

Pinning status: the reference is pure Python on top of jax/jaxlib (pinned
0.6.2-0.7.2 in its poetry.lock), and neither is installable in the build
image, so the reference cannot be executed.  The oracle is pinned by the
reference's own tests instead (tests/test_oracle_pins.py):

* the four neighbour-list known-answer edge lists of tests/test_neighbors.py:89-121,
* the kernel half-integral / sign tests of tests/test_kernel.py:33-44,
* the Poiseuille and Couette analytical velocity profiles of
  tests/test_pf2d.py:106-115 and tests/test_cf2d.py:110-119 (atol 1e-2).

Per-step rho / p / dudt values are NOT pinned by any stored vector in the
reference ("parity unpinned" for those); their authority is the
line-by-line restatement plus the pins above.

jax semantics that the restatement mimics on purpose: ``jnp.mod`` (sign of the
divisor), gather clamping / scatter dropping of the padding index N (padding
edges are stripped up front, which is equivalent), sequential scatter-add order
(edges sorted by sender, ``np.add.at``), ``x**5`` as ((x^2)^2)*x, weak-typed
Python scalars (all scalars are cast to the state dtype), and
``EPS = finfo(dtype).eps``.
"""

from . import cases, eos, integrator, kernel, partition, solver, space  # noqa: F401
