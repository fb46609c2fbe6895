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
* ``integrator`` -> jax_sph/integrator.py:22-56 (incl. the per-step ``nw_fn``, :33-34)
* ``cases``      -> jax_sph/case_setup.py:83-231, jax_sph/utils.py:22-118,169-277,
                    cases/{tgv,db,pf,cf,ht}.py

Pinning status: PINNED against output of the reference itself.  The reference is
pure Python on top of jax/jaxlib (pinned 0.6.2-0.7.2 in its poetry.lock); neither is
installable in the build image, so its unmodified sources are imported from
/root/reference against a torch-backed stand-in for the jax API
(tests/golden/jaxshim/) and run by tests/golden/make_reference_golden.py: case
setup, neighbour list, WCSPH.forward and 20 advance() calls for 13 cases in float32
and float64 -> tests/golden/ref_*.npz.  tests/test_reference_pins.py checks the
oracle against them: neighbour sets bit for bit, states to 1e-9 / 1e-8 relative in
float64 and within the parity tolerance in float32.  In addition
(tests/test_oracle_pins.py) the reference's own tests:

* the four neighbour-list known-answer edge lists of tests/test_neighbors.py:89-121,
* the kernel half-integral / sign tests of tests/test_kernel.py:33-44,
* the Poiseuille and Couette analytical velocity profiles of
  tests/test_pf2d.py:106-115 and tests/test_cf2d.py:110-119 (atol 1e-2).

What the stand-in cannot pin: bit patterns that depend on XLA:CPU's own summation
order / fusion (tolerance-level), and ``jax.random`` (lattice noise is an input).

jax semantics that the restatement mimics on purpose: ``jnp.mod`` (sign of the
divisor), gather clamping / scatter dropping of the padding index N (padding
edges are stripped up front, which is equivalent), sequential scatter-add order
(edges sorted by sender, ``np.add.at``), ``x**5`` as ((x^2)^2)*x, weak-typed
Python scalars (all scalars are cast to the state dtype), and
``EPS = finfo(dtype).eps``.
"""

from . import cases, eos, integrator, kernel, partition, solver, space  # noqa: F401
