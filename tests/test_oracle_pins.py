"""Pins the NumPy oracle against the reference's own tests (CPU only).

* neighbour-list known answers  <- reference tests/test_neighbors.py:86-121
* kernel half-integrals / signs <- reference tests/test_kernel.py:17-44
* Poiseuille / Couette profiles <- reference tests/test_pf2d.py:14-41,106-115 and
                                   tests/test_cf2d.py (same structure), atol 1e-2
"""

import numpy as np
import pytest

from oracle import cases, integrator, kernel, partition
from oracle.interp import interp_vel

# ---------------------------------------------------------------------------
# neighbour KATs (exact)
KAT = [
    dict(
        mask_self=False, cutoff=0.33, box=np.array([1.0, 1.0]),
        r=np.array([[0.1, 0.1], [0.1, 0.3], [0.1, 0.9], [0.6, 0.5]]),
        target=np.array([[0, 1, 2, 0, 1, 0, 2, 3], [0, 0, 0, 1, 1, 2, 2, 3]]),
    ),
    dict(
        mask_self=True, cutoff=0.33, box=np.array([1.0, 1.0]),
        r=np.array([[0.1, 0.1], [0.1, 0.3], [0.1, 0.9], [0.6, 0.5]]),
        target=np.array([[1, 2, 0, 0], [0, 0, 1, 2]]),
    ),
    dict(
        mask_self=False, cutoff=0.33, box=np.array([1.0, 1.0]),
        r=np.array([[0.5, 0.2], [0.2, 0.5], [0.5, 0.5], [0.8, 0.5], [0.5, 0.8]]),
        target=np.array(
            [[0, 2, 1, 2, 0, 1, 2, 3, 4, 2, 3, 2, 4], [0, 0, 1, 1, 2, 2, 2, 2, 2, 3, 3, 4, 4]]
        ),
    ),
    dict(
        mask_self=True, cutoff=0.33, box=np.array([1.0, 1.0]),
        r=np.array([[0.5, 0.2], [0.2, 0.5], [0.5, 0.5], [0.8, 0.5], [0.5, 0.8]]),
        target=np.array([[2, 2, 0, 1, 3, 4, 2, 2], [0, 1, 2, 2, 2, 2, 3, 4]]),
    ),
]


@pytest.mark.parametrize("kat", KAT)
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_neighbor_kat_reference_algorithm(kat, dtype):
    r = kat["r"].astype(dtype)
    N = len(r)
    nbrs = partition.neighbor_list_reference(r, kat["box"], kat["cutoff"], mask_self=kat["mask_self"])
    nbrs2 = partition.neighbor_list_reference(
        r, kat["box"], kat["cutoff"], mask_self=kat["mask_self"], neighbors=nbrs
    )
    assert not nbrs.did_buffer_overflow and not nbrs2.did_buffer_overflow
    assert (nbrs.idx == nbrs2.idx).all(), "allocate differs from update"
    assert ((nbrs.idx[0] == N) == (nbrs.idx[1] == N)).all(), "one sided edges"
    got = partition.canonical_pairs(nbrs.idx, N=N)
    assert got.shape == kat["target"].shape and (got == kat["target"]).all()


@pytest.mark.parametrize("kat", KAT)
def test_neighbor_kat_all_builders(kat):
    r = kat["r"].astype(np.float64)
    bf = partition.brute_force_pairs(r, kat["box"], kat["cutoff"], mask_self=kat["mask_self"])
    kd = partition.neighbor_pairs(r, kat["box"], kat["cutoff"], mask_self=kat["mask_self"])
    assert (bf == kat["target"]).all()
    assert (kd == kat["target"]).all()


@pytest.mark.parametrize("dim,n", [(2, 20), (3, 9)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_builders_agree_on_lattice_with_ties(dim, n, dtype):
    """A perfect lattice has many pairs at exactly 3*dx; all builders must agree."""
    dx = 1.0 / n
    setup = cases.make_case("tgv", dim=dim, dx=dx, dtype=dtype, box_override=[1.0] * dim)
    r = setup.state["r"]
    ref = partition.canonical_pairs(
        partition.neighbor_list_reference(r, setup.box_size, 3 * dx).idx, N=len(r)
    )
    bf = partition.brute_force_pairs(r, setup.box_size, 3 * dx)
    kd = partition.neighbor_pairs(r, setup.box_size, 3 * dx)
    assert ref.shape == bf.shape and (ref == bf).all()
    assert kd.shape == bf.shape and (kd == bf).all()


def test_overflow_bits():
    rng = np.random.default_rng(0)
    r = rng.random((200, 2))
    box = np.array([1.0, 1.0])
    nb = partition.neighbor_list_reference(r, box, 0.1)
    assert not nb.did_buffer_overflow
    r2 = r.copy()
    r2[:150] = 0.5 + 0.01 * rng.random((150, 2))  # clump -> cell and list overflow
    nb2 = partition.neighbor_list_reference(r2, box, 0.1, neighbors=nb)
    assert nb2.did_buffer_overflow
    assert nb2.idx.shape == nb.idx.shape


# ---------------------------------------------------------------------------
# kernels
@pytest.mark.parametrize(
    "name,dx_factor", [("CSK", 1), ("QSK", 1), ("WC2K", 1.3), ("WC4K", 1.3), ("WC6K", 1.3), ("GK", 1)]
)
def test_kernel_1d(name, dx_factor):
    N = 500
    for h in [1, 0.1, 0.01]:
        k = kernel.KERNELS[name](h=dx_factor * h, dim=1, dtype=np.float64)
        dx = k.cutoff / N
        x = np.linspace(dx / 2, k.cutoff + dx / 2, N)
        w = k.w(x)
        gw = k.grad_w(x)
        assert np.isclose(dx * np.sum(w), 0.5, atol=1e-2)
        assert (w >= 0).all()
        assert (gw <= 1e-12).all()


@pytest.mark.parametrize("name", ["QSK", "WC2K"])
@pytest.mark.parametrize("dim", [2, 3])
def test_kernel_grad_matches_finite_difference(name, dim):
    k = kernel.KERNELS[name](h=0.1, dim=dim, dtype=np.float64)
    x = np.linspace(1e-3, k.cutoff * 1.05, 400)
    e = 1e-7
    fd = (k.w(x + e) - k.w(x - e)) / (2 * e)
    assert np.allclose(k.grad_w(x), fd, rtol=1e-5, atol=1e-5 * abs(k.grad_w(x)).max())


@pytest.mark.parametrize("name", ["QSK", "WC2K"])
@pytest.mark.parametrize("dim", [2, 3])
def test_kernel_normalisation(name, dim):
    """Integral over the support is 1 (lattice quadrature)."""
    h = 1.0
    k = kernel.KERNELS[name](h=h, dim=dim, dtype=np.float64)
    n = 60
    ax = (np.arange(-n, n) + 0.5) * (k.cutoff / n)
    g = np.meshgrid(*([ax] * dim), indexing="ij")
    rr = np.sqrt(sum(x * x for x in g))
    # the reference's 3D quintic constant 3/(359 pi) integrates to 360/359 = 1.0028
    assert np.isclose(k.w(rr).sum() * (k.cutoff / n) ** dim, 1.0, atol=3e-3)


# ---------------------------------------------------------------------------
# Poiseuille / Couette analytical pins (full oracle: cases + integrator + solver)
def _u_series_pf(y, t_, n_max=10):
    eta, rho, u_max, d = 100.0, 1.0, 1.25, 1.0
    nu = eta / rho
    fx = -8 * nu * u_max / d**2
    res = fx / (2 * nu) * y * (y - d)
    for n in range(n_max):
        base = np.pi * (2 * n + 1) / d
        res = res + 4 * fx / (nu * base**3 * d) * np.sin(base * y) * np.exp(-(base**2) * nu * t_)
    return res


def _u_series_cf(y, t_, n_max=10):
    eta, rho, u_max, d = 100.0, 1.0, 1.25, 1.0
    nu = eta / rho
    res = u_max * y / d
    for n in range(1, n_max):
        base = np.pi * n / d
        res = res + 2 * u_max / (n * np.pi) * (-1) ** n * np.sin(base * y) * np.exp(-(base**2) * nu * t_)
    return res


@pytest.mark.slow
@pytest.mark.parametrize(
    "case,tvf,solver",
    [("pf", 0.0, "SPH"), ("pf", 1.0, "SPH"), ("pf", 0.0, "RIE"), ("cf", 0.0, "SPH"), ("cf", 0.0, "RIE")],
)
def test_channel_flow_matches_analytical_solution(case, tvf, solver):
    """reference tests/test_pf2d.py:53-62,106-115: dx=0.0333333, dt=2e-6, t_end=5e-3."""
    dx, dt = 0.0333333, 0.000002
    setup = cases.make_case(case, dim=2, dx=dx, dtype=np.float64, solver=solver, tvf=tvf, dt=dt)
    t_probe = [0.0005, 0.001, 0.005]
    probe_steps = {int(tp / dt): tp for tp in t_probe}
    y_axis = np.linspace(0, 1, 21)
    rs = 0.2 * np.ones([len(y_axis), 2])
    rs[:, 1] = y_axis + 3 * dx
    series = _u_series_pf if case == "pf" else _u_series_cf
    sols = {}

    def cb(step, state):
        # the reference writes frame k = state before step k (simulate.py:115)
        if (step + 1) in probe_steps:
            sols[probe_steps[step + 1]] = interp_vel(state, setup.box_size, dx, 2, rs)

    integrator.simulate(setup, max(probe_steps), fast_segment_sum=True, callback=cb)
    for tp in t_probe:
        assert np.allclose(sols[tp], series(y_axis, tp), atol=1e-2), (case, solver, tp)
