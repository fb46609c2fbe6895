"""GPU parity of the neighbour search: bit-exact neighbour SETS.

Reference behaviour: jax_sph/partition.py:492-571 + jax_md/partition.py:885-909
(Sparse list, row 0 receiver / row 1 sender, padding N, membership
d(r_sender, r_receiver)^2 < cutoff^2 in the position dtype).  Known answers:
reference tests/test_neighbors.py:89-121 (tests/golden/neighbors_kat.npz).
"""

import ctypes as C
import os

import numpy as np
import pytest

from _util import GOLDEN, canonical_pairs

pytestmark = pytest.mark.gpu


def _nl(r, box, cutoff, mask_self=False, **tuning):
    """Neighbour list through the mirror of the reference API."""
    import torch

    from jax_sph_b200 import partition, space

    disp, _ = space.periodic(np.asarray(box, dtype=np.float64))
    fns = partition.neighbor_list(disp, np.asarray(box, dtype=np.float64), cutoff,
                                  mask_self=mask_self)
    pos = torch.as_tensor(np.asarray(r, dtype=np.float32), device="cuda")
    nbrs = fns.allocate(pos)
    nbrs2 = nbrs.update(pos)
    return nbrs, nbrs2


@pytest.mark.parametrize("which,mask_self", [("1", False), ("1", True), ("2", False), ("2", True)])
def test_reference_known_answers(which, mask_self):
    z = np.load(os.path.join(GOLDEN, "neighbors_kat.npz"))
    r = z["r" + which]
    n = len(r)
    nbrs, nbrs2 = _nl(r, z["box"], float(z["cutoff"]), mask_self)
    assert not nbrs.did_buffer_overflow and not nbrs2.did_buffer_overflow  # test_neighbors.py:60-62
    idx, idx2 = nbrs.idx.cpu().numpy(), nbrs2.idx.cpu().numpy()
    assert (idx == idx2).all(), "allocate differs from update"  # :64
    assert ((idx[0] == n) == (idx[1] == n)).all(), "one sided edges"  # :66-68
    target = z[f"t{which}_{'mask' if mask_self else 'self'}"]
    got = canonical_pairs(idx, n)
    assert got.shape == target.shape and (got == target).all()  # :77-82
    assert (np.diff(idx[1][idx[1] < n]) >= 0).all(), "senders must ascend (solver.py:721)"


def _cloud(kind, n, dim, rng):
    box = np.array([1.0, 0.7, 0.5][:dim])
    if kind == "uniform":
        r = rng.random((n, dim)) * box
    elif kind == "clustered":  # ragged: most cells empty, a few crowded
        centers = rng.random((6, dim)) * box
        r = np.mod(centers[rng.integers(0, 6, n)] + 0.03 * rng.standard_normal((n, dim)), box)
    elif kind == "lattice":  # every pair at exactly 3 dx sits on the cutoff
        m = int(round(n ** (1.0 / dim)))
        box = np.ones(dim)
        g = (np.stack(np.meshgrid(*[np.arange(m)] * dim, indexing="ij"), -1).reshape(-1, dim) + 0.5) / m
        r = g
    elif kind == "faces":  # particles on the box faces / exact cell boundaries
        r = rng.integers(0, 24, (n, dim)) / 24.0 * box
    return r.astype(np.float32), box


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("kind", ["uniform", "clustered", "lattice", "faces"])
@pytest.mark.parametrize("sub", [1, 2])
def test_sets_are_bit_exact(kind, dim, sub):
    from jax_sph_b200 import Engine, make_config
    from oracle import partition as opart

    rng = np.random.default_rng(7)
    # lattice sizes are non-dyadic (1/50, 1/15) so that the 3 dx pairs round to either side
    n = {2: 3000, 3: 4096}[dim] if kind != "lattice" else {2: 2500, 3: 3375}[dim]
    r, box = _cloud(kind, n, dim, rng)
    n = len(r)
    h = 0.02 if dim == 2 else 0.03
    if kind == "lattice":
        h = 1.0 / round(n ** (1.0 / dim))
    cutoff = 3.0 * h
    ref = opart.neighbor_pairs(r, box, cutoff)
    eng = Engine(make_config(dim, box, h, 0.0, cell_sub=[sub] * 3), n)
    eng.upload({"r": r})
    idx, count = eng.neighbor_list(ref.shape[1] + 16)
    assert eng.error() == 0
    assert count == ref.shape[1], f"edge count {count} != {ref.shape[1]}"
    got = canonical_pairs(idx.cpu().numpy(), n)
    assert (got == ref).all()
    # tie band report: how many listed edges sit within 4 ulp of the cutoff
    band = opart.tie_band(r, box, cutoff, ref)
    print(f"{kind} {dim}D sub={sub}: {count} edges, tie band {band}")
    if kind == "lattice":
        assert band > 0  # the test really exercises the ties


def test_mask_self_and_overflow_flag():
    from jax_sph_b200 import Engine, _lib, make_config
    from oracle import partition as opart

    rng = np.random.default_rng(3)
    r = rng.random((2000, 2)).astype(np.float32)
    box = np.array([1.0, 1.0])
    ref = opart.neighbor_pairs(r, box, 0.06, mask_self=True)
    eng = Engine(make_config(2, box, 0.02, 0.0), len(r))
    eng.upload({"r": r})
    idx, count = eng.neighbor_list(ref.shape[1], mask_self=True)
    assert eng.error() == 0 and count == ref.shape[1]
    assert (canonical_pairs(idx.cpu().numpy(), len(r)) == ref).all()
    # too small a buffer: the overflow bit of PartitionErrorCode, the count is still reported
    idx, count = eng.neighbor_list(ref.shape[1] // 2, mask_self=True)
    assert count == ref.shape[1]
    assert eng.error() & _lib.ERR_NEIGHBOR_OVERFLOW


def test_stateless_c_abi_entry_point():
    """sphb200_neighbor_list with a caller-owned workspace (what the jax.ffi shim binds)."""
    import torch

    from jax_sph_b200 import _lib, make_config
    from oracle import partition as opart

    rng = np.random.default_rng(11)
    r = (rng.random((1500, 3)) * [1.0, 0.8, 0.6]).astype(np.float32)
    box = np.array([1.0, 0.8, 0.6])
    ref = opart.neighbor_pairs(r, box, 0.09)
    lib = _lib.load()
    cfg = make_config(3, box, 0.03, 0.0)
    nbytes = C.c_size_t()
    _lib.check(lib.sphb200_workspace_bytes(C.byref(cfg), len(r), C.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    rd = torch.as_tensor(r, device="cuda")
    cap = ref.shape[1] + 8
    idx = torch.empty((2, cap), dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.sphb200_neighbor_list(
        C.byref(cfg), len(r), C.c_void_p(rd.data_ptr()), C.c_void_p(idx.data_ptr()), cap, 0,
        C.c_void_p(cnt.data_ptr()), C.c_void_p(err.data_ptr()), C.c_void_p(ws.data_ptr()),
        nbytes.value, stream))
    torch.cuda.synchronize()
    assert int(err.item()) == 0 and int(cnt.item()) == ref.shape[1]
    assert (canonical_pairs(idx.cpu().numpy(), len(r)) == ref).all()
    # workspace too small -> loud error, no work
    rc = lib.sphb200_neighbor_list(
        C.byref(cfg), len(r), C.c_void_p(rd.data_ptr()), C.c_void_p(idx.data_ptr()), cap, 0,
        None, None, C.c_void_p(ws.data_ptr()), 1024, stream)
    assert rc == _lib.ENOMEM
