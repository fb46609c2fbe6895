"""CPU tests of the host side of the slab decomposition (jax_sph_b200/slab.py):
ownership arithmetic and the ring transport, the latter over `gloo` with 2 and 3 ranks."""

import os
import socket

import numpy as np
import pytest

from jax_sph_b200 import slab


def test_slab_ranges_tile_the_axis():
    for layers in (8, 15, 170, 171, 1023):
        for p in (2, 3, 4, 8):
            edges = [slab.slab_range(layers, r, p) for r in range(p)]
            assert edges[0][0] == 0 and edges[-1][1] == layers
            assert all(edges[i][1] == edges[i + 1][0] for i in range(p - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_every_particle_has_exactly_one_owner():
    rng = np.random.default_rng(0)
    box, layers = 2 * np.pi, 170
    inv_cell = np.float32(layers / box)
    x = rng.uniform(0, box, 200001).astype(np.float32)
    x[:4] = [0.0, np.float32(box), np.nextafter(np.float32(box), np.float32(0)), 1e-30]
    for p in (2, 3, 8):
        seen = np.zeros(len(x), dtype=int)
        for r in range(p):
            rows = slab.own_rows(x, inv_cell, layers, r, p)
            seen[rows] += 1
            z0, z1 = slab.slab_range(layers, r, p)
            lay = slab.layer_of(x[rows], inv_cell, layers)
            assert ((lay >= z0) & (lay < z1)).all()
        assert (seen == 1).all()
    # the clamp of common.cuh cell_of: x == box lands in the last layer, not in layer `layers`
    assert slab.layer_of(np.float32(box), inv_cell, layers) == layers - 1


def test_layer_of_is_the_float32_product():
    # the device computes (int)(r * inv_cell) in float32; a float64 product can land elsewhere
    inv_cell = np.float32(170 / (2 * np.pi))
    x = np.float32(0.036959913)  # ~ layer boundary 1
    want = int(np.float32(x * inv_cell))
    assert slab.layer_of(x, inv_cell, 170) == want


def test_assemble_detects_lost_and_duplicated_particles():
    a = ({"rho": np.ones(3, np.float32)}, np.array([0, 1, 2]))
    b = ({"rho": np.ones(2, np.float32)}, np.array([3, 4]))
    out = slab.assemble([a, b], 5)
    assert out["rho"].shape == (5,)
    with pytest.raises(Exception):
        slab.assemble([a, b], 6)
    with pytest.raises(Exception):
        slab.assemble([a, a], 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _ring_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for nbytes in (16, 4096):
        send_lo = torch.full((nbytes,), 10 * rank + 1, dtype=torch.uint8)
        send_hi = torch.full((nbytes,), 10 * rank + 2, dtype=torch.uint8)
        recv_lo = torch.zeros(nbytes, dtype=torch.uint8)
        recv_hi = torch.zeros(nbytes, dtype=torch.uint8)
        for _ in range(3):  # repeated exchanges reuse the buffers, as a step does
            slab.ring_exchange(send_lo, send_hi, recv_lo, recv_hi, rank, world)
        lo, hi = slab.ring_neighbours(rank, world)
        ok = ok and bool((recv_lo == 10 * lo + 2).all()) and bool((recv_hi == 10 * hi + 1).all())
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ring_exchange_over_gloo(world):
    """recv_lo must hold the LOWER neighbour's upward message, recv_hi the UPPER neighbour's
    downward one -- including the two-rank ring where both neighbours are the same peer."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ring_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert all(res[r] for r in range(world)), res


@pytest.mark.parametrize("nranks", [2, 3, 4, 8])
def test_direct_ring_layout_pairs_senders_with_receivers(nranks):
    """The direct transport (slab.py, DirectRing): what rank r packs as its downward / upward
    message of exchange x is exactly the buffer its lower / upper ring neighbour unpacks as
    recv_hi / recv_lo after that exchange; the two sets alternate, nothing overlaps, and the
    agreement flags lie behind the buffers of every rank."""
    from jax_sph_b200.slab import direct_ring_layout, ring_neighbours

    msg = 1000  # not a multiple of 256: the layout rounds up
    bases = [0x10000000 * (r + 1) for r in range(nranks)]
    lay = [direct_ring_layout(bases, r, nranks, msg) for r in range(nranks)]
    mb = lay[0][0]
    assert mb % 256 == 0 and mb >= msg
    for r in range(nranks):
        lo, hi = ring_neighbours(r, nranks)
        _, recv, send, flags = lay[r]
        for x in range(4):
            s = x & 1
            # run_phase before exchange x packs into set s; the phase after it unpacks set s
            assert send[s][0] == lay[lo][1][s][1]  # my send_lo == lower rank's recv_hi
            assert send[s][1] == lay[hi][1][s][0]  # my send_hi == upper rank's recv_lo
        spans = sorted((a, a + mb) for s in (0, 1) for a in recv[s])
        assert all(spans[i][1] <= spans[i + 1][0] for i in range(3))
        assert flags[r] >= spans[-1][1] and flags == [b + 4 * mb for b in bases]
