"""Slab-decomposed engine: one process per GPU, NCCL send/recv over NVLink.

The reference drives a single device (jax_sph/simulate.py:110-134); for the large
3D cases `north_star` asks for the periodic box to be cut into slabs across the
GPUs of one node, with a halo exchange every step and particle migration at every
neighbour rebuild.  (The reference rebuilds every step, simulate.py:75-86 passes no
dr_threshold; the engines re-sort and search only when a particle of ANY rank has moved
half the skin of the neighbour lists -- one 4-byte max-reduction per step decides --
and test every listed pair against the cutoff on every step, csrc/sweep.cuh.)

Device side: include/sphb200.h `sphb200_slab_*`, csrc/slab.cuh.  This module is
the transport and the host-side bookkeeping:

* `slab_range`, `layer_of`, `own_rows`   which rank owns which particle,
* `ring_exchange`                        the four messages of one exchange on the periodic
                                         ring (`torch.distributed` P2P: NCCL on GPUs, gloo in
                                         the CPU tests),
* `SlabEngine`                           upload / step / download / gather, mirroring `Engine`.

A step is a fixed sequence of phases; after each one the engine says how many
bytes of its send buffers the neighbours need (sizes depend only on capacities,
the real counts travel in the message headers), so the host never waits for the
device inside a step.
"""

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _lib
from .engine import STATE_KEYS, _stream_ptr, _torch

COUNT_NAMES = ("own", "immigrants", "emigrants_lo", "emigrants_hi", "halo_lo", "halo_hi",
               "sent_lo", "sent_hi")


def slab_range(layers: int, rank: int, nranks: int):
    """Global cell layers [z0, z1) of `rank` (same integer arithmetic as csrc/engine.cu)."""
    return layers * rank // nranks, layers * (rank + 1) // nranks


def layer_of(x, inv_cell: float, layers: int):
    """Cell layer of coordinate(s) x along the slab axis: the float32 product and the
    truncation of common.cuh `cell_of` (jax_md/partition.py:367), clamped to the grid."""
    x = np.asarray(x, dtype=np.float32)
    lay = (x * np.float32(inv_cell)).astype(np.int64)
    return np.clip(lay, 0, layers - 1)


def own_rows(x, inv_cell: float, layers: int, rank: int, nranks: int):
    """Indices of the particles whose layer lies in the slab of `rank`."""
    z0, z1 = slab_range(layers, rank, nranks)
    lay = layer_of(x, inv_cell, layers)
    return np.nonzero((lay >= z0) & (lay < z1))[0]


def ring_neighbours(rank: int, nranks: int):
    return (rank - 1) % nranks, (rank + 1) % nranks


def ring_exchange(send_lo, send_hi, recv_lo, recv_hi, rank: int, nranks: int, group=None):
    """send_lo -> rank-1, send_hi -> rank+1; recv_lo <- rank-1, recv_hi <- rank+1 (periodic).

    With two ranks both neighbours are the same peer, so the pairing relies on order (and, on
    backends that honour them, tags): what I send downwards is what the peer receives from above.
    """
    import torch.distributed as dist

    lo, hi = ring_neighbours(rank, nranks)
    ops = [dist.P2POp(dist.isend, send_lo, lo, group=group, tag=1),   # travelling downwards
           dist.P2POp(dist.isend, send_hi, hi, group=group, tag=0),   # travelling upwards
           dist.P2POp(dist.irecv, recv_hi, hi, group=group, tag=1),   # the upper rank's downward message
           dist.P2POp(dist.irecv, recv_lo, lo, group=group, tag=0)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()


SIGNAL_TIMEOUT_MS = 20000  # a neighbour that never answers traps the kernel instead of hanging it


def direct_ring_layout(ptrs, rank: int, nranks: int, msg_bytes: int):
    """Addresses of the direct transport.  Every rank's symmetric allocation (base `ptrs[r]` as
    mapped into THIS process) holds two sets of receive buffers, set s at `s * 2 * mb`:
    `[recv_lo | recv_hi]`, and the agreement flags behind them.  Returns (mb, recv, send, flags):
    recv[s] = (recv_lo, recv_hi) of this rank, send[s] = (send_lo, send_hi) = the lower
    neighbour's recv_hi and the upper neighbour's recv_lo of set s (my downward message is what
    the rank below receives from above), flags[r] = rank r's flag array."""
    mb = (int(msg_bytes) + 255) // 256 * 256
    lo, hi = ring_neighbours(rank, nranks)
    recv = [(ptrs[rank] + s * 2 * mb, ptrs[rank] + s * 2 * mb + mb) for s in (0, 1)]
    send = [(ptrs[lo] + s * 2 * mb + mb, ptrs[hi] + s * 2 * mb) for s in (0, 1)]
    flags = [p + 4 * mb for p in ptrs]
    return mb, recv, send, flags


class DirectRing:
    """The ring transport without collectives: every rank's receive buffers live in symmetric
    memory (torch.distributed._symmetric_memory: device memory mapped into all ranks of the node
    over NVLink / NVSwitch), the engine's pack kernels are handed the ring NEIGHBOURS' receive
    buffers as their send buffers (include/sphb200.h, sphb200_slab_set_agree) and store the live
    part of a message straight into them; an exchange is then two flags out and two flags in
    (sequence numbers in every rank's control block, one kernel: sphb200_slab_signal), no copy and
    no NCCL launch.

    Two sets of receive buffers alternate from exchange to exchange: a rank that signals
    exchange x has, in stream order, consumed exchange x - 1, so when its neighbour -- which
    waited for that signal -- stores exchange x + 1 into the same set, nobody reads it any more.
    The re-sort agreement is part of the same control block (words per step parity + arrival
    numbers): the phase kernels store and wait themselves.
    """

    def __init__(self, msg_bytes: int, rank: int, nranks: int, group=None):
        torch = _torch()
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.rank, self.nranks = rank, nranks
        self.group = group if group is not None else dist.group.WORLD
        self.mb = (int(msg_bytes) + 255) // 256 * 256
        total = 4 * self.mb + 4096
        dev = torch.device("cuda", torch.cuda.current_device())
        self.mem = symm_mem.empty(total, dtype=torch.uint8, device=dev)
        self.mem.zero_()
        self.hdl = symm_mem.rendezvous(self.mem, group=self.group)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.lo, self.hi = ring_neighbours(rank, nranks)
        _, self.recv, self.send, flags = direct_ring_layout(ptrs, rank, nranks, msg_bytes)
        self.flag_ptrs = (C.c_void_p * nranks)(*flags)
        self.x = 0  # exchanges so far
        torch.cuda.current_stream().synchronize()
        self.hdl.barrier(3, SIGNAL_TIMEOUT_MS)  # every rank's buffers are zeroed before anyone stores

    def buffers(self):
        """(send_lo, send_hi, recv_lo, recv_hi) device addresses for the next run_phase: it packs
        exchange x into the neighbours' set x % 2 and unpacks exchange x - 1 from the local other set."""
        s = self.x & 1
        return self.send[s] + self.recv[1 - s]

    def exchange(self):
        """The pack kernels of the last phase have stored exchange x into the neighbours' set
        x % 2; the engine's signal kernel (one launch: two flags out, two flags in) does the rest."""
        self.x += 1


class SlabEngine:
    """The slab of one rank, resident on the current CUDA device.  `transport`: "nccl"
    (batch_isend_irecv + all_reduce), "direct" (DirectRing: peer stores + flags), or "auto"
    (direct when the ranks can map each other's memory, else nccl; SPHB200_SLAB_TRANSPORT
    overrides)."""

    def __init__(self, cfg, rank: Optional[int] = None, nranks: Optional[int] = None, group=None,
                 own_cap: int = 0, halo_cap: int = 0, mig_cap: int = 0, transport: str = "auto"):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.Sphb200Error("no CUDA device: the engine has no CPU fallback")
        self.lib = _lib.load()
        self.cfg, self.group = cfg, group
        if rank is None or nranks is None:
            import torch.distributed as dist

            rank, nranks = dist.get_rank(group), dist.get_world_size(group)
        self.rank, self.nranks = int(rank), int(nranks)
        self.dim = int(cfg.dim)
        self._h = C.c_void_p()
        _lib.check(self.lib.sphb200_slab_create(C.byref(cfg), self.rank, self.nranks, int(own_cap),
                                                int(halo_cap), int(mig_cap), C.byref(self._h)))
        oi, od = (C.c_int64 * 16)(), (C.c_double * 4)()
        _lib.check(self.lib.sphb200_slab_info(self._h, C.byref(oi), C.byref(od)))
        self.axis, self.z0, self.z1, self.layers = int(oi[2]), int(oi[3]), int(oi[4]), int(oi[5])
        self.own_cap, self.halo_cap, self.mig_cap = int(oi[6]), int(oi[7]), int(oi[8])
        self.msg_bytes, self.arena_bytes = int(oi[9]), int(oi[12])
        self.inv_cell = float(od[0])
        self._buf = {k: torch.zeros(self.msg_bytes, dtype=torch.uint8, device="cuda")
                     for k in ("send_lo", "send_hi", "recv_lo", "recv_hi")}
        self.bytes_exchanged = 0
        self._keep = []
        self._pinned = {}
        self.time_exchanges = False  # CUDA events around every exchange (bench diagnostics)
        if getattr(cfg, "wall_layer", None):  # the integrator's nw_fn, see Engine.set_wall_layer
            from .engine import set_wall_layer

            set_wall_layer(self.lib, self._h, self.dim, **cfg.wall_layer)
        self._xev = []
        import os

        transport = os.environ.get("SPHB200_SLAB_TRANSPORT", transport)
        self.ring = None
        if transport not in ("nccl", "direct", "auto"):
            raise ValueError(f"transport {transport!r}")
        if transport == "auto":  # (a ring of engines inside one process has no transport at all)
            import torch.distributed as dist

            if not (dist.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl"
                    and dist.get_world_size(group) == self.nranks):
                transport = "nccl"
        if transport != "nccl" and self.nranks > 1:
            try:
                if self.nranks > 16:
                    raise RuntimeError("more than 16 ranks")
                self.ring = DirectRing(self.msg_bytes, self.rank, self.nranks, group)
                _lib.check(self.lib.sphb200_slab_set_agree(self._h, self.ring.flag_ptrs, self.nranks))
            except Exception as exc:  # no peer mapping here: the collectives still work
                if transport == "direct":
                    raise
                import warnings

                warnings.warn(f"sphb200 slab transport: symmetric memory unavailable ({exc}); using NCCL")
                self.ring = None
        self.transport = "direct" if self.ring is not None else "nccl"

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.sphb200_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- marshalling -----------------------------------------------------------
    def _state_struct(self, state: Dict, rows: int):
        torch = _torch()
        st = _lib.State()
        keep, on_host = [], None
        for k in STATE_KEYS + ("g_ext",):
            a = state.get(k)
            if a is None:
                continue
            if isinstance(a, np.ndarray):
                a = torch.from_numpy(np.ascontiguousarray(a))
            if a.dtype == torch.float64:
                raise _lib.Sphb200Error(f"state[{k!r}] is float64: float32 only (SPHB200_EDTYPE)")
            want = torch.int32 if k == "tag" else torch.float32
            if a.dtype != want or not a.is_contiguous():
                a = a.to(want).contiguous()
            n_expected = rows * (self.dim if k in _lib.VECTOR_FIELDS + ("g_ext",) else 1)
            if a.numel() != n_expected:
                raise _lib.Sphb200Error(f"state[{k!r}] has {a.numel()} elements, expected {n_expected}")
            host = not a.is_cuda
            if on_host is None:
                on_host = host
            elif on_host != host:
                raise _lib.Sphb200Error("state mixes host and device arrays")
            keep.append(a)
            setattr(st, k, a.data_ptr())
        return st, bool(on_host), keep

    # -- API -------------------------------------------------------------------
    def select_own(self, state: Dict):
        """(rows of the GLOBAL state this rank owns, their global indices)."""
        r = state["r"]
        r = r.numpy() if hasattr(r, "numpy") and not isinstance(r, np.ndarray) else np.asarray(r)
        ids = own_rows(r[:, self.axis], self.inv_cell, self.layers, self.rank, self.nranks)
        local = {}
        for k, v in state.items():
            if v is None:
                continue
            a = v.numpy() if hasattr(v, "numpy") and not isinstance(v, np.ndarray) else np.asarray(v)
            local[k] = np.ascontiguousarray(a[ids])
        return local, ids.astype(np.int32)

    def upload(self, state: Dict, ids=None):
        """`ids` None: `state` is the GLOBAL state (every rank passes the same arrays) and the rank
        keeps its own rows; else `state` holds this rank's particles and `ids` their global indices."""
        torch = _torch()
        if ids is None:
            state, ids = self.select_own(state)
        ids_t = torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int32)) if isinstance(ids, np.ndarray) else ids
        rows = int(ids_t.numel())
        if rows > self.own_cap:
            raise _lib.Sphb200Error(f"rank {self.rank}: {rows} own particles exceed own_cap {self.own_cap}")
        st, on_host, keep = self._state_struct(state, rows)
        if rows and ids_t.is_cuda == on_host:
            ids_t = ids_t.cpu() if on_host else ids_t.cuda()
        self._keep = keep + [ids_t]
        _lib.check(self.lib.sphb200_slab_upload(self._h, C.byref(st), C.c_void_p(ids_t.data_ptr()),
                                                rows, int(on_host), _stream_ptr()))
        if on_host:
            torch.cuda.current_stream().synchronize()

    def _exchange(self, nbytes: int, phase: int = 0):
        b = self._buf
        if self.time_exchanges:
            torch = _torch()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if self.ring is not None:
            _lib.check(self.lib.sphb200_slab_signal(self._h, _stream_ptr()))
            self.ring.exchange()
        else:
            ring_exchange(b["send_lo"][:nbytes], b["send_hi"][:nbytes], b["recv_lo"][:nbytes],
                          b["recv_hi"][:nbytes], self.rank, self.nranks, self.group)
        if self.time_exchanges:
            e1.record()
            self._xev.append((phase, nbytes, e0, e1))
        self.bytes_exchanged += 2 * nbytes

    def exchange_times(self):
        """{phase: (mean ms on the stream incl. waiting for the neighbour, bytes per message)}."""
        _torch().cuda.synchronize()
        acc = {}
        for phase, nb, e0, e1 in self._xev:
            a = acc.setdefault(phase, [0.0, 0, nb])
            a[0] += e0.elapsed_time(e1)
            a[1] += 1
        self._xev = []
        return {ph: (a[0] / a[1], a[2]) for ph, a in acc.items()}

    def run_phase(self, phase: int, dt: float, flags: int) -> int:
        """Enqueue one phase of the step; returns the bytes of the send buffers the ring
        neighbours need before the next phase (0: the step is complete)."""
        if self.ring is not None:
            ptrs = self.ring.buffers()
        else:
            b = self._buf
            ptrs = tuple(b[k].data_ptr() for k in ("send_lo", "send_hi", "recv_lo", "recv_hi"))
        nb = C.c_int64()
        _lib.check(self.lib.sphb200_slab_run(
            self._h, phase, float(dt), flags, C.c_void_p(ptrs[0]), C.c_void_p(ptrs[1]),
            C.c_void_p(ptrs[2]), C.c_void_p(ptrs[3]), _stream_ptr(), C.byref(nb)))
        return int(nb.value)

    def _agree(self):
        """The ranks' re-sort decisions (one int32 word each, written by phase 0 into the head of
        the send buffer) -> their maximum, in place: all ranks sort and search, or none does."""
        if self.ring is not None:  # (the agreement travels with the phase kernels' own flags)
            return
        import torch.distributed as dist

        word = self._buf["send_lo"][:4].view(_torch().int32)
        dist.all_reduce(word, op=dist.ReduceOp.MAX, group=self.group)

    def step(self, dt: float, nsteps: int = 1, integrate: bool = True, bc: bool = True):
        flags = step_flags(integrate, bc)
        for _ in range(nsteps):
            phase = 0
            while True:
                nb = self.run_phase(phase, dt, flags)
                if nb == 0:
                    break
                if nb < 0:
                    self._agree()
                else:
                    self._exchange(nb, phase)
                phase += 1

    def live_fields(self):
        """(read, written) of one advance() of this solver variant (engine.live_fields)."""
        from .engine import live_fields

        return live_fields(self.cfg)

    def advance_host(self, dt: float, local: Dict, ids):
        """One `advance(dt, state, neighbors)` on this rank's HOST-resident particles: copies
        host -> device only the entries advance() reads, steps (with the ring exchanges), and
        copies device -> host the entries it writes PLUS the ones it reads -- particles migrate
        between ranks during the step, so the constant entries (tag, mass, eta, ...) of the new
        particle set have to come back with it.  Returns (local state, ids) of the rank's
        particles after the step; entries advance() neither reads nor writes are dropped."""
        read, written = self.live_fields()
        self.upload({k: local[k] for k in read}, ids)
        self.step(dt, 1)
        keys = [k for k in STATE_KEYS if k in read or k in written]
        return self.download(keys)

    def counts(self) -> Dict[str, int]:
        out = (C.c_int32 * 8)()
        _lib.check(self.lib.sphb200_slab_counts(self._h, C.byref(out), _stream_ptr()))
        return dict(zip(COUNT_NAMES, [int(v) for v in out]))

    def download(self, keys=None):
        """This rank's own particles (host tensors, cell-sorted order) and their global indices."""
        torch = _torch()
        rows = self.counts()["own"]
        out = {}
        # pinned staging is allocated once at capacity (cudaHostAlloc is slow) and handed out as
        # views: the arrays of the previous download are overwritten by the next one
        pool = self._pinned
        for k in keys or STATE_KEYS:
            if k == "nw" and not (self.cfg.solver == 1 or self.cfg.flags & _lib.F_FREE_SLIP):
                continue
            if k in ("kappa", "Cp") and not self.cfg.flags & _lib.F_HEAT:
                continue
            width = self.dim if k in _lib.VECTOR_FIELDS else 1
            if k not in pool:
                pool[k] = torch.empty(self.own_cap * width, pin_memory=True,
                                      dtype=torch.int32 if k == "tag" else torch.float32)
            out[k] = pool[k][:rows * width].view((rows, self.dim) if width > 1 else (rows,))
        if "ids" not in pool:
            pool["ids"] = torch.empty(self.own_cap, dtype=torch.int32, pin_memory=True)
        ids = pool["ids"][:rows]
        st, on_host, _ = self._state_struct(out, rows)
        _lib.check(self.lib.sphb200_slab_download(self._h, C.byref(st), C.c_void_p(ids.data_ptr()),
                                                  rows, 1, _stream_ptr()))
        torch.cuda.current_stream().synchronize()
        return out, ids

    def gather(self, n_global: int, root: int = 0, keys=None):
        """Global state in the original particle order on `root` (None elsewhere)."""
        import torch.distributed as dist

        local, ids = self.download(keys)
        payload = ({k: v.numpy() for k, v in local.items()}, ids.numpy())
        parts = [None] * self.nranks if self.rank == root else None
        dist.gather_object(payload, parts, dst=root, group=self.group)
        if self.rank != root:
            return None
        return assemble(parts, n_global)

    def error(self, reduce: bool = True) -> int:
        """Device error word (read and cleared), OR-ed over the ranks when `reduce`."""
        torch = _torch()
        code = C.c_uint32()
        _lib.check(self.lib.sphb200_engine_error(self._h, C.byref(code), _stream_ptr()))
        if not reduce:
            return int(code.value)
        import torch.distributed as dist

        # NCCL has no bitwise OR: reduce the 32 bits separately with MAX
        bits = torch.tensor([(code.value >> i) & 1 for i in range(32)], dtype=torch.int32, device="cuda")
        dist.all_reduce(bits, op=dist.ReduceOp.MAX, group=self.group)
        return sum(int(v) << i for i, v in enumerate(bits.tolist()))

    def stats(self, reduce: bool = True):
        """Kinetic energy (sum) and max |u| (max) over the ranks (utils.py:128-166)."""
        torch = _torch()
        ek, um = C.c_double(), C.c_double()
        _lib.check(self.lib.sphb200_engine_stats(self._h, C.byref(ek), C.byref(um), _stream_ptr()))
        if not reduce:
            return ek.value, um.value
        import torch.distributed as dist

        a = torch.tensor([ek.value], dtype=torch.float64, device="cuda")
        b = torch.tensor([um.value], dtype=torch.float64, device="cuda")
        dist.all_reduce(a, op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(b, op=dist.ReduceOp.MAX, group=self.group)
        return float(a.item()), float(b.item())

    def launches(self) -> int:
        return int(self.lib.sphb200_engine_launches(self._h))

    def counters(self):
        """steps, searches among them, ... of this rank (Engine.counters)."""
        out = (C.c_int64 * 8)()
        _lib.check(self.lib.sphb200_engine_counters(self._h, C.byref(out), _stream_ptr()))
        v = list(out)
        return dict(steps=v[0], searches=v[1], list_rows=v[2], skin=v[3] * 1e-6, tiles=v[4],
                    tiles_without_lists=v[5])

    def profile(self, on: bool = True):
        _lib.check(self.lib.sphb200_engine_profile(self._h, int(on)))

    def last_times(self):
        ms = (C.c_float * 8)()
        _lib.check(self.lib.sphb200_engine_last_times(self._h, C.byref(ms)))
        return dict(cells=ms[1], density=ms[2], wall=ms[3], force=ms[4], total=ms[5])


def step_flags(integrate: bool = True, bc: bool = True) -> int:
    return (_lib.STEP_INTEGRATE if integrate else 0) | (_lib.STEP_BC if bc else 0)


def step_local_ring(engines, dt: float, nsteps: int = 1, integrate: bool = True, bc: bool = True):
    """Drive ALL ranks of a ring from one process on one device: the messages move by
    device-to-device copies instead of NCCL.  Same phases, same buffers, same kernels as the
    multi-process path -- used by the single-GPU tests of the decomposition."""
    flags = step_flags(integrate, bc)
    n = len(engines)
    for _ in range(nsteps):
        phase = 0
        while True:
            nbs = [e.run_phase(phase, dt, flags) for e in engines]
            assert len(set(nbs)) == 1, f"ranks disagree on the message size: {nbs}"
            nb = nbs[0]
            if nb == 0:
                break
            if nb < 0:  # the re-sort decision: maximum over the ranks (SlabEngine._agree)
                torch = _torch()
                words = [e._buf["send_lo"][:4].view(torch.int32) for e in engines]
                top = torch.stack(words).max(dim=0).values
                for w in words:
                    w.copy_(top)
                phase += 1
                continue
            for r, e in enumerate(engines):
                lo, hi = ring_neighbours(r, n)
                e._buf["recv_lo"][:nb].copy_(engines[lo]._buf["send_hi"][:nb])
                e._buf["recv_hi"][:nb].copy_(engines[hi]._buf["send_lo"][:nb])
                e.bytes_exchanged += 2 * nb
            phase += 1


def assemble(parts, n_global: int):
    """[(local dict, ids)] of all ranks -> global dict in original order; every particle must
    appear exactly once."""
    seen = np.zeros(n_global, dtype=np.int64)
    out = {}
    for local, ids in parts:
        np.add.at(seen, ids, 1)
        for k, v in local.items():
            if k not in out:
                out[k] = np.zeros((n_global,) + v.shape[1:], dtype=v.dtype)
            out[k][ids] = v
    if not (seen == 1).all():
        raise _lib.Sphb200Error(f"slab gather: {int((seen == 0).sum())} particles lost, "
                                f"{int((seen > 1).sum())} duplicated")
    return out
