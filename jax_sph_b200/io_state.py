"""State writer / reader either side of the hot path.

Mirrors jax_sph/io_state.py:
* ``write_h5`` / ``read_h5``   <- io_state.py:30-35, :69-85   (same keys, same arrays)
* ``write_vtk``                <- io_state.py:38-41 + dict2pyvista :119-144
* ``write_state``              <- io_state.py:44-66           (same step / mode rules, same names)
* ``io_setup``                 <- io_state.py:14-27           (directory name rule)

and adds the piece the reference does not need because its state lives in host-visible
jax arrays: ``TrajectoryWriter``, which takes a snapshot of a resident ``Engine`` WITHOUT
stalling the step loop -- the engine un-permutes the cell-sorted frame into device buffers in
the original particle order (one gather pass on the compute stream), the device-to-host copy
into pinned double buffers runs on a side stream, and a background thread formats the file
once the copy's event has fired.  The step loop only waits if both buffers are still in
flight.

File formats.  ``h5py`` and ``pyvista`` are not part of this image: ``.h5`` files are written
with h5py when it is importable and as NumPy archives with the SAME dataset names otherwise
(the file keeps the ``.h5`` name the reference's tooling globs for; ``read_h5`` sniffs the
magic bytes and reads either); ``.vtk`` files are legacy-VTK binary POLYDATA written directly
(readable by ParaView / pyvista: points padded to 3D, 2-component vectors padded with a zero
column, exactly what dict2pyvista does).
"""

import os
import queue
import threading
import time
from typing import Dict, Optional, Sequence

import numpy as np

_HDF5_MAGIC = b"\x89HDF\r\n\x1a\n"


def _get(cfg, path):
    """cfg.a.b for OmegaConf / namespaces, cfg["a"]["b"] for plain dicts."""
    cur = cfg
    for name in path.split("."):
        cur = cur[name] if isinstance(cur, dict) else getattr(cur, name)
    return cur


def _h5py():
    try:
        import h5py  # noqa: PLC0415

        return h5py if hasattr(h5py, "File") and hasattr(h5py, "version") else None
    except ImportError:
        return None


def _host(v):
    if hasattr(v, "detach"):  # torch tensor
        v = v.detach().cpu().numpy()
    return np.asarray(v)


def _plain(x):
    """NumPy scalars / arrays -> Python numbers / lists (YAML dump of a config dict)."""
    if isinstance(x, dict):
        return {str(k): _plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_plain(v) for v in x]
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, np.generic):
        return x.item()
    return x


def io_setup(cfg, stamp: Optional[str] = None) -> str:
    """io_state.py:14-27: `<data_path>/<dim>D_<CASE>_<solver>_<seed>_<time>` for simulation runs
    that write, else `<data_path>/`; the directory is created.  (The reference also dumps
    config.yaml with OmegaConf; here the config is dumped as YAML when it is a plain dict.)"""
    d = str(_get(cfg, "io.data_path"))
    if not d.endswith("/"):
        d += "/"
    if len(_get(cfg, "io.write_type")) > 0 and _get(cfg, "case.mode") == "sim":
        d += str(_get(cfg, "case.dim")) + "D_" + str(_get(cfg, "case.name")).upper()
        d += "_" + str(_get(cfg, "solver.name")) + "_" + str(_get(cfg, "seed"))
        d += "_" + (stamp or time.strftime("%Y%m%d-%H%M%S"))
    os.makedirs(d, exist_ok=True)
    if isinstance(cfg, dict):
        import yaml

        with open(os.path.join(d, "config.yaml"), "w") as f:
            yaml.safe_dump(_plain(cfg), f)
    return d


def write_h5(data_dict: Dict, path: str):
    """io_state.py:30-35: one dataset per key."""
    h5 = _h5py()
    if h5 is not None:
        with h5.File(path, "w") as hf:
            for k, v in data_dict.items():
                hf.create_dataset(k, data=_host(v))
        return
    tmp = path + ".part"
    with open(tmp, "wb") as f:
        np.savez(f, **{k: _host(v) for k, v in data_dict.items()})
    os.replace(tmp, path)  # readers never see a half-written snapshot


def read_h5(file_name: str, array_type: str = "numpy"):
    """io_state.py:69-85.  array_type "numpy" or "torch" (CUDA tensors, what Engine.upload
    takes without a host round trip); the reference's "jax" is not offered."""
    if array_type not in ("numpy", "torch"):
        raise ValueError('array_type must be either "numpy" or "torch"')
    with open(file_name, "rb") as f:
        magic = f.read(8)
    if magic == _HDF5_MAGIC:
        h5 = _h5py()
        if h5 is None:
            raise RuntimeError(f"{file_name} is an HDF5 file and h5py is not installed")
        with h5.File(file_name, "r") as hf:
            data = {k: np.array(v) for k, v in hf.items()}
    else:
        with np.load(file_name) as z:
            data = {k: z[k] for k in z.files}
    if array_type == "torch":
        import torch

        data = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in data.items()}
    return data


def write_vtk(data_dict: Dict, path: str):
    """io_state.py:38-41 without pyvista: legacy VTK, binary (big-endian), POLYDATA with one
    vertex per particle; arrays as dict2pyvista (:119-144) lays them out."""
    r = _host(data_dict["r"]).astype(np.float32)
    n, dim = r.shape
    if dim == 2:
        r = np.hstack([r, np.zeros((n, 1), dtype=np.float32)])
    tmp = path + ".part"
    with open(tmp, "wb") as f:
        f.write(b"# vtk DataFile Version 3.0\njax_sph_b200 state\nBINARY\nDATASET POLYDATA\n")
        f.write(f"POINTS {n} float\n".encode())
        f.write(r.astype(">f4").tobytes())
        f.write(f"\nVERTICES {n} {2 * n}\n".encode())
        cells = np.empty((n, 2), dtype=">i4")
        cells[:, 0] = 1
        cells[:, 1] = np.arange(n)
        f.write(cells.tobytes())
        f.write(f"\nPOINT_DATA {n}\n".encode())
        fields = [(k, _host(v)) for k, v in data_dict.items() if k != "r"]
        f.write(f"FIELD FieldData {len(fields)}\n".encode())
        for k, v in fields:
            if dim == 2 and v.ndim == 2:
                v = np.hstack([v, np.zeros((n, 1), dtype=v.dtype)])
            comps = 1 if v.ndim == 1 else v.shape[1]
            if np.issubdtype(v.dtype, np.integer):
                f.write(f"{k} {comps} {n} int\n".encode())
                f.write(v.astype(">i4").tobytes())
            else:
                f.write(f"{k} {comps} {n} float\n".encode())
                f.write(v.astype(">f4").tobytes())
            f.write(b"\n")
    os.replace(tmp, path)


def read_vtk(path: str) -> Dict:
    """Reader for the files write_vtk makes (round-trip checks; ParaView reads them too)."""
    with open(path, "rb") as f:
        buf = f.read()
    pos = 0

    def line():
        nonlocal pos
        end = buf.index(b"\n", pos)
        out = buf[pos:end].decode()
        pos = end + 1
        return out

    assert line().startswith("# vtk DataFile") and line() is not None
    assert line() == "BINARY" and line() == "DATASET POLYDATA"
    n = int(line().split()[1])
    out = {"r": np.frombuffer(buf, dtype=">f4", count=3 * n, offset=pos).reshape(n, 3).astype(np.float32)}
    pos += 12 * n + 1
    assert line().split()[0] == "VERTICES"
    pos += 8 * n + 1
    assert line() == f"POINT_DATA {n}"
    nf = int(line().split()[2])
    for _ in range(nf):
        name, comps, rows, kind = line().split()
        comps, rows = int(comps), int(rows)
        dt = ">i4" if kind == "int" else ">f4"
        v = np.frombuffer(buf, dtype=dt, count=comps * rows, offset=pos)
        pos += 4 * comps * rows + 1
        v = v.astype(np.int32 if kind == "int" else np.float32)
        out[name] = v if comps == 1 else v.reshape(rows, comps)
    return out


def snapshot_name(step: int, cfg) -> Optional[str]:
    """File stem `write_state` uses for this step, or None when the step is not written
    (io_state.py:46-63)."""
    step_max = _get(cfg, "solver.sequence_length")
    mode = _get(cfg, "case.mode")
    write_normal = (mode == "sim") and ((step % _get(cfg, "io.write_every")) == 0) and (step >= 0)
    write_relax = (mode == "rlx") and (step == (step_max - 1))
    if not (write_normal or write_relax):
        return None
    if mode == "rlx":
        name = [_get(cfg, "case.name"), str(_get(cfg, "case.dim")), str(_get(cfg, "case.dx")),
                str(_get(cfg, "seed"))]
        return "_".join(str(x) for x in name)  # e.g. "tgv_3_0.02_42"
    return "traj_" + str(step).zfill(len(str(step_max)))


REFERENCE_KEYS = ("r", "tag", "u", "v", "dudt", "dvdt", "drhodt", "rho", "p", "mass", "eta", "dTdt", "T",
                  "kappa", "Cp", "nw")  # the state dict of case_setup.py:152-181 / solver.py:930-947


def complete_state(state: Dict, cfg=None) -> Dict:
    """The reference's snapshots always hold all sixteen state entries (read_h5 with state0_keys,
    visualisation and validation scripts index them); an engine that does not carry `nw`
    (no Riemann / free-slip walls) or `kappa` / `Cp` (no heat conduction) downloads a state
    without them.  Fill those in as the reference's own initial state has them: zeros for `nw`,
    the case constants (case.kappa_ref / case.Cp_ref when `cfg` has them, else zeros) for the
    heat entries.  Entries that are present are never touched; a dict that lacks any other
    reference key is not an engine download and is written as it is."""
    optional = ("nw", "kappa", "Cp")
    if any(k not in state for k in REFERENCE_KEYS if k not in optional) or all(k in state for k in optional):
        return state  # not an engine download (a partial dict is written as it is), or complete
    r = state["r"]
    n, dim = int(r.shape[0]), int(r.shape[1])

    def const(path):
        try:
            return float(_get(cfg, path)) if cfg is not None else 0.0
        except (KeyError, AttributeError, TypeError):
            return 0.0

    out = dict(state)
    fill = {"nw": np.zeros((n, dim), dtype=np.float32),
            "kappa": np.full(n, const("case.kappa_ref"), dtype=np.float32),
            "Cp": np.full(n, const("case.Cp_ref"), dtype=np.float32)}
    for k, v in fill.items():
        if k not in out:
            out[k] = v
    return out


def _write_files(state: Dict, dir: str, name: str, write_type: Sequence[str], cfg=None):
    state = complete_state(state, cfg)
    if "h5" in write_type:
        write_h5(state, os.path.join(dir, name + ".h5"))
    if "vtk" in write_type:
        write_vtk(state, os.path.join(dir, name + ".vtk"))


def write_state(step: int, state: Dict, dir: str, cfg):
    """io_state.py:44-66: write `state` (host or device arrays) if this step is a write step."""
    name = snapshot_name(step, cfg)
    if name is not None:
        _write_files(state, dir, name, _get(cfg, "io.write_type"), cfg)


class TrajectoryWriter:
    """`write_state` for a resident Engine, off the step loop's critical path.

        writer = TrajectoryWriter(engine, dir, cfg)
        for step in range(n):                      # jax_sph/simulate.py:113-134
            writer.write(step - 1)                 # no-op unless a write step
            engine.step(dt, 1)
        writer.close()                             # drains; re-raises a writer-thread error

    Per snapshot: Engine.download into device buffers (un-permute gather on the compute
    stream), async copy to one of `depth` pinned host buffers on a side stream, file writing
    in a background thread.  `keys` defaults to every state entry the engine holds.
    """

    def __init__(self, engine, dir: str, cfg, keys: Optional[Sequence[str]] = None, depth: int = 2):
        import torch

        self.torch = torch
        self.engine, self.dir, self.cfg = engine, dir, cfg
        self.write_type = list(_get(cfg, "io.write_type"))
        probe = engine.download(keys=keys)  # device buffers in the original order, reused
        self.dev = probe
        self.host = [{k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in probe.items()}
                     for _ in range(depth)]
        self.free = queue.Queue()
        for i in range(depth):
            self.free.put(i)
        self.jobs = queue.Queue()
        self.copy_stream = torch.cuda.Stream()
        self.error = None
        self.written = []
        self.thread = threading.Thread(target=self._drain, daemon=True)
        self.thread.start()

    def write(self, step: int) -> bool:
        name = snapshot_name(step, self.cfg)
        if name is None or not self.write_type:
            return False
        if self.error is not None:
            raise self.error
        torch = self.torch
        slot = self.free.get()  # waits only when every pinned buffer is still being written
        cur = torch.cuda.current_stream()
        # the previous snapshot's D2H copy must have read self.dev before it is overwritten
        cur.wait_stream(self.copy_stream)
        self.engine.download(out=self.dev)
        self.copy_stream.wait_stream(cur)
        with torch.cuda.stream(self.copy_stream):
            for k, v in self.dev.items():
                self.host[slot][k].copy_(v, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        self.jobs.put((slot, name, done))
        return True

    def _drain(self):
        while True:
            job = self.jobs.get()
            if job is None:
                return
            slot, name, done = job
            try:
                done.synchronize()
                state = {k: v.numpy() for k, v in self.host[slot].items()}
                _write_files(state, self.dir, name, self.write_type, self.cfg)
                self.written.append(name)
            except Exception as exc:  # surfaced on the next write() / close()
                self.error = exc
            finally:
                self.free.put(slot)

    def close(self):
        self.jobs.put(None)
        self.thread.join()
        if self.error is not None:
            raise self.error
