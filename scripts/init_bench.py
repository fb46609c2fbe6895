"""Time the on-device lattice initialisation (csrc/init.cuh) against the NumPy generator +
host-to-device upload it replaces.  python scripts/init_bench.py [nx]  (tgv3d, default 256)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from jax_sph_b200 import case_setup  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
meta = bench.lattice_meta("tgv3d", nx)
lat = case_setup.lattice_spec(meta["box"], meta["dx"], velocity="tgv3d", eta=meta["viscosity"])
rows = case_setup.lattice_rows(lat)
for _ in range(3):
    st = case_setup.init_lattice(lat)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    st = case_setup.init_lattice(lat)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
nbytes = rows * (6 * 3 * 4 + 9 * 4 + 4)
t0 = time.time()
host, _ = bench.lattice_state("tgv3d", nx)
t_gen = time.time() - t0
t0 = time.time()
dev = {k: torch.from_numpy(v).cuda() for k, v in host.items()}
torch.cuda.synchronize()
t_up = time.time() - t0
same_r = bool(torch.equal(dev["r"], st["r"]))
print(json.dumps({"workload": f"tgv3d nx={nx}", "rows": rows, "device_ms": ms,
                  "device_write_GBps": nbytes / ms / 1e6, "bytes": nbytes,
                  "host_generate_s": t_gen, "host_upload_s": t_up, "r_bit_exact": same_r}))
