"""h5py stand-in backed by numpy archives: File(path, "w").create_dataset(k, data=...) / close()
and File(path, "r").items(), which is all jax_sph/io_state.py:30-75 uses.  Test infrastructure
only (lets the reference's own tests/test_pf2d.py, test_cf2d.py run their simulate() + read-back
through the jax stand-in)."""
import numpy as np


class File:
    def __init__(self, path, mode="r"):
        self.path, self.mode, self.data = path, mode, {}
        if mode.startswith("r"):
            with np.load(path) as z:
                self.data = {k: z[k] for k in z.files}

    def create_dataset(self, name, data=None, **kw):
        self.data[name] = np.asarray(data)

    def items(self):
        return self.data.items()

    def keys(self):
        return self.data.keys()

    def __getitem__(self, k):
        return self.data[k]

    def close(self):
        if self.mode.startswith("w"):
            with open(self.path, "wb") as f:
                np.savez(f, **self.data)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
