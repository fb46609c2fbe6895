"""Measured FP32 FMA peak of this GPU (sphb200_fp32_peak): the roofline denominator of the
interaction sweeps.  Prints one JSON line; run on the GPU box."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure():
    import torch

    from jax_sph_b200 import _lib

    lib = _lib.load()
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    out = {}
    for name, packed in (("ffma", 0), ("ffma2", 1)):
        tf, ms = C.c_double(), C.c_double()
        _lib.check(lib.sphb200_fp32_peak(packed, C.byref(tf), C.byref(ms), None))
        out[name] = {"tflops": tf.value, "ms": ms.value}
    out["nominal_tflops"] = 148 * 128 * 2 * 1.965e9 / 1e12
    return out


if __name__ == "__main__":
    print(json.dumps(measure()))
