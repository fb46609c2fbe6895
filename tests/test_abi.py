"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every
symbol include/sphb200.h declares, the ctypes structs match the header, the
host-only entry points validate their arguments, and compute entry points fail
loudly without a GPU (no fallback)."""

import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sphb200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    from jax_sph_b200 import _lib

    return _lib


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sphb200_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(lib):
    names = _declared()
    assert len(names) >= 20
    so = C.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(so, n), f"libsphb200.so does not export {n}"
        assert n in lib.SYMBOLS, f"_lib.py does not bind {n}"
    assert sorted(lib.SYMBOLS) == names, "binding table and header disagree"
    dyn = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True,
                         text=True).stdout
    exported = set(re.findall(r" T (sphb200_\w+)", dyn))
    assert set(names) <= exported


def test_struct_layout_matches_header(lib):
    cfg = lib.default_config()  # asserts sizeof(Config) == struct_size written by the C side
    assert cfg.dim == 3 and cfg.gamma == 1.0 and cfg.p_ref == 100.0
    # compile a one-liner against the header to cross-check the sizes with a C compiler
    code = ('#include <stdio.h>\n#include "sphb200.h"\nint main(){printf("%zu %zu %zu %zu",'
            "sizeof(sphb200_config),sizeof(sphb200_state),sizeof(sphb200_bc_rule),"
            "sizeof(sphb200_lattice));return 0;}")
    exe = os.path.join(ROOT, "build", "abi_sizes")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe],
                   input=code, text=True, check=True)
    sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    assert sizes == [C.sizeof(lib.Config), C.sizeof(lib.State), C.sizeof(lib.BcRule),
                     C.sizeof(lib.Lattice)]


def test_argument_validation_is_host_side(lib):
    so = lib.load()
    cfg = lib.default_config()
    nbytes = C.c_size_t()
    assert so.sphb200_engine_bytes(C.byref(cfg), 1000, C.byref(nbytes)) == 0 and nbytes.value > 0
    small = nbytes.value
    assert so.sphb200_engine_bytes(C.byref(cfg), 2000, C.byref(nbytes)) == 0 and nbytes.value > small
    assert so.sphb200_engine_bytes(C.byref(cfg), 0, C.byref(nbytes)) == lib.EINVAL
    cfg.dim = 4
    assert so.sphb200_engine_bytes(C.byref(cfg), 10, C.byref(nbytes)) == lib.EINVAL
    cfg = lib.default_config()
    cfg.solver = 2  # DELTA
    assert so.sphb200_engine_bytes(C.byref(cfg), 10, C.byref(nbytes)) == 0
    plain = nbytes.value
    cfg.flags |= lib.F_RHO_EVOL  # density diffusion keeps L matrices and gradient terms
    assert so.sphb200_engine_bytes(C.byref(cfg), 10, C.byref(nbytes)) == 0 and nbytes.value > plain
    cfg = lib.default_config()
    cfg.solver = 3
    assert so.sphb200_engine_bytes(C.byref(cfg), 10, C.byref(nbytes)) == lib.EUNSUP
    cfg = lib.default_config()
    cfg.struct_size = 8
    assert so.sphb200_engine_bytes(C.byref(cfg), 10, C.byref(nbytes)) == lib.EINVAL
    assert b"unsupported" in so.sphb200_strerror(lib.EUNSUP)
    assert so.sphb200_abi_version() == lib.ABI_VERSION


def test_no_gpu_means_loud_failure(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    so = lib.load()
    cfg = lib.default_config()
    h = C.c_void_p()
    assert so.sphb200_engine_create(C.byref(cfg), 100, C.byref(h)) == lib.ENODEV
    from jax_sph_b200 import Engine

    with pytest.raises(lib.Sphb200Error, match="no CPU fallback"):
        Engine(cfg, 100)


def test_product_never_imports_the_oracle():
    code = ("import sys; import jax_sph_b200; import jax_sph_b200.partition, jax_sph_b200.solver, "
            "jax_sph_b200.integrator; assert not any(m == 'oracle' or m.startswith('oracle.') "
            "for m in sys.modules), 'product imported the oracle'")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "jax_sph_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
