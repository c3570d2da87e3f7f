"""CPU-side checks: the three implementations of the synthetic-input generator agree bit for
bit, and the C-ABI library loads and exports every symbol include/psd_b200.h declares (no
compute calls: there is no GPU in the build container)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import psd_rng

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_generators_agree(oracle, psd):
    n, p, B, first = 7, 3, 5, 11
    a = oracle.gen_real(1234, n, p, B, first)
    b = psd_rng.gen_uniform(1234, n, p, B, first)
    assert a.tobytes() == b.tobytes()
    c = np.empty_like(a)
    psd.capi.check(psd.lib().psd_fill_uniform_host(1234, n, p, B, first, 0, C.c_void_p(c.ctypes.data)))
    assert a.tobytes() == c.tobytes()
    z = oracle.gen_complex(1234, n, p, B, first)
    zz = np.empty((B, p, n, n), dtype=np.complex128)
    psd.capi.check(psd.lib().psd_fill_uniform_host(1234, n, p, B, first, 1, C.c_void_p(zz.ctypes.data)))
    assert z.tobytes() == zz.tobytes()
    assert np.array_equal(z.real, b) and np.array_equal(z.imag, psd_rng.gen_uniform(1234, n, p, B, first, 1))
    assert 0.0 <= a.min() and a.max() < 1.0


def _declared_symbols():
    names = []
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names += re.findall(r"\b(psd_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(psd):
    L = psd.lib()
    declared = _declared_symbols()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/ but not exported"
    assert sorted(psd.capi.EXPORTED_SYMBOLS) == declared
    assert psd.version() == 100


def test_no_cpu_fallback(psd):
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    if psd.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(psd.PsdError) as ei:
        psd.Handle()
    assert ei.value.code == -2
    with pytest.raises(psd.PsdError):
        psd.pschur([np.eye(3)], "R")


def test_product_does_not_touch_oracle():
    """The product path may not import, include, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "periodicschurdecompositions.jl_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if not fn.endswith((".py", ".cu", ".cuh", ".h", ".jl", "Makefile")):
                continue
            for line in open(os.path.join(dp, fn), errors="ignore"):
                ls = line.strip()
                if ls.startswith("#include"):
                    assert "oracle" not in ls and "psdo" not in ls, (fn, ls)
                if ls.startswith(("import ", "from ")):
                    assert "oracle" not in ls, (fn, ls)
                assert "libpsdo" not in ls, (fn, ls)
