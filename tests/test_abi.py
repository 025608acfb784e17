"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/valle_b200.h declares; the ctypes table mirrors the header; failures are loud."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "valle_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from valle_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/valle_b200.h but not exported"
    assert set(_lib.PROTOTYPES) == set(syms), set(_lib.PROTOTYPES) ^ set(syms)


def test_abi_version_and_error_slot(lib):
    from valle_b200 import _lib
    assert lib.vb_abi_version() == _lib.ABI_VERSION
    assert isinstance(lib.vb_last_error(), bytes)
    assert lib.vb_launch_count() >= 0


def test_argument_errors_are_reported_not_thrown(lib):
    from valle_b200 import _lib
    # n_tables out of range -> VB_ERR_ARG with a message, no CUDA call made
    arr = (ctypes.c_void_p * 1)(0)
    st = lib.vb_embed_sum(0, 1, 0, arr, None, 9, 4, 256, 0, 256, 0, 0, 0, 0)
    assert st == 1
    assert b"n_tables" in lib.vb_last_error()
    with pytest.raises(_lib.VbError):
        _lib.check(st, "vb_embed_sum")


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA tensors / device."""
    import torch
    from conftest import build_model
    from valle_b200 import _lib
    m = build_model(dict(d_model=256, nhead=4, num_layers=2, prefix_mode=1, num_quantizers=8), 0)
    x = torch.randint(3, 100, (1, 8))
    y = torch.randint(0, 1024, (1, 20, 8))
    with pytest.raises(_lib.VbError):
        m.inference(x, torch.tensor([8], dtype=torch.int32), y, None, top_k=1)


def test_product_never_imports_oracle():
    """Nothing under valle_b200/ may import or call the oracle (test infrastructure)."""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "valle_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b|oracle\.", src, flags=re.M):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
