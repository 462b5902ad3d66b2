"""The C-ABI library loads on a machine without a GPU and exports every symbol include/pg_b200.h declares (no compute
calls here); and creating a context without a device fails loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest

from plonk_gadgets_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    path = _lib.build()
    lib = C.CDLL(path)
    for name in _declared():
        assert hasattr(lib, name), name
    _lib.bind(lib)
    assert lib.pg_abi_version() == _lib.ABI_VERSION
    assert lib.pg_strerror(1) == b"NonExistingInverse"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import plonk_gadgets_b200 as pg
    with pytest.raises(pg.EngineError) as e:
        pg.StandardComposer(device=0)
    assert e.value.code == -5          # PG_ERR_NO_DEVICE


def test_rust_sys_crate_declares_the_same_symbols():
    """bindings/rust/plonk-gadgets-b200-sys (source only: no rustc here) must stay in step with the header."""
    src = open(os.path.join(ROOT, "bindings", "rust", "plonk-gadgets-b200-sys", "src", "lib.rs")).read()
    assert sorted(set(re.findall(r"pub fn (pg_[a-z0-9_]+)", src))) == _declared()
