"""CPU: libhica.so loads and exports every symbol include/hica.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "hica.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hica_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    from hic_genome_assembler_b200 import _lib, build
    build.build()
    L = _lib.load_library()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), s
    assert sorted(_lib.ABI_SYMBOLS) == syms
    assert L.hica_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from hic_genome_assembler_b200._lib import Handle, HicaError
    with pytest.raises(HicaError):
        Handle(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "hic_genome_assembler_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "hica_oracle" not in src and "c_oracle" not in src and "ref_harness" not in src, f
