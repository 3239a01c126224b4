"""The C-ABI library loads without a GPU and exports every symbol include/qlb200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "qlb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qlb200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_functions():
    names = declared_functions()
    assert "qlb200_match_create" in names and "qlb200_execute" in names and len(names) >= 35


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "tensortoolkit_b200", "libqlb200.so"))
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_header():
    from tensortoolkit_b200 import _lib
    assert sorted(_lib.SYMBOLS) == declared_functions()
    assert _lib.lib.qlb200_version().startswith(b"qlb200")


def test_no_cpu_fallback_without_device():
    """On a box without a usable sm_100 GPU the execution entry points must fail loudly."""
    import tensortoolkit_b200 as tk
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(tk._lib.QLB200Error):
        tk.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tensortoolkit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                for bad in ("import oracle", "from oracle", "refbridge", "contract_np", "libqlref", '#include "oracle'):
                    assert bad not in src, f"{f} uses the oracle ({bad})"
    for f in ("qlb200.h", os.path.join("qlten_b200", "contract.h")):
        assert "oracle/" not in open(os.path.join(ROOT, "include", f)).read()
