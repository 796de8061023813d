"""The C-ABI library loads on a CPU-only box and exports every symbol include/walt_b200.h
declares; without a device it fails loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header="walt_b200.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(walt_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    import walt_b200
    L = walt_b200.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_every_header_has_its_library():
    """include/walt_host.h -> libwalthost.so (pure host code), include/walt_synth.h -> libwaltsynth.so (bench only)."""
    for header, lib, at_least in (("walt_host.h", "libwalthost.so", 25), ("walt_synth.h", "libwaltsynth.so", 5)):
        L = C.CDLL(os.path.join(ROOT, "walt_b200", "lib", lib))
        names = declared_symbols(header)
        assert len(names) >= at_least, (header, names)
        missing = [n for n in names if not hasattr(L, n)]
        assert not missing, (header, missing)
    # the product library exports no generator
    import walt_b200
    assert not [n for n in declared_symbols("walt_synth.h") if hasattr(walt_b200.load_library(), n)]


def test_no_cpu_fallback():
    import torch
    import walt_b200
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(walt_b200.WaltError) as ei:
        walt_b200.Engine(0)
    assert ei.value.code == 3
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    """Nothing under walt_b200/ may reference oracle/ or tests/."""
    for d, _, files in os.walk(os.path.join(ROOT, "walt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "walt_oracle" not in txt and "libwaltref" not in txt and "oracle/" not in txt, f
