"""The C-ABI library loads on a CPU-only box and exports every symbol include/fsb.h declares
(no compute calls here: there is no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fsb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fsb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built):
    lib = ctypes.CDLL(os.path.join(ROOT, "fluid_simulation_b200", "lib", "libfsb.so"))
    names = declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fsb.h but not exported"


def test_binding_covers_header(capi):
    assert sorted(capi.SIGNATURES) == declared_symbols()
    assert "sm_100a" in capi.version()


def test_no_cpu_fallback(capi):
    """Without a usable sm_100 device the product must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback|CUDA"):
        capi.Sim(16, 16)


def test_product_does_not_reference_oracle():
    """Nothing under the package or include/ may import, include or link the oracle."""
    bad = []
    for base in ("fluid_simulation_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    src = open(os.path.join(dp, f), errors="ignore").read()
                    for line in src.splitlines():
                        code = line.split("//")[0]
                        if re.search(r"#include.*oracle|import\s+oracle|from\s+oracle|libfsoracle|libfsref", code):
                            bad.append((f, line.strip()))
    assert not bad, bad
