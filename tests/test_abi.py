"""CPU: the C-ABI library builds, loads and exports every symbol include/bmb200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header="bmb200.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bmb200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    import bandedmatrices_b200 as bm

    lib = ctypes.CDLL(bm.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/bmb200.h but not exported by libbmb200.so"
    assert set(names) == set(bm.PROTOTYPES), "ctypes prototypes drifted from include/bmb200.h"
    assert bm.load().bmb200_version() == 100
    from bandedmatrices_b200 import _lib

    internal = _declared("bmb200_internal.h")
    for name in internal:
        assert hasattr(lib, name), f"{name} declared in include/bmb200_internal.h but not exported"
    assert set(internal) == set(_lib.INTERNAL_PROTOTYPES)


def test_no_cpu_fallback_without_gpu():
    import pytest
    import torch

    import bandedmatrices_b200 as bm

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(bm.BMB200Error):
        bm.Handle(0)
    with pytest.raises(RuntimeError):
        bm.brand(10, 10, 1, 1)
    A = bm.BandedMatrix(torch.zeros((4, 3), dtype=torch.float64), 4, 1, 1)
    with pytest.raises(TypeError):  # CPU tensors are refused, not silently computed
        bm.mul_(torch.zeros(4, dtype=torch.float64), A, torch.zeros(4, dtype=torch.float64))


def test_product_never_imports_oracle():
    """The product path may not route through oracle/ (tier rule): no import of it anywhere in the package."""
    pkg = os.path.join(ROOT, "bandedmatrices.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "bmoracle" not in txt, f
