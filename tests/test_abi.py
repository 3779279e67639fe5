"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol
include/tnb200.h declares; the ctypes table lists exactly those; no compute without a GPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "tnb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tnb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    import __graft_entry__ as g
    g.build()
    from itensorsgpu_b200 import tn
    lib = tn.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(tn._lib.SIGNATURES) == names


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from itensorsgpu_b200 import tn
    with pytest.raises(tn.TnbError):
        tn._lib.Handle()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "itensorsgpu.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
