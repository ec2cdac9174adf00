"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/sgicp_b200.h declares (no compute calls -- there is no GPU here), and fails loudly without a device."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sgicp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sgb_[a-z0-9_]+)\s*\(", text)))


def test_header_matches_python_binding():
    import small_gicp_b200 as sg

    assert declared_symbols() == sg.exported_symbols()


def test_library_exports_every_declared_symbol():
    import small_gicp_b200 as sg

    path = sg.library_path()
    assert os.path.exists(path), "build with __graft_entry__.build()"
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_no_cpu_fallback():
    """Without a CUDA device the product must raise, never compute on the CPU."""
    import torch

    import small_gicp_b200 as sg

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sg.SgbError):
        sg.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "small_gicp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "sgicp_oracle" not in text, f
