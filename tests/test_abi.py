"""CPU checks of the C-ABI boundary: the library builds, loads, and exports exactly what include/mts_b200.h
declares; the ctypes table agrees with the header; the product refuses to run without a GPU (no fallback)."""
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mts_b200.h")


def declared():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    decls = {}
    for m in re.finditer(r"(?:int|int64_t|const char \*)\s*\*?\s*(mts_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = [a.strip() for a in m.group(2).split(",")]
        decls[m.group(1)] = 0 if args == ["void"] else len(args)
    return decls


@pytest.fixture(scope="module")
def lib():
    from multimodaltopicsegmentation_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    return _lib


def test_header_symbols_exported(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    decl = declared()
    assert len(decl) >= 18
    missing = sorted(set(decl) - exported)
    assert not missing, f"declared in the header but not exported: {missing}"


def test_ctypes_table_matches_header(lib):
    decl = declared()
    for name, nargs in decl.items():
        assert name in lib.SIGNATURES, name
        assert len(lib.SIGNATURES[name][1]) == nargs, (name, nargs, len(lib.SIGNATURES[name][1]))


def test_library_loads_and_reports_version(lib):
    assert lib.load().mts_version() >= 1


def test_argument_errors_are_reported_not_swallowed(lib):
    with pytest.raises(lib.MtsError) as e:
        lib.call("mts_crf_viterbi", 0, 0, 0, 1, 1, 4, 0, 0, 0, 0)
    assert "null pointer" in str(e.value)
    with pytest.raises(lib.MtsError):
        lib.call("mts_gemm_tf32x3", 16, 16, 16, 16, 0, 16, 8, 8, 33, 8, 0, 0, 0)  # Kp not a multiple of 32


def test_no_cpu_fallback():
    """CPU tensors are refused: the product path never computes on the host."""
    from multimodaltopicsegmentation_b200 import BiLSTM, _lib

    m = BiLSTM(2, 12, 8, num_layers=1, loss_fn="FocalLoss")
    with pytest.raises((_lib.MtsError, RuntimeError)):
        m(torch.randn(2, 5, 12), torch.tensor([5, 3]))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "multimodaltopicsegmentation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src or f.endswith(".md"), f
