"""CPU-only checks of the drop-in boundary: the library loads and exports every symbol that
include/nwb200.h declares; without a GPU the engine refuses to start (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "nwb200.h")).read()
    return sorted(set(re.findall(r"\b(nwb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_match_binding_table():
    from gpuseqalign_b200 import capi
    assert _declared() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol():
    from gpuseqalign_b200 import capi
    if not os.path.exists(capi.lib_path()):
        pytest.skip("libnwb200.so not built (run __graft_entry__.build())")
    L = ctypes.CDLL(capi.lib_path())
    for name in _declared():
        assert hasattr(L, name), name
    capi.load_library()
    assert b"sm_100a" in capi.load_library().nwb200_version()


def test_no_cpu_fallback():
    """Without a CUDA device the engine must fail loudly instead of computing on the host."""
    import torch
    from gpuseqalign_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    if not os.path.exists(capi.lib_path()):
        pytest.skip("libnwb200.so not built")
    with pytest.raises(capi.NwB200Error) as ei:
        capi.Engine(0)
    assert ei.value.stat == capi.NwStat.errorCudaGeneral


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under gpuseqalign_b200/ or include/ may reference it."""
    bad = []
    for base in ("gpuseqalign_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", ".inc")):
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    if re.search(r"nw_oracle|pyoracle|libnworacle|libnwref|oracle/", txt):
                        bad.append(os.path.join(dp, fn))
    assert bad == []
