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


def test_pack5_round_trip():
    """synth.pack5 (the host side of nwb200_align_batch_packed5): 8 letters in 5 bytes, little-endian bit stream, every sequence on a byte
    boundary; the vectorised path (equal lengths, multiples of 8) and the sequence-by-sequence path agree with a bit-level unpack."""
    import numpy as np
    from gpuseqalign_b200 import synth
    rng = np.random.default_rng(5)
    pool = rng.integers(0, 25, 5000).astype(np.uint8)
    for lens in (np.full(12, 256), np.array([0, 1, 7, 8, 9, 255, 256, 257, 31, 64, 3, 100])):
        offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
        packed, noffs = synth.pack5(pool, offs, lens)
        assert packed.size == int(((lens * 5 + 7) // 8).sum()) + 64
        for k in range(lens.size):
            L = int(lens[k])
            raw = packed[int(noffs[k]): int(noffs[k]) + (5 * L + 7) // 8]
            bits = np.unpackbits(raw, bitorder="little")[: 5 * L].reshape(L, 5)
            assert np.array_equal((bits * (1 << np.arange(5))).sum(1), pool[int(offs[k]): int(offs[k]) + L])
