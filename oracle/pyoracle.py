"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU arm may import this
module; the product package ``gpuseqalign_b200`` never does.

Two libraries:

* ``oracle/libnworacle.so``    -- our C restatement (``nw_oracle.c``), built by ``make -C oracle oracle``.
* ``oracle/_ref/libnwref.so``  -- the unmodified reference compiled from ``/root/reference/src``
  plus ``ref_shim.cpp`` (``make -C oracle ref``); optional, present when it was prebuilt.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libnworacle.so")
REF_SO = os.path.join(_HERE, "_ref", "libnwref.so")
REF_RESRC = os.path.join(_HERE, "_ref", "resrc")

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")

_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            raise RuntimeError(f"{ORACLE_SO} missing: run `make -C oracle oracle` (or __graft_entry__.build())")
        L = C.CDLL(ORACLE_SO)
        L.nwo_hash_bytes.restype = C.c_uint32
        L.nwo_hash_bytes.argtypes = [C.c_char_p, C.c_size_t]
        L.nwo_fill_full.restype = C.c_int
        L.nwo_fill_full.argtypes = [_i32p, C.c_int64, _i32p, C.c_int64, _i32p, C.c_int, C.c_int, _i32p]
        L.nwo_fill_full_mt.restype = C.c_int
        L.nwo_fill_full_mt.argtypes = [_i32p, C.c_int64, _i32p, C.c_int64, _i32p, C.c_int, C.c_int, C.c_int, _i32p]
        L.nwo_score_hash_full.restype = C.c_uint32
        L.nwo_score_hash_full.argtypes = [_i32p, C.c_int64, C.c_int64]
        L.nwo_fill_rolling.restype = C.c_int
        L.nwo_fill_rolling.argtypes = [_i32p, C.c_int64, _i32p, C.c_int64, _i32p, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nwo_trace_full.restype = C.c_int
        L.nwo_trace_full.argtypes = [_i32p, _i32p, C.c_int64, _i32p, C.c_int64, C.c_char_p, C.c_size_t,
                                     C.POINTER(C.c_size_t), C.POINTER(C.c_uint32)]
        L.nwo_trace_sparse.restype = C.c_int
        L.nwo_trace_sparse.argtypes = [_i32p, _i32p, C.c_int, C.c_int, _i32p, C.c_int64, _i32p, C.c_int64,
                                       _i32p, C.c_int, C.c_int, C.c_char_p, C.c_size_t,
                                       C.POINTER(C.c_size_t), C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
        L.nwo_align_pair.restype = C.c_int
        L.nwo_align_pair.argtypes = [_i32p, C.c_int64, _i32p, C.c_int64, _i32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.c_int), C.c_void_p, C.c_char_p, C.c_size_t,
                                     C.POINTER(C.c_size_t), C.POINTER(C.c_uint32)]
        L.nwo_score_batch.restype = C.c_int
        L.nwo_score_batch.argtypes = [_u8p, _u64p, _u32p, _u64p, _u32p, C.c_size_t, _i32p, C.c_int, C.c_int, C.c_int, _i32p]
        L.nwo_score_batch_gotoh.restype = C.c_int
        L.nwo_score_batch_gotoh.argtypes = [_u8p, _u64p, _u32p, _u64p, _u32p, C.c_size_t, _i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p]
        L.nwo_synth_letters.restype = None
        L.nwo_synth_letters.argtypes = [C.c_uint64, _u8p, C.c_size_t]
        _lib = L
    return _lib


@dataclass
class PairResult:
    score: int
    score_hash: Optional[int] = None
    trace_hash: Optional[int] = None
    edit: Optional[str] = None


def _hdr(letters: np.ndarray) -> np.ndarray:
    out = np.zeros(letters.size + 1, dtype=np.int32)
    out[1:] = letters
    return out


def hash_bytes(s: bytes) -> int:
    return int(lib().nwo_hash_bytes(s, len(s)))


def align_pair(y: np.ndarray, x: np.ndarray, subst: np.ndarray, gap: int, *, want_hash=False, want_trace=True,
               threads: int = 1, blocksz: int = 256) -> PairResult:
    """Full-matrix oracle (cpu1/cpu4 + NwHash1_Plain + NwTrace1_Plain restatement). y, x: uint8 letters."""
    L = lib()
    sy, sx = _hdr(y), _hdr(x)
    substsz = int(round(len(subst) ** 0.5))
    score = C.c_int(0)
    sh = C.c_uint32(0)
    th = C.c_uint32(0)
    elen = C.c_size_t(0)
    cap = 2 * (sy.size + sx.size) + 16
    buf = C.create_string_buffer(cap) if want_trace else None
    rc = L.nwo_align_pair(sy, sy.size, sx, sx.size, np.ascontiguousarray(subst, dtype=np.int32), substsz, gap,
                          threads, blocksz, C.byref(score), C.cast(C.byref(sh), C.c_void_p) if want_hash else None,
                          buf, cap if want_trace else 0, C.byref(elen), C.byref(th))
    if rc != 0:
        raise RuntimeError(f"nwo_align_pair failed rc={rc}")
    return PairResult(score=score.value, score_hash=sh.value if want_hash else None,
                      trace_hash=th.value if want_trace else None,
                      edit=buf.raw[: elen.value].decode("ascii") if want_trace else None)


def fill_full(y: np.ndarray, x: np.ndarray, subst: np.ndarray, gap: int) -> np.ndarray:
    """The whole score matrix H, shape (len(y) + 1, len(x) + 1) (cpu1 restatement, nwalign_cpu1_st_row.cpp:39-62)."""
    sy, sx = _hdr(y), _hdr(x)
    H = np.zeros((sy.size, sx.size), dtype=np.int32)
    lib().nwo_fill_full(sy, sy.size, sx, sx.size, np.ascontiguousarray(subst, dtype=np.int32), int(round(len(subst) ** 0.5)), gap, H.reshape(-1))
    return H


def fill_rolling(y: np.ndarray, x: np.ndarray, subst: np.ndarray, gap: int, By: int = 0, Bx: int = 0,
                 want_hash: bool = False):
    """Rolling-row oracle: returns (score, hrow, hcol, score_hash); headers in the App. A-4 layout when By,Bx > 0."""
    L = lib()
    sy, sx = _hdr(y), _hdr(x)
    substsz = int(round(len(subst) ** 0.5))
    hrow = hcol = None
    pr = pc = None
    if By > 0 and Bx > 0:
        trows = max(1, -(-(sy.size - 1) // By))
        tcols = max(1, -(-(sx.size - 1) // Bx))
        hrow = np.zeros(trows * tcols * (1 + Bx), dtype=np.int32)
        hcol = np.zeros(trows * tcols * (1 + By), dtype=np.int32)
        pr, pc = hrow.ctypes.data_as(C.c_void_p), hcol.ctypes.data_as(C.c_void_p)
    sh = C.c_uint32(0)
    score = L.nwo_fill_rolling(sy, sy.size, sx, sx.size, np.ascontiguousarray(subst, dtype=np.int32), substsz, gap,
                               max(By, 1), max(Bx, 1), pr, pc, C.cast(C.byref(sh), C.c_void_p) if want_hash else None)
    return int(score), hrow, hcol, (sh.value if want_hash else None)


def trace_sparse(hrow: np.ndarray, hcol: np.ndarray, By: int, Bx: int, y: np.ndarray, x: np.ndarray,
                 subst: np.ndarray, gap: int) -> PairResult:
    L = lib()
    sy, sx = _hdr(y), _hdr(x)
    substsz = int(round(len(subst) ** 0.5))
    cap = 2 * (sy.size + sx.size) + 16
    buf = C.create_string_buffer(cap)
    elen = C.c_size_t(0)
    th = C.c_uint32(0)
    cost = C.c_int(0)
    rc = L.nwo_trace_sparse(np.ascontiguousarray(hrow, dtype=np.int32), np.ascontiguousarray(hcol, dtype=np.int32), By, Bx,
                            sy, sy.size, sx, sx.size, np.ascontiguousarray(subst, dtype=np.int32), substsz, gap,
                            buf, cap, C.byref(elen), C.byref(th), C.byref(cost))
    if rc != 0:
        raise RuntimeError(f"nwo_trace_sparse failed rc={rc}")
    return PairResult(score=cost.value, trace_hash=th.value, edit=buf.raw[: elen.value].decode("ascii"))


def score_batch(letters: np.ndarray, offY, lenY, offX, lenX, subst: np.ndarray, gap: int, threads: int = 0) -> np.ndarray:
    L = lib()
    n = len(lenY)
    scores = np.zeros(n, dtype=np.int32)
    substsz = int(round(len(subst) ** 0.5))
    rc = L.nwo_score_batch(np.ascontiguousarray(letters, dtype=np.uint8), np.ascontiguousarray(offY, dtype=np.uint64),
                           np.ascontiguousarray(lenY, dtype=np.uint32), np.ascontiguousarray(offX, dtype=np.uint64),
                           np.ascontiguousarray(lenX, dtype=np.uint32), n, np.ascontiguousarray(subst, dtype=np.int32),
                           substsz, gap, threads, scores)
    if rc != 0:
        raise RuntimeError(f"nwo_score_batch failed rc={rc}")
    return scores


def score_batch_gotoh(letters: np.ndarray, offY, lenY, offX, lenX, subst: np.ndarray, gapo: int, gape: int, local: bool, threads: int = 0) -> np.ndarray:
    """Affine-gap / Smith-Waterman scores (Gotoh; nw_oracle.c: parity unpinned -- the reference lists these as future work)."""
    L = lib()
    n = len(lenY)
    scores = np.zeros(n, dtype=np.int32)
    rc = L.nwo_score_batch_gotoh(np.ascontiguousarray(letters, dtype=np.uint8), np.ascontiguousarray(offY, dtype=np.uint64),
                                 np.ascontiguousarray(lenY, dtype=np.uint32), np.ascontiguousarray(offX, dtype=np.uint64),
                                 np.ascontiguousarray(lenX, dtype=np.uint32), n, np.ascontiguousarray(subst, dtype=np.int32),
                                 int(round(len(subst) ** 0.5)), gapo, gape, 1 if local else 0, threads, scores)
    if rc != 0:
        raise RuntimeError(f"nwo_score_batch_gotoh failed rc={rc}")
    return scores


def synth_letters(seed: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint8)
    lib().nwo_synth_letters(seed, out, n)
    return out


# --------------------------------------------------------------------------- reference shim
class _RefResult(C.Structure):
    _fields_ = [("stat", C.c_int), ("step", C.c_int), ("cuda_stat", C.c_int), ("align_cost", C.c_int),
                ("score_hash", C.c_uint), ("trace_hash", C.c_uint), ("edit_len", C.c_ulonglong),
                ("ms_align_alloc", C.c_float), ("ms_align_cpy_dev", C.c_float), ("ms_align_init_hdr", C.c_float),
                ("ms_align_calc", C.c_float), ("ms_align_cpy_host", C.c_float), ("ms_hash_calc", C.c_float),
                ("ms_trace_alloc", C.c_float), ("ms_trace_calc", C.c_float),
                ("tile_hdr_rows", C.c_int), ("tile_hdr_cols", C.c_int), ("tile_hrow_len", C.c_int), ("tile_hcol_len", C.c_int)]


def ref_available() -> bool:
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        if not ref_available():
            raise RuntimeError(f"{REF_SO} missing: run `make -C oracle ref` where /root/reference exists")
        R = C.CDLL(REF_SO)
        R.nwref_run.restype = C.c_int
        R.nwref_run.argtypes = [C.c_int, _i32p, C.c_int, _i32p, C.c_int, _i32p, C.c_int, C.c_int, _i32p, C.c_int,
                                C.c_int, C.c_int, C.c_char_p, C.c_ulonglong, C.POINTER(_RefResult)]
        R.nwref_trace_from_headers.restype = C.c_int
        R.nwref_trace_from_headers.argtypes = [_i32p, C.c_int, _i32p, C.c_int, _i32p, C.c_int, C.c_int, _i32p, _i32p,
                                               C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_ulonglong, C.POINTER(_RefResult)]
        R.nwref_has_gpu9.restype = C.c_int
        if hasattr(R, "nwref_batch_cpu"):
            R.nwref_batch_cpu.restype = C.c_int
            R.nwref_batch_cpu.argtypes = [C.c_int, _u8p, _u64p, _u32p, _u64p, _u32p, C.c_ulonglong, _i32p, C.c_int, C.c_int, C.c_int,
                                          C.c_int, _i32p, C.POINTER(C.c_double)]
        _ref = R
    return _ref


@dataclass
class RefRun:
    score: int
    score_hash: Optional[int]
    trace_hash: Optional[int]
    edit: Optional[str]
    laps_ms: dict
    tiles: tuple


def ref_run(alg: str, y: np.ndarray, x: np.ndarray, subst: np.ndarray, gap: int, *, params=None,
            want_hash=True, want_trace=True) -> RefRun:
    """alg in {'cpu1','cpu4','gpu9'}; runs the reference's own align (+hash, +trace)."""
    R = ref()
    code = {"cpu1": 1, "cpu4": 4, "gpu9": 9}[alg]
    sy, sx = _hdr(y), _hdr(x)
    substsz = int(round(len(subst) ** 0.5))
    if params is None:
        params = {"cpu1": [], "cpu4": [256], "gpu9": [128, 4, 4, 48]}[alg]
    p = np.asarray(params if len(params) else [0], dtype=np.int32)
    cap = 2 * (sy.size + sx.size) + 16
    buf = C.create_string_buffer(cap)
    out = _RefResult()
    rc = R.nwref_run(code, sy, sy.size, sx, sx.size, np.ascontiguousarray(subst, dtype=np.int32), substsz, gap,
                     p, len(params), int(want_hash), int(want_trace), buf, cap, C.byref(out))
    if rc != 0:
        raise RuntimeError(f"reference {alg} failed: NwStat={out.stat} step={out.step} cuda={out.cuda_stat}")
    laps = {k: getattr(out, "ms_" + k) for k in ("align_alloc", "align_cpy_dev", "align_init_hdr", "align_calc",
                                                  "align_cpy_host", "hash_calc", "trace_alloc", "trace_calc")}
    return RefRun(score=out.align_cost, score_hash=out.score_hash if want_hash else None,
                  trace_hash=out.trace_hash if want_trace else None,
                  edit=buf.raw[: out.edit_len].decode("ascii") if want_trace else None, laps_ms=laps,
                  tiles=(out.tile_hdr_rows, out.tile_hdr_cols, out.tile_hrow_len, out.tile_hcol_len))


def ref_trace_from_headers(y, x, subst, gap, hrow, hcol, By, Bx, want_hash=False) -> RefRun:
    R = ref()
    sy, sx = _hdr(y), _hdr(x)
    substsz = int(round(len(subst) ** 0.5))
    cap = 2 * (sy.size + sx.size) + 16
    buf = C.create_string_buffer(cap)
    out = _RefResult()
    rc = R.nwref_trace_from_headers(sy, sy.size, sx, sx.size, np.ascontiguousarray(subst, dtype=np.int32), substsz, gap,
                                    np.ascontiguousarray(hrow, dtype=np.int32), np.ascontiguousarray(hcol, dtype=np.int32),
                                    By, Bx, int(want_hash), buf, cap, C.byref(out))
    if rc != 0:
        raise RuntimeError(f"reference NwTrace2_Sparse failed: NwStat={out.stat} step={out.step}")
    return RefRun(score=out.align_cost, score_hash=out.score_hash if want_hash else None, trace_hash=out.trace_hash,
                  edit=buf.raw[: out.edit_len].decode("ascii"), laps_ms={}, tiles=())


def ref_batch_cpu(alg: str, letters, offY, lenY, offX, lenX, subst: np.ndarray, gap: int, *, threads: int = 1, blocksz: int = 256):
    """The reference's own cpu4 / cpu1 align called once per pair (benchmark.cpp:406 semantics), the pairs dealt to `threads`
    OpenMP threads by the shim.  Returns (scores, wall_ms)."""
    R = ref()
    n = len(lenY)
    scores = np.zeros(n, dtype=np.int32)
    ms = C.c_double(0.0)
    substsz = int(round(len(subst) ** 0.5))
    rc = R.nwref_batch_cpu({"cpu1": 1, "cpu4": 4}[alg], np.ascontiguousarray(letters, dtype=np.uint8),
                           np.ascontiguousarray(offY, dtype=np.uint64), np.ascontiguousarray(lenY, dtype=np.uint32),
                           np.ascontiguousarray(offX, dtype=np.uint64), np.ascontiguousarray(lenX, dtype=np.uint32), n,
                           np.ascontiguousarray(subst, dtype=np.int32), substsz, gap, blocksz, threads, scores, C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"reference {alg} batch failed rc={rc}")
    return scores, float(ms.value)
