"""ctypes binding of include/nwb200.h (the drop-in boundary) + a small object wrapper.

Mirrors the reference's plugin contract (nw_algorithm.hpp:11-13): ``align`` -> score,
``trace`` -> run-length transcript + trace hash, ``hash`` -> score hash, every failure reported
as an ``NwStat`` (run_types.hpp:12-24).  Nothing here computes alignments on the CPU.
"""
from __future__ import annotations

import ctypes as C
import enum
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class NwStat(enum.IntEnum):
    """run_types.hpp:12-24"""
    success = 0
    helpMenuRequested = 1
    errorCudaGeneral = 2
    errorMemoryAllocation = 3
    errorMemoryTransfer = 4
    errorKernelFailure = 5
    errorIoStream = 6
    errorInvalidFormat = 7
    errorInvalidValue = 8
    errorInvalidResult = 9


class NwB200Error(RuntimeError):
    def __init__(self, stat: int, msg: str, cuda: int = 0):
        super().__init__(f"{NwStat(stat).name}: {msg}" + (f" (cudaError {cuda})" if cuda else ""))
        self.stat = NwStat(stat)
        self.cuda = cuda


class _Params(C.Structure):
    _fields_ = [("rows_per_lane", C.c_int32), ("warps_per_block", C.c_int32), ("tile_cols", C.c_int32), ("reserved", C.c_int32)]


class _HdrInfo(C.Structure):
    _fields_ = [("tile_rows", C.c_int32), ("tile_cols", C.c_int32), ("trows", C.c_int32), ("tcols", C.c_int32),
                ("hrow_elems", C.c_int64), ("hcol_elems", C.c_int64)]


class _Timing(C.Structure):
    _fields_ = [("align_cpy_dev", C.c_float), ("align_calc", C.c_float), ("align_cpy_host", C.c_float),
                ("trace_calc", C.c_float), ("trace_cpy_host", C.c_float)]


class _MemUsage(C.Structure):
    _fields_ = [("device_bytes", C.c_uint64), ("pinned_host_bytes", C.c_uint64), ("shared_bytes", C.c_uint64), ("local_bytes", C.c_uint64),
                ("register_bytes", C.c_uint64), ("regs_per_thread", C.c_int32), ("threads_per_block", C.c_int32), ("blocks", C.c_int32),
                ("reserved", C.c_int32)]


@dataclass
class Params:
    rows_per_lane: int = 0
    warps_per_block: int = 0
    tile_cols: int = 0
    skew: int = 0          # 0 = automatic, 1 = dense systolic schedule, 2 = shuffle off the critical path

    def _c(self):
        return _Params(self.rows_per_lane, self.warps_per_block, self.tile_cols, self.skew)


@dataclass
class HeaderInfo:
    tile_rows: int
    tile_cols: int
    trows: int
    tcols: int
    hrow_elems: int
    hcol_elems: int


SCORE_ONLY = 0
KEEP_HEADERS = 1
WITH_TRACE = 2

# every symbol declared in include/nwb200.h (tests check the library exports all of them)
EXPORTS = [
    "nwb200_create", "nwb200_destroy", "nwb200_set_scoring", "nwb200_align_pair_i32", "nwb200_align_pair_u8",
    "nwb200_upload_pair_u8", "nwb200_fill_resident", "nwb200_trace_resident", "nwb200_fetch_score", "nwb200_fetch_trace",
    "nwb200_trace_pair", "nwb200_copy_headers", "nwb200_score_hash", "nwb200_align_batch", "nwb200_upload_batch",
    "nwb200_batch_resident", "nwb200_fetch_batch_scores", "nwb200_last_cuda_error", "nwb200_last_error",
    "nwb200_get_timing", "nwb200_stream", "nwb200_sync", "nwb200_kernel_launches", "nwb200_version",
    "nwb200_wave_upload", "nwb200_wave_export", "nwb200_wave_connect", "nwb200_wave_fill", "nwb200_wave_fetch",
    "nwb200_scan_upload", "nwb200_scan_fill", "nwb200_scan_fetch", "nwb200_batch_kernel_name",
    "nwb200_batch_resident_variant", "nwb200_align_batch_packed5", "nwb200_upload_batch_packed5", "nwb200_trace_info", "nwb200_wave_keep_headers", "nwb200_wave_export_headers", "nwb200_wave_connect_headers", "nwb200_wave_gather_headers",
    "nwb200_upload_pair_i32", "nwb200_get_hdr_info", "nwb200_score_rows", "nwb200_trace_values", "nwb200_get_memory_usage",
]

_lib = None


def lib_path() -> str:
    return os.path.join(_HERE, "libnwb200.so")


def load_library():
    """Load libnwb200.so; fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `make -C gpuseqalign_b200/csrc` "
                          f"(or __graft_entry__.build()). There is no CPU fallback.")
    L = C.CDLL(path)
    vp, i32, i64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
    P = C.POINTER
    L.nwb200_create.argtypes = [P(vp), C.c_int]
    L.nwb200_destroy.argtypes = [vp]; L.nwb200_destroy.restype = None
    L.nwb200_set_scoring.argtypes = [vp, P(i32), C.c_int, C.c_int]
    L.nwb200_align_pair_i32.argtypes = [vp, P(i32), i64, P(i32), i64, P(_Params), C.c_int, P(i32), P(_HdrInfo)]
    L.nwb200_align_pair_u8.argtypes = [vp, vp, i64, vp, i64, P(_Params), C.c_int, P(i32), P(_HdrInfo)]
    L.nwb200_upload_pair_u8.argtypes = [vp, vp, i64, vp, i64, P(_Params)]
    L.nwb200_fill_resident.argtypes = [vp, C.c_int]
    L.nwb200_trace_resident.argtypes = [vp]
    L.nwb200_fetch_score.argtypes = [vp, P(i32)]
    L.nwb200_fetch_trace.argtypes = [vp, vp, C.c_size_t, P(C.c_size_t), P(u32)]
    L.nwb200_trace_pair.argtypes = [vp, vp, C.c_size_t, P(C.c_size_t), P(u32)]
    L.nwb200_copy_headers.argtypes = [vp, vp, vp]
    L.nwb200_score_hash.argtypes = [vp, P(u32)]
    L.nwb200_align_batch.argtypes = [vp, vp, C.c_size_t, vp, vp, vp, vp, C.c_size_t, vp, vp, vp, vp, vp]
    L.nwb200_upload_batch.argtypes = [vp, vp, C.c_size_t, vp, vp, vp, vp, C.c_size_t]
    L.nwb200_batch_resident.argtypes = [vp]
    L.nwb200_fetch_batch_scores.argtypes = [vp, vp]
    L.nwb200_last_cuda_error.argtypes = [vp]
    L.nwb200_last_error.argtypes = [vp]; L.nwb200_last_error.restype = C.c_char_p
    L.nwb200_get_timing.argtypes = [vp, P(_Timing)]
    L.nwb200_stream.argtypes = [vp]; L.nwb200_stream.restype = vp
    L.nwb200_sync.argtypes = [vp]
    L.nwb200_kernel_launches.argtypes = [vp]
    L.nwb200_version.restype = C.c_char_p
    L.nwb200_batch_kernel_name.argtypes = [vp]; L.nwb200_batch_kernel_name.restype = C.c_char_p
    L.nwb200_upload_pair_i32.argtypes = [vp, P(i32), i64, P(i32), i64, P(_Params)]
    L.nwb200_get_hdr_info.argtypes = [vp, P(_HdrInfo)]
    L.nwb200_score_rows.argtypes = [vp, i64, i64, vp]
    L.nwb200_trace_values.argtypes = [vp, vp, C.c_size_t, P(C.c_size_t)]
    L.nwb200_get_memory_usage.argtypes = [vp, P(_MemUsage)]
    L.nwb200_trace_info.argtypes = [vp, P(C.c_int), P(C.c_int), P(C.c_int)]
    L.nwb200_batch_resident_variant.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.nwb200_align_batch_packed5.argtypes = [vp, vp, C.c_size_t, vp, vp, vp, vp, C.c_size_t, vp]
    L.nwb200_upload_batch_packed5.argtypes = [vp, vp, C.c_size_t, vp, vp, vp, vp, C.c_size_t]
    L.nwb200_wave_keep_headers.argtypes = [vp, C.c_int]
    L.nwb200_wave_export_headers.argtypes = [vp, vp, vp]
    L.nwb200_wave_connect_headers.argtypes = [vp, vp, vp]
    L.nwb200_wave_gather_headers.argtypes = [vp, C.c_int]
    L.nwb200_wave_upload.argtypes = [vp, vp, i64, vp, i64, P(_Params), C.c_int, C.c_int, C.c_int]
    L.nwb200_wave_export.argtypes = [vp, vp]
    L.nwb200_wave_connect.argtypes = [vp, vp]
    L.nwb200_wave_fill.argtypes = [vp, C.c_uint]
    L.nwb200_wave_fetch.argtypes = [vp, P(C.c_int), P(i32)]
    L.nwb200_scan_upload.argtypes = [vp, vp, i64, vp, i64, C.c_int, C.c_int]
    L.nwb200_scan_fill.argtypes = [vp, C.c_uint]
    L.nwb200_scan_fetch.argtypes = [vp, P(C.c_int), P(i32)]
    _lib = L
    return L


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Engine:
    """One context = one GPU, one stream, device buffers that persist across calls."""

    def __init__(self, device: int = 0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.nwb200_create(C.byref(h), device)
        if rc != 0:
            raise NwB200Error(rc, "nwb200_create failed (a B200 / sm_100 device is required; there is no CPU fallback)")
        self._h = h
        self.device = device
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self._L.nwb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise NwB200Error(rc, self._L.nwb200_last_error(self._h).decode(), self._L.nwb200_last_cuda_error(self._h))

    # ---- scoring ------------------------------------------------------------------------
    def set_scoring(self, subst: np.ndarray, gap: int):
        s = np.ascontiguousarray(subst, dtype=np.int32).ravel()
        n = int(round(s.size ** 0.5))
        if n * n != s.size:
            raise NwB200Error(NwStat.errorInvalidValue, "substitution matrix must be square")
        self._check(self._L.nwb200_set_scoring(self._h, s.ctypes.data_as(C.POINTER(C.c_int32)), n, int(gap)))

    # ---- one pair -----------------------------------------------------------------------
    def align(self, y: np.ndarray, x: np.ndarray, *, keep_headers: bool = True, with_trace: bool = False,
              params: Optional[Params] = None) -> int:
        """NwAlignFn: byte letters in, align_cost out (H2D + fill + D2H of the score).  with_trace: the traceback kernels and
        the copy of the move list are enqueued behind the fill in the same call (``trace()`` then only formats)."""
        y = np.ascontiguousarray(y, dtype=np.uint8); x = np.ascontiguousarray(x, dtype=np.uint8)
        score = C.c_int32(0)
        info = _HdrInfo()
        p = params._c() if params else None
        flags = (KEEP_HEADERS if keep_headers else SCORE_ONLY) | (WITH_TRACE if with_trace else 0)
        self._check(self._L.nwb200_align_pair_u8(self._h, _ptr(y), y.size, _ptr(x), x.size, C.byref(p) if p else None,
                                                 flags, C.byref(score), C.byref(info)))
        self.info = HeaderInfo(info.tile_rows, info.tile_cols, info.trows, info.tcols, info.hrow_elems, info.hcol_elems)
        return score.value

    def align_i32(self, seqY: np.ndarray, seqX: np.ndarray, *, keep_headers: bool = True, params: Optional[Params] = None) -> int:
        """Same, taking the reference's int vectors with the dummy header element 0."""
        seqY = np.ascontiguousarray(seqY, dtype=np.int32); seqX = np.ascontiguousarray(seqX, dtype=np.int32)
        score = C.c_int32(0)
        info = _HdrInfo()
        p = params._c() if params else None
        i32p = C.POINTER(C.c_int32)
        self._check(self._L.nwb200_align_pair_i32(self._h, seqY.ctypes.data_as(i32p), seqY.size, seqX.ctypes.data_as(i32p), seqX.size,
                                                  C.byref(p) if p else None, KEEP_HEADERS if keep_headers else SCORE_ONLY,
                                                  C.byref(score), C.byref(info)))
        self.info = HeaderInfo(info.tile_rows, info.tile_cols, info.trows, info.tcols, info.hrow_elems, info.hcol_elems)
        return score.value

    def trace(self, cap: Optional[int] = None) -> Tuple[str, int]:
        """NwTraceFn: (edit_trace, trace_hash) of the last align(keep_headers=True)."""
        if cap is None:
            cap = 1 << 16
        while True:
            buf = C.create_string_buffer(cap)
            n = C.c_size_t(0)
            h = C.c_uint32(0)
            rc = self._L.nwb200_trace_pair(self._h, buf, cap, C.byref(n), C.byref(h))
            if rc == NwStat.errorInvalidValue and n.value > cap:
                cap = n.value + 16
                continue
            self._check(rc)
            return buf.raw[: n.value].decode("ascii"), h.value

    def headers(self):
        """Tile headers of the last align in the reference's layout."""
        hrow = np.empty(self.info.hrow_elems, dtype=np.int32)
        hcol = np.empty(self.info.hcol_elems, dtype=np.int32)
        self._check(self._L.nwb200_copy_headers(self._h, _ptr(hrow), _ptr(hcol)))
        return hrow, hcol

    def score_rows(self, row0: int, nrows: int, adjcols: int) -> np.ndarray:
        """Rows [row0, row0 + nrows) of the full score matrix (NwPrintScore's data), shape (nrows, adjcols)."""
        out = np.empty((nrows, adjcols), dtype=np.int32)
        self._check(self._L.nwb200_score_rows(self._h, row0, nrows, _ptr(out)))
        return out

    def trace_values(self) -> np.ndarray:
        """calcDebugTrace: score-matrix values along the path, top-left -> bottom-right (after trace())."""
        n = C.c_size_t(0)
        self._L.nwb200_trace_values(self._h, None, 0, C.byref(n))
        out = np.empty(n.value, dtype=np.int32)
        self._check(self._L.nwb200_trace_values(self._h, _ptr(out), out.size, C.byref(n)))
        return out

    def trace_info(self) -> dict:
        """Corridor diagnostics of the last traceback: segments per band in the corridor pass, segments per band, miss flag."""
        cw, ns, miss = C.c_int(0), C.c_int(0), C.c_int(0)
        self._check(self._L.nwb200_trace_info(self._h, C.byref(cw), C.byref(ns), C.byref(miss)))
        return {"corridor_segments": cw.value, "segments": ns.value, "corridor_missed": bool(miss.value)}

    def memory_usage(self) -> dict:
        m = _MemUsage()
        self._check(self._L.nwb200_get_memory_usage(self._h, C.byref(m)))
        return {k: getattr(m, k) for k, _ in _MemUsage._fields_ if k != "reserved"}

    def score_hash(self) -> int:
        h = C.c_uint32(0)
        self._check(self._L.nwb200_score_hash(self._h, C.byref(h)))
        return h.value

    # ---- split form (inputs resident in HBM) --------------------------------------------
    def upload_pair(self, y: np.ndarray, x: np.ndarray, params: Optional[Params] = None):
        y = np.ascontiguousarray(y, dtype=np.uint8); x = np.ascontiguousarray(x, dtype=np.uint8)
        p = params._c() if params else None
        self._check(self._L.nwb200_upload_pair_u8(self._h, _ptr(y), y.size, _ptr(x), x.size, C.byref(p) if p else None))

    def fill_resident(self, keep_headers: bool = True):
        self._check(self._L.nwb200_fill_resident(self._h, KEEP_HEADERS if keep_headers else SCORE_ONLY))

    def trace_resident(self):
        self._check(self._L.nwb200_trace_resident(self._h))

    def fetch_score(self) -> int:
        s = C.c_int32(0)
        self._check(self._L.nwb200_fetch_score(self._h, C.byref(s)))
        return s.value

    def fetch_trace(self, cap: int = 1 << 16) -> Tuple[str, int]:
        while True:
            buf = C.create_string_buffer(cap)
            n = C.c_size_t(0)
            h = C.c_uint32(0)
            rc = self._L.nwb200_fetch_trace(self._h, buf, cap, C.byref(n), C.byref(h))
            if rc == NwStat.errorInvalidValue and n.value > cap:
                cap = n.value + 16
                continue
            self._check(rc)
            return buf.raw[: n.value].decode("ascii"), h.value

    # ---- batch --------------------------------------------------------------------------
    def upload_batch(self, letters, offY, lenY, offX, lenX):
        letters = np.ascontiguousarray(letters, dtype=np.uint8)
        offY = np.ascontiguousarray(offY, dtype=np.uint64); offX = np.ascontiguousarray(offX, dtype=np.uint64)
        lenY = np.ascontiguousarray(lenY, dtype=np.uint32); lenX = np.ascontiguousarray(lenX, dtype=np.uint32)
        self._npairs = lenY.size
        self._check(self._L.nwb200_upload_batch(self._h, _ptr(letters), letters.size, _ptr(offY), _ptr(lenY), _ptr(offX), _ptr(lenX), lenY.size))

    def upload_batch_packed5(self, packed, offY, lenY, offX, lenX):
        """Resident batch from 5-bit packed letters (synth.pack5): byte offsets of the sequences' bit streams, lengths in letters."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        offY = np.ascontiguousarray(offY, dtype=np.uint64); offX = np.ascontiguousarray(offX, dtype=np.uint64)
        lenY = np.ascontiguousarray(lenY, dtype=np.uint32); lenX = np.ascontiguousarray(lenX, dtype=np.uint32)
        self._npairs = lenY.size
        self._check(self._L.nwb200_upload_batch_packed5(self._h, _ptr(packed), packed.size, _ptr(offY), _ptr(lenY), _ptr(offX), _ptr(lenX), lenY.size))

    def align_batch_packed5(self, packed, offY, lenY, offX, lenX, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Scores of a batch whose letters arrive 5-bit packed, from host buffers (the slice pipeline of align_batch)."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        offY = np.ascontiguousarray(offY, dtype=np.uint64); offX = np.ascontiguousarray(offX, dtype=np.uint64)
        lenY = np.ascontiguousarray(lenY, dtype=np.uint32); lenX = np.ascontiguousarray(lenX, dtype=np.uint32)
        scores = out if out is not None else np.empty(lenY.size, dtype=np.int32)
        self._check(self._L.nwb200_align_batch_packed5(self._h, _ptr(packed), packed.size, _ptr(offY), _ptr(lenY), _ptr(offX), _ptr(lenX), lenY.size, _ptr(scores)))
        return scores

    def batch_resident(self):
        self._check(self._L.nwb200_batch_resident(self._h))

    VARIANTS = {"nw_affine": 1, "sw_linear": 2, "sw_affine": 3}

    def batch_resident_variant(self, variant: str, gap_open: int, gap_extend: int = 0):
        """The resident batch under an affine-gap / local variant (include/nwb200.h: NWB200_VARIANT_*); fetch_batch_scores afterwards."""
        self._check(self._L.nwb200_batch_resident_variant(self._h, self.VARIANTS[variant], gap_open, gap_extend))

    def align_batch_variant(self, letters, offY, lenY, offX, lenX, variant: str, gap_open: int, gap_extend: int = 0) -> np.ndarray:
        self.upload_batch(letters, offY, lenY, offX, lenX)
        self.batch_resident_variant(variant, gap_open, gap_extend)
        return self.fetch_batch_scores()

    def fetch_batch_scores(self) -> np.ndarray:
        out = np.empty(self._npairs, dtype=np.int32)
        self._check(self._L.nwb200_fetch_batch_scores(self._h, _ptr(out)))
        return out

    def align_batch(self, letters, offY, lenY, offX, lenX, want_trace: bool = False, out: Optional[np.ndarray] = None):
        """Scores of a batch of pairs from host buffers (H2D, kernels and D2H overlapped slice by slice); ``out``: a caller-owned
        int32 array for the scores (pinned memory saves the staging copy)."""
        letters = np.ascontiguousarray(letters, dtype=np.uint8)
        offY = np.ascontiguousarray(offY, dtype=np.uint64); offX = np.ascontiguousarray(offX, dtype=np.uint64)
        lenY = np.ascontiguousarray(lenY, dtype=np.uint32); lenX = np.ascontiguousarray(lenX, dtype=np.uint32)
        n = lenY.size
        if out is not None:
            if out.dtype != np.int32 or out.size != n or not out.flags.c_contiguous:
                raise NwB200Error(NwStat.errorInvalidValue, "out must be a contiguous int32 array with one entry per pair")
            scores = out
        else:
            scores = np.empty(n, dtype=np.int32)
        if not want_trace:
            self._check(self._L.nwb200_align_batch(self._h, _ptr(letters), letters.size, _ptr(offY), _ptr(lenY), _ptr(offX), _ptr(lenX), n,
                                                   _ptr(scores), None, None, None, None))
            return scores
        caps = 2 * (lenY.astype(np.uint64) + lenX.astype(np.uint64)) + 8
        edit_off = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(caps, out=edit_off[1:])
        edits = np.zeros(int(edit_off[-1]), dtype=np.uint8)
        edit_len = np.zeros(n, dtype=np.uint32)
        hashes = np.zeros(n, dtype=np.uint32)
        self._check(self._L.nwb200_align_batch(self._h, _ptr(letters), letters.size, _ptr(offY), _ptr(lenY), _ptr(offX), _ptr(lenX), n,
                                               _ptr(scores), _ptr(edits), _ptr(edit_off), _ptr(edit_len), _ptr(hashes)))
        strs = [edits[int(edit_off[p]): int(edit_off[p]) + int(edit_len[p])].tobytes().decode("ascii") for p in range(n)]
        return scores, strs, hashes

    # ---- one long pair as a cross-GPU wavefront (one Engine per rank) --------------------------
    def wave_upload(self, y, x, rank: int, world: int, block_cols: int, params: Optional[Params] = None) -> bytes:
        """Uploads the pair, plans the column blocks of this rank and returns the 64-byte IPC handle of its receive buffer."""
        y = np.ascontiguousarray(y, dtype=np.uint8); x = np.ascontiguousarray(x, dtype=np.uint8)
        p = params._c() if params else None
        self._check(self._L.nwb200_wave_upload(self._h, _ptr(y), y.size, _ptr(x), x.size, C.byref(p) if p else None, rank, world, block_cols))
        h = C.create_string_buffer(64)
        self._check(self._L.nwb200_wave_export(self._h, h))
        return h.raw

    def wave_connect(self, right_peer_handle: Optional[bytes]):
        if right_peer_handle is None:
            self._check(self._L.nwb200_wave_connect(self._h, None))
        else:
            buf = C.create_string_buffer(right_peer_handle, 64)
            self._check(self._L.nwb200_wave_connect(self._h, buf))

    def wave_fill(self, epoch: int):
        self._check(self._L.nwb200_wave_fill(self._h, epoch))

    def wave_fetch(self):
        has = C.c_int(0); s = C.c_int32(0)
        self._check(self._L.nwb200_wave_fetch(self._h, C.byref(has), C.byref(s)))
        return (s.value if has.value else None)

    def wave_keep_headers(self, on: bool = True):
        """Before wave_upload (world > 1): keep header rows and snapshots in the layout of the whole matrix, for a traceback."""
        self._check(self._L.nwb200_wave_keep_headers(self._h, 1 if on else 0))

    def wave_export_headers(self):
        """(64-byte IPC handle of the header rows, of the snapshots) of this rank."""
        hr = C.create_string_buffer(64); sn = C.create_string_buffer(64)
        self._check(self._L.nwb200_wave_export_headers(self._h, hr, sn))
        return hr.raw, sn.raw

    def wave_connect_headers(self, hr_handles, snap_handles):
        """Maps every other rank's header rows and snapshots (lists of 64-byte handles in rank order)."""
        a = C.create_string_buffer(b"".join(hr_handles), 64 * len(hr_handles))
        b = C.create_string_buffer(b"".join(snap_handles), 64 * len(snap_handles))
        self._check(self._L.nwb200_wave_connect_headers(self._h, a, b))

    def wave_gather_headers(self, full: bool = False):
        self._check(self._L.nwb200_wave_gather_headers(self._h, 1 if full else 0))

    def scan_upload(self, y, x, rank: int, world: int) -> bytes:
        y = np.ascontiguousarray(y, dtype=np.uint8); x = np.ascontiguousarray(x, dtype=np.uint8)
        self._check(self._L.nwb200_scan_upload(self._h, _ptr(y), y.size, _ptr(x), x.size, rank, world))
        h = C.create_string_buffer(64)
        self._check(self._L.nwb200_wave_export(self._h, h))
        return h.raw

    def scan_fill(self, epoch: int):
        self._check(self._L.nwb200_scan_fill(self._h, epoch))

    def scan_fetch(self):
        has = C.c_int(0); s = C.c_int32(0)
        self._check(self._L.nwb200_scan_fetch(self._h, C.byref(has), C.byref(s)))
        return (s.value if has.value else None)

    # ---- introspection ------------------------------------------------------------------
    def timing(self) -> dict:
        t = _Timing()
        self._L.nwb200_get_timing(self._h, C.byref(t))
        return {k: getattr(t, k) for k, _ in _Timing._fields_}

    def stream_ptr(self) -> int:
        return int(self._L.nwb200_stream(self._h) or 0)

    def sync(self):
        self._check(self._L.nwb200_sync(self._h))

    def launches(self) -> int:
        return int(self._L.nwb200_kernel_launches(self._h))

    def batch_kernel_name(self) -> str:
        return (self._L.nwb200_batch_kernel_name(self._h) or b"").decode()
