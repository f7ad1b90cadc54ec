"""On-disk formats of the reference, read bit-compatibly.

Host-side mirror of the reference's loaders so that the engine accepts the reference's
own ``subst.json``, FASTA, pair-list and parameter files unchanged:

* ``read_subst``   -- ``file_formats.hpp:80-96`` + checks in ``cmd_parser.cpp:316-355``;
  JSON with ``//`` and ``/* */`` comments (``io.hpp:29-33``).
* ``read_fasta``   -- ``file_formats.cpp:143-239``: first token after ``>`` is the id,
  sequences may span lines, whitespace ignored, unknown letters are an error
  (``file_formats.cpp:53-61``), duplicate ids rejected (``:101``).  Letters are returned as
  ``uint8`` index arrays *without* the dummy header element; ``with_header`` produces the
  reference's ``vector<int>`` convention (``file_formats.cpp:43-47``).
* ``read_pairs``   -- ``file_formats.cpp:241-399``: ``seqY_id seqX_id`` per line with optional
  ``[l:r]`` / ``[l:]`` / ``[:r]`` / ``[:]`` substring ranges (0-based, r exclusive).
* ``read_params``  -- ``file_formats.hpp:99-135``: ``{alg: {param: [values...]}}``.

Pure parsing; no alignment arithmetic lives here.
"""
from __future__ import annotations

import json
import re
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np


class FormatError(ValueError):
    """Mirrors NwStat::errorInvalidFormat / errorInvalidValue from the reference loaders."""


def _strip_json_comments(text: str) -> str:
    out = []
    i, n = 0, len(text)
    in_str = False
    while i < n:
        c = text[i]
        if in_str:
            out.append(c)
            if c == "\\" and i + 1 < n:
                out.append(text[i + 1])
                i += 1
            elif c == '"':
                in_str = False
        elif c == '"':
            in_str = True
            out.append(c)
        elif c == "/" and i + 1 < n and text[i + 1] == "/":
            while i < n and text[i] != "\n":
                i += 1
            continue
        elif c == "/" and i + 1 < n and text[i + 1] == "*":
            j = text.find("*/", i + 2)
            if j < 0:
                raise FormatError("unterminated /* comment")
            i = j + 2
            continue
        else:
            out.append(c)
        i += 1
    return "".join(out)


def read_json_with_comments(path: str):
    with open(path, "r") as f:
        text = f.read()
    try:
        return json.loads(_strip_json_comments(text))
    except json.JSONDecodeError as ex:  # same class of failure as io.hpp:43-47
        raise FormatError(f"{path}: {ex}") from ex


@dataclass
class SubstData:
    letter_map: Dict[str, int]            # insertion ordered, indices 0..N-1
    subst_map: Dict[str, np.ndarray]      # name -> int32[N*N], row-major subst[y*N + x]

    @property
    def size(self) -> int:
        return len(self.letter_map)

    @property
    def letters(self) -> str:
        return "".join(self.letter_map.keys())

    def matrix(self, name: str) -> np.ndarray:
        if name not in self.subst_map:
            raise FormatError(f"unknown substitution matrix name: {name}")
        return self.subst_map[name]

    def encode(self, s: str) -> np.ndarray:
        try:
            return np.fromiter((self.letter_map[c] for c in s), dtype=np.uint8, count=len(s))
        except KeyError as ex:
            raise FormatError(f"letter not found in substitution letters: {ex}") from ex

    def decode(self, idx: np.ndarray) -> str:
        letters = self.letters
        return "".join(letters[int(v)] for v in idx)


def read_subst(path: str) -> SubstData:
    obj = read_json_with_comments(path)
    if set(obj.keys()) != {"letterMap", "substMap"}:
        raise FormatError("expected exactly the keys 'letterMap' and 'substMap'")
    letter_map = obj["letterMap"]
    for k, (letter, idx) in enumerate(letter_map.items()):
        if len(letter) != 1:
            raise FormatError("letters must be single characters")
        if idx != k:
            raise FormatError("letter indices must be 0,1,2,... in file order")
    n = len(letter_map)
    subst_map = {}
    for name, flat in obj["substMap"].items():
        arr = np.asarray(flat, dtype=np.int32)
        if arr.size != n * n:
            raise FormatError(f"substitution matrix '{name}' must have {n * n} entries")
        subst_map[name] = arr
    return SubstData(letter_map=dict(letter_map), subst_map=subst_map)


def with_header(letters: np.ndarray) -> np.ndarray:
    """uint8 letters -> the reference's int32 vector with dummy element 0 in front."""
    out = np.zeros(letters.size + 1, dtype=np.int32)
    out[1:] = letters
    return out


@dataclass
class SeqData:
    ids: List[str]
    seqs: Dict[str, np.ndarray]           # id -> uint8 letter indices (no header element)


def read_fasta(path: str, subst: SubstData) -> SeqData:
    ids: List[str] = []
    chunks: Dict[str, List[str]] = {}
    cur: Optional[str] = None
    with open(path, "r") as f:
        for line_no, raw in enumerate(f, 1):
            line = raw.rstrip("\n")
            if line.strip() == "":
                continue
            if line.startswith(">"):
                toks = line[1:].split()
                if not toks:
                    raise FormatError(f"{path}:{line_no}: expected a sequence id after '>'")
                cur = toks[0]
                if cur in chunks:
                    raise FormatError(f"{path}:{line_no}: duplicate sequence id '{cur}'")
                ids.append(cur)
                chunks[cur] = []
            else:
                if cur is None:
                    raise FormatError(f"{path}:{line_no}: sequence data before the first header")
                chunks[cur].append("".join(line.split()))
    seqs = {}
    for sid in ids:
        s = "".join(chunks[sid])
        if len(s) == 0:
            raise FormatError(f"{path}: sequence '{sid}' is empty")
        seqs[sid] = subst.encode(s)
    return SeqData(ids=ids, seqs=seqs)


@dataclass
class SeqRange:
    l: Optional[int] = None               # inclusive
    r: Optional[int] = None               # exclusive

    def apply(self, n: int) -> Tuple[int, int]:
        lo = 0 if self.l is None else self.l
        hi = n if self.r is None else self.r
        if not (0 <= lo < hi <= n):
            raise FormatError(f"bad substring bounds [{lo}:{hi}] for length {n}")
        return lo, hi

    def suffix(self) -> str:
        if self.l is None and self.r is None:
            return ""
        return f"[{'' if self.l is None else self.l}:{'' if self.r is None else self.r}]"


@dataclass
class SeqPair:
    y_id: str
    x_id: str
    y_range: SeqRange
    x_range: SeqRange


_PAIR_TOKEN = re.compile(r"\s*([^\s\[\]]+)\s*(?:\[\s*(\d*)\s*:\s*(\d*)\s*\])?")


def read_pairs(path: str, seqs: SeqData) -> List[SeqPair]:
    pairs: List[SeqPair] = []
    with open(path, "r") as f:
        for line_no, raw in enumerate(f, 1):
            line = raw.strip()
            if not line:
                continue
            pos = 0
            parsed = []
            for _ in range(2):
                m = _PAIR_TOKEN.match(line, pos)
                if not m:
                    raise FormatError(f"{path}:{line_no}: expected 'seqY_id seqX_id'")
                sid, lo, hi = m.group(1), m.group(2), m.group(3)
                if sid not in seqs.seqs:
                    raise FormatError(f"{path}:{line_no}: unknown sequence id '{sid}'")
                rng = SeqRange(int(lo) if lo not in (None, "") else None,
                               int(hi) if hi not in (None, "") else None)
                if m.group(2) is not None:
                    # an explicit [:] keeps lNotDefault/rNotDefault false, like the reference
                    pass
                rng.apply(seqs.seqs[sid].size)
                parsed.append((sid, rng))
                pos = m.end()
            if line[pos:].strip():
                raise FormatError(f"{path}:{line_no}: trailing characters")
            pairs.append(SeqPair(parsed[0][0], parsed[1][0], parsed[0][1], parsed[1][1]))
    return pairs


def pair_letters(pair: SeqPair, seqs: SeqData) -> Tuple[np.ndarray, np.ndarray]:
    y = seqs.seqs[pair.y_id]
    x = seqs.seqs[pair.x_id]
    yl, yr = pair.y_range.apply(y.size)
    xl, xr = pair.x_range.apply(x.size)
    return y[yl:yr], x[xl:xr]


def read_params(path: str) -> Dict[str, Dict[str, List[int]]]:
    obj = read_json_with_comments(path)
    out: Dict[str, Dict[str, List[int]]] = {}
    for alg, params in obj.items():
        if not isinstance(params, dict):
            raise FormatError(f"{path}: parameters of '{alg}' must be an object")
        out[alg] = {}
        for name, values in params.items():
            if not isinstance(values, list) or not all(isinstance(v, int) for v in values):
                raise FormatError(f"{path}: parameter '{alg}.{name}' must be a list of integers")
            out[alg][name] = list(values)
    return out


def param_combinations(params: Dict[str, List[int]]):
    """Cartesian product, last key fastest (run_types.cpp:69-83)."""
    names = list(params.keys())
    if any(len(params[n]) == 0 for n in names):
        return
    idx = [0] * len(names)
    while True:
        yield {n: params[n][i] for n, i in zip(names, idx)}
        k = len(names) - 1
        while k >= 0:
            idx[k] += 1
            if idx[k] < len(params[names[k]]):
                break
            idx[k] = 0
            k -= 1
        if k < 0:
            return
