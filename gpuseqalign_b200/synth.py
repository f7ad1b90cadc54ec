"""Deterministic synthetic inputs of SURVEY.md 8(d): splitmix64 residues, mutated copies.

Input generation only (host side, numpy); nothing here aligns anything.
    state += 0x9E3779B97F4A7C15; z = state
    z = (z ^ z >> 30) * 0xBF58476D1CE4E5B9;  z = (z ^ z >> 27) * 0x94D049BB133111EB;  out = z ^ z >> 31
    residue = (out >> 33) % 20      -> letter index 0..19 = ARNDCQEGHILKMFPSTWYV (subst.json:5-24)
"""
from __future__ import annotations

import numpy as np

_G = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64(seed: int, n: int, start: int = 0) -> np.ndarray:
    """outputs start .. start+n-1 of the splitmix64 stream whose state was initialised to `seed`."""
    with np.errstate(over="ignore"):
        k = np.arange(start + 1, start + n + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + k * _G
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def letters(seed: int, n: int) -> np.ndarray:
    """n residues (uint8, 0..19)."""
    return ((splitmix64(seed, n) >> np.uint64(33)) % np.uint64(20)).astype(np.uint8)


def mutated_copy(x: np.ndarray, seed: int, target_len: int) -> np.ndarray:
    """SURVEY.md 8(d) 'mutated copy': per residue u = next() % 1000: u < 25 deletion, u < 50 insertion
    (random residue, then the original), u < 150 substitution, else copy; truncate / pad to target_len."""
    out = []
    s = seed & 0xFFFFFFFFFFFFFFFF
    # one stream of draws, consumed in order (decision, then the random residue when one is needed)
    draws = splitmix64(s, 2 * len(x) + 2 * target_len + 16)
    di = 0
    for r in x:
        u = int(draws[di] % np.uint64(1000)); di += 1
        if u < 25:
            continue
        if u < 50:
            out.append(int((draws[di] >> np.uint64(33)) % np.uint64(20))); di += 1
            out.append(int(r))
        elif u < 150:
            out.append(int((draws[di] >> np.uint64(33)) % np.uint64(20))); di += 1
        else:
            out.append(int(r))
    while len(out) < target_len:
        out.append(int((draws[di] >> np.uint64(33)) % np.uint64(20))); di += 1
    return np.array(out[:target_len], dtype=np.uint8)


def batch_pairs(first_pair: int, n_pairs: int, len_y: int, len_x: int):
    """cfg3 batch: pair p has X seed 3e6+2p and Y seed 3e6+2p+1.  Returns (letters, offY, lenY, offX, lenX); the byte pool
    holds the pairs one after the other ([X_p | Y_p]), so a prefix of the pool is a prefix of the batch."""
    stride = len_x + len_y
    pool = np.empty((n_pairs, stride), dtype=np.uint8)
    with np.errstate(over="ignore"):
        p = np.arange(first_pair, first_pair + n_pairs, dtype=np.uint64)
        for which, ln, col in ((0, len_x, 0), (1, len_y, len_x)):
            seeds = np.uint64(3_000_000) + np.uint64(2) * p + np.uint64(which)
            k = np.arange(1, ln + 1, dtype=np.uint64)
            chunk = max(1, (1 << 22) // max(ln, 1))
            for lo in range(0, n_pairs, chunk):
                hi = min(n_pairs, lo + chunk)
                z = seeds[lo:hi, None] + k[None, :] * _G
                z = (z ^ (z >> np.uint64(30))) * _M1
                z = (z ^ (z >> np.uint64(27))) * _M2
                z = z ^ (z >> np.uint64(31))
                pool[lo:hi, col:col + ln] = ((z >> np.uint64(33)) % np.uint64(20)).astype(np.uint8)
    offX = np.arange(n_pairs, dtype=np.uint64) * np.uint64(stride)
    offY = offX + np.uint64(len_x)
    lenX = np.full(n_pairs, len_x, dtype=np.uint32)
    lenY = np.full(n_pairs, len_y, dtype=np.uint32)
    return pool.reshape(-1), offY, lenY, offX, lenX


def pack5(letters: np.ndarray, offs: np.ndarray, lens: np.ndarray):
    """5-bit packing of the sequences of a byte pool (what nwb200_align_batch_packed5 takes): sequence k -- `lens[k]` letters at
    `letters[offs[k]:]` -- becomes a little-endian bit stream of 5 bits per letter that starts at byte `new_offs[k]` of the packed
    pool (8 letters in 5 bytes; every sequence starts on a byte).  Returns (packed pool with 64 bytes of slack, new_offs)."""
    letters = np.ascontiguousarray(letters, dtype=np.uint8)
    offs = np.asarray(offs, dtype=np.int64); lens = np.asarray(lens, dtype=np.int64)
    nbytes = (lens * 5 + 7) // 8
    new_offs = np.concatenate([[0], np.cumsum(nbytes)[:-1]]).astype(np.uint64) if lens.size else np.zeros(0, np.uint64)
    total = int(nbytes.sum())
    out = np.zeros(total + 64, dtype=np.uint8)
    if lens.size == 0:
        return out, new_offs
    if np.all(lens == lens[0]) and lens[0] % 8 == 0 and lens[0] > 0:      # equal lengths, whole groups of 8 letters: vectorised, in slabs
        L = int(lens[0])
        per = L // 8 * 5                                                   # packed bytes per sequence
        ar = np.arange(L, dtype=np.int64)[None, :]
        slab = max(1, (1 << 24) // L)                                       # ~16 M letters (a few hundred MB of temporaries) at a time
        for lo in range(0, lens.size, slab):
            hi = min(lens.size, lo + slab)
            v = letters[offs[lo:hi, None] + ar].reshape(hi - lo, L // 8, 8)
            word = np.zeros((hi - lo, L // 8), dtype=np.uint64)
            for k in range(8):
                word |= (v[:, :, k].astype(np.uint64) & np.uint64(31)) << np.uint64(5 * k)
            b = np.empty((hi - lo, L // 8, 5), dtype=np.uint8)
            for k in range(5):
                b[:, :, k] = ((word >> np.uint64(8 * k)) & np.uint64(255)).astype(np.uint8)
            out[lo * per: hi * per] = b.reshape(-1)
        return out, new_offs
    for k in range(lens.size):                                             # ragged: sequence by sequence
        L = int(lens[k])
        if L == 0:
            continue
        v = letters[int(offs[k]): int(offs[k]) + L].astype(np.uint64) & np.uint64(31)
        bits = ((v[:, None] >> np.arange(5, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.uint8).reshape(-1)
        pad = (-bits.size) % 8
        if pad:
            bits = np.concatenate([bits, np.zeros(pad, np.uint8)])
        out[int(new_offs[k]): int(new_offs[k]) + int(nbytes[k])] = np.packbits(bits, bitorder="little")
    return out, new_offs
