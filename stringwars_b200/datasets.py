"""Synthetic, seeded generators for the five BASELINE.json configurations.

There is no network and no dataset on the GPU box, so every workload is synthetic; seed 42 is the reference's
own convention (`STRINGWARS_SEED` default, similarities/bench.py:848). Each generator returns two BytesTapes
of equal length whose i-th strings form pair i.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .tape import BytesTape

SEED = 42


def _fixed(n: int, length: int, lo: int, hi: int, rng: np.random.Generator, pinned: bool) -> BytesTape:
    data = rng.integers(lo, hi, size=(n, length), dtype=np.uint8).reshape(-1)
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(length)
    return BytesTape.from_arrays(data, offsets, pinned=pinned)


def _ragged(lengths: np.ndarray, alphabet: np.ndarray, rng: np.random.Generator, pinned: bool) -> BytesTape:
    offsets = np.zeros(lengths.size + 1, dtype=np.uint64)
    np.cumsum(lengths, out=offsets[1:])
    data = alphabet[rng.integers(0, alphabet.size, size=int(offsets[-1]))].astype(np.uint8)
    return BytesTape.from_arrays(data, offsets, pinned=pinned)


def words_pairs(n: int = 1000, seed: int = SEED, pinned: bool = False) -> Tuple[BytesTape, BytesTape]:
    """Config 1: ASCII 'words', length ~ U{1..16}, alphabet a-z (bench.rs:404-423 diagonal-pairs shape)."""
    rng = np.random.default_rng(seed)
    alphabet = np.arange(ord("a"), ord("z") + 1, dtype=np.uint8)
    la, lb = rng.integers(1, 17, size=n), rng.integers(1, 17, size=n)
    return _ragged(la, alphabet, rng, pinned), _ragged(lb, alphabet, rng, pinned)


def ascii_pairs(n: int = 1 << 20, length: int = 64, seed: int = SEED, pinned: bool = False) -> Tuple[BytesTape, BytesTape]:
    """Configs 2 and 4: pairs of iid printable-ASCII (0x20-0x7E) strings, both exactly `length` bytes."""
    rng = np.random.default_rng(seed)
    return _fixed(n, length, 0x20, 0x7F, rng, pinned), _fixed(n, length, 0x20, 0x7F, rng, pinned)


def mutated_pairs(n: int, length: int = 64, rate: float = 0.1, seed: int = SEED) -> Tuple[BytesTape, BytesTape]:
    """Parity-diversity variant of config 2: b = a with ~`rate` random substitutions/insertions/deletions, so
    distances are small and alignments non-trivial (not used for headline numbers)."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0x20, 0x7F, size=(n, length), dtype=np.uint8)
    b_strings = []
    for row in a:
        out = []
        for ch in row:
            r = rng.random()
            if r < rate / 3:
                continue  # deletion
            if r < 2 * rate / 3:
                out.append(int(rng.integers(0x20, 0x7F)))  # substitution
                continue
            out.append(int(ch))
            if r > 1 - rate / 3:
                out.append(int(rng.integers(0x20, 0x7F)))  # insertion
        b_strings.append(bytes(out))
    ta = BytesTape.from_arrays(a.reshape(-1), np.arange(n + 1, dtype=np.uint64) * np.uint64(length))
    return ta, BytesTape.from_strings(b_strings)


def _vocabulary(rng: np.random.Generator, size: int = 50_000):
    """Mixed-script vocabulary: ASCII Latin, Latin-1 diacritics + Cyrillic (2-byte), CJK (3-byte)."""
    scripts = [
        (0.55, [chr(c) for c in range(ord("a"), ord("z") + 1)]),
        (0.15, [chr(c) for c in range(0xE0, 0xFF) if c != 0xF7] + [chr(c) for c in range(ord("a"), ord("z") + 1)]),
        (0.20, [chr(c) for c in range(0x430, 0x450)]),
        (0.10, [chr(c) for c in range(0x4E00, 0x4E00 + 2000)]),
    ]
    probs = np.array([s[0] for s in scripts])
    which = rng.choice(len(scripts), size=size, p=probs / probs.sum())
    lens = rng.integers(1, 11, size=size)
    words = []
    for w, l in zip(which, lens):
        letters = scripts[w][1]
        if w == 3:
            l = max(1, l // 3)
        words.append("".join(letters[i] for i in rng.integers(0, len(letters), size=l)).encode("utf-8"))
    return words


def alphabetic_words(n: int, seed: int = SEED, memory: str = None) -> BytesTape:
    """Words of 1..12 letters in alphabetic scripts — ASCII Latin, Latin-1 diacritics, Cyrillic (63 distinct two-byte code
    points) — the shape of a one- or two-language leipzig1M word list: the code-point engine's cross product over them
    fits one batch-wide byte alphabet."""
    rng = np.random.default_rng(seed)
    scripts = [[chr(c) for c in range(ord("a"), ord("z") + 1)],
               [chr(c) for c in range(0xE0, 0xFF) if c != 0xF7] + [chr(c) for c in range(ord("a"), ord("z") + 1)],
               [chr(c) for c in range(0x430, 0x450)]]
    which = rng.choice(3, size=n, p=[0.4, 0.2, 0.4])
    lens = rng.integers(1, 13, size=n)
    words = ["".join(scripts[w][i] for i in rng.integers(0, len(scripts[w]), size=l)) for w, l in zip(which, lens)]
    return BytesTape.from_strings(words, memory=memory)


def utf8_line_pairs(n: int = 100_000, lo: int = 128, hi: int = 512, seed: int = SEED,
                    pinned: bool = False) -> Tuple[BytesTape, BytesTape]:
    """Config 3: valid-UTF-8 'lines' of byte length ~ U{lo..hi} cut at codepoint boundaries, built from
    space-separated words of a Zipf(1.1) mixed-script vocabulary (leipzig1M-style)."""
    rng = np.random.default_rng(seed)
    vocab = _vocabulary(rng)
    ranks = np.arange(1, len(vocab) + 1, dtype=np.float64)
    p = ranks ** -1.1
    p /= p.sum()

    def one_tape() -> BytesTape:
        target = rng.integers(lo, hi + 1, size=n).astype(np.int64)
        need = int(target.sum()) + 8 * n
        ids = rng.choice(len(vocab), size=need // 5 + 16, p=p)
        blob = np.frombuffer(b" ".join(vocab[i] for i in ids), dtype=np.uint8)
        assert blob.size >= need, "vocabulary stream too short"
        is_lead = (blob & 0xC0) != 0x80
        starts = np.zeros(n, dtype=np.int64)
        ends = np.zeros(n, dtype=np.int64)
        pos = 0
        lead_idx = np.flatnonzero(is_lead)
        # snap every cut to the next codepoint boundary
        for i in range(n):
            s = pos
            e = s + int(target[i])
            e = int(lead_idx[np.searchsorted(lead_idx, e)])
            if e - s > hi:  # snapping forward overshot the cap: snap backwards instead
                e = int(lead_idx[np.searchsorted(lead_idx, s + int(target[i]), side="right") - 1])
            starts[i], ends[i] = s, e
            pos = e
        lengths = (ends - starts).astype(np.uint64)
        offsets = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(lengths, out=offsets[1:])
        data = blob[: int(ends[-1])]  # lines are consecutive slices of the blob
        return BytesTape.from_arrays(data, offsets, pinned=pinned)

    return one_tape(), one_tape()


def acgt_pairs(n: int = 10_000, length: int = 1024, seed: int = SEED, pinned: bool = False) -> Tuple[BytesTape, BytesTape]:
    """Config 5: iid ACGT strings of `length` bytes (mirrors `acgt_1k`, similarities/README.md:19)."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)

    def one() -> BytesTape:
        data = alphabet[rng.integers(0, 4, size=n * length)]
        return BytesTape.from_arrays(data, np.arange(n + 1, dtype=np.uint64) * np.uint64(length), pinned=pinned)

    return one(), one()


def cells_and_bytes(a: BytesTape, b: BytesTape) -> Tuple[int, int]:
    """Harness accounting for element-wise pairs: cells = sum len(a_i)*len(b_i) (bench.rs:413),
    bytes = sum len(a_i)+len(b_i) (bench.rs:414)."""
    la, lb = a.lengths().astype(np.uint64), b.lengths().astype(np.uint64)
    return int((la * lb).sum()), int(la.sum() + lb.sum())
