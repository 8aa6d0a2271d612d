"""Independent pure-Python restatements used ONLY to cross-check oracle/lev_oracle.c on small inputs.

TEST INFRASTRUCTURE ONLY. Written from the textbook definitions, deliberately in a different style
(full matrix, Python ints) from the C oracle so that a shared bug is unlikely.
"""
from __future__ import annotations


def levenshtein(a, b) -> int:
    """Full-matrix unit-cost edit distance over any two sequences (value at similarities/bench.rs:416-419)."""
    la, lb = len(a), len(b)
    D = [[0] * (lb + 1) for _ in range(la + 1)]
    for i in range(la + 1):
        D[i][0] = i
    for j in range(lb + 1):
        D[0][j] = j
    for i in range(1, la + 1):
        for j in range(1, lb + 1):
            D[i][j] = min(D[i - 1][j - 1] + (a[i - 1] != b[j - 1]), D[i - 1][j] + 1, D[i][j - 1] + 1)
    return D[la][lb]


def levenshtein_utf8(a: bytes, b: bytes) -> int:
    """Codepoint-level distance (similarities/bench.rs:437) via Python's own UTF-8 decoder."""
    return levenshtein(a.decode("utf-8"), b.decode("utf-8"))


def align_score(a: bytes, b: bytes, sub, first: int, nxt: int, local: bool = False) -> int:
    """Gotoh three-matrix score. `sub(x, y)` is the substitution score, a length-k gap costs first+(k-1)*nxt."""
    NEG = -(10 ** 12)
    la, lb = len(a), len(b)
    H = [[0] * (lb + 1) for _ in range(la + 1)]
    E = [[NEG] * (lb + 1) for _ in range(la + 1)]
    F = [[NEG] * (lb + 1) for _ in range(la + 1)]
    if not local:
        for i in range(1, la + 1):
            H[i][0] = first + (i - 1) * nxt
        for j in range(1, lb + 1):
            H[0][j] = first + (j - 1) * nxt
    best = 0
    for i in range(1, la + 1):
        for j in range(1, lb + 1):
            E[i][j] = max(E[i - 1][j] + nxt, H[i - 1][j] + first)
            F[i][j] = max(F[i][j - 1] + nxt, H[i][j - 1] + first)
            h = max(H[i - 1][j - 1] + sub(a[i - 1], b[j - 1]), E[i][j], F[i][j])
            if local:
                h = max(h, 0)
            H[i][j] = h
            best = max(best, h)
    return best if local else H[la][lb]
