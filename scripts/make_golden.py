"""Generates tests/golden/*.json — known-answer vectors for the Levenshtein path.

The reference (StringWars) ships NO golden vectors and its algorithm crates cannot be built or imported in this
container, so these vectors do not come from the reference. They come from a THIRD-PARTY
implementation that is present in the image and shares no code with this repo:
    torchaudio 2.11 `torchaudio.functional.edit_distance` (pure-Python DP over sequences)
plus literature known answers. Byte-level vectors feed the sequences as `bytes`, codepoint-level ones as `str`.
Run from the repo root:  python scripts/make_golden.py
"""
import json
import os

import numpy as np
import torchaudio.functional as F

from stringwars_b200 import datasets

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

LITERATURE = [  # (a, b, distance) — textbook examples
    ("kitten", "sitting", 3), ("flaw", "lawn", 2), ("saturday", "sunday", 3), ("", "abc", 3), ("abc", "", 3),
    ("", "", 0), ("a", "a", 0), ("a", "b", 1), ("intention", "execution", 5), ("gumbo", "gambol", 2),
    ("rosettacode", "raisethysword", 8), ("book", "back", 2), ("abcdef", "azced", 3),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = []
    for a, b, d in LITERATURE:
        assert F.edit_distance(a.encode(), b.encode()) == d, (a, b, d)
        cases.append({"a": a.encode().hex(), "b": b.encode().hex(), "bytes": d, "chars": F.edit_distance(a, b), "src": "literature"})
    gens = [
        ("words", datasets.words_pairs(64, seed=7)),
        ("ascii64", datasets.ascii_pairs(24, 64, seed=7)),
        ("ascii33x", datasets.ascii_pairs(16, 33, seed=8)),
        ("mutated64", datasets.mutated_pairs(32, 64, 0.15, seed=7)),
        ("mutated200", datasets.mutated_pairs(12, 200, 0.1, seed=9)),
        ("utf8lines", datasets.utf8_line_pairs(10, 40, 120, seed=7)),
        ("acgt300", datasets.acgt_pairs(6, 300, seed=7)),
    ]
    for name, (ta, tb) in gens:
        for i in range(len(ta)):
            a, b = ta[i], tb[i]
            case = {"a": a.hex(), "b": b.hex(), "bytes": F.edit_distance(a, b), "src": name}
            try:
                case["chars"] = F.edit_distance(a.decode("utf-8"), b.decode("utf-8"))
            except UnicodeDecodeError:
                pass
            cases.append(case)
    # ragged mix: different lengths on both sides, incl. empties and one-symbol strings
    rng = np.random.default_rng(11)
    for _ in range(60):
        la, lb = int(rng.integers(0, 90)), int(rng.integers(0, 90))
        a = bytes(rng.integers(97, 101, size=la, dtype=np.uint8))
        b = bytes(rng.integers(97, 101, size=lb, dtype=np.uint8))
        cases.append({"a": a.hex(), "b": b.hex(), "bytes": F.edit_distance(a, b), "chars": F.edit_distance(a.decode(), b.decode()), "src": "ragged"})
    with open(os.path.join(OUT, "levenshtein_torchaudio.json"), "w") as f:
        json.dump({"generator": "scripts/make_golden.py", "implementation": "torchaudio.functional.edit_distance 2.11.0",
                   "cases": cases}, f, indent=0)
    print(f"wrote {len(cases)} cases")


if __name__ == "__main__":
    main()
