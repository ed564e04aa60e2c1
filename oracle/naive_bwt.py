"""TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The BWT of a string collection straight from its definition, for tiny cases only
(pure Python, O(total^2 log total)).  Pins both the C restatement and the reference
binary to the textbook object ropebwt2 claims to build (README.md:76-101, SURVEY.md
section 4): sort all suffixes of P0 $0, P1 $1, ... with $0 < $1 < ... < A < C < G < T < N
and emit, for every suffix, the symbol in front of it ($ for a whole string).

``so`` = 0 keeps the input order; 1 (RLO) / 2 (RCLO) first sort the collection the way
README.md:10-25 states (reverse-lexicographic, or the same after complementing), which
the reference proves equivalent to its ``-s`` / ``-r`` modes.
"""
from __future__ import annotations

import numpy as np


def _sort_collection(strings, so):
    if so == 0:
        return list(strings)
    if so == 1:
        key = lambda s: tuple(s[::-1])
    else:  # RCLO: compare the complemented string read backwards (README.md:22-25)
        comp = lambda c: 5 - c if 1 <= c <= 4 else c
        key = lambda s: tuple(comp(c) for c in s[::-1])
    # ties keep input order, exactly like `sort` followed by equal strings being identical
    return sorted(strings, key=key)


def naive_bwt(strings, so: int = 0) -> np.ndarray:
    """``strings``: list of sequences of nt6 codes 1..5 (forward orientation, no sentinel).
    Returns the multi-string BWT as nt6 codes with every sentinel printed as 0."""
    strs = [list(map(int, s)) for s in _sort_collection([list(s) for s in strings], so)]
    n = len(strs)
    suffixes = []
    for i, s in enumerate(strs):
        # symbol k of string i; the sentinel of string i sorts as (0, i), bases as (c, 0)
        full = [(c, 0) for c in s] + [(0, i)]
        for k in range(len(full)):
            prev = s[k - 1] if k > 0 else 0
            suffixes.append((full[k:], prev))
    suffixes.sort(key=lambda t: t[0])
    return np.array([p for _, p in suffixes], dtype=np.uint8)
