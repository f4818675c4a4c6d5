"""Slow, obvious definitions used to cross-check the oracle (TEST INFRASTRUCTURE ONLY).

`naive_bwt` restates src/bwt_util.rs:154-171: the multi-string BWT is the last
column of the sorted, doubled rotations of every `s + "$"`.  `brute_count`
counts a k-mer the way the FM-index defines it (rotations that start with it);
`brute_rank` is the plain definition of rank used by the constrain_range KATs
(src/rle_bwt.rs:604-675).
"""
from __future__ import annotations


def naive_bwt(strings: list[str]) -> str:
    rotations: list[str] = []
    for s in strings:
        d = s + "$"
        for i in range(len(d)):
            # doubled so unequal lengths still break ties (bwt_util.rs:160-163)
            rotations.append(d[i:] + d + d[:i])
    rotations.sort()  # byte order: '$' < 'A' < 'C' < 'G' < 'N' < 'T'
    return "".join(r[-1] for r in rotations)


def brute_count(strings: list[str], kmer: str) -> int:
    """Occurrences of `kmer` (no '$' except possibly last) as a substring of the s+'$'."""
    total = 0
    for s in strings:
        d = s + "$"
        start = 0
        while True:
            i = d.find(kmer, start)
            if i < 0:
                break
            total += 1
            start = i + 1
    return total


def brute_rank(bwt_syms, sym: int, pos: int) -> int:
    return sum(1 for c in bwt_syms[:pos] if c == sym)
