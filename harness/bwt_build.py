"""Harness tooling (NOT part of the product path): builds the multi-string BWT of a set
of equal-length reads on whatever torch device the reads live on, and RLE-encodes it in
the msbwt byte format, so tests and bench.py can manufacture `.npy`-equivalent inputs
without `msbwt2-build` (Rust is not available here).

Order: exactly `naive_bwt`'s (src/bwt_util.rs:154-171) -- suffixes compared symbol by
symbol with '$' smallest, ties between equal suffixes broken by the lexicographic order
of the whole reads ("sorted insert", src/dynamic_bwt.rs:515-525).  Implemented as an LSD
radix sort over base-6 packed 24-symbol key words with stable sorts; the tie-break falls
out of starting from reads that are already sorted.

Encoding: src/bwt_converter.rs:52-56 (little-endian base-32 digits, one byte per digit).
"""
from __future__ import annotations

import torch

SYMS_PER_KEY = 24  # 6^24 < 2^63


def _pack_table(S: torch.Tensor) -> torch.Tensor:
    """S: [M, L1] uint8 reads with the trailing '$'(0).  Returns K [M, L1] int64 with
    K[r, o] = sum_{i<24, o+i<L1} S[r, o+i] * 6^(23-i)."""
    M, L1 = S.shape
    pad = torch.zeros((M, L1 + SYMS_PER_KEY), dtype=torch.uint8, device=S.device)
    pad[:, :L1] = S
    K = torch.zeros((M, L1), dtype=torch.int64, device=S.device)
    for i in range(SYMS_PER_KEY):
        K += pad[:, i:i + L1].to(torch.int64) * (6 ** (SYMS_PER_KEY - 1 - i))
    return K


def _stable_argsort(keys: torch.Tensor) -> torch.Tensor:
    return torch.sort(keys, stable=True)[1]


def sort_reads(reads: torch.Tensor) -> torch.Tensor:
    """Lexicographic order of equal-length reads (permutation)."""
    M, L = reads.shape
    words = (L + SYMS_PER_KEY - 1) // SYMS_PER_KEY
    perm = torch.arange(M, device=reads.device)
    for w in range(words - 1, -1, -1):
        chunk = reads[:, w * SYMS_PER_KEY:(w + 1) * SYMS_PER_KEY].to(torch.int64)
        key = torch.zeros(M, dtype=torch.int64, device=reads.device)
        for i in range(chunk.shape[1]):
            key += chunk[:, i] * (6 ** (SYMS_PER_KEY - 1 - i))
        idx = _stable_argsort(key[perm])
        perm = perm[idx]
    return perm


def build_msbwt(reads: torch.Tensor) -> torch.Tensor:
    """reads: [M, L] uint8 symbols in 1..5 (A,C,G,N,T).  Returns the BWT as a uint8
    symbol tensor of length M*(L+1) on the same device."""
    assert reads.dtype == torch.uint8 and reads.dim() == 2
    M, L = reads.shape
    L1 = L + 1
    dev = reads.device
    if M == 0:
        return torch.zeros(0, dtype=torch.uint8, device=dev)
    S = torch.zeros((M, L1), dtype=torch.uint8, device=dev)
    S[:, :L] = reads[sort_reads(reads)]
    K = _pack_table(S).reshape(-1)
    N = M * L1
    perm = torch.arange(N, device=dev)           # suffix id = r*L1 + l, already in tie-break order
    words = (L1 + SYMS_PER_KEY - 1) // SYMS_PER_KEY
    for w in range(words - 1, -1, -1):
        off = perm % L1 + w * SYMS_PER_KEY        # offset of this key word inside the read
        inside = off < L1
        key = torch.where(inside, K[torch.where(inside, perm + w * SYMS_PER_KEY, perm)], torch.zeros_like(perm))
        del off, inside
        idx = _stable_argsort(key)
        del key
        perm = perm[idx]
        del idx
    del K
    flat = S.reshape(-1)
    prev = torch.where(perm % L1 == 0, perm, perm - 1)
    bwt = torch.where(perm % L1 == 0, torch.zeros(1, dtype=torch.uint8, device=dev), flat[prev])
    return bwt


def rle_encode(bwt: torch.Tensor) -> torch.Tensor:
    """uint8 symbol tensor -> msbwt RLE byte tensor (uint8), same device."""
    n = bwt.numel()
    dev = bwt.device
    if n == 0:
        return torch.zeros(0, dtype=torch.uint8, device=dev)
    change = torch.ones(n, dtype=torch.bool, device=dev)
    change[1:] = bwt[1:] != bwt[:-1]
    starts = torch.nonzero(change).reshape(-1)
    ends = torch.cat([starts[1:], torch.tensor([n], device=dev, dtype=starts.dtype)])
    lens = ends - starts
    syms = bwt[starts].to(torch.int64)
    nd = torch.ones_like(lens)
    lim = 32
    while bool((lens >= lim).any()):
        nd += (lens >= lim).to(nd.dtype)
        lim *= 32
    offs = torch.cumsum(nd, 0) - nd
    out = torch.zeros(int(nd.sum().item()), dtype=torch.uint8, device=dev)
    for j in range(int(nd.max().item())):
        m = nd > j
        out[offs[m] + j] = (syms[m] | (((lens[m] >> (5 * j)) & 31) << 3)).to(torch.uint8)
    return out


def build_rle_bwt(reads: torch.Tensor) -> tuple[torch.Tensor, int]:
    """(RLE bytes on the reads' device, total symbol count)."""
    bwt = build_msbwt(reads)
    return rle_encode(bwt), int(bwt.numel())
