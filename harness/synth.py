"""Harness tooling (NOT part of the product path): synthetic genomes, reads and k-mer
query sets of the shapes BASELINE.json names.

Generators are torch (Philox on CUDA, mt19937 on CPU) with the seeds below stated in
every bench line; numpy variants exist for the small committed golden fixtures so that
those are reproducible bit for bit on any box.

  genome : iid uniform ACGT, length = reads * read_len / coverage
  reads  : start uniform in [0, G - read_len], forward strand; optional independent
           per-base substitution errors to one of the other three bases
  queries: read-sampled (read id uniform, offset uniform in [0, read_len - k]) and/or
           random iid ACGT
"""
from __future__ import annotations

import numpy as np
import torch

SEED_GENOME, SEED_READS, SEED_QUERIES_READ, SEED_QUERIES_RANDOM = 0x5EED0001, 0x5EED0002, 0x5EED0003, 0x5EED0004
_ACGT = (1, 2, 3, 5)  # symbol codes, src/msbwt_core.rs:4


def _gen(device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def make_reads(n_reads: int, read_len: int = 150, coverage: float = 30.0, error_rate: float = 0.0,
               device="cpu", seed_offset: int = 0) -> torch.Tensor:
    """[n_reads, read_len] uint8 symbols."""
    dev = torch.device(device)
    glen = max(read_len, int(n_reads * read_len / coverage))
    lut = torch.tensor(_ACGT, dtype=torch.uint8, device=dev)
    base = torch.randint(0, 4, (glen,), generator=_gen(dev, SEED_GENOME + seed_offset), device=dev, dtype=torch.uint8)
    g = _gen(dev, SEED_READS + seed_offset)
    starts = torch.randint(0, glen - read_len + 1, (n_reads,), generator=g, device=dev)
    reads = torch.empty((n_reads, read_len), dtype=torch.uint8, device=dev)
    step = max(1, (1 << 26) // read_len)
    ar = torch.arange(read_len, device=dev)
    for a in range(0, n_reads, step):
        b = min(n_reads, a + step)
        idx = base[starts[a:b, None] + ar]
        if error_rate > 0:
            hit = torch.rand((b - a, read_len), generator=g, device=dev) < error_rate
            bump = torch.randint(1, 4, (b - a, read_len), generator=g, device=dev, dtype=torch.uint8)
            idx = torch.where(hit, (idx + bump) % 4, idx)
        reads[a:b] = lut[idx.long()]
    return reads


def make_queries(reads: torch.Tensor, k: int, n_read_sampled: int, n_random: int, seed_offset: int = 0) -> torch.Tensor:
    """[n, k] uint8: read-sampled queries first, then random ACGT ones, then shuffled."""
    dev = reads.device
    M, L = reads.shape
    parts = []
    if n_read_sampled:
        g = _gen(dev, SEED_QUERIES_READ + seed_offset)
        rid = torch.randint(0, M, (n_read_sampled,), generator=g, device=dev)
        off = torch.randint(0, L - k + 1, (n_read_sampled,), generator=g, device=dev)
        out = torch.empty((n_read_sampled, k), dtype=torch.uint8, device=dev)
        flat = reads.reshape(-1)
        ar = torch.arange(k, device=dev)
        step = max(1, (1 << 26) // k)
        for a in range(0, n_read_sampled, step):
            b = min(n_read_sampled, a + step)
            out[a:b] = flat[(rid[a:b] * L + off[a:b])[:, None] + ar]
        parts.append(out)
    if n_random:
        g = _gen(dev, SEED_QUERIES_RANDOM + seed_offset)
        lut = torch.tensor(_ACGT, dtype=torch.uint8, device=dev)
        parts.append(lut[torch.randint(0, 4, (n_random, k), generator=g, device=dev)])
    q = torch.cat(parts) if len(parts) > 1 else parts[0]
    if len(parts) > 1:
        perm = torch.randperm(q.shape[0], generator=_gen(dev, SEED_QUERIES_RANDOM + 17 + seed_offset), device=dev)
        q = q[perm]
    return q.contiguous()


# ---- numpy twins for committed fixtures (bit-reproducible anywhere) ----

def np_make_reads(n_reads: int, read_len: int, coverage: float, error_rate: float, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    glen = max(read_len, int(n_reads * read_len / coverage))
    lut = np.array(_ACGT, dtype=np.uint8)
    base = rng.integers(0, 4, glen, dtype=np.uint8)
    starts = rng.integers(0, glen - read_len + 1, n_reads)
    idx = base[starts[:, None] + np.arange(read_len)]
    if error_rate > 0:
        hit = rng.random((n_reads, read_len)) < error_rate
        bump = rng.integers(1, 4, (n_reads, read_len), dtype=np.uint8)
        idx = np.where(hit, (idx + bump) % 4, idx)
    return lut[idx]


def np_make_queries(reads: np.ndarray, k: int, n_read_sampled: int, n_random: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    M, L = reads.shape
    rid = rng.integers(0, M, n_read_sampled)
    off = rng.integers(0, L - k + 1, n_read_sampled)
    a = reads[rid[:, None], off[:, None] + np.arange(k)]
    lut = np.array(_ACGT, dtype=np.uint8)
    b = lut[rng.integers(0, 4, (n_random, k))]
    q = np.concatenate([a, b])
    return np.ascontiguousarray(q[rng.permutation(q.shape[0])])
