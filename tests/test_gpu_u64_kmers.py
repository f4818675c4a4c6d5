"""`count_kmers_u64` (msbwt_count_kmers_u64): BWT::count_kmer (src/msbwt_core.rs:125-161) for k-mers the caller holds
as 2-bit-per-symbol integers, k <= 32 -- compared with the oracle's count_kmer on the same k-mers spelled out as
symbol bytes, over every image a query can be served from and every suffix-table depth class."""
import numpy as np
import pytest

import rust_msbwt_b200 as M
from oracle import oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

CODE = np.zeros(6, dtype=np.uint64)
CODE[[1, 2, 3, 5]] = [0, 1, 2, 3]


def encode(syms: np.ndarray) -> np.ndarray:
    """[n, k] symbol bytes in ACGT = {1,2,3,5} -> n integers, the first symbol in the most significant of the 2k bits"""
    n, k = syms.shape
    out = np.zeros(n, dtype=np.uint64)
    for j in range(k):
        out = (out << np.uint64(2)) | CODE[syms[:, j]]
    return out


@pytest.fixture(scope="module")
def midsize():
    from harness import bwt_build, synth
    reads = synth.make_reads(20000, read_len=100, coverage=25.0, error_rate=0.01, device="cuda")
    rle, _ = bwt_build.build_rle_bwt(reads)
    o = O.RleBWT()
    o.load_vector(rle.cpu().numpy())
    return reads, o


def test_encoding_helper_matches_the_documented_format():
    s = np.array([[1, 2, 3, 5]], dtype=np.uint8)          # ACGT
    assert int(encode(s)[0]) == 0b00011011
    assert int(encode(np.array([[5]], dtype=np.uint8))[0]) == 3


@pytest.mark.parametrize("opts", [
    dict(),                                                  # automatic policy (one-step blocks for this size)
    dict(suffix_table_s=0),
    dict(suffix_table_s=3, pair_index=1),
    dict(suffix_table_s=12, pair_index=1),
    dict(quad_index=1),
    dict(suffix_table_s=5, quad_index=1),
    dict(oct_index=1),
    dict(suffix_table_s=7, oct_index=1, oct_bucket_shift=18),
])
def test_u64_kmers_are_bit_exact(midsize, opts):
    from harness import synth
    reads, o = midsize
    g = M.RleBWT(**opts)
    g.load_vector(o.rle_bytes())
    for k in (1, 2, 3, 4, 7, 11, 12, 15, 16, 21, 22, 31, 32):
        q = synth.make_queries(reads, k, 9001, 6000).cpu().numpy()
        q = q[(np.isin(q, (1, 2, 3, 5))).all(axis=1)]      # integers cannot spell `$` or `N`
        got = g.count_kmers_u64(encode(q), k)
        want = o.count_kmers_fixed(q, k, threads=8)
        assert (got == want).all(), (opts, k, np.flatnonzero(got != want)[:5])
        assert (got == g.count_kmers_fixed(q, k)).all()
        if k >= 21:
            assert int((got > 0).sum()) >= 8000
    # bits above 2k are ignored
    q = synth.make_queries(reads, 31, 3000, 0).cpu().numpy()
    q = q[(np.isin(q, (1, 2, 3, 5))).all(axis=1)]
    assert (g.count_kmers_u64(encode(q) | (np.uint64(3) << np.uint64(62)), 31) == o.count_kmers_fixed(q, 31, threads=8)).all()


def test_u64_kmers_many_chunks_and_argument_checks(midsize):
    from harness import synth
    reads, o = midsize
    g = M.RleBWT(oct_index=1)
    g.load_vector(o.rle_bytes())
    q = synth.make_queries(reads, 31, 1_300_000, 300_000).cpu().numpy()   # three 512 Ki-query chunks and a tail
    q = q[(np.isin(q, (1, 2, 3, 5))).all(axis=1)]
    assert (g.count_kmers_u64(encode(q), 31) == o.count_kmers_fixed(q, 31, threads=8)).all()
    h2d, d2h = M.last_transfer_bytes()
    assert h2d == 8 * len(q) and d2h == 8 * len(q)
    got32 = g.count_kmers_u64(encode(q), 31, counts32=True)                # u32 counts: 4 bytes per query back
    assert got32.dtype == np.uint32 and (got32 == o.count_kmers_fixed(q, 31, threads=8)).all()
    h2d, d2h = M.last_transfer_bytes()
    assert h2d == 8 * len(q) and d2h == 4 * len(q)
    assert g.count_kmers_u64(np.zeros(0, dtype=np.uint64), 31).size == 0
    for bad_k in (0, 33):
        with pytest.raises(M.MsbwtError):
            g.count_kmers_u64(np.zeros(4, dtype=np.uint64), bad_k)
