"""End-to-end host path of msbwt_count_kmers_fixed: the packed route (worker pool packs all-ACGT k-mers
2 bits per symbol, seed_packed_kernel on the device, exceptions through the byte route) must return
exactly what the byte route and the CPU oracle return -- with and without the pair image."""
import numpy as np
import pytest

import rust_msbwt_b200 as M
from oracle import oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def midsize():
    from harness import bwt_build, synth
    reads = synth.make_reads(20000, read_len=100, coverage=25.0, error_rate=0.01, device="cuda")
    reads[17, 40:43] = 4
    rle, n = bwt_build.build_rle_bwt(reads)
    o = O.RleBWT()
    o.load_vector(rle.cpu().numpy())
    return reads, o


@pytest.mark.parametrize("pair,table_s,sb_shift", [(0, -1, 0), (1, -1, 0), (1, 0, 0), (1, 3, 3), (0, 4, 2), (1, 12, 0)])
def test_packed_route_equals_byte_route_and_oracle(midsize, monkeypatch, pair, table_s, sb_shift):
    from harness import synth
    reads, o = midsize
    monkeypatch.setenv("MSBWT_HOST_THREADS", "4")
    g = M.RleBWT(pair_index=pair, suffix_table_s=table_s, superblock_shift=sb_shift)
    g.load_vector(o.rle_bytes())
    assert g.pair_index == bool(pair)
    for k in (1, 2, 7, 11, 12, 13, 31, 32, 33, 64, 65, 100):
        q = synth.make_queries(reads, k, 9001, 6000).cpu().numpy()
        q[5, 0] = 4          # exceptions: N, $ ...
        q[7, k - 1] = 0
        q[4000:4100, k // 2] = 4
        want = o.count_kmers_fixed(q, k, threads=8)
        monkeypatch.setenv("MSBWT_HOST_PACK", "1")
        got = g.count_kmers_fixed(q, k)
        h2d, d2h = M.last_transfer_bytes()
        assert (got == want).all(), (pair, table_s, k, np.flatnonzero(got != want)[:5])
        assert d2h >= q.shape[0] * 8 and h2d < q.shape[0] * (8 * -(-k // 32)) + 200 * k + 4096
        monkeypatch.setenv("MSBWT_HOST_PACK", "0")
        assert (g.count_kmers_fixed(q, k) == want).all()
        assert M.last_transfer_bytes()[0] == q.size
    # a symbol >= 6 is refused on the packed route too (through the exception list)
    monkeypatch.setenv("MSBWT_HOST_PACK", "1")
    q = synth.make_queries(reads, 31, 9001, 6000).cpu().numpy()
    q[8000, 3] = 9
    with pytest.raises(M.MsbwtError) as e:
        g.count_kmers_fixed(q, 31)
    assert e.value.code == 1
    q[8000, 3] = 1
    assert (g.count_kmers_fixed(q, 31) == o.count_kmers_fixed(q, 31, threads=8)).all()


def test_packed_route_many_chunks_and_all_exceptions(midsize, monkeypatch):
    """More queries than one pipeline chunk (2^20) so that lanes and staging buffers are reused, and a batch
    in which every k-mer is an exception."""
    from harness import synth
    reads, o = midsize
    monkeypatch.setenv("MSBWT_HOST_PACK", "1")
    g = M.RleBWT(pair_index=1)
    g.load_vector(o.rle_bytes())
    k = 31
    q = synth.make_queries(reads, k, 2_500_000, 1_000_001).cpu().numpy()
    got = g.count_kmers_fixed(q, k)
    m = 300_000
    assert (got[:m] == o.count_kmers_fixed(q[:m], k, threads=8)).all()
    assert (got[-m:] == o.count_kmers_fixed(q[-m:], k, threads=8)).all()
    monkeypatch.setenv("MSBWT_HOST_PACK", "0")
    assert (g.count_kmers_fixed(q, k) == got).all()
    monkeypatch.setenv("MSBWT_HOST_PACK", "1")
    q2 = q[:50_000].copy()
    q2[:, 30] = 4
    assert (g.count_kmers_fixed(q2, k) == o.count_kmers_fixed(q2, k, threads=8)).all()


def test_hybrid_route_with_pinned_input(midsize, monkeypatch):
    """A pinned caller buffer enables the hybrid route: chunks go either through the host pool (2 bits per
    symbol) or, whenever the copy engine is idle, over the link as raw symbol bytes packed and validated on the
    device.  Whatever the split, the counts equal the byte route's and the oracle's."""
    from harness import synth
    reads, o = midsize
    monkeypatch.setenv("MSBWT_HOST_PACK", "1")
    monkeypatch.setenv("MSBWT_HOST_THREADS", "2")  # a slow pool: the link takes a visible share
    k = 31
    q_dev = synth.make_queries(reads, k, 2_500_000, 1_000_001)
    n = q_dev.shape[0]
    pinned = torch.empty((n, k), dtype=torch.uint8, pin_memory=True)
    pinned.copy_(q_dev)
    q = pinned.numpy()
    q[5, 0] = 4                      # exceptions in the first (raw) chunk and far into the batch
    q[7, k - 1] = 0
    q[3_000_000:3_000_100, k // 2] = 4
    for quad, pair in ((0, 0), (0, 1), (1, 0)):
        g = M.RleBWT(pair_index=pair, quad_index=quad)
        g.load_vector(o.rle_bytes())
        got = g.count_kmers_fixed(q, k)
        h2d, d2h = M.last_transfer_bytes()
        assert d2h >= n * 8
        assert n * 8 < h2d < n * k + 4096, "both routes should have carried chunks"
        m = 300_000
        assert (got[:m] == o.count_kmers_fixed(q[:m], k, threads=8)).all()
        assert (got[-m:] == o.count_kmers_fixed(q[-m:], k, threads=8)).all()
        monkeypatch.setenv("MSBWT_HYBRID", "0")
        assert (g.count_kmers_fixed(q, k) == got).all()
        assert M.last_transfer_bytes()[0] < n * 8 + 200 * k + 4096
        monkeypatch.delenv("MSBWT_HYBRID")
    # a symbol >= 6 is refused whichever route its chunk takes
    q[11, 3] = 9
    with pytest.raises(M.MsbwtError) as e:
        g.count_kmers_fixed(q, k)
    assert e.value.code == 1
    q[11, 3] = 1
    q[n - 3, 3] = 7
    with pytest.raises(M.MsbwtError):
        g.count_kmers_fixed(q, k)
