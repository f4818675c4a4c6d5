"""The one-request path (final_kernels.cu, pack_seed_final_kernel): k-mers that are a suffix-table entry + exactly 20
symbols (k = 31 / 32 on an index with a final-step image) are packed, seeded and answered from ONE final-step line by
one regular kernel; what it cannot answer (range over two buckets, overflowed line, `$` / `N`) is left to the general
kernels.  Every entry point that reaches it must return BWT::count_kmer's counts (src/msbwt_core.rs:125-161) bit for
bit: symbol bytes, caller-packed integers, host-packed words; whatever the batch size, whatever the fallback rate."""
import numpy as np
import pytest

import rust_msbwt_b200 as M
from oracle import oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

CODE = np.zeros(6, dtype=np.uint64)
CODE[[1, 2, 3, 5]] = [0, 1, 2, 3]


def encode(syms: np.ndarray) -> np.ndarray:
    out = np.zeros(syms.shape[0], dtype=np.uint64)
    for j in range(syms.shape[1]):
        out = (out << np.uint64(2)) | CODE[syms[:, j]]
    return out


@pytest.fixture(scope="module")
def midsize():
    from harness import bwt_build, synth
    reads = synth.make_reads(20000, read_len=100, coverage=25.0, error_rate=0.01, device="cuda")
    reads[17, 40:43] = 4
    rle, _ = bwt_build.build_rle_bwt(reads)
    o = O.RleBWT()
    o.load_vector(rle.cpu().numpy())
    return reads, o


def pack_stats_of(g, q, k):
    """runs the pack stage on a device copy of q and returns its counters + the counts after the search stage"""
    n = q.shape[0]
    dq = torch.from_numpy(q).cuda()
    d_packed = torch.empty(g.packed_bytes(k, n) // 8, dtype=torch.int64, device="cuda")
    d_out = torch.zeros(n, dtype=torch.int64, device="cuda")
    d_status = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    g.pack_kmers_device(dq.data_ptr(), k, n, d_packed.data_ptr(), d_out.data_ptr(), d_status.data_ptr(), st)
    stats = g.pack_stats(d_packed.data_ptr(), k, n)
    g.count_kmers_packed_device(d_packed.data_ptr(), k, n, d_out.data_ptr(), st)
    torch.cuda.synchronize()
    return stats, d_out.cpu().numpy().view(np.uint64), int(d_status.item())


@pytest.mark.parametrize("opts", [
    dict(),                                               # shift 16: one bucket for this index, no two-bucket ranges
    dict(final_bucket_shift=8),                           # 256-position buckets: many ranges span two
    dict(final_bucket_shift=10, final_lines_log2=12, keep_quad_index=0),
    dict(final_bucket_shift=12, suffix_table_s=12),       # deepest table level 12: k = 32 starts from it, k = 31 from 11
    dict(superblock_shift=4),                             # 64-bit positions (superblocks of 2048 symbols): the WIDE
    dict(superblock_shift=4, final_bucket_shift=8),       # instantiation -- 16-byte table entries, two seed words
])
def test_one_request_path_is_bit_exact(midsize, monkeypatch, opts):
    from harness import synth
    reads, o = midsize
    g = M.RleBWT(oct_index=1, final_index=1, **opts)
    g.load_vector(o.rle_bytes())
    assert g.final_index
    for k in (31, 32):
        for n_read, n_rand in ((1, 0), (20, 11), (33, 0), (4097, 3001), (150_001, 50_000)):
            q = synth.make_queries(reads, k, n_read, n_rand).cpu().numpy()
            if q.shape[0] > 100:
                q[5, 0] = 4          # exceptions: N, $ ... (list B)
                q[7, k - 1] = 0
                q[40:50, k // 2] = 4
            want = o.count_kmers_fixed(q, k, threads=8)
            stats, got, status = pack_stats_of(g, q, k)
            assert status == 0 and (got == want).all(), (opts, k, q.shape[0], np.flatnonzero(got != want)[:5])
            if g.table_depth_for_k(k) == k - 20:
                assert stats["final_lines"] > 0 or q.shape[0] < 50, "the one-request path did not run"
                if opts.get("final_bucket_shift", 16) <= 10 and q.shape[0] > 10000:
                    assert stats["two_buckets"] > 0 and stats["live_a"] >= stats["two_buckets"]
            # the same through the host entry points: byte route, u32 counts, caller-packed integers
            monkeypatch.setenv("MSBWT_HOST_PACK", "0")
            assert (g.count_kmers_fixed(q, k) == want).all()
            assert (g.count_kmers_fixed(q, k, counts32=True) == want).all()
            acgt = np.isin(q, (1, 2, 3, 5)).all(axis=1)
            assert (g.count_kmers_u64(encode(q[acgt]), k) == want[acgt]).all()
            if q.shape[0] >= 4096:   # host-packed words (seed_packed's input format)
                monkeypatch.setenv("MSBWT_HOST_PACK", "1")
                monkeypatch.setenv("MSBWT_HOST_THREADS", "4")
                assert (g.count_kmers_fixed(q, k) == want).all()
    # the general kernels on the same index (MSBWT_FINAL_FAST=0) agree, and their pack stage fetches no lines
    monkeypatch.setenv("MSBWT_FINAL_FAST", "0")
    q = synth.make_queries(reads, 31, 30_000, 10_000).cpu().numpy()
    stats, got, _ = pack_stats_of(g, q, 31)
    assert stats["final_lines"] == 0 and (got == o.count_kmers_fixed(q, 31, threads=8)).all()


def test_one_request_path_refuses_bad_symbols(midsize):
    from harness import synth
    reads, o = midsize
    g = M.RleBWT(oct_index=1)
    g.load_vector(o.rle_bytes())
    q = synth.make_queries(reads, 31, 10_000, 0).cpu().numpy()
    q[7777, 13] = 6
    stats, _, status = pack_stats_of(g, q, 31)
    assert status != 0
    with pytest.raises(M.MsbwtError) as e:
        g.count_kmers_fixed(q, 31)
    assert e.value.code == 1


def test_one_request_path_with_overflowed_lines():
    """a 2 kb genome at 1500x: few codes own every position, their final-step lines cannot hold the runs and the
    queries come back through the oct image (flagged: the oct kernel does not fetch the line again)"""
    from harness import bwt_build, synth
    reads = synth.make_reads(30000, read_len=100, coverage=1500.0, error_rate=0.01, device="cuda")
    rle = bwt_build.build_rle_bwt(reads)[0].cpu().numpy()
    o = O.RleBWT()
    o.load_vector(rle)
    for keep_quad, sb in ((1, 0), (0, 0), (-1, 5)):   # (sb = 5: 64-bit positions, the WIDE kernels, no quad image)
        g = M.RleBWT(oct_index=1, final_index=1, keep_quad_index=keep_quad, superblock_shift=sb)
        g.load_vector(rle)
        for k in (31, 32):
            q = synth.make_queries(reads, k, 60_000, 5_000).cpu().numpy()
            stats, got, status = pack_stats_of(g, q, k)
            assert status == 0 and stats["final_overflowed"] > 0 and stats["live_a"] >= stats["final_overflowed"]
            assert (got == o.count_kmers_fixed(q, k, threads=8)).all(), (keep_quad, k)
