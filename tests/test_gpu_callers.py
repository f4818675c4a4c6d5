"""Batched callers of the path: what a correction / assembly loop around msbwt2 does with
`constrain_range` and `count_kmer`, as batches through the C ABI.

  * `constrain_ranges_fanout`: for every range the four calls RleBWT::constrain_range(sym, [l,h)) for
    sym = A, C, G, T (src/rle_bwt.rs:202-287) -- compared with the oracle call by call, at every kind of
    boundary (block edges, l == h, [0,N), 64-bit positions);
  * `count_read_kmers`: BWT::count_kmer (src/msbwt_core.rs:125-161) of every k-mer window of every read,
    optionally summed with the count of its reverse complement (string_util::reverse_complement_i,
    src/string_util.rs:45-50) -- compared with the oracle's count_kmer on the windows laid out in numpy."""
import numpy as np
import pytest

import rust_msbwt_b200 as M
from oracle import naive
from oracle import oracle as O
from tests.test_gpu_pair_index import _random_rle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

ACGT = (1, 2, 3, 5)


def oracle_fanout(o, l, h):
    out_l = np.zeros((len(l), 4), dtype=np.uint64)
    out_h = np.zeros((len(l), 4), dtype=np.uint64)
    for i, (a, b) in enumerate(zip(l, h)):
        for j, s in enumerate(ACGT):
            out_l[i, j], out_h[i, j] = o.constrain_range(s, int(a), int(b))
    return out_l, out_h


@pytest.mark.parametrize("sb_shift", [0, 3])
def test_fanout_equals_four_constrain_range_calls(golden_dir, sb_shift):
    rng = np.random.default_rng(77)
    z = np.load(f"{golden_dir}/reads30x_k31.npz")
    streams = [
        z["rle"],
        O.convert_to_vec(naive.naive_bwt(["CCGTACGTA", "GGTACAGTA", "ACGACGACG", "ANNT"])),
        _random_rle(rng, 4000, [1, 1, 2, 3, 5, 9, 31, 32, 33, 127, 128, 129, 255, 1000]),
    ]
    for rle in streams:
        o = O.RleBWT()
        o.load_vector(rle)
        g = M.RleBWT(superblock_shift=sb_shift)
        g.load_vector(rle)
        n = g.get_total_size()
        l = rng.integers(0, n + 1, 3000)
        h = l + np.minimum(rng.integers(0, 300, 3000) * (rng.random(3000) < 0.8), n - l)
        edges = np.array([[0, 0], [0, n], [n, n], [127, 128], [128, 128], [128, 129], [0, 1], [n - 1, n], [n // 2, n]])
        edges = np.clip(edges, 0, n)
        l = np.concatenate([l, edges[:, 0], np.arange(0, min(n, 600))])
        h = np.concatenate([h, np.maximum(edges[:, 0], edges[:, 1]), np.minimum(np.arange(0, min(n, 600)) + 130, n)])
        got_l, got_h = g.constrain_ranges_fanout(l, h)
        want_l, want_h = oracle_fanout(o, l, h)
        assert (got_l == want_l).all() and (got_h == want_h).all()
        # the same through the one-symbol entry point: fan-out is exactly four constrain_range calls
        for j, s in enumerate(ACGT):
            a, b = g.constrain_ranges(np.full(l.size, s, np.uint8), l, h)
            assert (a == got_l[:, j]).all() and (b == got_h[:, j]).all()


def test_fanout_refuses_bad_ranges(golden_dir):
    z = np.load(f"{golden_dir}/reads30x_k31.npz")
    g = M.RleBWT()
    g.load_vector(z["rle"])
    n = g.get_total_size()
    with pytest.raises(M.MsbwtError):
        g.constrain_ranges_fanout([5, 9], [6, 8])       # l > h
    with pytest.raises(M.MsbwtError):
        g.constrain_ranges_fanout([0], [n + 1])         # h > N
    a, b = g.constrain_ranges_fanout([], [])
    assert a.shape == (0, 4) and b.shape == (0, 4)


@pytest.fixture(scope="module")
def readset():
    from harness import bwt_build, synth
    reads = synth.make_reads(20000, read_len=100, coverage=25.0, error_rate=0.01, device="cuda")
    reads[17, 40:43] = 4   # a few N
    rle, n = bwt_build.build_rle_bwt(reads)
    o = O.RleBWT()
    o.load_vector(rle.cpu().numpy())
    g = M.RleBWT()
    g.load_vector(o.rle_bytes())
    return reads.cpu().numpy(), o, g


def windows_of(reads: np.ndarray, k: int) -> np.ndarray:
    return np.ascontiguousarray(np.lib.stride_tricks.sliding_window_view(reads, k, axis=1)).reshape(-1, k)


@pytest.mark.parametrize("k", [1, 5, 16, 31, 32, 33, 64, 100])
def test_count_read_kmers_equals_count_kmer_of_every_window(readset, k):
    reads, o, g = readset
    sub = reads[:1500]
    w = windows_of(sub, k)
    want = o.count_kmers_fixed(w, k, threads=8).reshape(sub.shape[0], -1)
    got = g.count_read_kmers(sub, k)
    assert got.shape == want.shape and (got == want).all()
    assert (got >= 1).all()   # every window of a read of the set occurs in it
    rc = M.reverse_complement_i  # element-wise on rows: reverse + complement
    w_rc = np.stack([rc(row) for row in w[:: max(1, w.shape[0] // 20000)]])
    want_rc = o.count_kmers_fixed(w_rc, k, threads=8)
    both = g.count_read_kmers(sub, k, both_strands=True).reshape(-1)
    assert (both[:: max(1, w.shape[0] // 20000)] == want.reshape(-1)[:: max(1, w.shape[0] // 20000)] + want_rc).all()


def test_count_read_kmers_over_several_chunks_and_foreign_reads(readset):
    reads, o, g = readset
    rng = np.random.default_rng(5)
    foreign = rng.choice(np.array([1, 2, 3, 5], dtype=np.uint8), size=(500, 100))   # reads that are not in the set
    foreign[3, 50] = 0
    foreign[4, 10:12] = 4
    batch = np.concatenate([reads, foreign])                    # 20500 reads x 70 windows x 2 strands = 2.9 M queries
    k = 31
    got = g.count_read_kmers(batch, k, both_strands=True)
    w = windows_of(batch, k)
    sel = rng.choice(w.shape[0], 200000, replace=False)
    fwd = o.count_kmers_fixed(w[sel], k, threads=8)
    rc = np.stack([M.reverse_complement_i(r) for r in w[sel]])
    rev = o.count_kmers_fixed(rc, k, threads=8)
    assert (got.reshape(-1)[sel] == fwd + rev).all()
    assert (g.count_read_kmers(batch[-500:], k) == o.count_kmers_fixed(windows_of(batch[-500:], k), k, threads=8).reshape(500, -1)).all()


def test_count_read_kmers_refuses_bad_input(readset):
    reads, o, g = readset
    with pytest.raises(M.MsbwtError):
        g.count_read_kmers(reads[:4], 0)
    with pytest.raises(M.MsbwtError):
        g.count_read_kmers(reads[:4], 101)
    bad = reads[:4].copy()
    bad[2, 7] = 6
    with pytest.raises(M.MsbwtError):
        g.count_read_kmers(bad, 31)
    assert g.count_read_kmers(reads[:0], 31).shape == (0, 70)


def test_ragged_batches_of_one_length_equal_the_fixed_call(readset):
    """`count_kmers(&[Vec<u8>])` with k-mers of one length (the usual k-mer counting loop) is routed to the
    fixed-k path (suffix table, packed queries); a batch of mixed lengths keeps the byte-wise path."""
    reads, o, g = readset
    q = windows_of(reads[:300], 31).copy()
    q[5, 0] = 4
    q[7, 30] = 0
    kmers = [row for row in q]
    want = o.count_kmers_fixed(q, 31, threads=8)
    assert (g.count_kmers(kmers) == want).all()
    assert (g.count_kmers_fixed(q, 31) == want).all()
    mixed = kmers[:5000] + [q[0][:17], q[1][:1], q[2][:0]] + kmers[5000:6000]
    assert (g.count_kmers(mixed) == o.count_kmers(mixed, threads=8)).all()
    one = [q[3]]
    assert g.count_kmers(one)[0] == want[3]
