"""The OCT image (layout.h: one 128-byte line of explicit occurrence runs per (m-symbol code, 2^b-position
bucket), m = oct_symbols() = 10; one line answers m constrain_range steps, src/rle_bwt.rs:202-287 composed m
times) and the kernel that walks it, falling back to quad and one-symbol steps where a line overflowed.

  * the image built on the device is compared with a numpy brute-force construction from the decoded BWT
    (LF by counting, m-symbol codes, runs cut at chunk boundaries, per-bucket run sets, checkpoints);
  * every count through the oct path must equal the CPU oracle's (the reference's algorithm), for every
    suffix-table depth / k remainder combination, with $ / N inside the k-mers, on typical read sets and on
    low-complexity ones where most lines overflow."""
import numpy as np
import pytest

import rust_msbwt_b200 as M
from oracle import naive
from oracle import oracle as O
from tests.test_gpu_pair_index import _random_rle, decode

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

ACGT = np.array([1, 2, 3, 5])
CAP = 30
M_SYMS = None  # oct_symbols(), read from the library on first use


def m_syms() -> int:
    global M_SYMS
    if M_SYMS is None:
        M_SYMS = M.oct_symbols()
    return M_SYMS


def brute_oct(bwt: np.ndarray):
    """(m-symbol code per position or -1, Cm[4^m])"""
    m = m_syms()
    n = bwt.size
    cnt = np.bincount(bwt, minlength=6)
    cstart = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    lf = np.zeros(n, dtype=np.int64)
    occ = np.zeros((6, n + 1), dtype=np.int64)
    for s in range(6):
        is_s = bwt == s
        occ[s, 1:] = np.cumsum(is_s)
        at = np.flatnonzero(is_s)
        lf[at] = cstart[s] + np.arange(at.size)
    idx = np.full(8, -1)
    idx[ACGT] = np.arange(4)
    j = np.arange(n)
    code = np.zeros(n, dtype=np.int64)
    valid = np.ones(n, dtype=bool)
    for _ in range(m):
        b = idx[bwt[j]] if n else np.zeros(0, dtype=np.int64)
        valid &= b >= 0
        code = code * 4 + np.maximum(b, 0)
        j = lf[j] if n else j
    codes = np.arange(4 ** m)
    pos = np.zeros(4 ** m, dtype=np.int64)
    for r in range(m):
        sym = ACGT[(codes >> (2 * (m - 1 - r))) & 3]
        pos = cstart[sym] + occ[sym, pos]
    return np.where(valid, code, -1), pos


def auto_shift(runs: int, n: int) -> int:
    """layout.h: the largest bucket shift in 16..24 that keeps the mean number of runs per line <= 6"""
    s = 24
    while s > 16 and runs * (1 << s) > 6 * (4 ** m_syms()) * max(n, 1):
        s -= 1
    return s


def check_oct_image(g, bwt, wide=False):
    """wide: the index has 64-bit positions -- word 1 of a line = min(runs, 31) | (checkpoint >> 32) << 8 (layout.h)"""
    lines = g.oct_image()
    n = bwt.size
    shift = g.oct_bucket_shift
    cs = min(31 - shift, 10, shift)
    nb = (n >> shift) + 1
    ncodes = 4 ** m_syms()
    assert lines.shape == (ncodes, nb, 32)
    code, c8 = brute_oct(bwt)
    # runs of equal codes, cut at every multiple of 2^cs (so no run crosses a bucket, len <= 2^cs)
    bound = np.ones(n, dtype=bool)
    bound[1:] = code[1:] != code[:-1]
    bound |= (np.arange(n) & ((1 << cs) - 1)) == 0
    starts = np.flatnonzero(bound)
    lens = np.diff(np.r_[starts, n])
    keep = code[starts] >= 0 if n else np.zeros(0, dtype=bool)
    starts, lens = starts[keep], lens[keep]
    rc, rb = code[starts], starts >> shift
    plain_heads = int((bound & (code >= 0) & np.r_[True, code[1:] != code[:-1]]).sum()) if n else 0
    assert g.oct_runs == plain_heads
    per = np.zeros((ncodes, nb), dtype=np.int64)
    nruns = np.zeros((ncodes, nb), dtype=np.int64)
    np.add.at(per, (rc, rb), lens)
    np.add.at(nruns, (rc, rb), 1)
    before = np.cumsum(per, axis=1) - per
    assert (lines[:, :, 1] == (np.minimum(nruns, CAP + 1) if wide else nruns)).all()  # (checkpoints below 2^32 here)
    assert (lines[:, :, 0] == (before + c8[:, None]).astype(np.uint32)).all()
    over = nruns > CAP
    assert g.oct_overflow_lines == int(over.sum())
    assert g.oct_overflow_occurrences == int(per[over].sum())
    # stored runs `(len << shift) | offset`: every non-overflowed line holds exactly its runs, the rest is 0
    ent = np.sort(lines[:, :, 2:], axis=2)
    entry = ((lens << shift) | (starts & ((1 << shift) - 1))).astype(np.uint32)
    want = np.zeros((ncodes, nb, CAP), dtype=np.uint32)
    order = np.lexsort((entry, rb, rc))
    c_s, b_s, e_s = rc[order], rb[order], entry[order]
    first = np.flatnonzero(np.r_[True, (c_s[1:] != c_s[:-1]) | (b_s[1:] != b_s[:-1])]) if e_s.size else np.zeros(0, int)
    rank_in_line = np.arange(e_s.size) - np.repeat(first, np.diff(np.r_[first, e_s.size])) if e_s.size else np.zeros(0, int)
    ok_run = ~over[c_s, b_s] if e_s.size else np.zeros(0, dtype=bool)
    want[c_s[ok_run], b_s[ok_run], rank_in_line[ok_run]] = e_s[ok_run]
    want.sort(axis=2)
    assert (ent[~over] == want[~over]).all()
    # overflowed lines hold some 30 of their runs (which ones depends on the emit kernel's timing)
    for c, b in np.argwhere(over)[:50]:
        mine = set(entry[(rc == c) & (rb == b)].tolist())
        assert set(ent[c, b].tolist()) <= mine and len(set(ent[c, b].tolist())) == CAP


@pytest.mark.parametrize("shift", [0, 16, 17])
def test_oct_image_equals_brute_force(shift):
    rng = np.random.default_rng(2020)
    from harness import bwt_build, synth
    reads = synth.make_reads(400, 60, 15.0, 0.02, device="cuda")
    reads[3, 10:12] = 4
    low = synth.make_reads(3000, 50, 600.0, 0.03, device="cuda")   # 250-base genome, 600x with errors: the lines of its codes overflow
    streams = [
        O.convert_to_vec(naive.naive_bwt(["CCGTACGTA", "GGTACAGTA", "ACGACGACG", "ANNT"])),
        _random_rle(rng, 3000, [1, 1, 1, 2, 3, 5, 9, 31, 32, 33, 95, 96, 97, 223, 224, 225, 255]),
        O.convert_to_vec("ACGT" * 56 + "T"),
        bwt_build.build_rle_bwt(reads)[0].cpu().numpy(),
        bwt_build.build_rle_bwt(low)[0].cpu().numpy(),
        np.zeros(0, np.uint8),
    ]
    for rle in streams:
        g = M.RleBWT(oct_index=1, oct_bucket_shift=shift)
        g.load_vector(rle)
        assert g.oct_index and g.quad_index and not g.pair_index
        assert g.oct_bucket_shift == (shift or auto_shift(g.oct_runs, g.get_total_size()))
        bwt = decode(np.asarray(rle, dtype=np.uint8))
        assert bwt.size == g.get_total_size()
        check_oct_image(g, bwt)


@pytest.fixture(scope="module")
def midsize():
    from harness import bwt_build, synth
    reads = synth.make_reads(20000, read_len=100, coverage=25.0, error_rate=0.01, device="cuda")
    reads[17, 40:43] = 4  # a few N
    rle, n = bwt_build.build_rle_bwt(reads)
    o = O.RleBWT()
    o.load_vector(rle.cpu().numpy())
    return reads, o


def test_oct_image_over_several_buckets(midsize):
    reads, o = midsize
    g = M.RleBWT(oct_index=1, oct_bucket_shift=20)
    g.load_vector(o.rle_bytes())
    assert g.oct_index and g.oct_bucket_shift == 20 and (g.get_total_size() >> 20) >= 1
    check_oct_image(g, decode(o.rle_bytes()))


@pytest.mark.parametrize("table_s,shift", [(-1, 0), (0, 18), (1, 17), (2, 0), (3, 19), (4, 24), (5, 20), (7, 18), (11, 19), (12, 20)])
def test_oct_path_is_bit_exact(midsize, table_s, shift):
    from harness import synth
    reads, o = midsize
    g = M.RleBWT(suffix_table_s=table_s, oct_index=1, oct_bucket_shift=shift)
    g.load_vector(o.rle_bytes())
    assert g.oct_index
    rng = np.random.default_rng(11 + 10 * (table_s + 1))
    for k in (1, 2, 3, 4, 5, 7, 8, 9, 11, 12, 13, 15, 16, 17, 23, 24, 25, 31, 32, 33, 39, 40, 41, 47, 48, 63, 64, 65, 66, 67, 71, 72, 100):
        q = synth.make_queries(reads, k, 12001, 8000).cpu().numpy()
        q[5, 0] = 4
        q[7, k - 1] = 0
        q[11, k // 2] = 4
        got = g.count_kmers_fixed(q, k)
        want = o.count_kmers_fixed(q, k, threads=8)
        assert (got == want).all(), (table_s, k, np.flatnonzero(got != want)[:5])
        if k >= 29:
            assert int((got > 0).sum()) >= 11990
    ragged = [rng.integers(0, 6, int(rng.integers(0, 40))).astype(np.uint8) for _ in range(3000)]
    assert (g.count_kmers(ragged) == o.count_kmers(ragged)).all()


def test_oct_path_on_low_complexity_reads_where_lines_overflow(monkeypatch):
    """A 2 kb genome at 1500x: a handful of codes own every position of a bucket, most lines in use hold far
    more than 30 runs and are answered through the quad image -- the counts must not notice."""
    from harness import bwt_build, synth
    reads = synth.make_reads(30000, read_len=100, coverage=1500.0, error_rate=0.01, device="cuda")
    rle = bwt_build.build_rle_bwt(reads)[0].cpu().numpy()
    o = O.RleBWT()
    o.load_vector(rle)
    g = M.RleBWT(oct_index=1)
    g.load_vector(rle)
    assert g.oct_index and g.oct_overflow_lines > 1000 and g.oct_overflow_occurrences > g.get_total_size() // 4
    for k in (8, 16, 31, 32, 41, 64):
        q = synth.make_queries(reads, k, 20000, 5000).cpu().numpy()
        assert (g.count_kmers_fixed(q, k) == o.count_kmers_fixed(q, k, threads=8)).all(), k
    monkeypatch.setenv("MSBWT_HOST_PACK", "1")
    q = synth.make_queries(reads, 31, 50000, 5000).cpu().numpy()
    assert (g.count_kmers_fixed(q, 31) == o.count_kmers_fixed(q, 31, threads=8)).all()


@pytest.mark.parametrize("shift", [0, 17])
def test_wide_oct_image_equals_brute_force(shift):
    """64-bit positions (here: superblocks of 1024 symbols): the codes come from walking LF through the one-step
    blocks (no pair / quad image on the way), the lines carry 40-bit checkpoints"""
    rng = np.random.default_rng(2021)
    from harness import bwt_build, synth
    reads = synth.make_reads(400, 60, 15.0, 0.02, device="cuda")
    reads[3, 10:12] = 4
    low = synth.make_reads(3000, 50, 600.0, 0.03, device="cuda")
    streams = [
        O.convert_to_vec(naive.naive_bwt(["CCGTACGTA", "GGTACAGTA", "ACGACGACG", "ANNT"])),
        _random_rle(rng, 3000, [1, 1, 1, 2, 3, 5, 9, 31, 32, 33, 95, 96, 97, 223, 224, 225, 255]),
        bwt_build.build_rle_bwt(reads)[0].cpu().numpy(),
        bwt_build.build_rle_bwt(low)[0].cpu().numpy(),
    ]
    for rle in streams:
        g = M.RleBWT(oct_index=1, oct_bucket_shift=shift, superblock_shift=3)
        g.load_vector(rle)
        wide = (g.get_total_size() >> 7) + 1 > 8  # more than one superblock of 8 blocks (the 4-string BWT is not)
        assert g.oct_index and g.quad_index == (not wide) and not g.pair_index
        assert g.oct_bucket_shift == (shift or auto_shift(g.oct_runs, g.get_total_size()))
        bwt = decode(np.asarray(rle, dtype=np.uint8))
        check_oct_image(g, bwt, wide=wide)


@pytest.mark.parametrize("table_s,shift,final", [(-1, 0, -1), (0, 18, 0), (3, 19, 1), (7, 18, 0), (11, 19, -1), (12, 20, 0), (14, 17, 1)])
def test_wide_oct_path_is_bit_exact(midsize, table_s, shift, final):
    """the WIDE instantiation of the oct kernel (wide_kernels.cu) against the reference's algorithm, with and
    without final-step lines, every table depth / remainder combination, `$` / N inside the k-mers"""
    from harness import synth
    reads, o = midsize
    g = M.RleBWT(suffix_table_s=table_s, oct_index=1, oct_bucket_shift=shift, final_index=final, superblock_shift=4)
    g.load_vector(o.rle_bytes())
    assert g.oct_index and not g.quad_index and g.final_index == (final != 0)
    rng = np.random.default_rng(12 + 10 * (table_s + 1))
    for k in (1, 2, 3, 9, 10, 11, 13, 20, 21, 24, 25, 30, 31, 32, 33, 34, 35, 40, 41, 44, 45, 51, 54, 63, 64, 65, 71, 100):
        q = synth.make_queries(reads, k, 12001, 8000).cpu().numpy()
        q[5, 0] = 4
        q[7, k - 1] = 0
        q[11, k // 2] = 4
        got = g.count_kmers_fixed(q, k)
        want = o.count_kmers_fixed(q, k, threads=8)
        assert (got == want).all(), (table_s, k, np.flatnonzero(got != want)[:5])
        if k >= 29:
            assert int((got > 0).sum()) >= 11990
        if k <= 32:
            packed = np.zeros(q.shape[0], dtype=np.uint64)
            ok = np.isin(q, ACGT).all(axis=1)
            idx = np.zeros(8, dtype=np.uint64)
            idx[ACGT] = np.arange(4, dtype=np.uint64)
            for i in range(k):
                packed = (packed << np.uint64(2)) | idx[q[:, i]]
            assert (g.count_kmers_u64(packed[ok], k) == want[ok]).all(), (table_s, k)
    ragged = [rng.integers(0, 6, int(rng.integers(0, 40))).astype(np.uint8) for _ in range(3000)]
    assert (g.count_kmers(ragged) == o.count_kmers(ragged)).all()


def test_wide_oct_path_on_low_complexity_reads_where_lines_overflow():
    """most lines in use overflow: the WIDE kernel answers them with one-symbol steps (no quad image beside it)"""
    from harness import bwt_build, synth
    reads = synth.make_reads(30000, read_len=100, coverage=1500.0, error_rate=0.01, device="cuda")
    rle = bwt_build.build_rle_bwt(reads)[0].cpu().numpy()
    o = O.RleBWT()
    o.load_vector(rle)
    g = M.RleBWT(oct_index=1, superblock_shift=5)
    g.load_vector(rle)
    assert g.oct_index and g.final_index and g.oct_overflow_lines > 1000
    for k in (8, 16, 31, 32, 41, 64):
        q = synth.make_queries(reads, k, 20000, 5000).cpu().numpy()
        assert (g.count_kmers_fixed(q, k) == o.count_kmers_fixed(q, k, threads=8)).all(), k


def test_oct_on_golden_fixture_also_with_64_bit_positions(golden_dir):
    z = np.load(f"{golden_dir}/reads30x_k31.npz")
    g = M.RleBWT(oct_index=1)
    g.load_vector(z["rle"])
    assert g.oct_index
    assert (g.count_kmers_fixed(z["queries"], int(z["k"])) == z["counts"]).all()
    assert (g.count_kmers_fixed(z["queries_k12"], 12) == z["counts_k12"]).all()
    # 64-bit positions (several superblocks): oct and final-step lines built by the LF walk, no quad image
    w = M.RleBWT(oct_index=1, superblock_shift=3)
    w.load_vector(z["rle"])
    assert w.oct_index and w.final_index and not w.quad_index
    assert (w.count_kmers_fixed(z["queries"], int(z["k"])) == z["counts"]).all()
    assert (w.count_kmers_fixed(z["queries_k12"], 12) == z["counts_k12"]).all()
    # ... and with the oct image switched off the quad image serves alone, as before
    w4 = M.RleBWT(oct_index=0, quad_index=1, superblock_shift=3)
    w4.load_vector(z["rle"])
    assert w4.quad_index and not w4.oct_index
    assert (w4.count_kmers_fixed(z["queries"], int(z["k"])) == z["counts"]).all()
