"""The QUAD image (layout.h: one occurrence bit-vector per 4-symbol code, in self-contained 32-byte
sectors of 224 BWT positions; one sector answers FOUR constrain_range steps, src/rle_bwt.rs:202-287
composed four times) and the thread-per-query kernel that walks it.

  * the image built on the device is compared word for word with a numpy brute-force construction
    from the decoded BWT (LF by counting, four-symbol codes, occurrence bits, checkpoints);
  * every count through the quad path must equal the CPU oracle's (the reference's algorithm), for
    every suffix-table depth / k remainder combination, with $ / N inside the k-mers, with 32-bit and
    64-bit positions (small superblocks), through the host-packed and byte routes."""
import numpy as np
import pytest

import rust_msbwt_b200 as M
from oracle import naive
from oracle import oracle as O
from tests.test_gpu_pair_index import _random_rle, decode

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

ACGT = np.array([1, 2, 3, 5])


def brute_quad_image(bwt: np.ndarray, sb_shift4: int, wide: bool):
    n = bwt.size
    cnt = np.bincount(bwt, minlength=6)
    cstart = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    lf = np.zeros(n, dtype=np.int64)
    for s in range(6):
        at = np.flatnonzero(bwt == s)
        lf[at] = cstart[s] + np.arange(at.size)
    idx = np.full(8, -1)
    idx[ACGT] = np.arange(4)
    j = np.arange(n)
    code = np.zeros(n, dtype=np.int64)
    valid = np.ones(n, dtype=bool)
    for _ in range(4):
        b = idx[bwt[j]] if n else np.zeros(0, dtype=np.int64)
        valid &= b >= 0
        code = code * 4 + np.maximum(b, 0)
        j = lf[j] if n else j
    nsec4 = n // 224 + 2
    sectors = np.zeros((256, nsec4, 8), dtype=np.uint32)
    pos = np.flatnonzero(valid)
    np.bitwise_or.at(sectors, (code[pos], pos // 224, 1 + (pos % 224) // 32), (np.uint32(1) << (pos % 32).astype(np.uint32)))
    per = np.zeros((256, nsec4), dtype=np.int64)
    np.add.at(per, (code[pos], pos // 224), 1)
    before = np.cumsum(per, axis=1) - per
    c4 = np.zeros(256, dtype=np.int64)
    for c in range(256):
        p = 0
        for r in range(4):
            sym = ACGT[(c >> (2 * (3 - r))) & 3]
            p = cstart[sym] + int((bwt[:p] == sym).sum())
        c4[c] = p
    c4base = np.zeros((0, 256), dtype=np.uint64)
    if wide:
        first = (np.arange(nsec4) >> sb_shift4) << sb_shift4
        ck = before - before[:, first]
        n_super4 = ((nsec4 - 1) >> sb_shift4) + 1
        c4base = (before[:, np.arange(n_super4) << sb_shift4] + c4[:, None]).T.astype(np.uint64)
    else:
        ck = before + c4[:, None]
    sectors[:, :, 0] = ck.astype(np.uint64).astype(np.uint32)
    return sectors, c4base


@pytest.mark.parametrize("sb_shift", [0, 1, 3])
def test_quad_image_equals_brute_force(sb_shift):
    rng = np.random.default_rng(224 + sb_shift)
    from harness import bwt_build, synth
    reads = synth.make_reads(400, 60, 15.0, 0.02, device="cuda")
    reads[3, 10:12] = 4
    streams = [
        O.convert_to_vec(naive.naive_bwt(["CCGTACGTA", "GGTACAGTA", "ACGACGACG", "ANNT"])),
        _random_rle(rng, 3000, [1, 1, 1, 2, 3, 5, 9, 31, 32, 33, 95, 96, 97, 223, 224, 225, 255]),
        O.convert_to_vec("A" * 224),         # N a multiple of 224: the sector of position N is all padding
        O.convert_to_vec("ACGT" * 56 + "T"),
        O.convert_to_vec("ACGT" * 24),       # N = 96: pair lines and quad sectors end at different places
        bwt_build.build_rle_bwt(reads)[0].cpu().numpy(),
        np.zeros(0, np.uint8),
    ]
    for rle in streams:
        g = M.RleBWT(superblock_shift=sb_shift, quad_index=1)
        g.load_vector(rle)
        assert g.quad_index and not g.pair_index
        bwt = decode(np.asarray(rle, dtype=np.uint8))
        assert bwt.size == g.get_total_size()
        got, got_c4 = g.quad_image()
        wide = sb_shift != 0 and ((bwt.size >> 7) >> sb_shift) >= 1
        want, want_c4 = brute_quad_image(bwt, min(sb_shift if sb_shift else 25, 24), wide)
        assert got.shape == want.shape
        assert (got == want).all(), np.argwhere(got != want)[:5]
        assert got_c4.shape == want_c4.shape and (got_c4 == want_c4).all()


@pytest.fixture(scope="module")
def midsize():
    from harness import bwt_build, synth
    reads = synth.make_reads(20000, read_len=100, coverage=25.0, error_rate=0.01, device="cuda")
    reads[17, 40:43] = 4  # a few N
    rle, n = bwt_build.build_rle_bwt(reads)
    o = O.RleBWT()
    o.load_vector(rle.cpu().numpy())
    return reads, o


@pytest.mark.parametrize("sb_shift,table_s", [(0, -1), (0, 0), (0, 1), (0, 2), (0, 3), (0, 4), (0, 7), (3, 5), (3, 6), (2, 0)])
def test_quad_path_is_bit_exact(midsize, sb_shift, table_s):
    from harness import synth
    reads, o = midsize
    g = M.RleBWT(superblock_shift=sb_shift, suffix_table_s=table_s, quad_index=1)
    g.load_vector(o.rle_bytes())
    assert g.quad_index
    rng = np.random.default_rng(7 + sb_shift + 10 * (table_s + 1))
    for k in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 29, 30, 31, 32, 33, 34, 35, 36, 63, 64, 65, 66, 67, 97, 100):
        q = synth.make_queries(reads, k, 12001, 8000).cpu().numpy()
        q[5, 0] = 4           # N at the far end: table usable, quad path not
        q[7, k - 1] = 0       # $ as the first consumed symbol
        q[11, k // 2] = 4
        got = g.count_kmers_fixed(q, k)
        want = o.count_kmers_fixed(q, k, threads=8)
        assert (got == want).all(), (sb_shift, table_s, k, np.flatnonzero(got != want)[:5])
        if k >= 29:
            assert int((got > 0).sum()) >= 11990
    ragged = [rng.integers(0, 6, int(rng.integers(0, 40))).astype(np.uint8) for _ in range(3000)]
    assert (g.count_kmers(ragged) == o.count_kmers(ragged)).all()


def test_quad_path_on_golden_fixture_and_both_host_routes(golden_dir, monkeypatch):
    z = np.load(f"{golden_dir}/reads30x_k31.npz")
    for host_pack in ("0", "1"):
        monkeypatch.setenv("MSBWT_HOST_PACK", host_pack)
        g = M.RleBWT(quad_index=1)
        g.load_vector(z["rle"])
        assert g.quad_index
        assert (g.count_kmers_fixed(z["queries"], int(z["k"])) == z["counts"]).all()
        assert (g.count_kmers_fixed(z["queries_k12"], 12) == z["counts_k12"]).all()


def test_quad_path_device_entry_and_invalid_symbols(midsize):
    from harness import synth
    reads, o = midsize
    g = M.RleBWT(quad_index=1)
    g.load_vector(o.rle_bytes())
    k = 31
    q = synth.make_queries(reads, k, 50000, 50000)
    want = o.count_kmers_fixed(q.cpu().numpy(), k, threads=8)
    out = torch.zeros(q.shape[0], dtype=torch.int64, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    g.count_kmers_fixed_device(q.data_ptr(), k, q.shape[0], out.data_ptr(), status.data_ptr(),
                               torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    assert (out.cpu().numpy().astype(np.uint64) == want).all()
    sym = np.array([1, 2, 3, 5, 0, 4], dtype=np.uint8)
    n = o.get_total_size()
    gl, gh = g.constrain_ranges(sym, np.zeros(6, np.uint64), np.full(6, n, np.uint64))
    for i in range(6):
        assert (int(gl[i]), int(gh[i])) == o.constrain_range(int(sym[i]), 0, n)
    q[123, 5] = 6
    g.count_kmers_fixed_device(q.data_ptr(), k, q.shape[0], out.data_ptr(), status.data_ptr(),
                               torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert int(status.item()) != 0
    with pytest.raises(M.MsbwtError):
        g.count_kmers_fixed(q.cpu().numpy(), k)


def test_quad_stats_replay_matches_counts(midsize):
    """The oracle's accounting replay of the quad path (bench.py's algorithmic bytes) walks the same
    ranges as count_kmer: its step counts are bounded by k and its table lookups by the batch size."""
    from harness import synth
    reads, o = midsize
    k = 31
    q = synth.make_queries(reads, k, 4000, 4000).cpu().numpy()
    st = o.count_kmers_stats_quad(q, k, 7)
    assert st["queries"] == 8000 and st["table_hits"] == 8000
    assert 0 < st["quad_steps"] <= 8000 * ((k - 7) // 4)
    assert st["two_line_quad_steps"] <= st["two_sector_quad_steps"] <= st["quad_steps"]
    assert st["one_steps"] == 0  # 31 - 7 = 24 is a multiple of four
    st = o.count_kmers_stats_quad(q, k, 0)
    assert st["table_hits"] == 0 and st["one_steps"] > 0  # 31 = 7 quads + 3 one-steps
