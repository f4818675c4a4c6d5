"""The PAIR image (layout.h: one 128-byte line per 96 BWT positions answers TWO constrain_range steps,
src/rle_bwt.rs:202-287 composed with itself) and the quad-per-query kernel that walks it.

  * the image built on the device is compared word for word with a numpy brute-force construction
    from the decoded BWT (LF by counting, codes, bit-planes, checkpoints);
  * every count through the pair path must equal the CPU oracle's (the reference's algorithm),
    for every suffix-table depth / k parity combination, with $ / N inside the k-mers, with 32-bit
    and 64-bit positions, and at N > 2^32."""
import numpy as np
import pytest

import rust_msbwt_b200 as M
from oracle import naive
from oracle import oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

ACGT = np.array([1, 2, 3, 5])


def decode(rle: np.ndarray) -> np.ndarray:
    """RLE byte stream -> one symbol per position (msbwt_core.rs:4-14)."""
    syms, counts, prev, j = [], [], -1, 0
    for v in rle.tolist():
        c, d = v & 7, v >> 3
        if c == prev:
            j += 1
            counts[-1] += d << (5 * j)
        else:
            prev, j = c, 0
            syms.append(c)
            counts.append(d)
    return np.repeat(np.array(syms, dtype=np.uint8), np.array(counts, dtype=np.int64))


def brute_pair_image(bwt: np.ndarray, sb_shift: int, wide: bool):
    n = bwt.size
    cnt = np.bincount(bwt, minlength=6)
    cstart = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    # LF(j) = C[b] + rank(b, j)
    lf = np.zeros(n, dtype=np.int64)
    for s in range(6):
        at = np.flatnonzero(bwt == s)
        lf[at] = cstart[s] + np.arange(at.size)
    idx = np.full(8, -1)
    idx[ACGT] = np.arange(4)
    b = idx[bwt]
    a = idx[bwt[lf]] if n else b
    valid = (b >= 0) & (a >= 0)
    code = np.where(valid, 4 * b + a, 0)
    npair = n // 96 + 1
    pad = npair * 96
    codep = np.zeros(pad, dtype=np.int64)
    validp = np.zeros(pad, dtype=bool)
    codep[:n], validp[:n] = code, valid
    lines = np.zeros((npair, 32), dtype=np.uint64)
    cq = codep.reshape(npair, 4, 24)
    vq = validp.reshape(npair, 4, 24)
    w = (1 << np.arange(24, dtype=np.uint64))
    for p in range(4):
        lines[:, 4 + p::8][:, :4] = ((((cq >> p) & 1) * vq).astype(np.uint64) * w).sum(axis=2)
    vbits = (vq.astype(np.uint64) * w).sum(axis=2)
    for t in range(4):
        for p in range(3):
            lines[:, t * 8 + 4 + p] |= ((vbits[:, t] >> np.uint64(8 * p)) & np.uint64(0xFF)) << np.uint64(24)
    # checkpoints
    onehot = np.zeros((pad, 16), dtype=np.int64)
    onehot[np.arange(pad)[validp], codep[validp]] = 1
    per_line = onehot.reshape(npair, 96, 16).sum(axis=1)
    before = np.cumsum(per_line, axis=0) - per_line
    c2 = np.zeros(16, dtype=np.int64)
    for c in range(16):
        sb_, sa = ACGT[c >> 2], ACGT[c & 3]
        c2[c] = cstart[sa] + int((bwt[:cstart[sb_]] == sa).sum())
    c2base = np.zeros((0, 16), dtype=np.uint64)
    if wide:
        first = (np.arange(npair) >> sb_shift) << sb_shift
        ck = before - before[first]
        n_super2 = ((npair - 1) >> sb_shift) + 1
        c2base = (before[np.arange(n_super2) << sb_shift] + c2).astype(np.uint64)
    else:
        ck = before + c2
    for c in range(16):
        lines[:, (c >> 2) * 8 + (c & 3)] = ck[:, c].astype(np.uint64)
    return lines.astype(np.uint32), c2base


def _random_rle(rng, nruns, choices):
    syms, counts, prev = [], [], -1
    for _ in range(nruns):
        s = int(rng.choice(6, p=[0.05, 0.27, 0.25, 0.25, 0.03, 0.15]))
        if s == prev:
            continue
        prev = s
        syms.append(s)
        counts.append(int(rng.choice(choices)))
    return O.encode_runs(syms, counts)


@pytest.mark.parametrize("sb_shift", [0, 1, 3])
def test_pair_image_equals_brute_force(sb_shift):
    rng = np.random.default_rng(96 + sb_shift)
    from harness import bwt_build, synth
    reads = synth.make_reads(400, 60, 15.0, 0.02, device="cuda")
    reads[3, 10:12] = 4
    streams = [
        O.convert_to_vec(naive.naive_bwt(["CCGTACGTA", "GGTACAGTA", "ACGACGACG", "ANNT"])),
        _random_rle(rng, 3000, [1, 1, 1, 2, 3, 5, 9, 31, 32, 33, 95, 96, 97, 255]),
        O.convert_to_vec("A" * 96),          # N a multiple of 96: the line for position N is all padding
        O.convert_to_vec("ACGT" * 24 + "T"),
        bwt_build.build_rle_bwt(reads)[0].cpu().numpy(),
        np.zeros(0, np.uint8),
    ]
    for rle in streams:
        g = M.RleBWT(superblock_shift=sb_shift, pair_index=1)
        g.load_vector(rle)
        assert g.pair_index
        bwt = decode(np.asarray(rle, dtype=np.uint8))
        assert bwt.size == g.get_total_size()
        got_lines, got_c2 = g.pair_image()
        # positions are 64-bit ("wide") as soon as there is more than one superblock of one-step blocks
        wide = sb_shift != 0 and ((bwt.size >> 7) >> sb_shift) >= 1
        want_lines, want_c2 = brute_pair_image(bwt, sb_shift if sb_shift else 25, wide)
        assert got_lines.shape == want_lines.shape
        assert (got_lines == want_lines).all(), np.argwhere(got_lines != want_lines)[:5]
        assert got_c2.shape == want_c2.shape and (got_c2 == want_c2).all()


@pytest.fixture(scope="module")
def midsize():
    from harness import bwt_build, synth
    reads = synth.make_reads(20000, read_len=100, coverage=25.0, error_rate=0.01, device="cuda")
    reads[17, 40:43] = 4  # a few N
    rle, n = bwt_build.build_rle_bwt(reads)
    o = O.RleBWT()
    o.load_vector(rle.cpu().numpy())
    return reads, o


@pytest.mark.parametrize("sb_shift,table_s", [(0, -1), (0, 0), (0, 1), (0, 2), (0, 7), (3, 4), (3, 5), (2, 0)])
def test_pair_path_is_bit_exact(midsize, sb_shift, table_s):
    from harness import synth
    reads, o = midsize
    g = M.RleBWT(superblock_shift=sb_shift, suffix_table_s=table_s, pair_index=1)
    g.load_vector(o.rle_bytes())
    assert g.pair_index
    rng = np.random.default_rng(5 + sb_shift + 10 * (table_s + 1))
    for k in (1, 2, 3, 4, 6, 7, 8, 9, 30, 31, 32, 33, 64, 65, 66, 99, 100):
        q = synth.make_queries(reads, k, 12001, 8000).cpu().numpy()
        q[5, 0] = 4           # N at the far end: table usable, pair path not
        q[7, k - 1] = 0       # $ as the first consumed symbol
        q[11, k // 2] = 4
        got = g.count_kmers_fixed(q, k)
        want = o.count_kmers_fixed(q, k, threads=8)
        assert (got == want).all(), (sb_shift, table_s, k, np.flatnonzero(got != want)[:5])
        if k >= 30:
            assert int((got > 0).sum()) >= 11990
    ragged = [rng.integers(0, 6, int(rng.integers(0, 40))).astype(np.uint8) for _ in range(3000)]
    assert (g.count_kmers(ragged) == o.count_kmers(ragged)).all()


def test_pair_path_equals_one_step_path_on_golden_fixture(golden_dir):
    z = np.load(f"{golden_dir}/reads30x_k31.npz")
    for lanes in (1, 2):
        g = M.RleBWT(pair_index=1, kernel_lanes=lanes)
        g.load_vector(z["rle"])
        assert g.pair_index and g.kernel_lanes == lanes
        assert (g.count_kmers_fixed(z["queries"], int(z["k"])) == z["counts"]).all()
        assert (g.count_kmers_fixed(z["queries_k12"], 12) == z["counts_k12"]).all()


def test_pair_path_device_entry_and_invalid_symbols(midsize):
    from harness import synth
    reads, o = midsize
    g = M.RleBWT(pair_index=1)
    g.load_vector(o.rle_bytes())
    k = 31
    q = synth.make_queries(reads, k, 50000, 50000)
    want = o.count_kmers_fixed(q.cpu().numpy(), k, threads=8)
    out = torch.zeros(q.shape[0], dtype=torch.int64, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    # an unaligned device pointer exercises the head/tail of the pack kernel's staged copy
    buf = torch.zeros(q.numel() + 64, dtype=torch.uint8, device="cuda")
    for shift in (0, 1, 7, 16, 33):
        view = buf[shift:shift + q.numel()]
        view.copy_(q.reshape(-1))
        out.zero_()
        g.count_kmers_fixed_device(view.data_ptr(), k, q.shape[0], out.data_ptr(), status.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert int(status.item()) == 0
        assert (out.cpu().numpy().astype(np.uint64) == want).all(), shift
    q[123, 5] = 6
    g.count_kmers_fixed_device(q.data_ptr(), k, q.shape[0], out.data_ptr(), status.data_ptr(),
                               torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert int(status.item()) != 0
    with pytest.raises(M.MsbwtError):
        g.count_kmers_fixed(q.cpu().numpy(), k)


def test_pair_path_beyond_2_pow_32_symbols():
    """N > 2^32: 64-bit positions, relative checkpoints + c2base, two pair superblocks."""
    rng = np.random.default_rng(2033)
    nruns = 4_600_000
    syms = rng.choice(np.array([0, 1, 2, 3, 4, 5], dtype=np.uint8), size=nruns, p=[0.02, 0.26, 0.24, 0.24, 0.02, 0.22])
    keep = np.ones(nruns, dtype=bool)
    keep[1:] = syms[1:] != syms[:-1]
    syms = syms[keep]
    counts = rng.integers(1, 2800, size=syms.size).astype(np.uint64)
    rle = O.encode_runs(syms, counts)
    g = M.RleBWT(pair_index=1)
    g.load_vector(rle)
    o = O.RleBWT()
    o.load_vector(rle)
    n = o.get_total_size()
    assert n > (1 << 32) and g.get_total_size() == n and g.pair_index
    for k in (2, 3, 8, 9, 24, 25):
        q = rng.choice(np.array([1, 2, 3, 5], dtype=np.uint8), size=(100_000, k))
        q[::97, 0] = 4
        # long runs: make most queries homopolymer-ish so that ranges stay non-empty for many steps
        q[: 50_000] = q[: 50_000, :1]
        assert (g.count_kmers_fixed(q, k) == o.count_kmers_fixed(q, k, threads=8)).all(), k
