"""CPU check of the final-step image specification (oracle/final_step.py): for the last m
symbols a count_kmer consumes, counting the positions of [l, h) whose m-symbol code matches equals what the
reference's loop of constrain_range calls returns (src/msbwt_core.rs:125-161, src/rle_bwt.rs:202-287) -- from one
hashed line, or an explicit fallback (None), never a wrong count."""
import numpy as np
import pytest

from oracle import final_step as F
from oracle import oracle as O


def decode(rle: np.ndarray) -> np.ndarray:
    """RLE bytes -> one symbol per position (msbwt_core.rs:4-14: sym | digit << 3, base-32 little-endian digits)"""
    out, i = [], 0
    rle = np.asarray(rle, dtype=np.uint8)
    while i < rle.size:
        s, cnt, p = int(rle[i]) & 7, 0, 0
        while i < rle.size and (int(rle[i]) & 7) == s:
            cnt += (int(rle[i]) >> 3) << (5 * p)
            p += 1
            i += 1
        out.append(np.full(cnt, s, dtype=np.uint8))
    return np.concatenate(out) if out else np.zeros(0, dtype=np.uint8)


@pytest.fixture(scope="module")
def reads_index():
    from harness import bwt_build, synth
    reads = synth.make_reads(3000, read_len=100, coverage=25.0, error_rate=0.01, device="cpu")
    reads[5, 40:42] = 4  # N
    rle, n = bwt_build.build_rle_bwt(reads)
    o = O.RleBWT()
    o.load_vector(rle.numpy())
    bwt = decode(rle.numpy())
    assert bwt.size == n == o.get_total_size()
    return reads, o, bwt


def test_mix40_is_a_bijection_on_what_we_can_check():
    a = np.arange(1 << 20, dtype=np.uint64)
    vals = np.concatenate([a, a << np.uint64(20), np.random.default_rng(3).integers(0, 1 << 40, 1 << 20, dtype=np.uint64)])
    vals = np.unique(vals)
    out = F.mix40(vals)
    assert out.max() < (1 << 40) and np.unique(out).size == vals.size


@pytest.mark.parametrize("b,lb", [(16, 12), (12, 12), (10, 13)])
def test_final_count_equals_the_reference_loop(reads_index, b, lb):
    from harness import synth
    reads, o, bwt = reads_index
    lines, stats = F.build_final_image(bwt, b=b, lb=lb)
    assert stats["runs"] > 0 and stats["lines"] == ((bwt.size >> b) + 1) << lb
    k, m, ts = 31, F.M_SYMS, 11
    q = synth.make_queries(reads, k, 1500, 300).numpy()
    q = q[np.isin(q, (1, 2, 3, 5)).all(axis=1)]
    answered = fell_back = 0
    for kmer in q:
        l, h = 0, int(bwt.size)
        for t in range(ts):                                    # the last `ts` symbols: what the suffix table holds
            l, h = o.constrain_range(int(kmer[k - 1 - t]), l, h)
        got = F.final_count(lines, F.query_code(kmer[:m]), l, h, b=b, lb=lb)
        want = o.count_kmer(kmer)
        if got is None:
            fell_back += 1
        else:
            answered += 1
            assert got == want, (kmer, l, h, got, want)
    assert answered > 0.8 * len(q), (answered, fell_back)


def test_lines_compare_as_sets_of_groups_and_overflow_is_explicit(reads_index):
    reads, o, bwt = reads_index
    lines, stats = F.build_final_image(bwt, b=16, lb=12)
    used = lines[lines[:, 0] != 0]
    assert used.shape[0] > 0
    g = F.line_groups(used[0])
    assert g == sorted(g) and all(len(r) >= 1 for _, r in g) and len({t for t, _ in g}) == len(g)
    # a low-complexity text: few codes own every position, their lines cannot hold the runs -> marked, not truncated
    rng = np.random.default_rng(5)
    unit = rng.choice(np.array([1, 2, 3, 5], dtype=np.uint8), 40)
    noisy = np.tile(unit, 600)
    noisy[rng.integers(0, noisy.size, 400)] = 1
    from oracle import naive
    text = "".join("$ACGNT"[s] for s in noisy[:6000])
    bw = decode(O.convert_to_vec(naive.naive_bwt([text[i:i + 100] for i in range(0, 5900, 7)])))
    lines2, stats2 = F.build_final_image(bw, b=16, lb=12)
    assert stats2["overflowed_lines"] == int((lines2[:, 0] == F.OVERFLOW).sum())
    for ln in lines2[lines2[:, 0] != F.OVERFLOW]:
        assert int(ln[0]) <= F.LINE_WORDS - 1
