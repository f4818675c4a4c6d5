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


def _device_lines_from_records(keys, vals, nlines):
    """fin_lines_kernel (fin_builder.cu) statement by statement: records sorted by key, the first record of a line
    writes the whole line"""
    TAG_BITS, WORDS = 28, F.LINE_WORDS
    lines = np.zeros((nlines, WORDS), dtype=np.uint32)
    n = len(keys)
    for i in range(n):
        line = int(keys[i]) >> TAG_BITS
        if i and (int(keys[i - 1]) >> TAG_BITS) == line:
            continue
        w = lines[line]
        used = header = in_group = 0
        cur_key, over = None, False
        r = i
        while r < n and (int(keys[r]) >> TAG_BITS) == line:
            if int(keys[r]) != cur_key or in_group == 15:
                if used + 2 > WORDS - 1:
                    over = True
                    break
                cur_key = int(keys[r])
                used += 1
                header = used
                in_group = 0
                w[header] = (cur_key & ((1 << TAG_BITS) - 1)) << 4
            elif used + 1 > WORDS - 1:
                over = True
                break
            used += 1
            w[used] = vals[r]
            in_group += 1
            w[header] = (int(w[header]) & ~15) | in_group
            r += 1
        w[0] = F.OVERFLOW if over else used
    return lines


def _device_count(line, tag, pl, ph):
    """the is_fin CONSUME step of count_kmers_oct_kernel (oct_kernel.cuh) statement by statement: eight uint4 of the
    staged line, word 0 = words in use"""
    used = int(line[0])
    if used == F.OVERFLOW:
        return None
    state = {"cnt": 0, "left": 0, "match": False}

    def eat(w, idx):
        if idx > used:
            return
        if state["left"] == 0:
            state["match"] = (int(w) >> 4) == tag
            state["left"] = int(w) & 15
        else:
            off, ln = int(w) & 0xFFFF, int(w) >> 16
            if state["match"]:
                state["cnt"] += min(max(ph - off, 0), ln) - min(max(pl - off, 0), ln)
            state["left"] -= 1

    eat(line[1], 1)
    eat(line[2], 2)
    eat(line[3], 3)
    v = 1
    while v < 8 and 4 * v <= used:
        for t in range(4):
            eat(line[4 * v + t], 4 * v + t)
        v += 1
    return state["cnt"]


@pytest.mark.parametrize("b,lb", [(16, 12), (11, 12)])
def test_device_algorithms_replayed_on_the_cpu_agree_with_the_specification(reads_index, b, lb):
    """The builder's record -> line loop and the kernel's line parse, transcribed from the CUDA sources, give the
    specification's lines and counts (what can be checked of the experimental kernels without a GPU)."""
    reads, o, bwt = reads_index
    want, stats = F.build_final_image(bwt, b=b, lb=lb)
    key = F.position_codes(bwt)
    n = bwt.size
    bmask = (1 << b) - 1
    keys, vals = [], []
    j = 0
    while j < n:                                               # fin_emit_kernel: heads, runs cut at 65535 and at buckets
        if key[j] < 0:
            j += 1
            continue
        ln = 1
        while j + ln < n and ((j + ln) & bmask) != 0 and key[j + ln] == key[j]:
            ln += 1
        mixed = int(F.mix40(np.uint64(key[j])))
        line = ((j >> b) << lb) | (mixed & ((1 << lb) - 1))
        at, left = j, ln
        while left:
            piece = min(left, 65535)
            keys.append((line << 28) | (mixed >> lb))
            vals.append((piece << 16) | (at & bmask))
            at += piece
            left -= piece
        j += ln
    order = np.argsort(np.array(keys, dtype=np.uint64), kind="stable")   # the device sorts with a CUB radix sort
    keys = np.array(keys, dtype=np.uint64)[order]
    vals = np.array(vals, dtype=np.uint32)[order]
    got = _device_lines_from_records(keys, vals, want.shape[0])
    assert ((got[:, 0] == F.OVERFLOW) == (want[:, 0] == F.OVERFLOW)).all()
    for i in np.flatnonzero(want[:, 0] != 0):
        assert F.line_groups(got[i]) == F.line_groups(want[i]), i
    # the kernel's parse of those lines against final_count
    from harness import synth
    q = synth.make_queries(reads, 31, 400, 50).numpy()
    q = q[np.isin(q, (1, 2, 3, 5)).all(axis=1)]
    checked = 0
    for kmer in q:
        l, h = 0, int(n)
        for t in range(11):
            l, h = o.constrain_range(int(kmer[30 - t]), l, h)
        if l == h or (l >> b) != (h >> b):
            continue
        code = F.query_code(kmer[:20])
        mixed = int(F.mix40(np.uint64(code)))
        line = got[((l >> b) << lb) | (mixed & ((1 << lb) - 1))]
        dev = _device_count(line, mixed >> lb, l & bmask, h & bmask)
        assert dev == F.final_count(want, code, l, h, b=b, lb=lb)
        if dev is not None:
            assert dev == o.count_kmer(kmer)
            checked += 1
    assert checked > 100


def test_accounting_replay_with_the_final_step(reads_index):
    """bench.py's algorithmic-bytes replay (orc_count_kmers_stats_*): with the final-step image a 31-mer that
    survives the depth-11 table entry costs ONE final line instead of two oct lines; without it nothing changes."""
    from harness import synth
    reads, o, bwt = reads_index
    q = synth.make_queries(reads, 31, 3000, 500).numpy()
    base = o.count_kmers_stats_quad(q, 31, 14, oct_bucket_shift=16, oct_syms=10)
    assert base["final_steps"] == 0 and base["two_bucket_final_steps"] == 0
    fin = o.count_kmers_stats_quad(q, 31, 14, oct_bucket_shift=16, oct_syms=10, fin_bucket_shift=16)
    for key in ("queries", "table_hits"):
        assert fin[key] == base[key], key
    for key in ("quad_steps", "one_steps"):   # fallbacks of a SECOND oct step no longer happen: the query has ended
        assert fin[key] <= base[key], key
    assert fin["final_steps"] > 0
    # every query that took a final step would have taken one oct step, and a second one unless it died on the first
    assert base["oct_steps"] - fin["oct_steps"] >= fin["final_steps"]
    assert base["oct_steps"] - fin["oct_steps"] <= 2 * fin["final_steps"]
    # a bucket so small that ranges straddle it: the replay falls back to the oct steps and says so
    tiny = o.count_kmers_stats_quad(q, 31, 14, oct_bucket_shift=16, oct_syms=10, fin_bucket_shift=3)
    assert tiny["two_bucket_final_steps"] > 0
    assert tiny["final_steps"] + tiny["two_bucket_final_steps"] == fin["final_steps"] + fin["two_bucket_final_steps"]
    # other lengths: k = 41 (table 11 + one oct step) and k = 30 (no table: three oct steps cost no more) reach the
    # final step with exactly 20 symbols left, k = 29 (table 11, 18 left) never does
    for k, reaches in ((41, True), (30, True), (29, False)):
        qk = synth.make_queries(reads, k, 2000, 0).numpy()
        st = o.count_kmers_stats_quad(qk, k, 14, oct_bucket_shift=16, oct_syms=10, fin_bucket_shift=16)
        assert (st["final_steps"] > 0) == reaches, (k, st)
