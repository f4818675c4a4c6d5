"""The host-side 2-bit packer of the end-to-end path (rust-msbwt_b200/csrc/hostpack.cpp) against a numpy
restatement: word format, exception list (any symbol outside ACGT, including >= 6), every k up to several
words, the tail of the caller's buffer (no read past the end) and the worker split.  Needs no GPU."""
import numpy as np
import pytest

import rust_msbwt_b200 as M

CODE = {1: 0, 2: 1, 3: 2, 5: 3}


def reference_pack(q: np.ndarray):
    n, k = q.shape
    nw = -(-k // 32)
    words = np.zeros((nw, n), dtype=np.uint64)
    exc = []
    for i in range(n):
        row = q[i].tolist()
        if any(s not in CODE for s in row):
            exc.append(i)
        for t in range(k):                      # consumption step t takes the k-mer's symbol k-1-t
            c = CODE.get(row[k - 1 - t], 0)
            words[t // 32, i] |= np.uint64(c) << np.uint64(62 - 2 * (t % 32))
    return words, np.array(exc, dtype=np.uint64)


@pytest.mark.parametrize("k", [1, 2, 5, 16, 31, 32, 33, 47, 63, 64, 65, 101])
def test_pack_matches_reference(k):
    rng = np.random.default_rng(k)
    n = 257
    q = rng.choice(np.array([1, 2, 3, 5], dtype=np.uint8), size=(n, k))
    bad_rows = rng.choice(n, 20, replace=False)
    for j, r in enumerate(bad_rows):
        q[r, rng.integers(0, k)] = [0, 4, 6, 7, 16, 17, 128, 255, 21, 0x81][j % 10]
    want_words, want_exc = reference_pack(q)
    for threads in (1, 3, 8):
        # exact-size buffer: the packer must not read past the last symbol
        buf = np.ascontiguousarray(q.reshape(-1).copy())
        words, exc = M.debug_host_pack(buf, k, threads)
        assert (exc == np.sort(want_exc)).all()
        ok = np.ones(n, dtype=bool)
        ok[want_exc.astype(np.int64)] = False
        assert (words[:, ok] == want_words[:, ok]).all(), (k, threads)


def test_tiny_batches_and_zero_queries():
    w, e = M.debug_host_pack(np.zeros(0, np.uint8), 31, 4)
    assert w.shape == (1, 0) and e.size == 0
    q = np.array([[1, 2, 3, 5, 5, 3, 2]], dtype=np.uint8)
    w, e = M.debug_host_pack(q, 7, 2)
    assert e.size == 0
    # last symbol (2 = C -> 1) on top, then 3 (G -> 2), 5 (T -> 3), 5, 3 (G -> 2), 2, 1 (A -> 0)
    want = 0
    for t, c in enumerate([1, 2, 3, 3, 2, 1, 0]):
        want |= c << (62 - 2 * t)
    assert int(w[0, 0]) == want
    assert M.host_pack_threads() >= 1


def test_suffix_table_depth_policy():
    """Which kept suffix-table level an all-ACGT k-mer starts from (kernel_common.cuh; no device needed): with a
    pair / quad image the level that leaves a multiple of 2 / 4 symbols; with the oct image the level of the
    four kept ones that leaves the cheapest walk (oct steps of m symbols, quad steps of four, one-symbol steps
    at two accesses each) -- e.g. depth 11 for k = 31 under a depth-14 table: two oct steps."""
    import rust_msbwt_b200 as M
    L = M.load_library()
    m = L.msbwt_oct_symbols()
    assert m == 10

    def cost(rest):
        r = rest % m
        return rest // m + r // 4 + 2 * (r % 4)

    for ts in range(0, 17):
        for k in range(1, 140):
            got = L.msbwt_debug_table_depth(k, ts, m)
            if ts == 0:
                want = 0
            elif k < ts:
                want = k if k + 4 > ts else 0
            else:
                want, best = 0, cost(k)
                for back in range(min(4, ts)):
                    c = cost(k - (ts - back))
                    if c < best:
                        best, want = c, ts - back
            assert got == want, (k, ts, got, want)
            for stride in (1, 2, 4):
                d = L.msbwt_debug_table_depth(k, ts, stride)
                assert 0 <= d <= min(k, ts)
                if d and k >= ts:
                    assert (k - d) % stride == 0 and ts - d < stride
    assert L.msbwt_debug_table_depth(31, 14, m) == 11
    assert L.msbwt_debug_table_depth(31, 15, 4) == 15
    assert L.msbwt_debug_table_depth(31, 14, 3) == -1


def test_pack_kernel_table_arithmetic_is_exact_for_every_symbol_word():
    """The pack / seed kernels validate and 2-bit-pack four symbol bytes at a time through an 8-entry PRMT byte
    table and one multiply (swar_lut_pack4_top / swar_gather4, kernel_common.cuh).  The constants are read from
    the source and the arithmetic is replayed with PTX `prmt.b32` semantics (generic mode: the 3 low bits of a
    selector nibble pick one of the 8 source bytes, the top bit replicates that byte's sign) over every word of
    four bytes drawn from a set that holds all symbols 0..9 and bytes with every high bit: a word is flagged
    iff some byte is outside ACGT = {1,2,3,5} (the reference's count_kmer takes any symbol < 6,
    msbwt_core.rs:127; the others go to the byte-wise path), and a clean word packs to byte i at bits 2i."""
    import itertools
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                            "rust-msbwt_b200", "csrc", "kernel_common.cuh")).read()
    body = src[src.index("swar_lut_pack4_top(uint32_t x"):src.index("inline int sm_count")]
    bad_mask = int(re.search(r"kSwarBadMask = (0x[0-9A-Fa-f]+)u", src).group(1), 16)
    sel_nibbles = int(re.search(r"prmt_b32\(t, 0u, (0x[0-9A-Fa-f]+)u\)", body).group(1), 16)
    lut_lo, lut_hi = (int(v, 16) for v in re.search(r"prmt_b32\((0x[0-9A-Fa-f]+)u, (0x[0-9A-Fa-f]+)u, sel\)", body).groups())
    mul = int(re.search(r"return y \* (0x[0-9A-Fa-f]+)u", body).group(1), 16)
    pair_sel = int(re.search(r"prmt_b32\(m0, m1, (0x[0-9A-Fa-f]+)u\)", body).group(1), 16)
    join_sel = int(re.search(r"\), (0x[0-9A-Fa-f]+)u\);\n\}", body).group(1), 16)

    def prmt(a, b, sel):
        srcb = [(a >> (8 * i)) & 0xFF for i in range(4)] + [(b >> (8 * i)) & 0xFF for i in range(4)]
        out = 0
        for i in range(4):
            nib = (sel >> (4 * i)) & 0xF
            byte = srcb[nib & 7]
            if nib & 8:
                byte = 0xFF if byte & 0x80 else 0
            out |= byte << (8 * i)
        return out

    def top(x):
        t = (x + (x >> 4)) & 0xFFFFFFFF
        y = prmt(lut_lo, lut_hi, prmt(t, 0, sel_nibbles))
        return (y * mul) & 0xFFFFFFFF, x | y

    code = {1: 0, 2: 1, 3: 2, 5: 3}
    vals = list(range(10)) + [0x0F, 0x10, 0x11, 0x15, 0x21, 0x51, 0x80, 0xF0, 0xFF]
    for bs in itertools.product(vals, repeat=4):
        x = bs[0] | bs[1] << 8 | bs[2] << 16 | bs[3] << 24
        m, acc = top(x)
        clean = all(b in code for b in bs)
        assert ((acc & bad_mask) != 0) == (not clean), bs
        if clean:
            assert m >> 24 == sum(code[b] << (2 * i) for i, b in enumerate(bs)), bs
    ms = [0x12345678, 0x9ABCDEF0, 0x0F1E2D3C, 0xC0FFEE11]
    g = prmt(prmt(ms[0], ms[1], pair_sel), prmt(ms[2], ms[3], pair_sel), join_sel)
    assert g == sum((ms[i] >> 24) << (8 * i) for i in range(4))
