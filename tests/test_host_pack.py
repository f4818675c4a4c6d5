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
