"""The product's host codec (rust-msbwt_b200/csrc/codec.cpp): `convert_to_vec`, `save_bwt_numpy`, `save_bwt_runs_numpy`
of src/bwt_converter.rs:26-184 -- the data format either side of the query path.  Pinned by the reference's own KATs
(src/bwt_converter.rs:195-321, the same literals as tests/test_oracle_kat.py holds for the oracle), compared byte for
byte with the oracle's writer on random inputs, and with the reference's fixture.  No GPU involved."""
import numpy as np
import pytest

import rust_msbwt_b200 as M
from oracle import naive
from oracle import oracle as O


# ---- bwt_converter.rs:195-256 ----
def test_convert_to_vec_kats():
    assert list(M.convert_to_vec("ACGNT$")) == [9, 10, 11, 12, 13, 8]
    assert list(M.convert_to_vec("\n$$\n$$\nAAA\n")) == [32, 25]
    assert list(M.convert_to_vec("A" * 3104)) == [1, 9, 25]
    assert list(M.convert_to_vec("A" * 31 + "C" * 31)) == [249, 250]
    assert list(M.convert_to_vec("N" * 32767)) == [252, 252, 252]
    assert list(M.convert_to_vec("GTN$$ACCC$G")) == [11, 13, 12, 16, 9, 26, 8, 11]
    assert len(M.convert_to_vec("AAAACCCGGGGNTTTTT$$")) == 6
    assert M.convert_to_vec("").size == 0 and M.convert_to_vec("\n\n").size == 0
    with pytest.raises(M.MsbwtError) as e:        # the reference panics (bwt_converter.rs:43-46)
        M.convert_to_vec("ACGX")
    assert e.value.code == 3 and "offset 3" in str(e.value)
    with pytest.raises(M.MsbwtError):
        M.convert_to_vec("ACGa")                   # lower case is not part of the text format


def test_convert_to_vec_equals_the_oracle_on_random_texts():
    rng = np.random.default_rng(11)
    alphabet = np.frombuffer(b"$ACGNT\n", dtype=np.uint8)
    for n in (1, 2, 31, 32, 33, 1000, 100_000):
        for p_same in (0.0, 0.7, 0.999):
            idx = rng.integers(0, 7, n)
            keep = rng.random(n) < p_same           # long runs: repeat the previous symbol
            for i in range(1, n):
                if keep[i]:
                    idx[i] = idx[i - 1]
            text = alphabet[idx].tobytes()
            assert (M.convert_to_vec(text) == O.convert_to_vec(text)).all(), (n, p_same)


# ---- bwt_converter.rs:259-321 ----
def test_save_bwt_numpy_is_byte_exact(tmp_path, two_string_npy):
    head = b"\x93NUMPY\x01\x00\x56\x00{'descr': '|u1', 'fortran_order': False, 'shape': (3, ), }"
    expect = head + b" " * (95 - len(head)) + b"\n" + bytes([1, 9, 25])
    p = str(tmp_path / "a.npy")
    M.save_bwt_numpy(M.convert_to_vec("A" * 3104), p)
    assert open(p, "rb").read() == expect
    assert list(np.load(p)) == [1, 9, 25]          # numpy itself reads what we wrote
    # run form (bwt_converter.rs:287-321): (A, 3104), ($, 1) -> the same header with 4 bytes
    r = str(tmp_path / "r.npy")
    M.save_bwt_runs_numpy([(1, 3104), (0, 1)], r)
    assert list(np.load(r)) == [1, 9, 25, 8]
    assert open(r, "rb").read()[:96] == expect[:96].replace(b"(3, )", b"(4, )")
    M.save_bwt_runs_numpy([(1, 0), (2, 5)], r)      # a zero count writes nothing
    assert list(np.load(r)) == [2 | (5 << 3)]
    # the reference's own fixture, regenerated: test_data/two_string.npy (a byte-identical copy is tests/golden/)
    want = open(two_string_npy, "rb").read()
    g = str(tmp_path / "two.npy")
    M.save_bwt_numpy(np.frombuffer(want[96:], dtype=np.uint8), g)
    assert open(g, "rb").read() == want
    # and against the oracle's writer on a multi-string BWT
    rle = M.convert_to_vec(naive.naive_bwt(["CCGT", "N", "ACG", "ACGTTTTTTTTGA"]))
    a, b = str(tmp_path / "x.npy"), str(tmp_path / "y.npy")
    M.save_bwt_numpy(rle, a)
    O.save_bwt_numpy(rle, b)
    assert open(a, "rb").read() == open(b, "rb").read()
    M.save_bwt_numpy(np.zeros(0, np.uint8), a)      # an empty BWT is a header alone
    assert len(open(a, "rb").read()) == 96 and np.load(a).size == 0


def test_save_reports_file_errors_as_eio(tmp_path):
    with pytest.raises(OSError):
        M.save_bwt_numpy([9, 10], str(tmp_path / "no_such_dir" / "a.npy"))
    with pytest.raises(M.MsbwtError) as e:
        M.save_bwt_runs_numpy([(7, 3)], str(tmp_path / "b.npy"))
    assert e.value.code == 1


def test_written_file_loads_back_through_the_oracle(tmp_path):
    """save -> load_numpy_file round trip (the oracle stands in for the loader here; the GPU round trip is in
    tests/test_gpu_parity.py)"""
    p = str(tmp_path / "b.npy")
    M.save_bwt_numpy(M.convert_to_vec(naive.naive_bwt(["CCGT", "N", "ACG"])), p)
    b = O.RleBWT()
    b.load_numpy_file(p)
    assert [b.get_symbol_count(i) for i in range(6)] == [3, 1, 3, 2, 1, 1]
    assert b.get_total_size() == 11
