"""The device-side multi-string BWT builder (bwt_build.cu, `msbwt_build_rle_bwt`) must produce what
`msbwt2-build` produces with its default sorted insertion: naive_bwt's order (src/bwt_util.rs:154-171,
src/dynamic_bwt.rs:515-525) in the RLE byte format of src/bwt_converter.rs:52-56."""
import numpy as np
import pytest

import rust_msbwt_b200 as M
from oracle import naive
from oracle import oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _as_strings(reads):
    return [O.convert_itos(r) for r in reads]


def test_reference_shapes_match_naive_bwt():
    # dynamic_bwt.rs:551-577 test_sampled_bwt: 36 windows x 32 copies of 20-mers
    genome = "ACCGTGTTGCCGTAGTGAAAAGTGACGACGTGAGATGGCCAAAGTGGGTCTCTGTG"
    data = [genome[s:s + 20] for s in range(len(genome) - 20) for _ in range(32)]
    reads = np.stack([O.convert_stoi(s) for s in data])
    rle, total = M.build_rle_bwt(reads)
    assert total == len(data) * 21
    assert (rle == O.convert_to_vec(naive.naive_bwt(data))).all()
    # rle_bwt.rs:76-79 / README.md:62-70: the two strings of test_data/two_string.npy
    rle, total = M.build_rle_bwt(np.stack([O.convert_stoi("ACGT"), O.convert_stoi("TGCA")]))
    assert total == 10 and (rle == O.convert_to_vec(naive.naive_bwt(["ACGT", "TGCA"]))).all()


def test_two_string_fixture_is_reproduced(two_string_npy):
    o = O.RleBWT()
    o.load_numpy_file(two_string_npy)
    rle, total = M.build_rle_bwt(np.stack([O.convert_stoi("ACGT"), O.convert_stoi("TGCA")]))
    g = M.RleBWT.new()
    g.load_vector(rle)
    assert g.get_total_size() == o.get_total_size() == total
    for s in range(6):
        assert g.get_symbol_count(s) == o.get_symbol_count(s)
    assert (rle == o.rle_bytes()).all()


def test_random_reads_with_n_and_duplicates_match_naive_bwt():
    rng = np.random.default_rng(5)
    for L in (1, 5, 20, 21, 22, 41, 42, 43, 49, 63, 64):
        reads = rng.choice(np.array([1, 2, 3, 4, 5], dtype=np.uint8), size=(60, L), p=[0.3, 0.2, 0.2, 0.05, 0.25])
        reads[7] = reads[3]
        rle, total = M.build_rle_bwt(reads)
        assert total == 60 * (L + 1)
        assert (rle == O.convert_to_vec(naive.naive_bwt(_as_strings(reads)))).all(), L
    one = np.array([[2, 2, 3, 5]], dtype=np.uint8)
    assert (M.build_rle_bwt(one)[0] == O.convert_to_vec(naive.naive_bwt(["CCGT"]))).all()
    assert M.build_rle_bwt(np.zeros((0, 4), np.uint8)) [1] == 0


def test_long_runs_and_midsize_equal_the_harness_builder():
    from harness import bwt_build, synth
    same = np.full((5000, 30), 1, dtype=np.uint8)      # one run of 150 000 'A': four base-32 digits
    rle, total = M.build_rle_bwt(same)
    assert (rle == bwt_build.build_rle_bwt(torch.from_numpy(same).cuda())[0].cpu().numpy()).all()
    reads = synth.make_reads(200_000, read_len=150, coverage=30.0, error_rate=0.01, device="cuda")
    reads[17, 40:43] = 4
    want, n = bwt_build.build_rle_bwt(reads)
    got, total = M.build_rle_bwt(reads.data_ptr(), 0, reads.shape[0], reads.shape[1])   # device pointer form
    assert total == n == 200_000 * 151
    assert got.size == want.numel() and (got == want.cpu().numpy()).all()
    got2, _ = M.build_rle_bwt(reads.cpu().numpy())                                      # host form
    assert (got2 == got).all()


def test_reads_of_different_lengths_match_naive_bwt():
    """N2: create_from_fastx takes reads as they come (src/dynamic_bwt.rs:453-473); `msbwt_build_rle_bwt_ragged` must
    order their suffixes as naive_bwt does (src/bwt_util.rs:154-171: doubled rotations, '$' smallest)"""
    # the reference's own mixed-length examples: bwt_util.rs doc-test and dynamic_bwt.rs tests
    for data, want in ((["CCGT", "N", "ACG"], "GTN$$ACCC$G"), (["ACA", "CA"], "AACC$A$")):
        rle, total = M.build_rle_bwt_ragged(data)
        assert total == sum(len(s) + 1 for s in data)
        assert (rle == O.convert_to_vec(want)).all(), data
    rng = np.random.default_rng(9)
    for trial in range(6):
        lens = rng.integers(0, 70, 80)                  # empty reads, lengths around the 21-symbol key word, prefixes
        reads = [rng.choice(np.array([1, 2, 3, 4, 5], dtype=np.uint8), size=int(n), p=[0.3, 0.2, 0.2, 0.05, 0.25]) for n in lens]
        reads[5] = reads[9][: len(reads[9]) // 2].copy()   # a read that is a prefix of another
        reads[6] = reads[9].copy()                         # a duplicate
        reads[7] = reads[9][len(reads[9]) // 3:].copy()    # a read that is a suffix of another
        rle, total = M.build_rle_bwt_ragged(reads)
        assert total == sum(len(r) + 1 for r in reads)
        assert (rle == O.convert_to_vec(naive.naive_bwt(_as_strings(reads)))).all(), trial
    # equal lengths through the ragged entry = the fixed-length builder
    same = rng.choice(np.array([1, 2, 3, 5], dtype=np.uint8), size=(300, 37))
    assert (M.build_rle_bwt_ragged(list(same))[0] == M.build_rle_bwt(same)[0]).all()
    assert M.build_rle_bwt_ragged([])[1] == 0
    # the index built from it answers like the oracle on the same bytes
    reads = [rng.choice(np.array([1, 2, 3, 5], dtype=np.uint8), size=int(n)) for n in rng.integers(20, 120, 3000)]
    rle, total = M.build_rle_bwt_ragged(reads)
    o = O.RleBWT()
    o.load_vector(rle)
    g = M.RleBWT.new()
    g.load_vector(rle)
    assert g.get_total_size() == o.get_total_size() == total
    q = np.stack([r[:12] for r in reads if len(r) >= 12][:2000])
    assert (g.count_kmers_fixed(q, 12) == o.count_kmers_fixed(q, 12)).all()
    assert int((g.count_kmers_fixed(q, 12) > 0).all())


def test_bad_symbols_are_refused():
    reads = np.full((10, 8), 2, dtype=np.uint8)
    reads[3, 4] = 0
    with pytest.raises(M.MsbwtError) as e:
        M.build_rle_bwt(reads)
    assert e.value.code == 1
    reads[3, 4] = 6
    with pytest.raises(M.MsbwtError):
        M.build_rle_bwt(reads)
    with pytest.raises(M.MsbwtError):
        M.build_rle_bwt_ragged([np.array([1, 2, 0, 3], dtype=np.uint8), np.array([1], dtype=np.uint8)])
