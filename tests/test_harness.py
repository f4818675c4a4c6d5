"""The harness's BWT builder must equal `naive_bwt` (src/bwt_util.rs:154-171), its RLE
encoder must equal `convert_to_vec` (src/bwt_converter.rs:26-80).  CPU only."""
import numpy as np
import torch

from harness import bwt_build, synth
from oracle import naive
from oracle import oracle as O


def _as_strings(reads):
    return [O.convert_itos(r) for r in reads]


def test_builder_matches_naive_on_reference_shapes():
    # dynamic_bwt.rs:551-577 test_sampled_bwt: 36 windows x 32 copies of 20-mers
    genome = "ACCGTGTTGCCGTAGTGAAAAGTGACGACGTGAGATGGCCAAAGTGGGTCTCTGTG"
    data = [genome[s:s + 20] for s in range(len(genome) - 20) for _ in range(32)]
    reads = torch.from_numpy(np.stack([O.convert_stoi(s) for s in data]))
    bwt = bwt_build.build_msbwt(reads)
    assert O.convert_itos(bwt.numpy()) == naive.naive_bwt(data)


def test_builder_matches_naive_random_with_n_and_dups():
    rng = np.random.default_rng(5)
    for L in (1, 5, 23, 24, 25, 49):
        reads = rng.choice(np.array([1, 2, 3, 4, 5], dtype=np.uint8), size=(60, L), p=[0.3, 0.2, 0.2, 0.05, 0.25])
        reads[7] = reads[3]
        t = torch.from_numpy(reads)
        bwt = bwt_build.build_msbwt(t)
        assert O.convert_itos(bwt.numpy()) == naive.naive_bwt(_as_strings(reads)), L
        rle = bwt_build.rle_encode(bwt).numpy()
        assert (rle == O.convert_to_vec(naive.naive_bwt(_as_strings(reads)))).all()


def test_rle_encode_long_runs():
    for text in ("A" * 3104, "A" * 31 + "C" * 31, "N" * 32767, "$" * 32 + "T" * 1024 + "G"):
        got = bwt_build.rle_encode(torch.from_numpy(O.convert_stoi(text))).numpy()
        assert list(got) == list(O.convert_to_vec(text))


def test_synthetic_read_sampled_queries_are_found():
    reads = synth.make_reads(300, read_len=40, coverage=10.0, error_rate=0.02)
    q = synth.make_queries(reads, 12, 200, 200)
    assert q.shape == (400, 12) and int(q.min()) >= 1
    rle, n = bwt_build.build_rle_bwt(reads)
    b = O.RleBWT()
    b.load_vector(rle.numpy())
    assert b.get_total_size() == n == 300 * 41
    counts = b.count_kmers_fixed(q.numpy(), 12)
    # every read-sampled k-mer occurs at least once; random 12-mers over a 1.2 kb genome mostly do not
    strings = _as_strings(reads.numpy())
    for i in range(0, 400, 7):
        assert counts[i] == naive.brute_count(strings, O.convert_itos(q[i].numpy()))
    assert (counts > 0).sum() >= 200
    r2 = synth.np_make_reads(50, 30, 8.0, 0.01, 1)
    q2 = synth.np_make_queries(r2, 9, 20, 20, 2)
    assert r2.shape == (50, 30) and q2.shape == (40, 9)
