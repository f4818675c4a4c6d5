"""Final-step image (layout.h, fin_builder.cu, the `is_fin` step of count_kmers_oct_kernel): the last 20 symbols of a
k-mer answered from ONE hashed line.  Checked against its numpy specification (oracle/final_step.py) line by line, and
against the oracle's count_kmer (src/msbwt_core.rs:125-161) for the lengths that reach the step
(k = table depth + 10 a + 20) and for some that do not -- with the quad image kept beside it and dropped."""
import numpy as np
import pytest

import rust_msbwt_b200 as M
from oracle import final_step as F
from oracle import oracle as O
from tests.test_oracle_final_step import decode

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def make_index(rle, monkeypatch, shift=16, lb=12, **kw):
    g = M.RleBWT(oct_index=1, final_index=1, final_bucket_shift=shift, final_lines_log2=lb, **kw)
    g.load_vector(rle)
    assert g.final_index and g.oct_index
    return g


@pytest.fixture(scope="module")
def midsize():
    from harness import bwt_build, synth
    reads = synth.make_reads(20000, read_len=100, coverage=25.0, error_rate=0.01, device="cuda")
    reads[17, 40:43] = 4
    rle, _ = bwt_build.build_rle_bwt(reads)
    o = O.RleBWT()
    o.load_vector(rle.cpu().numpy())
    return reads, o


@pytest.mark.parametrize("shift,lb", [(16, 12), (12, 12), (10, 13)])
def test_device_image_equals_the_specification(midsize, monkeypatch, shift, lb):
    reads, o = midsize
    g = make_index(o.rle_bytes(), monkeypatch, shift, lb)
    lines, b, l2, over = g.final_image()
    assert (b, l2) == (shift, lb)
    want, stats = F.build_final_image(decode(o.rle_bytes()), b=shift, lb=lb)
    assert lines.shape == want.shape and over == stats["overflowed_lines"]
    assert ((lines[:, 0] == F.OVERFLOW) == (want[:, 0] == F.OVERFLOW)).all()
    for i in np.flatnonzero((want[:, 0] != 0) | (lines[:, 0] != 0)):
        assert F.line_groups(lines[i]) == F.line_groups(want[i]), i


@pytest.mark.parametrize("shift,lb", [(16, 12), (11, 13)])
def test_device_image_of_a_wide_index_equals_the_specification(midsize, monkeypatch, shift, lb):
    """64-bit positions: the 20-symbol codes come from walking LF through the one-step blocks (fin_builder.cu
    build_fin_codes_by_walk); the image has no positions in it, so it must equal the same specification"""
    reads, o = midsize
    g = make_index(o.rle_bytes(), monkeypatch, shift, lb, superblock_shift=4)
    assert not g.quad_index
    lines, b, l2, over = g.final_image()
    assert (b, l2) == (shift, lb)
    want, stats = F.build_final_image(decode(o.rle_bytes()), b=shift, lb=lb)
    assert lines.shape == want.shape and over == stats["overflowed_lines"]
    assert ((lines[:, 0] == F.OVERFLOW) == (want[:, 0] == F.OVERFLOW)).all()
    for i in np.flatnonzero((want[:, 0] != 0) | (lines[:, 0] != 0)):
        assert F.line_groups(lines[i]) == F.line_groups(want[i]), i


@pytest.mark.parametrize("keep_quad", [1, 0])
@pytest.mark.parametrize("shift,lb,table_s", [(16, 12, -1), (12, 12, -1), (10, 13, 7), (16, 12, 0), (14, 12, 12)])
def test_counts_are_bit_exact_with_the_final_step(midsize, monkeypatch, shift, lb, table_s, keep_quad):
    from harness import synth
    reads, o = midsize
    g = make_index(o.rle_bytes(), monkeypatch, shift, lb, suffix_table_s=table_s, keep_quad_index=keep_quad)
    assert g.quad_index == bool(keep_quad)
    for k in (20, 21, 24, 30, 31, 32, 33, 40, 41, 42, 51, 61, 64, 71, 100):
        q = synth.make_queries(reads, k, 12001, 6000).cpu().numpy()
        q[5, 0] = 4
        q[11, k // 2] = 4
        got = g.count_kmers_fixed(q, k)
        want = o.count_kmers_fixed(q, k, threads=8)
        assert (got == want).all(), (shift, lb, table_s, k, np.flatnonzero(got != want)[:5])


def test_overflowed_lines_fall_back(monkeypatch):
    """a 2 kb genome at 1500x: few codes own every position, their lines cannot hold the runs"""
    from harness import bwt_build, synth
    reads = synth.make_reads(30000, read_len=100, coverage=1500.0, error_rate=0.01, device="cuda")
    rle = bwt_build.build_rle_bwt(reads)[0].cpu().numpy()
    o = O.RleBWT()
    o.load_vector(rle)
    for keep_quad in (1, 0):
        g = make_index(rle, monkeypatch, keep_quad_index=keep_quad)
        assert g.final_image()[3] > 0 and g.oct_overflow_lines > 0
        for k in (31, 37, 41, 64):
            q = synth.make_queries(reads, k, 20000, 5000).cpu().numpy()
            assert (g.count_kmers_fixed(q, k) == o.count_kmers_fixed(q, k, threads=8)).all(), (keep_quad, k)


def test_automatic_policy_builds_it_with_the_oct_image_and_can_be_switched_off(midsize, monkeypatch):
    from harness import synth
    reads, o = midsize
    g = M.RleBWT(oct_index=1)
    g.load_vector(o.rle_bytes())
    assert g.final_index and g.final_bucket_shift == 16
    off = M.RleBWT(oct_index=1, final_index=0)
    off.load_vector(o.rle_bytes())
    assert not off.final_index and off.oct_index
    monkeypatch.setenv("MSBWT_FINAL_INDEX", "0")
    env_off = M.RleBWT(oct_index=1)
    env_off.load_vector(o.rle_bytes())
    assert not env_off.final_index
    q = synth.make_queries(reads, 31, 50001, 20000).cpu().numpy()
    want = o.count_kmers_fixed(q, 31, threads=8)
    for b in (g, off, env_off):
        assert (b.count_kmers_fixed(q, 31) == want).all()
        assert (b.count_kmers_fixed(q, 31, counts32=True) == want).all()
