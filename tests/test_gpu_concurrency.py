"""Concurrent callers (SURVEY §8b threading: `RleBWT` is `Send + Sync`, queries borrow `&self`, src/rle_bwt.rs:14-24).
ADVICE r1: (1) two threads querying two DIFFERENT handles used to share one process-wide packer pool and could hang or
pack wrongly -- every replica now owns its pool; (2) the asynchronous device entry point used to share its scratch
between calls on different streams -- calls are now ordered by an event.  Both are exercised here, against the oracle."""
import threading

import numpy as np
import pytest

import rust_msbwt_b200 as M
from oracle import oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def midsize():
    from harness import bwt_build, synth
    reads = synth.make_reads(20000, read_len=100, coverage=25.0, error_rate=0.01, device="cuda")
    rle, _ = bwt_build.build_rle_bwt(reads)
    o = O.RleBWT()
    o.load_vector(rle.cpu().numpy())
    return reads, o


def test_two_threads_two_handles_and_one_shared_handle(midsize, monkeypatch):
    from harness import synth
    reads, o = midsize
    monkeypatch.setenv("MSBWT_HOST_PACK", "1")       # the packed route: the one that uses the worker pools
    monkeypatch.setenv("MSBWT_HOST_THREADS", "4")
    a = M.RleBWT(oct_index=1)
    a.load_vector(o.rle_bytes())
    b = M.RleBWT(pair_index=1)
    b.load_vector(o.rle_bytes())
    qs = [synth.make_queries(reads, 31, 300_001, 100_000, seed_offset=s).cpu().numpy() for s in range(4)]
    wants = [o.count_kmers_fixed(q[:50_000], 31, threads=8) for q in qs]
    errors = []

    def worker(handle, i, rounds):
        try:
            for _ in range(rounds):
                got = handle.count_kmers_fixed(qs[i], 31)
                if not (got[:50_000] == wants[i]).all():
                    errors.append(("counts differ", i))
        except Exception as e:   # noqa: BLE001
            errors.append((repr(e), i))

    # two handles, one thread each; then four threads over the two handles (two of them share a handle)
    for plan in ([(a, 0), (b, 1)], [(a, 0), (b, 1), (a, 2), (b, 3)]):
        threads = [threading.Thread(target=worker, args=(h, i, 6)) for h, i in plan]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=300)
        assert not any(t.is_alive() for t in threads), "a caller hangs"
        assert not errors, errors


def test_device_entry_point_on_two_streams_shares_scratch_safely(midsize):
    """back-to-back asynchronous calls on DIFFERENT streams: the second must not start rewriting the replica's
    pack / seed scratch while the first one's kernels still read it"""
    from harness import synth
    reads, o = midsize
    g = M.RleBWT(oct_index=1, final_index=0)       # the general kernels: pack kernel + search kernel share the scratch
    g.load_vector(o.rle_bytes())
    k = 31
    qa = synth.make_queries(reads, k, 2_000_001, 500_000, seed_offset=7)
    qb = synth.make_queries(reads, k, 2_000_001, 500_000, seed_offset=8)
    wa = o.count_kmers_fixed(qa[:100_000].cpu().numpy(), k, threads=8)
    wb = o.count_kmers_fixed(qb[:100_000].cpu().numpy(), k, threads=8)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(5):
        oa = torch.zeros(qa.shape[0], dtype=torch.int64, device="cuda")
        ob = torch.zeros(qb.shape[0], dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        g.count_kmers_fixed_device(qa.data_ptr(), k, qa.shape[0], oa.data_ptr(), 0, s1.cuda_stream)
        g.count_kmers_fixed_device(qb.data_ptr(), k, qb.shape[0], ob.data_ptr(), 0, s2.cuda_stream)
        torch.cuda.synchronize()
        assert (oa[:100_000].cpu().numpy().view(np.uint64) == wa).all()
        assert (ob[:100_000].cpu().numpy().view(np.uint64) == wb).all()
