"""ctypes front-end for the CPU oracle (oracle/msbwt_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/msbwt_oracle.h.  Importable only from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package never imports this module.

The class below mirrors the reference's `RleBWT` (src/rle_bwt.rs:14-322) +
the `BWT` trait (src/msbwt_core.rs:28-162) so parity tests can drive the
oracle and the CUDA path through the same calls.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmsbwt_oracle.so")

OK, ERR_IO, ERR_SHORT_HEADER, ERR_SIZE_MISMATCH = 0, 1, 2, 3
PANIC_SHORT_FILE, PANIC_HEADER_PARSE, PANIC_BAD_SYMBOL = 10, 11, 12


class OraclePanic(Exception):
    """The reference would `panic!` here (code says where)."""


class OracleIoError(OSError):
    """The reference returns `Err(io::Error)` here."""


def _cpu_signature() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def build(force: bool = False) -> str:
    """Compile the C restatement in place (gcc -O3 -march=native; seconds).  The library is
    rebuilt when the sources are newer or when it was built on a different CPU model
    (-march=native code from the build container must not run on the GPU box's host)."""
    src = os.path.join(_HERE, "msbwt_oracle.c")
    hdr = os.path.join(_HERE, "msbwt_oracle.h")
    tag = _SO + ".cpu"
    sig = _cpu_signature()
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    try:
        stale = stale or open(tag).read() != sig
    except OSError:
        stale = True
    if force or stale:
        subprocess.check_call(["make", "-s", "-B", "-C", _HERE, "libmsbwt_oracle.so"])
        with open(tag, "w") as f:
            f.write(sig)
    return _SO


class _Range(C.Structure):
    _fields_ = [("l", C.c_uint64), ("h", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    u8p, u64p, vp = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.c_void_p
    sig = {
        "orc_new": (vp, [C.c_uint]),
        "orc_free": (None, [vp]),
        "orc_load_vector": (C.c_int, [vp, vp, C.c_uint64]),
        "orc_load_numpy_file": (C.c_int, [vp, C.c_char_p]),
        "orc_get_symbol_count": (C.c_uint64, [vp, C.c_uint8]),
        "orc_get_total_size": (C.c_uint64, [vp]),
        "orc_start_index": (C.c_uint64, [vp, C.c_uint8]),
        "orc_end_index": (C.c_uint64, [vp, C.c_uint8]),
        "orc_index_length": (C.c_uint64, [vp]),
        "orc_ref_index": (u64p, [vp]),
        "orc_fm_index": (u64p, [vp, C.c_uint8]),
        "orc_rle_len": (C.c_uint64, [vp]),
        "orc_rle_bytes": (u8p, [vp]),
        "orc_constrain_range": (_Range, [vp, C.c_uint8, _Range]),
        "orc_count_kmer": (C.c_int, [vp, vp, C.c_uint64, u64p]),
        "orc_count_kmers_fixed": (C.c_int, [vp, vp, C.c_uint32, C.c_uint64, vp, C.c_int]),
        "orc_count_kmers": (C.c_int, [vp, vp, vp, C.c_uint64, vp, C.c_int]),
        "orc_count_kmers_stats": (C.c_int, [vp, vp, C.c_uint32, C.c_uint64, C.c_uint, u64p, u64p]),
        "orc_count_kmers_stats_skip": (C.c_int, [vp, vp, C.c_uint32, C.c_uint64, C.c_uint, C.c_uint32, u64p, u64p, u64p]),
        "orc_count_kmers_stats_pair": (C.c_int, [vp, vp, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint, u64p]),
        "orc_count_kmers_stats_quad": (C.c_int, [vp, vp, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint, u64p]),
        "orc_count_kmers_stats_oct": (C.c_int, [vp, vp, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint, C.c_uint, C.c_uint, u64p]),
        "orc_count_kmers_stats_fin": (C.c_int, [vp, vp, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, u64p, u64p]),
        "orc_convert_to_vec": (C.c_uint64, [vp, C.c_uint64, vp, C.c_uint64]),
        "orc_encode_runs": (C.c_uint64, [vp, vp, C.c_uint64, vp, C.c_uint64]),
        "orc_save_bwt_numpy": (C.c_int, [vp, C.c_uint64, C.c_char_p]),
        "orc_string_to_int": (C.c_uint8, [C.c_uint8]),
        "orc_convert_stoi": (None, [vp, C.c_uint64, vp]),
        "orc_convert_itos": (None, [vp, C.c_uint64, vp]),
        "orc_reverse_complement_i": (None, [vp, C.c_uint64, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint8))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _raise(rc: int, what: str):
    if rc == OK:
        return
    if rc >= 10:
        raise OraclePanic(f"{what}: reference would panic (code {rc})")
    raise OracleIoError(rc, f"{what}: reference returns io::Error (code {rc})")


# ---- string_util.rs ----
def convert_stoi(s: str | bytes) -> np.ndarray:
    b = s.encode() if isinstance(s, str) else bytes(s)
    src = np.frombuffer(b, dtype=np.uint8)
    out = np.empty(len(b), dtype=np.uint8)
    lib().orc_convert_stoi(_ptr(_u8(src)), len(b), _ptr(out))
    return out


def convert_itos(syms) -> str:
    a = _u8(syms)
    out = np.empty(a.size, dtype=np.uint8)
    lib().orc_convert_itos(_ptr(a), a.size, _ptr(out))
    return out.tobytes().decode()


def reverse_complement_i(syms) -> np.ndarray:
    a = _u8(syms)
    out = np.empty(a.size, dtype=np.uint8)
    lib().orc_reverse_complement_i(_ptr(a), a.size, _ptr(out))
    return out


# ---- bwt_converter.rs ----
def convert_to_vec(ascii_bwt: str | bytes) -> np.ndarray:
    b = ascii_bwt.encode() if isinstance(ascii_bwt, str) else bytes(ascii_bwt)
    src = _u8(np.frombuffer(b, dtype=np.uint8))
    n = lib().orc_convert_to_vec(_ptr(src), src.size, None, 0)
    if n == 2**64 - 1:
        raise OraclePanic("convert_to_vec: unexpected symbol")
    out = np.empty(n, dtype=np.uint8)
    lib().orc_convert_to_vec(_ptr(src), src.size, _ptr(out), n)
    return out


def encode_runs(syms, counts) -> np.ndarray:
    s = _u8(syms)
    c = np.ascontiguousarray(np.asarray(counts, dtype=np.uint64))
    n = lib().orc_encode_runs(_ptr(s), _ptr(c), s.size, None, 0)
    out = np.empty(n, dtype=np.uint8)
    lib().orc_encode_runs(_ptr(s), _ptr(c), s.size, _ptr(out), n)
    return out


def save_bwt_numpy(rle, path: str) -> None:
    a = _u8(rle)
    _raise(lib().orc_save_bwt_numpy(_ptr(a), a.size, path.encode()), "save_bwt_numpy")


class RleBWT:
    """Oracle `RleBWT`: same method names/semantics as the reference."""

    def __init__(self, bin_power: int = 8):
        self._h = lib().orc_new(bin_power)
        self.bin_power = bin_power

    @classmethod
    def with_bin_power(cls, bin_power: int) -> "RleBWT":
        return cls(bin_power)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.orc_free(h)

    def load_vector(self, rle) -> None:
        a = _u8(rle)
        _raise(lib().orc_load_vector(self._h, _ptr(a), a.size), "load_vector")

    def load_numpy_file(self, path: str) -> None:
        _raise(lib().orc_load_numpy_file(self._h, os.fsencode(path)), "load_numpy_file")

    def get_symbol_count(self, sym: int) -> int:
        return int(lib().orc_get_symbol_count(self._h, sym))

    def get_total_size(self) -> int:
        return int(lib().orc_get_total_size(self._h))

    def start_index(self, sym: int) -> int:
        return int(lib().orc_start_index(self._h, sym))

    def end_index(self, sym: int) -> int:
        return int(lib().orc_end_index(self._h, sym))

    @property
    def ref_index(self) -> list[int]:
        n = lib().orc_index_length(self._h)
        p = lib().orc_ref_index(self._h)
        return [int(p[i]) for i in range(n)]

    def fm_index(self, sym: int) -> list[int]:
        n = lib().orc_index_length(self._h)
        p = lib().orc_fm_index(self._h, sym)
        return [int(p[i]) for i in range(n)]

    def rle_bytes(self) -> np.ndarray:
        n = lib().orc_rle_len(self._h)
        p = lib().orc_rle_bytes(self._h)
        return np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.empty(0, np.uint8)

    def constrain_range(self, sym: int, l: int, h: int) -> tuple[int, int]:
        r = lib().orc_constrain_range(self._h, sym, _Range(l, h))
        return int(r.l), int(r.h)

    def count_kmer(self, kmer) -> int:
        a = _u8(kmer)
        out = C.c_uint64(0)
        _raise(lib().orc_count_kmer(self._h, _ptr(a), a.size, C.byref(out)), "count_kmer")
        return int(out.value)

    def count_kmers_fixed(self, syms, k: int, threads: int = 1) -> np.ndarray:
        a = _u8(syms).reshape(-1)
        n = a.size // k if k else 0
        out = np.empty(n, dtype=np.uint64)
        _raise(lib().orc_count_kmers_fixed(self._h, _ptr(a), k, n, _ptr(out), threads), "count_kmers")
        return out

    def count_kmers(self, kmers, threads: int = 1) -> np.ndarray:
        """Batched form over a list of variable-length k-mers."""
        lens = np.fromiter((len(q) for q in kmers), dtype=np.uint64, count=len(kmers))
        offs = np.zeros(len(kmers) + 1, dtype=np.uint64)
        np.cumsum(lens, out=offs[1:])
        flat = _u8(np.concatenate([_u8(q) for q in kmers]) if len(kmers) else np.empty(0, np.uint8))
        out = np.empty(len(kmers), dtype=np.uint64)
        _raise(lib().orc_count_kmers(self._h, _ptr(flat), _ptr(offs), len(kmers), _ptr(out), threads),
               "count_kmers")
        return out

    def count_kmers_stats_skip(self, syms, k: int, block_shift: int, skip: int) -> tuple[int, int, int]:
        """(steps, two_block_steps, table_hits) for an engine whose suffix table answers the first `skip` steps."""
        a = _u8(syms).reshape(-1)
        n = a.size // k
        st, tb, th = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        _raise(lib().orc_count_kmers_stats_skip(self._h, _ptr(a), k, n, block_shift, skip, C.byref(st), C.byref(tb),
                                                C.byref(th)), "count_kmers_stats_skip")
        return int(st.value), int(tb.value), int(th.value)

    def count_kmers_stats_pair(self, syms, k: int, table_s: int, line_syms: int = 96, block_shift: int = 7) -> dict:
        """Accounting replay of the engine's pair path: index lines a batch must touch."""
        a = _u8(syms).reshape(-1)
        n = a.size // k
        out = (C.c_uint64 * 6)()
        _raise(lib().orc_count_kmers_stats_pair(self._h, _ptr(a), k, n, table_s, line_syms, block_shift, out),
               "count_kmers_stats_pair")
        keys = ("pair_steps", "two_line_pair_steps", "one_steps", "two_block_one_steps", "table_hits", "queries")
        return dict(zip(keys, (int(v) for v in out)))

    def count_kmers_stats_quad(self, syms, k: int, table_s: int, sector_syms: int = 224, line_sectors: int = 4,
                               block_shift: int = 7, oct_bucket_shift: int = 0, oct_syms: int = 8,
                               fin_bucket_shift: int = 0, fin_syms: int = 20) -> dict:
        """Accounting replay of the engine's quad path (oct_bucket_shift != 0: with the oct image on top;
        fin_bucket_shift != 0: and the experimental final-step image): index sectors / lines a batch must touch."""
        a = _u8(syms).reshape(-1)
        n = a.size // k
        out = (C.c_uint64 * 9)()
        fin = (C.c_uint64 * 2)()
        _raise(lib().orc_count_kmers_stats_fin(self._h, _ptr(a), k, n, table_s, sector_syms, line_sectors,
                                               block_shift, oct_bucket_shift, oct_syms, fin_bucket_shift, fin_syms, out, fin),
               "count_kmers_stats_quad")
        keys = ("quad_steps", "two_sector_quad_steps", "two_line_quad_steps", "one_steps", "two_block_one_steps",
                "table_hits", "queries", "oct_steps", "two_bucket_oct_steps")
        d = dict(zip(keys, (int(v) for v in out)))
        d["final_steps"], d["two_bucket_final_steps"] = int(fin[0]), int(fin[1])
        return d

    def count_kmers_stats(self, syms, k: int, block_shift: int = 8) -> tuple[int, int]:
        a = _u8(syms).reshape(-1)
        n = a.size // k
        st, tb = C.c_uint64(0), C.c_uint64(0)
        _raise(lib().orc_count_kmers_stats(self._h, _ptr(a), k, n, block_shift, C.byref(st), C.byref(tb)),
               "count_kmers_stats")
        return int(st.value), int(tb.value)
