"""ctypes binding of the CPU oracle (oracle/herro_oracle.cpp).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package herro_b200/ never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libherro_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "herro_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, u32p, u64p, i32p, f32p = (C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_int32), C.POINTER(C.c_float))
        L.ho_last_error.restype = C.c_char_p
        L.ho_encode.argtypes = [u8p, C.c_uint64, u64p]
        L.ho_decode.argtypes = [u64p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, u8p]
        L.ho_cigar_iter.argtypes = [u8p, C.c_uint64, u32p, C.c_uint64]
        L.ho_extract_windows.argtypes = [u32p, u8p, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, u32p, C.c_uint64]
        L.ho_reads_new.restype = C.c_void_p
        L.ho_reads_new.argtypes = [C.c_uint32, u8p, u8p, u64p, u8p, u64p, u8p, u64p, u8p]
        L.ho_reads_free.argtypes = [C.c_void_p]
        L.ho_features.restype = C.c_void_p
        L.ho_features.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, u32p, u8p, u64p, C.c_uint32, C.c_uint32]
        L.ho_target_free.argtypes = [C.c_void_p]
        L.ho_n_windows.argtypes = [C.c_void_p]
        L.ho_n_windows.restype = C.c_uint32
        L.ho_window_info.argtypes = [C.c_void_p, C.c_uint32, u32p]
        L.ho_window_get.argtypes = [C.c_void_p, C.c_uint32, u8p, u8p, u32p, u32p, u32p]
        L.ho_n_batches.argtypes = [C.c_void_p]
        L.ho_n_batches.restype = C.c_uint32
        L.ho_batch_shape.argtypes = [C.c_void_p, C.c_uint32, u32p]
        L.ho_batch_get.argtypes = [C.c_void_p, C.c_uint32, u8p, u8p, i32p, u32p, i32p]
        L.ho_set_logits.argtypes = [C.c_void_p, C.c_uint32, f32p, f32p, C.c_uint32]
        L.ho_consensus.argtypes = [C.c_void_p]
        L.ho_seg_len.argtypes = [C.c_void_p, C.c_uint32]
        L.ho_seg_len.restype = C.c_uint64
        L.ho_seg_get.argtypes = [C.c_void_p, C.c_uint32, u8p]
        L.ho_fasta.argtypes = [C.c_void_p, u8p, C.c_uint64]
        L.ho_fasta.restype = C.c_int64
        _lib = L
    return _lib


class OraclePanic(RuntimeError):
    """The restated reference code hit one of its assert!/panic!/unwrap sites."""


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _err():
    return OraclePanic(lib().ho_last_error().decode())


# --------------------------------------------------------------------------- codec
def encode(seq: bytes) -> np.ndarray:
    s = np.frombuffer(seq, dtype=np.uint8)
    out = np.zeros((len(seq) + 31) // 32, dtype=np.uint64)
    if lib().ho_encode(_p(s, C.c_uint8), len(seq), _p(out, C.c_uint64)) != 0:
        raise _err()
    return out


def decode(words: np.ndarray, length: int, start: int = 0, end: int | None = None, rc: bool = False) -> bytes:
    end = length if end is None else end
    words = np.ascontiguousarray(words, dtype=np.uint64)
    out = np.zeros(max(end - start, 0), dtype=np.uint8)
    if lib().ho_decode(_p(words, C.c_uint64), length, start, end, int(rc), _p(out, C.c_uint8)) != 0:
        raise _err()
    return out.tobytes()


def cigar_iter(cigar: bytes):
    c = np.frombuffer(cigar, dtype=np.uint8)
    cap = len(cigar) // 2 + 1
    out = np.zeros((cap, 4), dtype=np.uint32)
    n = lib().ho_cigar_iter(_p(c, C.c_uint8), len(cigar), _p(out, C.c_uint32), cap)
    if n < 0:
        raise _err()
    kinds = {0: "M", 2: "I", 3: "D"}
    return [(kinds[int(k)], int(l), int(s), int(e)) for k, l, s, e in out[:n]]


def extract_windows(ovl9, cigar: bytes, window_size: int, n_windows: int, is_target: bool = True):
    """-> list of (window_idx, tstart, qstart, qend, csi, cso, cei, ceo)"""
    o = np.asarray(ovl9, dtype=np.uint32)
    c = np.frombuffer(cigar, dtype=np.uint8)
    cap = n_windows + 4
    out = np.zeros((cap, 8), dtype=np.uint32)
    n = lib().ho_extract_windows(_p(o, C.c_uint32), _p(c, C.c_uint8), len(cigar), int(is_target), window_size,
                                 n_windows, _p(out, C.c_uint32), cap)
    if n < 0:
        raise _err()
    return [tuple(int(x) for x in r) for r in out[:n]]


# --------------------------------------------------------------------------- read store
class Reads:
    """HAECRecord store (src/haec_io.rs:19-24).  seqs/quals: list[bytes]."""

    def __init__(self, ids, seqs, quals, descs=None):
        n = len(seqs)
        self.n = n
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(s) for s in seqs])
        sq = np.frombuffer(b"".join(seqs), dtype=np.uint8)
        ql = np.frombuffer(b"".join(quals), dtype=np.uint8)
        ids_b = [i if isinstance(i, bytes) else i.encode() for i in ids]
        ioff = np.zeros(n + 1, dtype=np.uint64)
        ioff[1:] = np.cumsum([len(s) for s in ids_b])
        idb = np.frombuffer(b"".join(ids_b) or b"\0", dtype=np.uint8)
        if descs is None:
            descs = [None] * n
        d_b = [(d if isinstance(d, bytes) else d.encode()) if d is not None else b"" for d in descs]
        doff = np.zeros(n + 1, dtype=np.uint64)
        doff[1:] = np.cumsum([len(s) for s in d_b])
        db = np.frombuffer(b"".join(d_b) or b"\0", dtype=np.uint8)
        has = np.array([d is not None for d in descs], dtype=np.uint8)
        self._h = lib().ho_reads_new(n, _p(sq, C.c_uint8), _p(ql, C.c_uint8), _p(off, C.c_uint64),
                                     _p(idb, C.c_uint8), _p(ioff, C.c_uint64), _p(db, C.c_uint8),
                                     _p(doff, C.c_uint64), _p(has, C.c_uint8))
        if not self._h:
            raise _err()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ho_reads_free(self._h)
            self._h = None


@dataclass
class Window:
    wid: int
    n_total_wins: int
    n_alns: int
    bases: np.ndarray        # [L',31] u8 tokens (BASES_MAP)
    quals: np.ndarray        # [L',31] u8 raw
    supported: np.ndarray    # [n,2] (pos, ins)
    sup_rows: np.ndarray     # [n] row index = indices[pos]+ins
    qids: np.ndarray         # all n overlaps, final rank order


@dataclass
class Batch:
    bases: np.ndarray        # [B,Lmax,31] u8
    quals: np.ndarray        # [B,Lmax,31] u8
    lens: np.ndarray         # [B] i32
    win_index: np.ndarray    # [B] flat window index
    indices: list = field(default_factory=list)  # list of i32 arrays


class Target:
    """extract_features + InferenceOutput + prepare_examples for one target read."""

    def __init__(self, reads: Reads, rid: int, overlaps: np.ndarray, cigars: list, window_size: int = 4096,
                 batch_size: int = 64):
        self.reads = reads
        o = np.ascontiguousarray(overlaps, dtype=np.uint32).reshape(-1, 9)
        n = o.shape[0]
        coff = np.zeros(n + 1, dtype=np.uint64)
        coff[1:] = np.cumsum([len(c) for c in cigars])
        cg = np.frombuffer(b"".join(cigars) or b"\0", dtype=np.uint8)
        self._h = lib().ho_features(reads._h, rid, n, _p(o, C.c_uint32), _p(cg, C.c_uint8), _p(coff, C.c_uint64),
                                    window_size, batch_size)
        if not self._h:
            raise _err()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ho_target_free(self._h)
            self._h = None

    @property
    def n_windows(self):
        return lib().ho_n_windows(self._h)

    def window(self, w: int) -> Window:
        info = np.zeros(6, dtype=np.uint32)
        if lib().ho_window_info(self._h, w, _p(info, C.c_uint32)) != 0:
            raise IndexError(w)
        L, n_alns, ns, nq, wid, ntw = (int(x) for x in info)
        b = np.zeros((L, 31), dtype=np.uint8)
        q = np.zeros((L, 31), dtype=np.uint8)
        sup = np.zeros((max(ns, 1), 2), dtype=np.uint32)
        rows = np.zeros(max(ns, 1), dtype=np.uint32)
        qids = np.zeros(max(nq, 1), dtype=np.uint32)
        lib().ho_window_get(self._h, w, _p(b, C.c_uint8), _p(q, C.c_uint8), _p(sup, C.c_uint32),
                            _p(rows, C.c_uint32), _p(qids, C.c_uint32))
        return Window(wid, ntw, n_alns, b, q, sup[:ns], rows[:ns], qids[:nq])

    def windows(self):
        return [self.window(w) for w in range(self.n_windows)]

    @property
    def n_batches(self):
        return lib().ho_n_batches(self._h)

    def batch(self, b: int) -> Batch:
        sh = np.zeros(3, dtype=np.uint32)
        if lib().ho_batch_shape(self._h, b, _p(sh, C.c_uint32)) != 0:
            raise IndexError(b)
        B, L, R = (int(x) for x in sh)
        bases = np.zeros((B, L, R), dtype=np.uint8)
        quals = np.zeros((B, L, R), dtype=np.uint8)
        lens = np.zeros(B, dtype=np.int32)
        wi = np.zeros(B, dtype=np.uint32)
        lib().ho_batch_get(self._h, b, _p(bases, C.c_uint8), _p(quals, C.c_uint8), _p(lens, C.c_int32),
                           _p(wi, C.c_uint32), None)
        flat = np.zeros(max(int(lens.sum()), 1), dtype=np.int32)
        lib().ho_batch_get(self._h, b, None, None, None, None, _p(flat, C.c_int32))
        idx, o = [], 0
        for l in lens:
            idx.append(flat[o:o + int(l)].copy())
            o += int(l)
        return Batch(bases, quals, lens, wi, idx)

    def set_logits(self, w: int, info: np.ndarray, bases5: np.ndarray):
        info = np.ascontiguousarray(info, dtype=np.float32)
        bases5 = np.ascontiguousarray(bases5, dtype=np.float32).reshape(-1, 5)
        assert info.shape[0] == bases5.shape[0]
        lib().ho_set_logits(self._h, w, _p(info, C.c_float), _p(bases5, C.c_float), info.shape[0])

    def consensus(self):
        """-> list[bytes] or None (consensus() returned None: the read is omitted)."""
        n = lib().ho_consensus(self._h)
        if n == -2:
            return None
        if n < 0:
            raise _err()
        segs = []
        for s in range(n):
            ln = int(lib().ho_seg_len(self._h, s))
            buf = np.zeros(max(ln, 1), dtype=np.uint8)
            lib().ho_seg_get(self._h, s, _p(buf, C.c_uint8))
            segs.append(buf[:ln].tobytes())
        return segs

    def fasta(self) -> bytes:
        n = lib().ho_fasta(self._h, None, 0)
        buf = np.zeros(max(int(n), 1), dtype=np.uint8)
        lib().ho_fasta(self._h, _p(buf, C.c_uint8), int(n))
        return buf[:int(n)].tobytes()
