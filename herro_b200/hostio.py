"""ctypes view of the native host data plane (herro_b200/host/io.cpp): FASTQ -> packed read store, `--read-alns` batches ->
alignments grouped by target, FASTA writer, and the whole `herro inference` pipeline (hbh_inference).  In the deployed layout
the Rust host owns these stages (haec_io.rs, overlaps.rs, lib.rs); here they are C++ threads over the public C ABI."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import api


def _lib():
    H = api._host_lib()
    if getattr(H, "_io_ready", False):
        return H
    vp, u32 = C.c_void_p, C.c_uint32
    H.hbh_last_error.restype = C.c_char_p
    H.hbh_reads_load.argtypes = [C.c_char_p, u32, vp, u32, vp, u32, C.c_int, C.POINTER(vp)]
    H.hbh_reads_free.argtypes = [vp]
    H.hbh_reads_free.restype = None
    H.hbh_reads_count.argtypes = [vp]
    H.hbh_reads_count.restype = u32
    for f in ("hbh_reads_lens", "hbh_reads_word_ptrs", "hbh_reads_qual_ptrs", "hbh_reads_names"):
        getattr(H, f).argtypes = [vp]
        getattr(H, f).restype = vp
    H.hbh_reads_description.argtypes = [vp, u32]
    H.hbh_reads_description.restype = C.c_char_p
    H.hbh_reads_stats.argtypes = [vp, vp]
    H.hbh_reads_stats.restype = None
    H.hbh_alns_load.argtypes = [C.c_char_p, vp, vp, u32, C.c_int, C.POINTER(vp)]
    H.hbh_alns_free.argtypes = [vp]
    H.hbh_alns_free.restype = None
    H.hbh_alns_targets.argtypes = [vp]
    H.hbh_alns_targets.restype = u32
    for f in ("hbh_alns_target_rids", "hbh_alns_target_offsets", "hbh_alns_overlaps"):
        getattr(H, f).argtypes = [vp]
        getattr(H, f).restype = vp
    H.hbh_alns_stats.argtypes = [vp, vp]
    H.hbh_alns_stats.restype = None
    H.hbh_fasta_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    H.hbh_fasta_write.argtypes = [vp, C.c_char_p, C.c_char_p, vp, vp, u32]
    H.hbh_fasta_close.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    H.hbh_inference.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, u32, u32, C.c_int, vp, C.c_int, vp, u32, vp, u32,
                                C.c_int, vp, vp]
    H._io_ready = True
    return H


def _strs(names):
    if names is None:
        return None, 0, None
    buf = [n if isinstance(n, bytes) else str(n).encode() for n in names]
    arr = (C.c_char_p * len(buf))(*buf)
    return arr, len(buf), buf


class Reads:
    """haec_io::get_reads (src/haec_io.rs:37-75): FASTQ file / directory -> ids, descriptions, 2-bit words, qualities."""

    def __init__(self, path: str, min_len: int = 4096, core=None, neighbour=None, threads: int = 0):
        H = _lib()
        self._H, self._h = H, C.c_void_p()
        ca, nc, self._k1 = _strs(core)
        na, nn, self._k2 = _strs(neighbour)
        rc = H.hbh_reads_load(path.encode(), min_len, ca, nc, na, nn, threads or min(os.cpu_count() or 1, 32), C.byref(self._h))
        if rc != 0:
            raise api.HerroError(rc, H.hbh_last_error().decode())
        self.n = H.hbh_reads_count(self._h)
        n = self.n
        self.lens = np.ctypeslib.as_array(C.cast(H.hbh_reads_lens(self._h), C.POINTER(C.c_uint32)), (n,)) if n else np.zeros(0, np.uint32)
        names = C.cast(H.hbh_reads_names(self._h), C.POINTER(C.c_char_p))
        self.ids = [names[i] for i in range(n)]
        self.descriptions = [H.hbh_reads_description(self._h, i) for i in range(n)]

    def words(self, i) -> np.ndarray:
        p = C.cast(self._H.hbh_reads_word_ptrs(self._h), C.POINTER(C.c_void_p))[i]
        nw = (int(self.lens[i]) + 31) // 32
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), (max(nw, 1),))[:nw].copy()

    def qual(self, i) -> bytes:
        p = C.cast(self._H.hbh_reads_qual_ptrs(self._h), C.POINTER(C.c_void_p))[i]
        return C.string_at(p, int(self.lens[i]))

    def stats(self) -> dict:
        s = np.zeros(4)
        self._H.hbh_reads_stats(self._h, s.ctypes.data)
        return dict(load_s=s[0], pack_s=s[1], skipped_short=int(s[2]), bases=int(s[3]))

    def upload(self, ctx: "api.Context"):
        """hb_upload_reads straight from the packed store (no copy through Python)."""
        H = self._H
        ctx._check(ctx._L.hb_upload_reads(ctx._h, self.n, H.hbh_reads_word_ptrs(self._h), H.hbh_reads_lens(self._h),
                                          H.hbh_reads_qual_ptrs(self._h)))
        ctx.read_len = self.lens.copy()

    def close(self):
        if self._h:
            self._H.hbh_reads_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Alignments:
    """overlaps::read_batches + parse_paf (src/overlaps.rs:288-323,117-202), one worker per batch file."""

    def __init__(self, alns_dir: str, reads: Reads, core=None, threads: int = 0):
        H = _lib()
        self._H, self._h, self.reads = H, C.c_void_p(), reads
        ca, nc, self._k = _strs(core)
        rc = H.hbh_alns_load(alns_dir.encode(), reads._h, ca, nc, threads or min(os.cpu_count() or 1, 32), C.byref(self._h))
        if rc != 0:
            raise api.HerroError(rc, H.hbh_last_error().decode())
        nt = H.hbh_alns_targets(self._h)
        self.n_targets = nt
        self.target_rids = (np.ctypeslib.as_array(C.cast(H.hbh_alns_target_rids(self._h), C.POINTER(C.c_uint32)), (nt,))
                            if nt else np.zeros(0, np.uint32))
        self.offsets = np.ctypeslib.as_array(C.cast(H.hbh_alns_target_offsets(self._h), C.POINTER(C.c_uint64)), (nt + 1,))
        na = int(self.offsets[-1])
        buf = (C.c_char * (max(na, 1) * api.OVERLAP_DTYPE.itemsize)).from_address(H.hbh_alns_overlaps(self._h) or 0) if na else None
        self.overlaps = np.frombuffer(buf, dtype=api.OVERLAP_DTYPE, count=na) if na else np.zeros(0, api.OVERLAP_DTYPE)

    def target(self, k):
        """-> (rid, hb_overlap[] view) of the k-th target group."""
        return int(self.target_rids[k]), self.overlaps[int(self.offsets[k]):int(self.offsets[k + 1])]

    def cigar(self, a) -> bytes:
        o = self.overlaps[a]
        return C.string_at(int(o["cigar"]), int(o["cigar_len"]))

    def stats(self) -> dict:
        s = np.zeros(6)
        self._H.hbh_alns_stats(self._h, s.ctypes.data)
        return dict(decode_s_sum=s[0], parse_s_sum=s[1], lines=int(s[2]), kept=int(s[3]), compressed_bytes=int(s[4]), text_bytes=int(s[5]))

    def close(self):
        if self._h:
            self._H.hbh_alns_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FastaWriter:
    """correction_writer / write_sequence (src/lib.rs:267-317)."""

    def __init__(self, path: str):
        H = _lib()
        self._H, self._h = H, C.c_void_p()
        if H.hbh_fasta_open(path.encode(), C.byref(self._h)) != 0:
            raise api.HerroError(-1, H.hbh_last_error().decode())

    def write(self, read_id: bytes, description, segments: list):
        seq = b"".join(segments)
        lens = np.array([len(s) for s in segments], dtype=np.uint32)
        buf = np.frombuffer(seq, dtype=np.uint8) if seq else np.zeros(1, np.uint8)
        self._H.hbh_fasta_write(self._h, read_id, description, buf.ctypes.data, lens.ctypes.data, len(segments))

    def close(self):
        rec, bases = C.c_uint64(), C.c_uint64()
        if self._h:
            self._H.hbh_fasta_close(self._h, C.byref(rec), C.byref(bases))
            self._h = C.c_void_p()
        return rec.value, bases.value


def inference(reads_path: str, alns_dir: str, model: str, output: str, window: int = 4096, batch: int = 64, threads: int = 1,
              devices=(0,), core=None, neighbour=None, io_threads: int = 0) -> dict:
    """The reference's `herro inference --read-alns` command, natively (hbh_inference) -> stage times and counts."""
    H = _lib()
    dev = np.asarray(list(devices), dtype=np.int32)
    ca, nc, k1 = _strs(core)
    na, nn, k2 = _strs(neighbour)
    t8, c4 = np.zeros(8), np.zeros(4, np.uint64)
    rc = H.hbh_inference(reads_path.encode(), alns_dir.encode(), model.encode(), output.encode(), window, batch, threads,
                         dev.ctypes.data, len(dev), ca, nc, na, nn, io_threads or min(os.cpu_count() or 1, 32), t8.ctypes.data, c4.ctypes.data)
    if rc != 0:
        raise api.HerroError(rc, H.hbh_last_error().decode())
    return dict(fastq_load_s=t8[0], pack_s=t8[1], alignment_ingest_s=t8[2], read_store_upload_s=t8[3], correction_s=t8[4],
                fasta_close_s=t8[5], total_s=t8[6], corrected_bases=int(t8[7]), reads=int(c4[0]), targets=int(c4[1]),
                records=int(c4[2]), failed_targets=int(c4[3]))
