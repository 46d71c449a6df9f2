"""No-GPU checks of the drop-in boundary: the shared library loads, exports every symbol that
include/herro_b200.h declares, fails loudly (no CPU fallback) and its host-only utility agrees
with the oracle.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers
from herro_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "herro_b200.h")).read()
    declared = set(re.findall(r"\b(hb_[a-z_]+)\s*\(", hdr))
    declared -= {"hb_ctx"}
    assert declared, "no declarations parsed"
    lib = C.CDLL(api.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in include/herro_b200.h but not exported"
    assert declared == set(api.EXPORTED_SYMBOLS)


def test_struct_layouts_match_header():
    assert api.OVERLAP_DTYPE.itemsize == 56 and api.OVERLAP_DTYPE.fields["cigar"][1] == 40
    assert api.OVERLAP_WINDOW_DTYPE.itemsize == 36
    assert C.sizeof(api.HbOptions) == 20


def _no_cuda():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_cuda(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(api.HerroError) as e:
        api.Context(helpers.model_path(seed=3))
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_null_arguments_are_errors_not_crashes():
    L = api.load_library()
    assert L.hb_create(None, 0, b"x", None) == -1
    assert L.hb_flush(None) == -1
    assert L.hb_get_stats(None, None) == -1
    n = C.c_uint32()
    assert L.hb_extract_windows(None, 1, 4096, 4, None, 0, C.byref(n)) == -1


def test_host_windowing_matches_oracle_and_rejects_bad_cigars():
    from oracle import pyoracle as po
    rs = helpers.small_readset(n_reads=25, mean_len=8000, seed=21)
    checked = 0
    for W in (4096, 256):
        for t in range(rs.n):
            a0, a1 = int(rs.aln_off[t]), int(rs.aln_off[t + 1])
            if a1 == a0:
                continue
            ovl = api.Context.make_overlaps(rs.ovl9[a0:a1], rs.cigars, rs.cig_off[a0:a1 + 1])
            nw = (int(rs.off[t + 1] - rs.off[t]) + W - 1) // W
            got = [tuple(int(x) for x in r) for r in api.extract_windows(ovl, W, nw)]
            want = []
            for k in range(a1 - a0):
                want += [(k,) + w for w in po.extract_windows(rs.ovl9[a0 + k], rs.cigar(a0 + k), W, nw)]
            assert got == want
            checked += len(want)
    assert checked > 1000
    bad = np.frombuffer(b"12M0I30M5X", dtype=np.uint8).copy()
    o = api.Context.make_overlaps(np.array([[0, 9000, 0, 9000, 0, 1, 9000, 0, 9000]], np.uint32), bad, np.array([0, len(bad)], np.uint64))
    with pytest.raises(api.HerroError):
        api.extract_windows(o, 4096, 3)


def test_pack_2bit_matches_reference_known_answers():
    assert api.pack_2bit(np.frombuffer(b"ACGT", np.uint8)).tolist() == [0b11100100]          # src/haec_io.rs:191-197
    assert api.pack_2bit(np.frombuffer(b"ACGTACG", np.uint8)).tolist() == [0b10010011100100]   # src/haec_io.rs:199-203
    from oracle import pyoracle as po
    rng = np.random.default_rng(3)
    s = bytes(rng.choice(list(b"ACGT"), 1000).tolist())
    assert np.array_equal(api.pack_2bit(np.frombuffer(s, np.uint8)), po.encode(s))
    with pytest.raises(ValueError):
        api.pack_2bit(np.frombuffer(b"ACNGT", np.uint8))
