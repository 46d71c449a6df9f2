"""Pins the oracle's 2-bit codec and token tables to the reference's own known answers
(src/haec_io.rs:191-299; tables of src/features.rs:24-42, src/inference.rs:23-31)."""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as po

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "haec_io_known_answers.json")))


@pytest.mark.parametrize("case", G["encode"], ids=lambda c: c["seq"])
def test_encode_known_answers(case):
    assert po.encode(case["seq"].encode()).tolist() == case["words"]


@pytest.mark.parametrize("case", G["decode"], ids=lambda c: c["out"])
def test_decode_known_answers(case):
    w = np.array(case["words"], dtype=np.uint64)
    assert po.decode(w, case["len"], case["start"], case["end"], case["rc"]) == case["out"].encode()


@pytest.mark.parametrize("case", G["subseq"], ids=lambda c: c["ref"])
def test_subseq_known_answers(case):
    seq = case["seq"].encode()
    w = po.encode(seq)
    assert po.decode(w, len(seq), case["start"], case["end"], case["rc"]) == case["out"].encode()


def test_decode_out_of_bounds_panics():
    w = po.encode(b"ACGT")
    with pytest.raises(po.OraclePanic):
        po.decode(w, 4, 0, 5)


def test_roundtrip_random():
    rng = np.random.default_rng(0)
    for n in (1, 31, 32, 33, 64, 1000, 4097):
        s = bytes(rng.choice(list(b"ACGT"), n).tolist())
        w = po.encode(s)
        assert len(w) == (n + 31) // 32
        assert po.decode(w, n) == s
        rc = bytes({65: 84, 67: 71, 71: 67, 84: 65}[c] for c in reversed(s))
        assert po.decode(w, n, 0, n, True) == rc
        a, b = n // 3, n - n // 4
        assert po.decode(w, n, a, b) == s[a:b]
        assert po.decode(w, n, a, b, True) == bytes({65: 84, 67: 71, 71: 67, 84: 65}[c] for c in reversed(s[a:b]))


def test_token_tables():
    L = po.lib()
    # BASES_MAP src/inference.rs:23-31
    for ch, tok in zip(b"ACGT*acgt#.", range(11)):
        assert L.ho_bases_map(ch) == tok
    assert L.ho_bases_map(ord("N")) == 255
    # BASE_LOWER src/features.rs:24-32, BASE_FORWARD :34-42
    for u, l in zip(b"ACGT", b"acgt"):
        assert L.ho_base_lower(u) == l
        assert L.ho_base_forward(u) == u and L.ho_base_forward(l) == u
    assert L.ho_base_forward(ord("#")) == ord("*") and L.ho_base_forward(ord("*")) == ord("*")
    assert L.ho_base_forward(ord(".")) == 255


def test_cigar_iter_ranges():
    # src/aligners.rs:252-293 — byte ranges relative to the slice
    assert po.cigar_iter(b"3M1I5M2D6M") == [("M", 3, 0, 2), ("I", 1, 2, 4), ("M", 5, 4, 6), ("D", 2, 6, 8), ("M", 6, 8, 10)]
    assert po.cigar_iter(b"5000M") == [("M", 5000, 0, 5)]
    assert po.cigar_iter(b"") == []
    for bad in (b"0M", b"3X", b"M", b"12"):
        with pytest.raises(po.OraclePanic):
            po.cigar_iter(bad)
