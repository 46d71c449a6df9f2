"""extract_windows (src/windowing.rs:44-273) — hand-derived cases of SURVEY.md App. E plus
structural invariants on generated CIGARs.  Not reference-pinned (see golden file header)."""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as po

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "windowing_app_e.json")))


@pytest.mark.parametrize("case", G["cases"], ids=lambda c: c["name"])
def test_app_e_cases(case):
    nw = (case["tlen"] + case["W"] - 1) // case["W"]
    got = po.extract_windows(case["ovl"], case["cigar"].encode(), case["W"], nw)
    assert [list(g) for g in got] == case["windows"]


def _random_cigar(rng, tspan_target):
    ops, t, q = [], 0, 0
    last = None
    while t < tspan_target:
        kind = "M" if last != "M" else rng.choice(["I", "D"])
        n = int(rng.integers(1, 40)) if kind == "M" else int(rng.integers(1, 4))
        if kind != "I" and t + n > tspan_target:
            n = tspan_target - t
        ops.append((kind, n))
        if kind != "I":
            t += n
        if kind != "D":
            q += n
        last = kind
    if ops[-1][0] != "M":
        ops.append(("M", 3)); t += 3; q += 3
    return "".join(f"{n}{k}" for k, n in ops).encode(), t, q


@pytest.mark.parametrize("W", [8, 16, 64])
def test_window_invariants(W):
    rng = np.random.default_rng(W)
    for _ in range(200):
        tlen = int(rng.integers(3 * W, 12 * W))
        ts = int(rng.integers(0, tlen - 2 * W))
        cigar, tspan, qspan = _random_cigar(rng, int(rng.integers(2 * W, tlen - ts + 1)) - 3)
        te = ts + tspan
        if te > tlen:
            continue
        ovl = [0, qspan + 5, 2, 2 + qspan, 0, 1, tlen, ts, te]
        nw = (tlen + W - 1) // W
        try:
            wins = po.extract_windows(ovl, cigar, W, nw)
        except po.OraclePanic:
            continue
        items = po.cigar_iter(cigar)
        for (wi, wts, qs, qe, csi, cso, cei, ceo) in wins:
            sl = po.cigar_iter(cigar[csi:cei])
            tl = ql = 0
            for k, (kind, n, s, e) in enumerate(sl):
                if len(sl) == 1:
                    eff = ceo - cso
                elif k == 0:
                    eff = n - cso
                elif k == len(sl) - 1:
                    eff = ceo
                else:
                    eff = n
                assert eff > 0
                if kind != "I":
                    tl += eff
                if kind != "D":
                    ql += eff
            # every overlap-window covers target [wts, end of window or overlap end)
            wend = min((wi + 1) * W, te)
            assert wts + tl == wend, (cigar, wi)
            assert qe - qs == ql
            assert wi * W <= wts < (wi + 1) * W


def test_window_range_matches_extract_windows():
    """hb_window_range (the coordinate-only skeleton the device windowing is laid out from, ctx.cu) names exactly the
    windows extract_windows emits, for every alignment of synthetic sets at several window sizes."""
    from herro_b200 import Context, api
    from tools import synth
    n_checked = 0
    for W, seed in ((512, 1), (1024, 2), (4096, 3)):
        rs = synth.generate(60, 6000 if W < 4096 else 12000, seed=seed, coverage=15.0, min_ovl=max(600, W // 2), sd_frac=0.3)
        for t in range(rs.n):
            a0, a1 = int(rs.aln_off[t]), int(rs.aln_off[t + 1])
            if a1 == a0:
                continue
            nw = (int(rs.off[t + 1] - rs.off[t]) + W - 1) // W
            ovl = Context.make_overlaps(rs.ovl9[a0:a1], rs.cigars, rs.cig_off[a0:a1 + 1])
            for k in range(a1 - a0):
                ows = api.extract_windows(ovl[k:k + 1], W, nw)
                got = sorted(int(x) for x in ows["window_idx"])
                first, end = api.window_range(ovl[k:k + 1], W, nw)
                assert got == list(range(first, end)), (W, t, k, got, first, end)
                n_checked += 1
    assert n_checked > 1000
