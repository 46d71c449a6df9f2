"""Shared test plumbing: run the CPU oracle (+ torch fp32 forward) and the CUDA product on the
same read set and compare.  oracle/ is imported here and only here-abouts (tests/, smoke, bench)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tools import synth  # noqa: E402
from herro_b200 import weights as hbw  # noqa: E402

TMP = os.path.join(ROOT, "tests", "_tmp")


def model_path(seed=3, cfg=None):
    os.makedirs(TMP, exist_ok=True)
    cfg = cfg or hbw.NetConfig()
    tag = "_".join(str(v) for v in hbw.config_dict(cfg).values())
    p = os.path.join(TMP, f"model_s{seed}_{tag}.hbw")
    if not os.path.exists(p):
        hbw.save_blob(p, cfg, hbw.random_weights(cfg, seed))
    return p


def small_readset(n_reads=40, mean_len=9000, seed=1, profile="r10", coverage=25.0, **kw):
    kw.setdefault("min_ovl", 1024)
    return synth.generate(n_reads, mean_len, profile=profile, seed=seed, coverage=coverage, **kw)


def run_oracle(rs, model, window_size=4096, batch_size=64, targets=None, with_forward=True):
    """-> dict(windows={(rid,wid): Window}, logits={(rid,wid): (info, bases)}, segments={rid: [bytes]|None})"""
    from oracle import pyoracle as po
    reads = po.Reads(rs.ids, [rs.seq(i) for i in range(rs.n)], [rs.qual(i) for i in range(rs.n)])
    net = None
    if with_forward:
        from oracle import forward_ref
        cfg, tensors = hbw.load_blob(model)
        net = forward_ref.from_weights(cfg, tensors)
    out = dict(windows={}, logits={}, segments={}, targets={}, win_index={})
    targets = range(rs.n) if targets is None else targets
    for t in targets:
        ovl, cigs = rs.target_alns(t)
        if len(ovl) == 0:
            continue  # reads that never appear as a PAF target produce no record (H6)
        T = po.Target(reads, t, ovl, cigs, window_size, batch_size)
        wins = T.windows()
        out["targets"][t] = T
        for i, w in enumerate(wins):
            out["windows"][(t, w.wid)] = w
            out["win_index"][(t, w.wid)] = i
        if with_forward:
            for b in range(T.n_batches):
                B = T.batch(b)
                info, bl = forward_ref.run_batch(net, B.bases, B.quals, B.lens, B.indices)
                for k, wi in enumerate(B.win_index):
                    T.set_logits(int(wi), info[k], bl[k])
                    out["logits"][(t, wins[int(wi)].wid)] = (info[k], bl[k])
            out["segments"][t] = T.consensus()
    return out


def run_product(rs, model, window_size=4096, batch_size=64, targets=None, keep_debug=False, launch_targets=0,
                use_submit_target=False, dump=True):
    from herro_b200 import Context
    ctx = Context(model, 0, window_size, batch_size, launch_targets=launch_targets or 1 << 20, keep_debug=keep_debug)
    ctx.upload_reads(rs.seqs, rs.quals, rs.off)
    targets = list(range(rs.n) if targets is None else targets)
    submitted = []
    for t in targets:
        a0, a1 = int(rs.aln_off[t]), int(rs.aln_off[t + 1])
        if a1 == a0:
            continue
        ovl = Context.make_overlaps(rs.ovl9[a0:a1], rs.cigars, rs.cig_off[a0:a1 + 1])
        if use_submit_target:
            from oracle import pyoracle as po
            nw = (int(rs.off[t + 1] - rs.off[t]) + window_size - 1) // window_size
            ows = []
            for k in range(a1 - a0):
                for (wi, ts, qs, qe, csi, cso, cei, ceo) in po.extract_windows(rs.ovl9[a0 + k], rs.cigar(a0 + k), window_size, nw):
                    ows.append((k, wi, ts, qs, qe, csi, cso, cei, ceo))
            import herro_b200.api as api
            ctx.submit_target(t, nw, ovl, np.array(ows, dtype=api.OVERLAP_WINDOW_DTYPE))
        else:
            ctx.submit_alignments(t, ovl)
        submitted.append(t)
    ctx.flush()
    out = dict(windows={}, logits={}, segments={}, stats=ctx.stats())
    for r in ctx.drain():
        out["segments"][r.rid] = r.segments if r.segments else None
    if keep_debug and dump:
        for t in submitted:
            nw = (int(rs.off[t + 1] - rs.off[t]) + window_size - 1) // window_size
            for w in range(nw):
                out["windows"][(t, w)] = ctx.debug_window(t, w)
    out["ctx"] = ctx
    return out


def compare(ora, got, logits_tol=1e-3, check_windows=True):
    """Asserts parity; returns dict(reads, windows, tie_reads, max_logit_diff)."""
    # segments: byte-identical per read; None (read omitted) on both sides
    assert set(ora["segments"].keys()) == set(got["segments"].keys())
    max_diff = 0.0
    if check_windows and got["windows"]:
        for key, w in ora["windows"].items():
            g = got["windows"][key]
            assert g["L"] == w.bases.shape[0], (key, g["L"], w.bases.shape)
            assert g["n_alns"] == w.n_alns, key
            assert np.array_equal(g["bases"], w.bases), key
            assert np.array_equal(g["quals"], w.quals), key
            assert np.array_equal(g["supported"], w.supported.reshape(-1, 2)), key
            assert np.array_equal(g["sup_rows"], w.sup_rows), key
            if key in ora["logits"] and logits_tol is not None:
                info, bl = ora["logits"][key]
                assert np.allclose(g["info_logits"], info, atol=logits_tol, rtol=0), (key, np.abs(g["info_logits"] - info).max())
                assert np.allclose(g["bases_logits"], bl, atol=logits_tol, rtol=0), (key, np.abs(g["bases_logits"] - bl).max())
                if len(bl):
                    max_diff = max(max_diff, float(np.abs(g["bases_logits"] - bl).max()), float(np.abs(g["info_logits"] - info).max()))
    n_tie_reads = 0
    for rid, segs in ora["segments"].items():
        if got["segments"][rid] == segs:
            continue
        # The corrected bases are argmaxes of logits that agree only to `logits_tol`: where the oracle's two largest logits
        # of a position are closer than that, the two sides may legitimately pick different bases.  Accept such a read iff
        # (a) it has such a near-tie and (b) the oracle's own consensus() fed with the product's logits reproduces the
        # product's segments byte for byte (so everything but the tie-break is identical).
        assert logits_tol is not None and got["windows"] and rid in ora.get("targets", {}), f"segments differ for read {rid}"
        T = ora["targets"][rid]
        gap = np.inf
        for (t, wid), i in ora["win_index"].items():
            if t != rid:
                continue
            if (t, wid) in ora["logits"]:
                bl = np.sort(ora["logits"][(t, wid)][1], axis=1)
                if len(bl):
                    gap = min(gap, float((bl[:, -1] - bl[:, -2]).min()))
                g = got["windows"][(t, wid)]
                T.set_logits(i, g["info_logits"], g["bases_logits"])
        assert gap < 2 * logits_tol, f"segments differ for read {rid} without a near-tie in the logits (min top-2 gap {gap})"
        assert T.consensus() == got["segments"][rid], f"segments differ for read {rid} beyond argmax ties"
        n_tie_reads += 1
    assert n_tie_reads <= max(1, len(ora["segments"]) // 20), f"{n_tie_reads} reads differ through logit near-ties: too many to be ties"
    return dict(reads=len(ora["segments"]), windows=len(ora["windows"]), tie_reads=n_tie_reads, max_logit_diff=max_diff)
