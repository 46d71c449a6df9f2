"""Edge cases of the hot path through the C ABI (GPU): inputs the reference would panic on, targets without
alignments, reads shorter than a window, more overlaps than the 30 the model takes, ABI call-sequence errors."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _ctx(rs, model, **kw):
    from herro_b200 import Context
    ctx = Context(model, 0, kw.pop("window_size", 4096), kw.pop("batch_size", 64), **kw)
    ctx.upload_reads(rs.seqs, rs.quals, rs.off)
    return ctx


def _overlaps(rs, t):
    from herro_b200 import Context
    a0, a1 = int(rs.aln_off[t]), int(rs.aln_off[t + 1])
    return Context.make_overlaps(rs.ovl9[a0:a1], rs.cigars, rs.cig_off[a0:a1 + 1])


def _drain_all(ctx):
    """-> ({rid: segments}, {rid: error code}); a panicking target is reported, the others still arrive"""
    from herro_b200.api import HerroError
    ok, bad = {}, {}
    while True:
        try:
            r = ctx.poll()
        except HerroError as e:
            bad[len(bad)] = e.code
            continue
        if r is None:
            return ok, bad
        ok[r.rid] = r.segments


@pytest.mark.parametrize("host_windowing", [False, True])
def test_malformed_cigar_fails_only_its_target(host_windowing, monkeypatch):
    """A CIGAR that overruns the target is the reference's panic in extract_windows (src/windowing.rs).  hb_submit_alignments
    windows on the device: the target polls HB_ERR_INPUT, every other target is corrected as usual.  With the host windowing
    (HERRO_B200_HOST_WINDOWING, the A-B path) the same input is rejected by the submit call itself."""
    from herro_b200.api import HerroError
    rs = helpers.small_readset(n_reads=20, mean_len=6000, seed=21)
    model = helpers.model_path(seed=3)
    good = helpers.run_product(rs, model, 4096, 64)["segments"]
    if host_windowing:
        monkeypatch.setenv("HERRO_B200_HOST_WINDOWING", "1")
    ctx = _ctx(rs, model, launch_targets=1 << 20)
    victim = next(t for t in range(rs.n) if rs.aln_off[t + 1] - rs.aln_off[t] >= 2)
    bad_cig = np.frombuffer(b"999999M", dtype=np.uint8).copy()
    ovl = _overlaps(rs, victim)
    ovl["cigar"][0] = bad_cig.ctypes.data
    ovl["cigar_len"][0] = len(bad_cig)
    if host_windowing:
        with pytest.raises(HerroError) as ei:
            ctx.submit_alignments(victim, ovl)
        assert ei.value.code == -4  # HB_ERR_INPUT
    else:
        ctx.submit_alignments(victim, ovl)
    for t in range(rs.n):
        if t != victim and rs.aln_off[t + 1] > rs.aln_off[t]:
            ctx.submit_alignments(t, _overlaps(rs, t))
    ctx.flush()
    ok, bad = _drain_all(ctx)
    assert list(bad.values()) == ([] if host_windowing else [-4])
    for t, segs in ok.items():
        assert (segs or None) == good[t]
    assert victim not in ok


@pytest.mark.parametrize("garbage", [b"12M3X4M", b"M", b"12", b"0M", b"5M5", b"99999999999M"])
def test_unparsable_cigars_fail_only_their_target(garbage):
    """CigarIter panics on anything but [0-9]+[MID] with non-zero lengths (src/aligners.rs:252-293): reported per target."""
    rs = helpers.small_readset(n_reads=12, mean_len=6000, seed=27)
    model = helpers.model_path(seed=3)
    from herro_b200 import api
    ctx = _ctx(rs, model, launch_targets=1 << 20)

    def windowed(t):  # alignments of t that contribute to a window: only their CIGARs are ever parsed (src/windowing.rs:53-57)
        ovl = _overlaps(rs, t)
        nw = (int(rs.off[t + 1] - rs.off[t]) + 4095) // 4096
        return [k for k in range(len(ovl)) if (lambda r: r[1] > r[0])(api.window_range(ovl[k:k + 1], 4096, nw))]
    victim = next(t for t in range(rs.n) if len(windowed(t)) >= 2)
    g = np.frombuffer(garbage, dtype=np.uint8).copy()
    ovl = _overlaps(rs, victim)
    k = windowed(victim)[1]
    ovl["cigar"][k] = g.ctypes.data
    ovl["cigar_len"][k] = len(g)
    ctx.submit_alignments(victim, ovl)
    others = [t for t in range(rs.n) if t != victim and rs.aln_off[t + 1] > rs.aln_off[t]]
    for t in others:
        ctx.submit_alignments(t, _overlaps(rs, t))
    ctx.flush()
    ok, bad = _drain_all(ctx)
    assert list(bad.values()) == [-4] and set(ok) == set(others)


def test_inconsistent_overlap_window_fails_only_its_target():
    """hb_submit_target trusts the host's OverlapWindows; a descriptor whose query range lies outside the read is the
    reference's slice-out-of-bounds panic in get_features_for_ol_window -> that target polls HB_ERR_INPUT with no
    record, every other target of the same launch is corrected as usual."""
    from oracle import pyoracle as po
    import herro_b200.api as api
    rs = helpers.small_readset(n_reads=20, mean_len=6000, seed=22)
    model = helpers.model_path(seed=3)
    good = helpers.run_product(rs, model, 4096, 64)["segments"]
    ctx = _ctx(rs, model, launch_targets=1 << 20)
    victim = next(t for t in range(rs.n) if rs.aln_off[t + 1] - rs.aln_off[t] >= 2)
    submitted = []
    for t in range(rs.n):
        a0, a1 = int(rs.aln_off[t]), int(rs.aln_off[t + 1])
        if a1 == a0:
            continue
        nw = (int(rs.off[t + 1] - rs.off[t]) + 4095) // 4096
        ows = []
        for k in range(a1 - a0):
            for (wi, ts, qs, qe, csi, cso, cei, ceo) in po.extract_windows(rs.ovl9[a0 + k], rs.cigar(a0 + k), 4096, nw):
                ows.append((k, wi, ts, qs, qe, csi, cso, cei, ceo))
        ows = np.array(ows, dtype=api.OVERLAP_WINDOW_DTYPE)
        if t == victim:
            ows["qend"][0] = 0x7fffffff  # far past the query read
        ctx.submit_target(t, nw, _overlaps(rs, t), ows)
        submitted.append(t)
    ctx.flush()
    ok, bad = _drain_all(ctx)
    assert list(bad.values()) == [-4]
    assert set(ok) == set(submitted) - {victim}
    for t, segs in ok.items():
        assert (segs or None) == good[t]


def test_target_without_alignments_yields_no_record():
    """consensus() returns None when no window has >= 2 alignments (src/consensus.rs:104-111): the target is
    answered (so the caller can count it) with zero segments."""
    rs = helpers.small_readset(n_reads=12, mean_len=6000, seed=23)
    model = helpers.model_path(seed=3)
    ctx = _ctx(rs, model, launch_targets=4)
    from herro_b200 import Context
    empty = Context.make_overlaps(np.zeros((0, 9), np.uint32), rs.cigars, np.zeros(1, np.uint64))
    ctx.submit_alignments(0, empty)
    ctx.submit_alignments(1, _overlaps(rs, 1)[:1])  # a single alignment: every window has n_alns < 2
    ctx.flush()
    ok, bad = _drain_all(ctx)
    assert not bad and ok == {0: [], 1: []}


def test_reads_shorter_than_one_window_and_exact_multiples():
    """tlen < W (one short window), tlen == k*W (no partial last window): the window count and the last-window
    length follow ceil(len / W) (src/features.rs:343, :476-480)."""
    rs = helpers.small_readset(n_reads=50, mean_len=1500, seed=24, coverage=20.0, min_ovl=400, sd_frac=0.5, min_len=600)
    model = helpers.model_path(seed=4)
    lens = np.diff(rs.off)
    assert (lens < 1024).any() and (lens > 2048).any()
    ora = helpers.run_oracle(rs, model, 1024, 8)
    got = helpers.run_product(rs, model, 1024, 8, keep_debug=True)
    helpers.compare(ora, got, 1e-3)


def test_more_overlaps_than_the_model_takes():
    """coverage 70x: windows have up to ~70 overlapping reads, of which the 30 best by the ln-weighted agreement
    ratio are kept (src/features.rs:494-520) after the stable accuracy sort — exercises ranking ties and the
    recomputed max_ins over the kept columns."""
    rs = helpers.small_readset(n_reads=60, mean_len=5000, seed=25, coverage=70.0, min_ovl=1500)
    model = helpers.model_path(seed=3)
    n_alns = np.diff(rs.aln_off)
    assert n_alns.max() > 45
    ora = helpers.run_oracle(rs, model, 1024, 16)   # W = 1024: interior windows are spanned by ~60 overlaps
    got = helpers.run_product(rs, model, 1024, 16, keep_debug=True)
    helpers.compare(ora, got, 1e-3)
    assert max(w["n_alns"] for w in got["windows"].values()) == 30


def test_call_sequence_errors():
    from herro_b200 import Context
    from herro_b200.api import HerroError
    rs = helpers.small_readset(n_reads=8, mean_len=5000, seed=26)
    model = helpers.model_path(seed=3)
    ctx = Context(model, 0, 4096, 64)
    with pytest.raises(HerroError) as ei:
        ctx.submit_alignments(0, _overlaps(rs, 0))
    assert ei.value.code == -6  # HB_ERR_STATE: no read store yet
    ctx.upload_reads(rs.seqs, rs.quals, rs.off)
    with pytest.raises(HerroError) as ei:
        ctx.submit_alignments(rs.n + 5, _overlaps(rs, 0))
    assert ei.value.code == -1  # HB_ERR_ARG: rid out of range
    wrong = _overlaps(rs, 1) if rs.aln_off[2] > rs.aln_off[1] else None
    if wrong is not None and len(wrong):
        with pytest.raises(HerroError) as ei:
            ctx.submit_alignments(0, wrong)  # alignments of another target
        assert ei.value.code == -1
    ctx.flush()
    assert ctx.drain() == []
