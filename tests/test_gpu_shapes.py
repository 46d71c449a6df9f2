"""Parity at the shapes the benchmark is quoted on (BASELINE.json configs 2, 4, 5) and of the paths the small
tests never reach: the bit-parallel pileup kernel against the former position-walk kernel, a forced row-arena
overflow, windows with more overlaps than the on-chip sort holds, maximal insertions, and shard invariance
(SURVEY.md §7.3 T2/T3).  Everything goes through the C ABI; the oracle (oracle/) is the checker."""
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

LOGITS_TOL = 1e-3


class _env:
    """Environment toggles are read once, in hb_create: set them around the creation of a context."""

    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update({k: str(v) for k, v in self.kw.items()})

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _targets_with_alns(rs, lo, hi):
    return [t for t in range(lo, hi) if rs.aln_off[t + 1] > rs.aln_off[t]]


def _same_windows(a, b):
    assert set(a["windows"]) == set(b["windows"])
    for key, wa in a["windows"].items():
        wb = b["windows"][key]
        assert wa["L"] == wb["L"] and wa["n_alns"] == wb["n_alns"], key
        for f in ("bases", "quals", "supported", "sup_rows"):
            assert np.array_equal(wa[f], wb[f]), (key, f)


# ------------------------------------------------------------------------------------------ pileup kernel A-B
@pytest.mark.parametrize("profile,W,b", [("r10", 4096, 64), ("r9", 4096, 128), ("r10", 512, 4)])
def test_bit_parallel_pileup_equals_position_walk_kernel(profile, W, b):
    """pileup.cu (consume bitmaps, 4-row groups) against features.cu:k_pass2b on every matrix byte, SupportedPos and segment."""
    rs = helpers.small_readset(n_reads=50, mean_len=9000 if W == 4096 else 5000, seed=31, profile=profile,
                               min_ovl=1024 if W == 4096 else 600)
    model = helpers.model_path(seed=3)
    with _env(HERRO_B200_PILEUP_V1=1):
        old = helpers.run_product(rs, model, W, b, keep_debug=True)
    new = helpers.run_product(rs, model, W, b, keep_debug=True)
    assert old["segments"] == new["segments"]
    _same_windows(old, new)


def test_windows_longer_than_one_pileup_chunk():
    """The pileup kernel handles a window in chunks of 5120 rows (pileup.cu: P_CH).  W = 8192 makes every full
    window span two chunks (carried prefix counts, ranges clipped at the chunk boundary)."""
    rs = helpers.small_readset(n_reads=30, mean_len=20000, seed=32, profile="r9", coverage=20.0, min_ovl=9000)
    model = helpers.model_path(seed=3)
    ora = helpers.run_oracle(rs, model, 8192, 64, with_forward=False)
    got = helpers.run_product(rs, model, 8192, 64, keep_debug=True)
    assert max(w["L"] for w in got["windows"].values()) > 5120 + 3000  # at least two chunks, the second one well filled
    helpers.compare({"windows": ora["windows"], "logits": {}, "segments": {}}, {**got, "segments": {}}, None)


# ------------------------------------------------------------------------------------------ device windowing A-B
@pytest.mark.parametrize("profile,W,b,n,ml,cov", [("r10", 4096, 64, 50, 12000, 25.0), ("r9", 1024, 8, 60, 6000, 15.0), ("r10", 512, 4, 60, 5000, 12.0)])
def test_device_windowing_equals_host_windowing(profile, W, b, n, ml, cov):
    """hb_submit_alignments windows on the device (windowing_dev.cu: CIGAR parse, boundary search, op-slot scan); with
    HERRO_B200_HOST_WINDOWING it runs the C++ restatement of extract_windows on the calling thread.  Same matrices, same
    segments — and both equal the oracle (whose extract_windows is a third implementation)."""
    rs = helpers.small_readset(n_reads=n, mean_len=ml, seed=71, profile=profile, coverage=cov, min_ovl=max(600, W // 2), sd_frac=0.3)
    model = helpers.model_path(seed=3)
    with _env(HERRO_B200_HOST_WINDOWING=1):
        host = helpers.run_product(rs, model, W, b, keep_debug=True)
    dev = helpers.run_product(rs, model, W, b, keep_debug=True)
    assert host["segments"] == dev["segments"]
    _same_windows(host, dev)
    assert host["stats"]["overlap_windows"] == dev["stats"]["overlap_windows"]
    ora = helpers.run_oracle(rs, model, W, b, with_forward=False)
    helpers.compare({"windows": ora["windows"], "logits": {}, "segments": {}}, {**dev, "segments": {}}, None)


# ------------------------------------------------------------------------------------------ benchmark shapes
def test_cfg2_shape_15kb_40x_b64():
    """BASELINE.json configs[1]: 15 kb reads, R10 profile, 40x, W = 4096, -b 64 — 200 target reads against the oracle."""
    rs = helpers.synth.generate(1200, 15000, profile="r10", seed=41, coverage=40.0, min_ovl=2048, targets=(0, 200))
    model = helpers.model_path(seed=3)
    tg = _targets_with_alns(rs, 0, 200)
    assert len(tg) >= 195
    ora = helpers.run_oracle(rs, model, 4096, 64, targets=tg)
    got = helpers.run_product(rs, model, 4096, 64, targets=tg, keep_debug=True)
    helpers.compare(ora, got, LOGITS_TOL)
    n_alns = [w.n_alns for w in ora["windows"].values()]
    assert max(n_alns) == 30 and np.mean(n_alns) > 25  # the shape really is >= 30 overlaps per window


def test_cfg4_shape_100kb_reads_b128():
    """BASELINE.json configs[3]: ultra-long reads (100 kb = 25 windows per read, -b 128 groups all of them into one
    reference batch: k_ref_lmax over 25 windows, long CIGARs, stitch over 25 windows)."""
    rs = helpers.synth.generate(160, 100000, profile="r10", seed=42, coverage=40.0, min_ovl=2048, sd_frac=0.05, targets=(0, 20))
    model = helpers.model_path(seed=3)
    tg = _targets_with_alns(rs, 0, 20)
    assert len(tg) == 20 and int(np.diff(rs.off)[tg].min()) > 80000
    feat = helpers.run_oracle(rs, model, 4096, 128, targets=tg, with_forward=False)
    ora = helpers.run_oracle(rs, model, 4096, 128, targets=tg[:4])  # the torch forward of a 25-window batch is slow on the host
    got = helpers.run_product(rs, model, 4096, 128, targets=tg, keep_debug=True)
    assert max(k[1] for k in got["windows"]) >= 22
    helpers.compare({"windows": feat["windows"], "logits": ora["logits"], "segments": {}}, {**got, "segments": {}}, LOGITS_TOL)
    for t in tg[:4]:
        assert got["segments"][t] == ora["segments"][t]


def test_cfg5_shape_r9_profile_b128():
    """BASELINE.json configs[4]: R9.4.1 error profile (5.5 % errors: ~3x the insertion rows, longer CIGARs), 15 kb, -b 128,
    with a second weights file."""
    rs = helpers.synth.generate(800, 15000, profile="r9", seed=43, coverage=40.0, min_ovl=2048, targets=(0, 100))
    model = helpers.model_path(seed=9)  # "second-model weights path": another blob
    tg = _targets_with_alns(rs, 0, 100)
    ora = helpers.run_oracle(rs, model, 4096, 128, targets=tg)
    got = helpers.run_product(rs, model, 4096, 128, targets=tg, keep_debug=True)
    helpers.compare(ora, got, LOGITS_TOL)


# ------------------------------------------------------------------------------------------ capacity paths
def test_row_arena_overflow_regrows_and_relaunches():
    """ctx.cu sizes the row arena at 1.5 W rows per window and re-launches after an overflow; HERRO_B200_ARENA_ROWS shrinks the
    initial arena so that the path runs (it never does at R10/R9 error rates)."""
    rs = helpers.small_readset(n_reads=30, mean_len=7000, seed=51)
    model = helpers.model_path(seed=3)
    ref = helpers.run_product(rs, model, 4096, 64, keep_debug=True)
    with _env(HERRO_B200_ARENA_ROWS=64):
        got = helpers.run_product(rs, model, 4096, 64, keep_debug=True)                 # one launch, overflows once
        many = helpers.run_product(rs, model, 4096, 64, launch_targets=7)              # several launches, growing arena
    assert got["segments"] == ref["segments"] == many["segments"]
    _same_windows(ref, got)
    assert got["stats"]["kernel_launches"] > ref["stats"]["kernel_launches"]  # the feature kernels ran twice


def test_more_than_1024_overlaps_in_a_window():
    """The reference ranks any number of overlaps (src/features.rs:376-418, 502-525).  Beyond 1024 per window the sort keys
    live in HBM instead of shared memory (features.cu: big_key / big_cand / big_score)."""
    from herro_b200 import Context, api
    rs = helpers.synth.generate(2200, 5000, profile="r10", seed=52, coverage=1000.0, min_ovl=1100, sd_frac=0.05, targets=(0, 3))
    model = helpers.model_path(seed=3)
    tg = _targets_with_alns(rs, 0, 3)
    per_window = 0
    for t in tg:  # overlap-windows per window, from the host windowing
        a0, a1 = int(rs.aln_off[t]), int(rs.aln_off[t + 1])
        nw = (int(rs.off[t + 1] - rs.off[t]) + 1023) // 1024
        ows = api.extract_windows(Context.make_overlaps(rs.ovl9[a0:a1], rs.cigars, rs.cig_off[a0:a1 + 1]), 1024, nw)
        per_window = max(per_window, int(np.bincount(ows["window_idx"]).max()))
    assert per_window > 1024
    ora = helpers.run_oracle(rs, model, 1024, 16, targets=tg)
    got = helpers.run_product(rs, model, 1024, 16, targets=tg, keep_debug=True)
    helpers.compare(ora, got, LOGITS_TOL)


def _edit_cigar_add_indel_pair(cig: bytes, n: int) -> bytes:
    """Replace the first M op of length x >= 2n + 30 by 10M nI 10M nD (x - 20 - n)M: same target and query span, any bases."""
    import re
    ops = re.findall(rb"(\d+)([MID])", cig)
    out, done = [], False
    for ln, op in ops:
        x = int(ln)
        if not done and op == b"M" and x >= 2 * n + 30:
            out += [b"10M", b"%dI" % n, b"10M", b"%dD" % n, b"%dM" % (x - 20 - n)]
            done = True
        else:
            out.append(ln + op)
    return b"".join(out) if done else None


@pytest.mark.parametrize("n,kept", [(50, True), (51, False)])
def test_longest_insertion_the_filter_lets_through(n, kept):
    """overlap_window_filter drops an overlap-window with any I/D longer than 50 (src/features.rs:315-324), so 50 is the
    longest insertion run a pileup can hold: SupportedPos.ins (u8, H13) never wraps.  An alignment edited to carry a 50I/50D
    pair stays in (50 insertion rows after one position); with 51 its window drops it."""
    rs = helpers.small_readset(n_reads=30, mean_len=7000, seed=53)
    model = helpers.model_path(seed=3)
    t = next(t for t in range(rs.n) if rs.aln_off[t + 1] - rs.aln_off[t] >= 6)
    a0, a1 = int(rs.aln_off[t]), int(rs.aln_off[t + 1])
    cigs = [rs.cigar(a) for a in range(a0, a1)]
    edited = 0
    for k in range(len(cigs)):  # 50: every alignment that has a long enough M op carries one; 51: two do
        e = _edit_cigar_add_indel_pair(cigs[k], n)
        if e is not None and (kept or edited < 2):
            cigs[k] = e
            edited += 1
    assert edited >= 2
    # a read set whose target t carries the edited alignments (other targets untouched)
    newc = np.frombuffer(b"".join(cigs), dtype=np.uint8)
    cig_off = rs.cig_off.copy()
    lens = np.array([len(c) for c in cigs], dtype=np.uint64)
    delta = int(lens.sum()) - int(rs.cig_off[a1] - rs.cig_off[a0])
    cig_off[a0 + 1:a1 + 1] = rs.cig_off[a0] + np.cumsum(lens)
    cig_off[a1 + 1:] = (rs.cig_off[a1 + 1:].astype(np.int64) + delta).astype(np.uint64)
    rs.cigars = np.concatenate([rs.cigars[:int(rs.cig_off[a0])], newc, rs.cigars[int(rs.cig_off[a1]):]])
    rs.cig_off = cig_off
    ora = helpers.run_oracle(rs, model, 4096, 64, targets=[t])
    got = helpers.run_product(rs, model, 4096, 64, targets=[t], keep_debug=True)
    helpers.compare(ora, got, LOGITS_TOL)
    longest = 0
    for (rid, wid), w in ora["windows"].items():
        run = 0
        for tok in w.bases[:, 0]:
            run = run + 1 if tok == 4 else 0  # target '*' = insertion row
            longest = max(longest, run)
    assert longest == 50 if kept else longest < 50
    assert max(int(w["supported"][:, 1].max()) if len(w["supported"]) else 0 for w in got["windows"].values()) <= 50


# ------------------------------------------------------------------------------------------ shard invariance
@pytest.mark.parametrize("world", [2, 4])
def test_shard_invariance_of_the_corrected_set(world):
    """north_star: target reads shard across GPUs by read id with a replicated read store and no collective.  The union of
    the per-shard outputs (one context per shard, each with its own store replica) is the 1-context output, record for
    record (FASTA order is unspecified in the reference, F8)."""
    from herro_b200 import shard
    rs = helpers.small_readset(n_reads=60, mean_len=8000, seed=61)
    model = helpers.model_path(seed=3)
    whole = helpers.run_product(rs, model, 4096, 64)["segments"]
    lens = np.diff(rs.off).astype(np.int64)
    union = {}
    sizes = []
    for r in range(world):
        mine = [int(t) for t in shard.shard_targets(lens, 4096, r, world)]
        part = helpers.run_product(rs, model, 4096, 64, targets=mine)["segments"]
        assert not (set(part) & set(union))
        union.update(part)
        sizes.append(len(mine))
    assert union == whole
    assert sum(sizes) == rs.n


# ------------------------------------------------------------------------------------------ `herro features` dump
def test_feature_dump_matches_golden_files(tmp_path):
    """hb_dump_features writes the reference's `herro features` files (src/features.rs:724-764): [2,L',31] ASCII pileup +
    qualities, SupportedPos records, ranked query ids.  tests/golden/features_dump/ holds the same files produced by the CPU
    oracle through numpy (tools/make_feature_fixture.py): byte-identical, headers included."""
    import glob
    sys_path_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_feature_fixture", os.path.join(sys_path_root, "tools", "make_feature_fixture.py"))
    fx = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fx)
    rs = fx.readset()
    model = helpers.model_path(seed=3)
    got = helpers.run_product(rs, model, fx.W, 4, targets=list(fx.TARGETS), keep_debug=True, dump=False)
    ctx = got["ctx"]
    for t in fx.TARGETS:
        ctx.dump_features(t, str(tmp_path), rs.ids)
    golden = os.path.join(sys_path_root, "tests", "golden", "features_dump")
    n = 0
    for t in fx.TARGETS:
        want = sorted(glob.glob(os.path.join(golden, rs.ids[t], "*")))
        have = sorted(glob.glob(os.path.join(str(tmp_path), rs.ids[t], "*")))
        assert [os.path.basename(p) for p in want] == [os.path.basename(p) for p in have] and want
        for a, b in zip(want, have):
            assert open(a, "rb").read() == open(b, "rb").read(), os.path.basename(a)
            n += 1
    assert n >= 30
    f = np.load(os.path.join(str(tmp_path), rs.ids[0], "0.features.npy"))
    assert f.dtype == np.uint8 and f.shape[0] == 2 and f.shape[2] == 31 and set(np.unique(f[0])) <= set(b"ACGTacgt*#.")
