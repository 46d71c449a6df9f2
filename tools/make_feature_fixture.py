"""Writes tests/golden/features_dump/: the `herro features` files (src/features.rs:724-764) of a tiny synthetic read set,
produced by the CPU ORACLE (oracle/) with numpy's own .npy writer.  The GPU test test_feature_dump_matches_golden_files
compares hb_dump_features' files with these byte for byte; anyone with the reference binary can diff them against
`herro features -w 256 --read-alns <alns> <reads.fastq> <out>` on the inputs this script also writes (reads.fastq,
alns/0.oec.zst), after np.load (the header padding of the npyz crate may differ from numpy's).

    python tools/make_feature_fixture.py          # regenerates the fixture (deterministic)
"""
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "features_dump")
W = 256
ASCII = np.frombuffer(b"ACGT*acgt#..", dtype=np.uint8)
SUP_DTYPE = np.dtype([("pos", "<u2"), ("ins", "u1")])
PARAMS = dict(n_reads=24, mean_len=900, seed=77, coverage=10.0, min_ovl=300, sd_frac=0.15, min_len=600)
TARGETS = (0, 1, 2)


def readset():
    return synth.generate(PARAMS["n_reads"], PARAMS["mean_len"], seed=PARAMS["seed"], coverage=PARAMS["coverage"],
                          min_ovl=PARAMS["min_ovl"], sd_frac=PARAMS["sd_frac"], min_len=PARAMS["min_len"])


def write_window(dirpath, w, ids):
    os.makedirs(dirpath, exist_ok=True)
    feats = np.stack([ASCII[w.bases], w.quals]).astype(np.uint8)            # [2, L', 31]
    np.save(os.path.join(dirpath, f"{w.wid}.features.npy"), feats)
    sup = np.zeros(len(w.supported), dtype=SUP_DTYPE)
    sup["pos"] = w.supported.reshape(-1, 2)[:, 0]
    sup["ins"] = w.supported.reshape(-1, 2)[:, 1]
    np.save(os.path.join(dirpath, f"{w.wid}.supported.npy"), sup)
    with open(os.path.join(dirpath, f"{w.wid}.ids.txt"), "wb") as f:
        for q in w.qids:
            f.write(ids[int(q)].encode() + b"\n")


def main():
    from oracle import pyoracle as po
    rs = readset()
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    reads = po.Reads(rs.ids, [rs.seq(i) for i in range(rs.n)], [rs.qual(i) for i in range(rs.n)])
    n = 0
    for t in TARGETS:
        ovl, cigs = rs.target_alns(t)
        if not len(ovl):
            continue
        T = po.Target(reads, t, ovl, cigs, W, 4)
        for w in T.windows():
            write_window(os.path.join(OUT, rs.ids[t]), w, rs.ids)
            n += 1
    synth.write_fastq(rs, os.path.join(OUT, "reads.fastq"))
    synth.write_oec_batches(rs, os.path.join(OUT, "alns"))
    print(f"wrote {n} windows of {len(TARGETS)} targets under {OUT}")


if __name__ == "__main__":
    main()
