"""`python -m herro_b200.cli inference ...` — the reference's `herro inference` command line
(src/main.rs:64-112, README.md:75-96) on top of the C ABI, for plumbing runs (BASELINE.json configs[0]):

    python -m herro_b200.cli inference --read-alns <dir> -t 4 -d 0 -m model.hbw -b 64 reads.fastq out.fasta

Host data plane in Python (FASTQ, `*.oec.zst` batches via pyarrow's zstd, FASTA writer): in deployment
that part is the unchanged Rust binary.  Semantics kept from the reference: reads shorter
than `-w` are not loaded (src/haec_io.rs:48), unknown names / self overlaps / repeated (query,target) pairs
are skipped (src/overlaps.rs:137-185), `-c` cluster files restrict targets to core reads (:154-159).
"""
from __future__ import annotations

import argparse
import glob
import gzip
import os
import sys

import numpy as np

from . import api


def read_fastq(path, min_len):
    """-> ids, descriptions, seqs(list[bytes]), quals(list[bytes]) — get_reads (src/haec_io.rs:37-75)."""
    op = gzip.open if path.endswith(".gz") else open
    ids, descs, seqs, quals = [], [], [], []
    with op(path, "rb") as f:
        while True:
            h = f.readline()
            if not h:
                break
            s = f.readline().rstrip(b"\r\n")
            f.readline()
            q = f.readline().rstrip(b"\r\n")
            if len(s) < min_len:
                continue
            hdr = h[1:].rstrip(b"\r\n")
            cut = min([i for i in (hdr.find(b" "), hdr.find(b"\t")) if i >= 0], default=-1)
            ids.append(hdr if cut < 0 else hdr[:cut])
            descs.append(None if cut < 0 else hdr[cut + 1:])
            seqs.append(s)
            quals.append(q)
    return ids, descs, seqs, quals


def read_cluster(path):
    core, neigh = set(), set()
    for line in open(path, "rb"):
        f = line.rstrip(b"\n").split(b"\t")
        (core if f[0] == b"0" else neigh).add(f[1])
    return core, neigh


def read_oec_batches(alns_dir, name_to_id, core=None):
    """parse_paf over every `*.oec.zst` (src/overlaps.rs:117-202,288-323) -> {tid: [(ovl9, cigar)]}"""
    import pyarrow as pa
    codec = pa.Codec("zstd")
    out = {}
    for p in sorted(glob.glob(os.path.join(alns_dir, "*.oec.zst"))):
        raw = open(p, "rb").read()
        # single-frame streams written without a content size are handled by the streaming reader
        try:
            data = pa.CompressedInputStream(pa.BufferReader(raw), "zstd").read()
        except Exception:
            data = codec.decompress(raw, asbytes=True)
        lines = data.split(b"\n")
        n_targets = int(lines[0])
        seen = set()
        for line in lines[1 + n_targets:]:
            if not line:
                continue
            f = line.split(b"\t")
            qid = name_to_id.get(f[0])
            if qid is None:
                continue
            if core is not None and f[5] not in core:
                continue
            tid = name_to_id.get(f[5])
            if tid is None or tid == qid or (qid, tid) in seen:
                continue
            seen.add((qid, tid))
            ovl = [qid, int(f[1]), int(f[2]), int(f[3]), 0 if f[4][:1] == b"+" else 1, tid, int(f[6]), int(f[7]), int(f[8])]
            out.setdefault(tid, []).append((ovl, f[-1][5:]))
    return out


def inference(args):
    core = neigh = None
    if args.cluster:
        core, neigh = read_cluster(args.cluster)
    files = [args.reads] if os.path.isfile(args.reads) else sorted(
        p for p in glob.glob(os.path.join(args.reads, "*")) if p.endswith(".fastq") or p.endswith(".fastq.gz"))
    ids, descs, seqs, quals = [], [], [], []
    for p in files:
        a, b, c, d = read_fastq(p, args.window_size)
        for i in range(len(a)):
            if core is not None and a[i] not in core and a[i] not in neigh:
                continue
            ids.append(a[i]); descs.append(b[i]); seqs.append(c[i]); quals.append(d[i])
    name_to_id = {n: i for i, n in enumerate(ids)}
    alns = read_oec_batches(args.read_alns, name_to_id, core)
    off = np.zeros(len(ids) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    seq_arr = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    qual_arr = np.frombuffer(b"".join(quals), dtype=np.uint8)
    devices = [int(d) for d in str(args.devices).split(",")]
    ctxs = [api.Context(args.model, d, args.window_size, args.batch_size) for d in devices]
    for c in ctxs:
        c.upload_reads(seq_arr, qual_arr, off)
    # targets are dealt to the devices like the reference's per-device workers pulling one channel
    n_rec = 0
    with open(args.output, "wb") as out:
        for k, (tid, lst) in enumerate(alns.items()):
            ovl9 = np.array([o for o, _ in lst], dtype=np.uint32)
            cig = np.frombuffer(b"".join(c for _, c in lst), dtype=np.uint8)
            coff = np.zeros(len(lst) + 1, dtype=np.uint64)
            coff[1:] = np.cumsum([len(c) for _, c in lst])
            ctxs[k % len(ctxs)].submit_alignments(tid, api.Context.make_overlaps(ovl9, cig, coff))
        for c in ctxs:
            c.flush()
            for r in c.drain(skip_failed=True):
                if r.segments:
                    out.write(api.fasta_records(ids[r.rid], descs[r.rid], r.segments))
                    n_rec += len(r.segments)
            for rid, code, msg in c.failed:  # the reference would have aborted the whole run here; log the read and go on
                print(f"skipped read {ids[rid].decode() if rid is not None else '?'}: {msg}", file=sys.stderr)
    print(f"Processed {len(alns)} reads, wrote {n_rec} records.", file=sys.stderr)


def main(argv=None):
    ap = argparse.ArgumentParser(prog="herro_b200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    inf = sub.add_parser("inference")
    inf.add_argument("--read-alns", required=True, help="directory with *.oec.zst alignment batches")
    inf.add_argument("-w", dest="window_size", type=int, default=4096)
    inf.add_argument("-t", dest="feat_gen_threads", type=int, default=1)
    inf.add_argument("-m", dest="model", required=True)
    inf.add_argument("-d", dest="devices", default="0")
    inf.add_argument("-b", dest="batch_size", type=int, required=True)
    inf.add_argument("-c", dest="cluster", default="")
    inf.add_argument("reads")
    inf.add_argument("output")
    args = ap.parse_args(argv)
    if args.cmd == "inference":
        inference(args)


if __name__ == "__main__":
    main()
