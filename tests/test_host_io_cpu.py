"""The native host data plane (herro_b200/host/io.cpp) against the reference's semantics, on CPU:
FASTQ -> packed read store (src/haec_io.rs:37-75,121-136), *.oec.zst batches -> alignments grouped by target
(src/overlaps.rs:288-323,117-202), FASTA records (src/lib.rs:294-317)."""
import gzip
import os

import numpy as np
import pytest

import helpers
from herro_b200 import api, hostio
from tools import synth


@pytest.fixture(scope="module")
def rs():
    return helpers.small_readset(n_reads=40, mean_len=3000, seed=91, coverage=12.0, min_ovl=500, sd_frac=0.4, min_len=300)


@pytest.mark.parametrize("gz", [False, True])
def test_fastq_reader_matches_haec_io(rs, tmp_path, gz):
    descs = [None if i % 3 else f"ch={i}\tstart=1.5 x" for i in range(rs.n)]
    fq = str(tmp_path / ("reads.fastq.gz" if gz else "reads.fastq"))
    plain = str(tmp_path / "plain.fastq")
    synth.write_fastq(rs, plain, descs)
    if gz:
        with open(plain, "rb") as f, gzip.open(fq, "wb") as g:
            g.write(f.read())
    else:
        fq = plain
    W = 2048
    R = hostio.Reads(fq, min_len=W, threads=3)
    keep = [i for i in range(rs.n) if int(rs.off[i + 1] - rs.off[i]) >= W]      # reads < window_size are dropped (src/haec_io.rs:48)
    assert R.n == len(keep) and 0 < R.n < rs.n
    assert R.stats()["skipped_short"] == rs.n - len(keep)
    for k, i in enumerate(keep):
        assert R.ids[k] == rs.ids[i].encode()
        want = None if descs[i] is None else descs[i].encode()
        assert R.descriptions[k] == want                                         # everything after the first space / tab (:52-54)
        assert int(R.lens[k]) == int(rs.off[i + 1] - rs.off[i])
        assert R.qual(k) == rs.qual(i)
        assert np.array_equal(R.words(k), api.pack_2bit(rs.seqs[int(rs.off[i]):int(rs.off[i + 1])]))  # HAECSeq layout (:121-136)


def test_fastq_reader_cluster_filter_and_directory(rs, tmp_path):
    d = tmp_path / "fq"
    d.mkdir()
    half = rs.n // 2
    for name, lo, hi in (("a.fastq", 0, half), ("b.fastq", half, rs.n)):
        with open(d / name, "wb") as f:
            for i in range(lo, hi):
                f.write(b"@" + rs.ids[i].encode() + b"\n" + rs.seq(i) + b"\n+\n" + rs.qual(i) + b"\n")
    (d / "ignored.txt").write_text("x")
    R = hostio.Reads(str(d), min_len=1, threads=2)
    assert sorted(R.ids) == sorted(x.encode() for x in rs.ids)
    core, neigh = [rs.ids[0], rs.ids[5]], [rs.ids[7]]
    R2 = hostio.Reads(str(d), min_len=1, core=core, neighbour=neigh)             # src/haec_io.rs:64-70
    assert sorted(R2.ids) == sorted(x.encode() for x in core + neigh)


def test_fastq_reader_rejects_non_acgt_and_fasta(tmp_path):
    p = tmp_path / "n.fastq"
    p.write_bytes(b"@r1\nACGTNACGT\n+\nIIIIIIIII\n")
    with pytest.raises(api.HerroError):
        hostio.Reads(str(p), min_len=1)
    p.write_bytes(b">r1\nACGT\n")
    with pytest.raises(api.HerroError):
        hostio.Reads(str(p), min_len=1)


def _python_groups(rs, names, batch_of, core=None):
    """The reference's grouping, restated in Python: per batch file, first line of an ordered (q,t) pair wins, self overlaps and
    unknown names are dropped, the core filter applies to the target name."""
    out = {}
    for b, targets in batch_of.items():
        seen = set()
        for t in targets:
            for a in range(int(rs.aln_off[t]), int(rs.aln_off[t + 1])):
                q = int(rs.ovl9[a][0])
                if rs.ids[q] not in names or rs.ids[t] not in names or (core is not None and rs.ids[t] not in core):
                    continue
                if q == t or (q, t) in seen:
                    continue
                seen.add((q, t))
                out.setdefault((b, t), []).append(a)
    return out


def test_oec_reader_matches_parse_paf(rs, tmp_path):
    fq = str(tmp_path / "reads.fastq")
    synth.write_fastq(rs, fq)
    R = hostio.Reads(fq, min_len=1500)           # some reads vanish: their alignments must be skipped (name lookup fails)
    names = {i.decode() for i in R.ids}
    alns = tmp_path / "alns"
    synth.write_oec_batches(rs, str(alns), batch_size=9)
    nb = len(list(alns.glob("*.oec.zst")))
    assert nb >= 4
    batch_of = {b: list(range(b * 9, min((b + 1) * 9, rs.n))) for b in range(nb)}
    A = hostio.Alignments(str(alns), R, threads=3)
    want = _python_groups(rs, names, batch_of)
    got = {}
    name_of = [i.decode() for i in R.ids]
    for k in range(A.n_targets):
        rid, ov = A.target(k)
        got[name_of[rid]] = (k, ov)
    assert sorted(got) == sorted(rs.ids[t] for (_, t) in want)
    for (b, t), alist in want.items():
        k, ov = got[rs.ids[t]]
        assert len(ov) == len(alist)
        base = int(A.offsets[k])
        for j, a in enumerate(alist):
            q, ql, qs, qe, st, tt, tl, ts, te = (int(x) for x in rs.ovl9[a])
            o = ov[j]
            assert name_of[int(o["qid"])] == rs.ids[q] and name_of[int(o["tid"])] == rs.ids[t]
            assert (int(o["qlen"]), int(o["qstart"]), int(o["qend"]), int(o["strand"]), int(o["tlen"]), int(o["tstart"]), int(o["tend"])) == \
                   (ql, qs, qe, st, tl, ts, te)
            assert A.cigar(base + j) == rs.cigar(a)
    st = A.stats()
    assert st["kept"] == sum(len(v) for v in want.values()) and st["lines"] >= st["kept"] and st["text_bytes"] > st["compressed_bytes"]


def test_oec_reader_duplicates_self_overlaps_core(rs, tmp_path):
    import pyarrow as pa
    fq = str(tmp_path / "reads.fastq")
    synth.write_fastq(rs, fq)
    R = hostio.Reads(fq, min_len=1)
    t = next(t for t in range(rs.n) if rs.aln_off[t + 1] - rs.aln_off[t] >= 3)
    lines = list(synth.paf_lines(rs, [t]))
    dup = lines[0].replace(b"\t60\t", b"\t13\t")                    # same (q,t) pair again, later: ignored
    selfl = (f"{rs.ids[t]}\t100\t0\t100\t+\t{rs.ids[t]}\t100\t0\t100\t0\t0\t60\tcg:Z:100M\n").encode()
    unknown = lines[1].replace(rs.ids[int(rs.ovl9[int(rs.aln_off[t]) + 1][0])].encode(), b"nobody")
    body = b"1\n" + rs.ids[t].encode() + b"\n" + lines[0] + selfl + unknown + dup + b"".join(lines[2:])
    d = tmp_path / "alns"
    d.mkdir()
    (d / "0.oec.zst").write_bytes(pa.Codec("zstd").compress(body, asbytes=True))
    A = hostio.Alignments(str(d), R)
    assert A.n_targets == 1
    rid, ov = A.target(0)
    assert R.ids[rid] == rs.ids[t].encode() and len(ov) == len(lines) - 1    # line 1 lost its query, the duplicate and the self overlap are dropped
    assert A.cigar(0) == rs.cigar(int(rs.aln_off[t]))
    A2 = hostio.Alignments(str(d), R, core=["someone_else"])
    assert A2.n_targets == 0                                                   # core filter on the target name (src/overlaps.rs:155-160)


def test_fasta_writer_format(tmp_path):
    p = str(tmp_path / "o.fasta")
    w = hostio.FastaWriter(p)
    w.write(b"r1", None, [b"ACGT"])
    w.write(b"r2", b"ch=1 x", [b"AAA", b"CC"])
    w.write(b"r3", b"d", [])
    rec, bases = w.close()
    assert (rec, bases) == (3, 9)
    # `>id ` always carries the space; `:k` only when a read has several segments (src/lib.rs:282-314, H7)
    assert open(p, "rb").read() == b">r1 \nACGT\n>r2:0 ch=1 x\nAAA\n>r2:1 ch=1 x\nCC\n"
    assert api.fasta_records(b"r2", b"ch=1 x", [b"AAA", b"CC"]) == b">r2:0 ch=1 x\nAAA\n>r2:1 ch=1 x\nCC\n"
