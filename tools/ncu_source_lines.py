#!/usr/bin/env python
"""Attribute the per-instruction samples of `ncu --page source --csv --print-source sass` to CUDA source lines, using the line
table of the in-tree object files (nvcc -lineinfo; nvdisasm -g).  Usage:
    python tools/ncu_source_lines.py gpurun_out/src_r02.csv [kernel-substring] [top-N]
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_tables():
    """mangled kernel name -> {offset: (file, line)}"""
    out = {}
    tmp = tempfile.mkdtemp()
    for obj in glob.glob(os.path.join(ROOT, "herro_b200", "csrc", "*.o")):
        subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, capture_output=True)
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        kern, line = None, None
        for l in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", l)
            if m:
                kern = m.group(1)
                out.setdefault(kern, {})
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.search(r"/\*([0-9a-f]{4,})\*/", l)
            if m and kern:
                out[kern][int(m.group(1), 16)] = line
    return out


def main(path, want="", top=30):
    tabs = line_tables()
    rows = list(csv.reader(open(path)))
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name = rows[i][1]
            hdr = rows[i + 1]
            idx = {h: k for k, h in enumerate(hdr)}
            j = i + 2
            body = []
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                if len(rows[j]) == len(hdr):
                    body.append(rows[j])
                j += 1
            i = j
            if want not in name or not body:
                continue
            short = re.sub(r"\(.*", "", name)
            key = next((k for k in tabs if re.sub(r"[^A-Za-z0-9_]", "", short.split("::")[-1].split("<")[0]) in k and len(tabs[k]) == len(body)), None)
            base = int(body[0][0], 16)
            stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
            T = I = 0
            for r in body:
                s, n = int(r[idx["# Samples"]] or 0), int(r[idx["Instructions Executed"]] or 0)
                ln = tabs[key].get(int(r[0], 16) - base) if key else None
                a = agg[ln]
                a[0] += s
                a[1] += n
                T += s
                I += n
                for st in stalls:
                    a[2][st] += int(r[idx[st]] or 0)
            tot = collections.Counter()
            for a in agg.values():
                tot.update(a[2])
            print(f"== {short}: {T} samples, {I} warp instructions; stalls: " + ", ".join(f"{k[6:]} {100 * v / max(T, 1):.0f}%" for k, v in tot.most_common(6)))
            src_cache = {}
            for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(top)]:
                text = ""
                if ln:
                    f = os.path.join(ROOT, "herro_b200", "csrc", ln[0])
                    if os.path.exists(f):
                        src_cache.setdefault(f, open(f).read().split("\n"))
                        text = src_cache[f][ln[1] - 1].strip()[:100]
                t2 = ", ".join(f"{k[6:]} {100 * v / max(a[0], 1):.0f}%" for k, v in a[2].most_common(2))
                print(f"  {100 * a[0] / max(T, 1):5.1f}% smp {100 * a[1] / max(I, 1):5.1f}% inst  {str(ln):28s} [{t2}] {text}")
        else:
            i += 1


if __name__ == "__main__":
    main(*sys.argv[1:4])
