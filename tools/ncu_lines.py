#!/usr/bin/env python3
"""Per-source-line view of an ncu report: joins `ncu --page source --csv` (SASS level) with the
line table of the cubin (`nvdisasm -g`), because the CUDA-C view of ncu's CSV export carries no
metrics.  Usage:

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep decode_blocks_v4 [--top 40] [--so bo_lz4_ada_b200/liblz4b200.so]
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    out = []
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, on = None, False
        for line in txt.splitlines():
            if line.startswith("//--------------------- .text."):
                on = kernel in line
                continue
            if not on:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", line)
            if m:
                out.append((int(m.group(1), 16), cur, m.group(2).strip()))
        if out:
            break
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--so", default="bo_lz4_ada_b200/liblz4b200.so")
    ap.add_argument("--by", default="inst", choices=["inst", "samples"])
    ap.add_argument("--cubin-kernel", default=None, help="substring of the MANGLED name in the cubin (template instantiations)")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # several kernels may be in the report: take the first section whose name matches
    start = None
    for i, r in enumerate(rows):
        if r and r[0] == "Kernel Name" and args.kernel in r[1]:
            start = i
            break
    if start is None:
        sys.exit("kernel not in report")
    hdr = rows[start + 1]
    col = {h: k for k, h in enumerate(hdr)}
    body = []
    for r in rows[start + 2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) == len(hdr):
            body.append(r)
    sass = sass_lines(args.so, args.cubin_kernel or args.kernel)
    if len(sass) != len(body):
        print("warning: %d SASS instructions in the cubin, %d in the report" % (len(sass), len(body)), file=sys.stderr)
    agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    tot_i = tot_s = 0
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for (addr, loc, text), r in zip(sass, body):
        n_i = int(float(r[col["Instructions Executed"]] or 0))
        n_t = int(float(r[col["Thread Instructions Executed"]] or 0))
        n_s = int(float(r[col["# Samples"]] or 0))
        a = agg[loc]
        a[0] += n_i
        a[1] += n_t
        a[2] += n_s
        for h in stall_cols:
            v = int(float(r[col[h]] or 0))
            if v:
                a[3][h[6:]] += v
        tot_i += n_i
        tot_s += n_s
    key = (lambda kv: -kv[1][0]) if args.by == "inst" else (lambda kv: -kv[1][2])
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    src_cache = {}
    for loc, a in sorted(agg.items(), key=key)[:args.top]:
        text = ""
        if loc:
            path = os.path.join("bo_lz4_ada_b200/csrc", loc[0])
            if path not in src_cache and os.path.exists(path):
                src_cache[path] = open(path).read().splitlines()
            if path in src_cache and loc[1] - 1 < len(src_cache[path]):
                text = src_cache[path][loc[1] - 1].strip()[:70]
        top = ",".join("%s:%d" % kv for kv in a[3].most_common(3))
        print("%5.1f%% inst %5.1f%% smp  thr/inst %4.1f  %s:%s  %-70s %s" % (
            100.0 * a[0] / max(tot_i, 1), 100.0 * a[2] / max(tot_s, 1), a[1] / max(a[0], 1),
            loc[0] if loc else "?", loc[1] if loc else 0, text, top))


if __name__ == "__main__":
    main()
