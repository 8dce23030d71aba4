#!/usr/bin/env python
"""Per-source-line roll-up of an ncu source page (SASS) using nvdisasm line info of the same build.

    ncu -i X.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all libfastdiff_b200.so; nvdisasm --print-line-info fd_api.sm_100a.cubin > lines.txt
    python tools/ncu_lines.py src.csv lines.txt '<mangled kernel name>' [top]
"""
import csv
import re
import sys


def sass_lines(path, fn):
    out, cur, on = [], None, False
    for l in open(path):
        if l.startswith(".text."):
            on = l.strip() == f".text.{fn}:"
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)), "inlined" in m.group(3))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            out.append((int(m.group(1), 16), cur, m.group(2)))
    return out


def main():
    src, lines, fn = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    sl = sass_lines(lines, fn)
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[hi + 1:]:
        if r and r[0] == "Kernel Name":   # a second block (ncu repeats the listing): keep the first
            break
        if len(r) == len(hdr):
            data.append(r)
    assert len(data) == len(sl), (len(data), len(sl))
    agg = {}
    tot_i = tot_s = 0
    for r, (off, cur, txt) in zip(data, sl):
        inst, samp = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        exc = int(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
        key = cur[:2] if cur else ("?", 0)
        a = agg.setdefault(key, [0, 0, 0, 0])
        a[0] += inst; a[1] += samp; a[2] += exc; a[3] += 1
        tot_i += inst; tot_s += samp
    print(f"# total warp-instructions {tot_i}, samples {tot_s}, sass instructions {len(sl)}")
    srcs = {}
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        if f not in srcs:
            try:
                srcs[f] = open(f"/root/repo/fastdiff_b200/csrc/{f}").read().split("\n")
            except Exception:
                srcs[f] = []
        text = srcs[f][ln - 1].strip()[:110] if 0 < ln <= len(srcs[f]) else ""
        print(f"{100 * a[1] / max(tot_s, 1):5.1f}% samp {100 * a[0] / max(tot_i, 1):5.1f}% inst  excess_wf {a[2]:>9d}  sass {a[3]:4d}  {f}:{ln}: {text}")


if __name__ == "__main__":
    main()
