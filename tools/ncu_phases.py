#!/usr/bin/env python
"""Split a kernel's ncu source page at its BAR.SYNC instructions (= the phases of the tile loop of the LVC kernels) and list the
SASS instructions that hold the most warp-stall samples.

    ncu -i X.ncu-rep --page source --csv --kernel-id :::N > src.csv
    cuobjdump -xelf all libfastdiff_b200.so; nvdisasm --print-line-info fd_api.sm_100a.cubin > lines.txt
    python tools/ncu_phases.py src.csv lines.txt '<mangled kernel name>' [min % of samples for the instruction list]
"""
import csv
import sys

import ncu_lines


def main():
    src, lines, fn = sys.argv[1:4]
    floor = float(sys.argv[4]) if len(sys.argv) > 4 else 1.5
    sl = ncu_lines.sass_lines(lines, fn)
    rows = list(csv.reader(open(src)))
    print("#", rows[0][1][:100])
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[hi + 1:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) == len(hdr):
            data.append(r)
    assert len(data) == len(sl), (len(data), len(sl))
    num = lambda r, k: int(float(r[ix[k]] or 0))
    tot = sum(num(r, "# Samples") for r in data)
    toti = sum(num(r, "Instructions Executed") for r in data)
    print(f"# {tot} stall samples, {toti} warp-instructions, {len(sl)} SASS instructions")
    stall = {h: sum(num(r, h) for r in data) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
    print("# stall reasons:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:7]))
    seg, cur = [], dict(start=0, samp=0, inst=0, wf=0, ex=0, lines=set())
    for i, (r, (off, c, txt)) in enumerate(zip(data, sl)):
        cur["samp"] += num(r, "# Samples"); cur["inst"] += num(r, "Instructions Executed")
        cur["wf"] += num(r, "L1 Wavefronts Shared"); cur["ex"] += num(r, "L1 Wavefronts Shared Excessive")
        if c:
            cur["lines"].add(c[1])
        if "BAR.SYNC" in txt:
            cur["end"] = i; seg.append(cur); cur = dict(start=i + 1, samp=0, inst=0, wf=0, ex=0, lines=set())
    cur["end"] = len(data); seg.append(cur)
    print("# segments between BAR.SYNC instructions")
    for s in seg:
        ls = sorted(s["lines"])
        print(f"sass {s['start']:5d}-{s['end']:5d}  samples {100 * s['samp'] / tot:5.1f}%  instructions {s['inst'] / 1e6:7.2f}M ({100 * s['inst'] / toti:4.1f}%)"
              f"  smem wavefronts {s['wf'] / 1e6:6.2f}M (excess {s['ex'] / 1e6:5.2f}M)  source lines {ls[0] if ls else ''}..{ls[-1] if ls else ''}")
    print(f"# instructions with >= {floor}% of the samples")
    for i, (r, (off, c, txt)) in enumerate(zip(data, sl)):
        s = num(r, "# Samples")
        if 100 * s / tot >= floor:
            print(f"{i:5d} {100 * s / tot:5.1f}%  {txt[:64]:64s} line {c[1] if c else 0}")


if __name__ == "__main__":
    main()
