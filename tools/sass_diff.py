#!/usr/bin/env python
"""Compare two builds of libfastdiff_b200.so kernel by kernel (cuobjdump -sass, addresses and line info stripped).
Used to show that a source change (refactor, new optional kernel) left the machine code of the existing kernels untouched when no
GPU is at hand:  python tools/sass_diff.py old.so new.so"""
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    res, name, body = {}, None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                res[name] = body
            name, body = m.group(1), []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?)\s*/\*", line)
        if m and name:
            body.append(m.group(1))
    if name:
        res[name] = body
    return res


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    same = [k for k in a if k in b and a[k] == b[k]]
    diff = [k for k in a if k in b and a[k] != b[k]]
    print(f"{len(same)} kernels identical, {len(diff)} differ, {len(set(a) - set(b))} removed, {len(set(b) - set(a))} added")
    for k in diff:
        print("  DIFF", k, len(a[k]), "->", len(b[k]), "instructions")
    gone = sorted(set(a) - set(b))
    for k in sorted(set(b) - set(a)):
        twin = [g for g in gone if a[g] == b[k]]     # e.g. a template that gained a defaulted parameter: new name, same code
        print("  NEW ", k, len(b[k]), "instructions" + (f"  == body of removed {twin[0]}" if twin else ""))
    for k in gone:
        print("  GONE", k, "(identical body present under a new name)" if any(a[k] == v for v in b.values()) else "")
    return 1 if diff else 0


if __name__ == "__main__":
    sys.exit(main())
