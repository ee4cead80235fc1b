#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: instruction mix by opcode and the hottest lines."""
import collections
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    hdr = rows[h]
    si, ii, st = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    tot = 0
    byop, stall, data = collections.Counter(), collections.Counter(), []
    for r in rows[h + 1:]:
        if len(r) <= max(ii, st) or not r[ii].isdigit():
            continue
        n, s, src = int(r[ii]), int(r[st] or 0), r[si].strip()
        toks = src.split()
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.split(".")[0]
        byop[op] += n
        stall[op] += s
        tot += n
        data.append((n, s, src))
    print("total warp instructions", tot, " stall samples", sum(stall.values()))
    for op, n in byop.most_common(top):
        print(f"{op:14s} {n:10d} {100 * n / tot:5.1f}%   stall {stall[op]:7d} {100 * stall[op] / max(1, sum(stall.values())):5.1f}%")
    print("--- top stall lines")
    for n, s, src in sorted(data, key=lambda t: -t[1])[:top]:
        print(f"{s:7d} {n:9d}  {src[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
