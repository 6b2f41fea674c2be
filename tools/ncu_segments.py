#!/usr/bin/env python3
"""Group an ncu report's SASS instructions into runs of equal execution count (= basic-block paths) and print
the heaviest runs: python tools/ncu_segments.py REPORT.ncu-rep [N]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    h = rows[hi]
    ci, cs, ca = h.index("Instructions Executed"), h.index("Source"), h.index("Address")
    data = []
    for r in rows[hi + 1:]:
        try:
            data.append((int(r[ca], 16), r[cs], int(r[ci] or 0)))
        except (ValueError, IndexError):
            pass
    base = data[0][0]
    tot = sum(d[2] for d in data)
    segs, prev, start, acc = [], None, 0, 0
    for i, d in enumerate(data):
        c = d[2]
        if prev is None or abs(c - prev) > 0.02 * max(prev, 1):
            if prev is not None:
                segs.append((start, i - 1, prev, acc))
            start, acc = i, 0
        acc += c
        prev = c
    segs.append((start, len(data) - 1, prev, acc))
    segs.sort(key=lambda s: -s[3])
    print("total warp-instructions", tot, "SASS instructions", len(data))
    for s in segs[:top]:
        print("%5d-%5d (%4d insts) x %10d  share %5.1f%%  off 0x%x  %s" % (
            s[0], s[1], s[1] - s[0] + 1, s[2], 100 * s[3] / tot, data[s[0]][0] - base, data[s[0]][1][:60]))


if __name__ == "__main__":
    main()
