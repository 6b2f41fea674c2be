#!/usr/bin/env python3
"""Attribute an ncu report's per-SASS-instruction counters to CUDA source lines.

    python tools/ncu_hot.py REPORT.ncu-rep KERNEL_SUBSTRING [cubin-name-substring]

ncu's CSV export of the source page carries metrics only in SASS view; this joins it with
nvdisasm's line info (the library is built with -lineinfo) and prints the hottest source lines.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    cub = sys.argv[3] if len(sys.argv) > 3 else "fast_kernels"
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(os.environ.get("UGX_LIB") or os.path.join(ROOT, "ugrep_b200", "libugrep_b200.so"))], cwd=d,
                   capture_output=True)
    cubin = [f for f in os.listdir(d) if f.startswith(cub + ".")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    # walk the disassembly: track current function, current source line, instruction offsets
    line_of = {}
    func = None
    cur = ("?", 0)
    for ln in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            func = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and func and kern in func:
            line_of[(func, int(m.group(1), 16))] = (cur, m.group(2).strip())
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[hdr_i]
    col = {n: hdr.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed",
                                     "Thread Instructions Executed", "L1 Wavefronts Shared")}
    kname = rows[0][1] if rows[0] and rows[0][0] == "Kernel Name" else ""
    funcs = sorted({f for f, _ in line_of})
    # choose the function whose instruction count matches best: use the first containing kern
    agg = {}
    tot = [0, 0, 0, 0]
    base = None
    n_inst = 0
    for r in rows[hdr_i + 1:]:
        if len(r) <= col["Instructions Executed"]:
            continue
        try:
            addr = int(r[col["Address"]], 16)
        except ValueError:
            continue
        if base is None:
            base = addr
        off = addr - base
        ie = int(r[col["Instructions Executed"]] or 0)
        te = int(r[col["Thread Instructions Executed"]] or 0)
        sm = int(r[col["# Samples"]] or 0)
        wf = int(r[col["L1 Wavefronts Shared"]] or 0)
        n_inst += 1
        key = None
        for f in funcs:
            if (f, off) in line_of and line_of[(f, off)][1].split()[0] in r[col["Source"]]:
                key = line_of[(f, off)][0]
                break
        if key is None:
            key = ("?", 0)
        a = agg.setdefault(key, [0, 0, 0, 0])
        a[0] += ie
        a[1] += te
        a[2] += sm
        a[3] += wf
        tot[0] += ie
        tot[1] += te
        tot[2] += sm
        tot[3] += wf
    print("kernel:", kname[:120])
    print("total warp-inst %d, thread-inst %d, samples %d, smem wavefronts %d, SASS insts %d" % (*tot, n_inst))
    per_file = {}
    for (f, l), a in agg.items():
        per_file[f] = per_file.get(f, 0) + a[0]
    print("per file: " + ", ".join("%s %.1f%%" % (f, 100.0 * v / max(1, tot[0])) for f, v in sorted(per_file.items(), key=lambda kv: -kv[1])))
    srcs = {}
    top = int(os.environ.get("NCU_HOT_TOP", "40"))
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in srcs:
            p = os.path.join(ROOT, "ugrep_b200", "csrc", f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = srcs[f][l - 1].strip() if 0 < l <= len(srcs[f]) else ""
        print("%5.1f%% inst %5.1f%% smp %5.1f%% wf  act=%4.1f  %s:%d  %s" % (
            100.0 * a[0] / max(1, tot[0]), 100.0 * a[2] / max(1, tot[2]), 100.0 * a[3] / max(1, tot[3]),
            a[1] / max(1, a[0]), f, l, text[:90]))


if __name__ == "__main__":
    main()
