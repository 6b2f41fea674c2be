#!/usr/bin/env python3
"""Run a few scans of one config on cuda:0 (short driver for ncu / quick timing).

    python tools/prof_one.py --config c2 --gib 1 --reps 3 [--generic]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from ugrep_b200 import api, corpus  # noqa: E402

CFG = {"c3bl": ("c3b", "c3", "lines"), "c1": ("c1", "c1", "lines"), "c2": ("c2", "c2", "lines"), "c2s": ("c2", "c2s", "lines"), "c3b": ("c3b", "c3", "list"),
       "c3c": ("c3c", "c3", "list"), "c4": ("c4", "c4", "lines"), "c5": ("c5", "c5", "matches"),
       "c5l": ("c5", "c5", "lines"), "c3bm": ("c3b", "c3", "matches")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--gib", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--generic", action="store_true")
    ap.add_argument("--pattern", default=None, help="a .ugxp file instead of the config's pattern")
    ap.add_argument("--mode", default=None, choices=["lines", "matches", "list"])
    ap.add_argument("--opt", action="append", default=[], help="scanner option name=value")
    a = ap.parse_args()
    pname, cname, mode = CFG[a.config]
    block = corpus.block(cname, 64 << 20)
    reps = max(1, int(a.gib * (1 << 30)) // block.size)
    dev = torch.from_numpy(block).cuda().repeat(reps)
    pat = api.Pattern.load(a.pattern or os.path.join(ROOT, "ugrep_b200", "patterns", pname + ".ugxp"), 0)
    mode = a.mode or mode
    sc = api.Scanner(0, torch.cuda.current_stream().cuda_stream)
    if a.generic:
        sc.set_option("force_generic", 1)
    for o in a.opt:
        k, v = o.split("=")
        sc.set_option(k, int(v))
    best = None
    for _ in range(a.reps):
        if mode == "lines":
            t = sc.count_lines(pat, dev)
        elif mode == "matches":
            t = sc.count_matches(pat, dev)
        else:
            t = sc.find_all_device(pat, dev)
        best = t.kernel_ms if best is None else min(best, t.kernel_ms)
    print("%s %s: %.3f GiB, best %.3f ms = %.1f GB/s, result %d, newlines %d, launches %d, info %s"
          % (a.config, mode, dev.numel() / (1 << 30), best, dev.numel() / best / 1e6, t.matches, t.newlines, t.launches,
             pat.info))


if __name__ == "__main__":
    main()
