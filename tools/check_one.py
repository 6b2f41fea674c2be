#!/usr/bin/env python3
"""Quick GPU-vs-oracle check of one config on cuda:0 (development aid; the real parity tests are tests/ -m gpu).

    python tools/check_one.py c2 16 [--opt no_cover=1]      # config, MiB of corpus
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from ugrep_b200 import api, corpus  # noqa: E402
import oracle_lib as O  # noqa: E402

CFG = {"c1": ("c1", "c1", "lines"), "c2": ("c2", "c2", "lines"), "c2s": ("c2", "c2s", "lines"), "c3b": ("c3b", "c3", "list"),
       "c4": ("c4", "c4", "lines"), "c5": ("c5", "c5", "matches"), "c5l": ("c5", "c5", "lines"), "c3bm": ("c3b", "c3", "matches"),
       "c2m": ("c2", "c2", "matches"), "c2r": ("c2", "c2", "list"), "c2sm": ("c2", "c2s", "matches")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("mib", type=int)
    ap.add_argument("--opt", action="append", default=[])
    a = ap.parse_args()
    pname, cname, mode = CFG[a.config]
    block = corpus.block(cname, a.mib << 20)
    path = os.path.join(ROOT, "ugrep_b200", "patterns", pname + ".ugxp")
    pat = api.Pattern.load(path, 0)
    sc = api.Scanner(0, torch.cuda.current_stream().cuda_stream)
    for o in a.opt:
        k, v = o.split("=")
        sc.set_option(k, int(v))
    dev = torch.from_numpy(block).cuda()
    op = O.OraclePattern(path)
    if mode == "lines":
        want, got = op.count_lines(block), [sc.count_lines(pat, dev).matches for _ in range(3)]
    elif mode == "matches":
        want, got = op.count_matches(block), [sc.count_matches(pat, dev).matches for _ in range(3)]
    else:
        want = len(op.find_all(block))
        got = [sc.find_all_device(pat, dev).matches for _ in range(3)]
    print(a.config, mode, a.opt, "gpu", got, "oracle", want, "OK" if all(g == want for g in got) else "MISMATCH")


if __name__ == "__main__":
    main()
