#!/usr/bin/env python3
"""Word lists of growing size, compiled by the library itself (ugx_compile_words), counted over the dense c2 corpus:
what happens when the transition table no longer fits in shared memory (tables > ~190 KiB are read from global / L2).

    python tools/big_list.py [--gib 1] [--counts 1000,3000,10000]      (on the GPU box)
    python tools/big_list.py --dry                                     (CPU: compile + table sizes only)
"""
import argparse
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ugrep_b200 import api, corpus  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=1.0)
    ap.add_argument("--counts", default="1000,3000,10000")
    ap.add_argument("--dry", action="store_true")
    a = ap.parse_args()
    block = corpus.block("c2", 64 << 20)
    for count in [int(x) for x in a.counts.split(",")]:
        words = corpus.words_list(11, count)
        try:
            opc, pf = api.compile_words(words)
        except api.UgxError as e:
            print("%d words: compiler refused (%s)" % (count, e))
            continue
        if a.dry:
            print("%d words: %d opcode words" % (count, len(opc)))
            continue
        import torch
        import oracle_lib as O
        try:
            pat = api.Pattern.words(words, 0)
        except api.UgxError as e:
            print("%d words: refused at upload (%s)" % (count, e))
            continue
        info = pat.info
        sc = api.Scanner(0, torch.cuda.current_stream().cuda_stream)
        reps = max(1, int(a.gib * (1 << 30)) // block.size)
        dev = torch.from_numpy(block).cuda().repeat(reps)
        best = min(sc.count_lines(pat, dev).kernel_ms for _ in range(3))
        got = sc.count_lines(pat, torch.from_numpy(block[:4 << 20].copy()).cuda()).matches
        with tempfile.TemporaryDirectory() as d:
            p = os.path.join(d, "w.ugxp")
            api.write_ugxp(p, opc, pf)
            want = O.OraclePattern(p).count_lines(block[:4 << 20])
        print("%5d words: table %7d B, in smem %d, %7.1f GB/s, 4 MiB check gpu %d oracle %d %s"
              % (count, info["table_bytes"], info["table_in_smem"], dev.numel() / best / 1e6, got, want,
                 "OK" if got == want else "MISMATCH"))


if __name__ == "__main__":
    main()
