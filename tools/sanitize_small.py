#!/usr/bin/env python3
"""Small run of every kernel route on odd-sized device buffers, each its own allocation (run with
PYTORCH_NO_CUDA_MEMORY_CACHING=1 so that a read far past `n` leaves the allocation); results are checked against
the oracle.  Written for `compute-sanitizer --tool memcheck`, which is closed on this pool: it still serves as a
quick all-routes consistency run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import oracle_lib as O  # noqa: E402
from ugrep_b200 import api, corpus  # noqa: E402


def main():
    pats = {"c1": "c1", "c2": "c2", "c3b": "c3", "c3c": "c3", "c4": "c4", "c5": "c5"}
    routes = [{}, {"stream_dfa": 1, "count_newlines": 1}, {"legacy_any": 1}, {"force_generic": 1}, {"two_pass_records": 1}]
    bad = 0
    for pname, cname in pats.items():
        path = os.path.join(ROOT, "ugrep_b200", "patterns", pname + ".ugxp")
        pat = api.Pattern.load(path, 0)
        op = O.OraclePattern(path)
        base = corpus.block(cname, 70000)
        for n in (1, 15, 17, 511, 2049, 16385, 33001, len(base)):
            data = base[:n].copy()
            # cudaMalloc directly (256-byte granularity) so that reads past n + 255 fault under memcheck
            dev = torch.empty(n, dtype=torch.uint8, device="cuda")
            dev.copy_(torch.from_numpy(data))
            for r in routes:
                sc = api.Scanner(0)
                for k, v in r.items():
                    sc.set_option(k, v)
                got = (sc.count_lines(pat, dev).matches, sc.count_matches(pat, dev).matches, len(sc.find_all(pat, dev)[0]),
                       sc.count_newlines(dev).newlines)
                want = (op.count_lines(data), op.count_matches(data), len(op.find_all(data)), int((data == 10).sum()))
                if got != want:
                    bad += 1
                    print("MISMATCH", pname, n, r, got, want)
                sc.close()
    print("sanitize_small: %s" % ("ok" if bad == 0 else "%d mismatches" % bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
