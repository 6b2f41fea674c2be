#!/usr/bin/env python3
"""Randomised differential run on the GPU: every golden pattern x random texts (corpus slices spliced with random
bytes, long runs of one byte, very long and empty lines, missing final newline), all three modes through the default
routes, against the oracle.  python tools/fuzz_gpu.py [seconds] [seed]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import golden_lib as G  # noqa: E402
import oracle_lib as O  # noqa: E402
from ugrep_b200 import api, corpus  # noqa: E402


def make_text(rng, blocks):
    parts = []
    total = int(rng.choice([200, 5000, 40000, 200000, 700000]))
    while sum(len(p) for p in parts) < total:
        kind = int(rng.integers(0, 9))
        if kind <= 3:
            b = blocks[int(rng.integers(0, len(blocks)))]
            lo = int(rng.integers(0, len(b) - 1))
            parts.append(b[lo:lo + int(rng.integers(1, 60000))])
        elif kind == 4:
            parts.append(bytes(rng.integers(1, 256, size=int(rng.integers(1, 3000)), dtype=np.uint8)).replace(b"\n", b" "))
        elif kind == 5:
            c = bytes([int(rng.choice(list(b"a0E-x \xce\xb1e9")))])
            parts.append(c * int(rng.choice([3, 100, 600, 5000, 40000])))
        elif kind == 6:
            parts.append(b"\n" * int(rng.integers(1, 5)))
        elif kind == 7:
            b = blocks[int(rng.integers(0, len(blocks)))]
            lo = int(rng.integers(0, len(b) - 1))
            parts.append(b[lo:lo + int(rng.integers(1, 30000))].replace(b"\n", b" "))   # one long line
        else:
            parts.append(rng.choice([b"ERROR", b"WARN 555-1234", b"Running", b"the", b"Sherlock Holmes", b"id=7", "naïve".encode()]))
    data = b"".join(parts)
    if rng.random() < 0.5 and not data.endswith(b"\n"):
        data += b"\n"
    return data


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    blocks = [corpus.block(c, 300000).tobytes() for c in ("c1", "c2", "c3", "c4", "c5")]
    sc = api.Scanner(0)
    names = G.pattern_names()
    pats = {}
    t0 = time.time()
    n = bad = 0
    kernels = {}
    while time.time() - t0 < budget:
        name = names[int(rng.integers(0, len(names)))]
        if name not in pats:
            pats[name] = (api.Pattern.load(G.pattern_path(name), 0), O.OraclePattern(G.pattern_path(name)))
        pat, op = pats[name]
        data = make_text(rng, blocks)
        if os.environ.get("UGX_FUZZ_TRACE"):
            print("case %d %s %d bytes t=%.1f" % (n, name, len(data), time.time() - t0), flush=True)
        want = op.find_all(data)
        rec, tot = sc.find_all(pat, data)
        kernels[tot.kernel] = kernels.get(tot.kernel, 0) + 1
        ok = len(rec) == len(want) and bool(np.all(rec == want))
        ok = ok and sc.count_matches(pat, data).matches == len(want)
        ok = ok and sc.count_lines(pat, data).matches == op.count_lines(data)
        n += 1
        if not ok:
            bad += 1
            path = os.path.join(ROOT, "gpurun_out", "fuzz_fail_%s_%d.bin" % (name, n))
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "wb") as f:
                f.write(data)
            print("MISMATCH", name, len(data), "records", len(rec), len(want), "saved", path)
            if bad >= 5:
                break
    print("fuzz: %d cases, %d mismatches, kernels %r, seed %d" % (n, bad, kernels, seed))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
