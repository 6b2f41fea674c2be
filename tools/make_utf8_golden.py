#!/usr/bin/env python3
"""Golden vectors for reflex::isutf8 (lib/simd.cpp:169-421; ugrep's binary-file test is !isutf8,
src/ugrep.cpp:699-711), made with the UNMODIFIED reference through oracle/_ref/refscan isutf8.

    python tools/make_utf8_golden.py        (needs /root/reference built into oracle/_ref)

Writes tests/golden/utf8.json: base64 inputs + the reference's verdict.  Inputs: every class of invalid byte, cut-off
sequences at every alignment of the SIMD loops (16 / 32 bytes), overlong forms and surrogates (which the reference
lets pass), NUL, valid mixed text, and random corruptions of it."""
import base64
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ugrep_b200 import corpus  # noqa: E402

REFSCAN = os.path.join(ROOT, "oracle", "_ref", "refscan")


def inputs():
    out = [b"", b"a", b"hello world\n", b"\x00", b"abc\x00def", b"a" * 100 + b"\x00", b"\x00" + b"a" * 100]
    seqs = [b"\xc2\x80", b"\xdf\xbf", b"\xe0\xa0\x80", b"\xef\xbf\xbf", b"\xf0\x90\x80\x80", b"\xf4\x8f\xbf\xbf",
            b"\xe0\x80\x80", b"\xed\xa0\x80", b"\xf0\x80\x80\x80", b"\xf4\x90\x80\x80",       # overlong / surrogate / > 10FFFF
            b"\xc0\x80", b"\xc1\xbf", b"\xf5\x80\x80\x80", b"\xff", b"\xfe", b"\x80", b"\xbf",  # bytes that never occur
            b"\xc2", b"\xe0\xa0", b"\xf0\x90\x80", b"\xc2\x41", b"\xe0\x41\x80", b"\xe0\xa0\x41", b"\xf0\x90\x41\x80",
            b"\xc2\x80\x80", b"\xe0\xa0\x80\x80", b"\xf0\x90\x80\x80\x80", b"\xc2\xc2\x80", b"\xe0\xc2\x80"]
    for s in seqs:
        for pad in (0, 1, 13, 14, 15, 16, 17, 29, 30, 31, 32, 33, 47, 63, 64, 65, 100):
            out.append(b"x" * pad + s)
            out.append(b"x" * pad + s + b"y" * 40)
            out.append("é".encode() * 20 + b"x" * pad + s + b"tail")
    text = corpus.block("c4", 6000).tobytes()
    out.append(text)
    for n in list(range(0, 70)) + [255, 256, 257, 1023, 1024, 1025, 4095, 4096, 4097]:
        out.append(text[:n])
    rng = np.random.default_rng(1)
    for _ in range(300):
        a = bytearray(text[:int(rng.integers(40, 3000))])
        for _ in range(int(rng.integers(1, 3))):
            a[int(rng.integers(0, len(a)))] = int(rng.integers(0, 256))
        out.append(bytes(a))
    return out


def main():
    ins = inputs()
    verdicts = []
    with tempfile.TemporaryDirectory() as d:
        paths = []
        for i, data in enumerate(ins):
            p = os.path.join(d, "%05d" % i)
            with open(p, "wb") as f:
                f.write(data)
            paths.append(p)
        for i in range(0, len(paths), 200):
            r = subprocess.run([REFSCAN, "isutf8", *paths[i:i + 200]], capture_output=True, text=True, check=True)
            verdicts += [int(x) for x in r.stdout.split()]
    assert len(verdicts) == len(ins)
    out = {"generator": "tools/make_utf8_golden.py", "reference": "reflex::isutf8, GerHobbelt/ugrep 7.4.2 (oracle/_ref/refscan isutf8)",
           "cases": [[base64.b64encode(d).decode(), v] for d, v in zip(ins, verdicts)]}
    with open(os.path.join(ROOT, "tests", "golden", "utf8.json"), "w") as f:
        json.dump(out, f)
    print("wrote %d cases, %d valid" % (len(ins), sum(verdicts)))


if __name__ == "__main__":
    main()
