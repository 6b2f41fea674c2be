#!/usr/bin/env python3
"""tests/golden/words_icase.ugxp: what the unmodified reference compiles for `-F -i -f LIST` (oracle/_ref/refscan dump),
for the CPU test of ugx_compile_words_ex(UGX_COMPILE_ICASE).  The list is WORDS below.

    python tools/make_icase_golden.py            (needs the built reference, oracle/_ref)
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORDS = [b"Error", b"WARN", b"fatal", b"Sherlock Holmes", b"na\xc3\xafve", b"id=7", b"timeout", b"TimeOut2", b"x"]


def main():
    out = os.path.join(ROOT, "tests", "golden", "words_icase.ugxp")
    with tempfile.TemporaryDirectory() as d:
        wf = os.path.join(d, "w.txt")
        with open(wf, "wb") as f:
            f.write(b"\n".join(WORDS) + b"\n")
        r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "refscan"), "dump", "-F", "-i", "-f", wf, "-o", out],
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.exit("refscan dump failed: " + r.stderr)
    print("wrote", out, r.stderr.strip())


if __name__ == "__main__":
    main()
