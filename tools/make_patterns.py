#!/usr/bin/env python3
"""Compile the BASELINE.json config patterns with the UNMODIFIED reference
(oracle/_ref/refscan, built from /root/reference by oracle/Makefile) and store
the compiled form (opcode words + prefilter fields, UGXP container) under
ugrep_b200/patterns/.  The pattern compiler (reflex::Pattern::init,
lib/pattern.cpp:171-4639) is out of this path's scope: the
scan path starts from the compiled tables, exactly as reflex::Matcher does.

Run here (needs /root/reference):  python tools/make_patterns.py
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ugrep_b200 import corpus  # noqa: E402

REFSCAN = os.path.join(ROOT, "oracle", "_ref", "refscan")
OUT = os.path.join(ROOT, "ugrep_b200", "patterns")

# name -> (refscan pattern options, the ugrep command line it stands for)
CONFIGS = {
    "c1": (["-F", "-e", "Sherlock Holmes"], "ugrep -c -F 'Sherlock Holmes'"),
    "c2": (["-F", "-f", "@WORDS@"], "ugrep -c -F -f words.txt"),
    "c3": (["-e", "[A-Z][a-z]+ing\\b"], "ugrep -n -b -o '[A-Z][a-z]+ing\\b'"),
    "c3b": (["-e", "[A-Z][a-z]+ing"], "ugrep -n -b -o '[A-Z][a-z]+ing'  (companion: no \\b)"),
    "c3c": (["-e", "[A-Z][a-z]{1,9}ing\\b"], "ugrep -n -b -o '[A-Z][a-z]{1,9}ing\\b'  (companion: bounded)"),
    "c4": (["-i", "-e", "\\p{Greek}+|naïve\\w*"], "ugrep -i -c '\\p{Greek}+|naïve\\w*'"),
    "c5": (["-e", "ERROR|WARN", "-e", "\\d{3}-\\d{4}"], "ugrep -c -o -e 'ERROR|WARN' -e '\\d{3}-\\d{4}'"),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    words = os.path.join(OUT, "words.txt")
    with open(words, "wb") as f:
        f.write(b"\n".join(corpus.words_list()) + b"\n")
    with open(os.path.join(OUT, "README.txt"), "w") as readme:
        readme.write("Compiled patterns (UGXP, include/ugrep_b200.h) written by tools/make_patterns.py\n"
                     "with the unmodified reference's pattern compiler.\n\n")
        for name, (popts, cmd) in CONFIGS.items():
            popts = [words if p == "@WORDS@" else p for p in popts]
            out = os.path.join(OUT, name + ".ugxp")
            r = subprocess.run([REFSCAN, "dump", *popts, "-o", out], capture_output=True, text=True)
            if r.returncode != 0:
                raise SystemExit("refscan failed for %s: %s" % (name, r.stderr))
            readme.write("%-4s %s\n     %s\n" % (name, cmd, r.stderr.strip().replace("refscan: ", "")))
            print(name, r.stderr.strip())


if __name__ == "__main__":
    main()
