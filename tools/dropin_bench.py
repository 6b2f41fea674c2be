#!/usr/bin/env python3
"""Wall-clock of the drop-in CLI (integration/_build/ugrep-b200: the reference's CLI with B200Matcher) next to the
unmodified reference CLI (oracle/_ref/ugrep) on the same file in /dev/shm — the application-level view of the path.

    python tools/dropin_bench.py [--gib 1]          (on the GPU box)
"""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ugrep_b200 import corpus  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ugrep")
NEW = os.path.join(ROOT, "integration", "_build", "ugrep-b200")
WORDS = os.path.join(ROOT, "ugrep_b200", "patterns", "words.txt")

CASES = [
    ("c1  -c -F 'Sherlock Holmes'", "c1", ["-c", "-F", "Sherlock Holmes"]),
    ("c2s -c -F -f words.txt", "c2s", ["-c", "-F", "-f", WORDS]),
    ("c2  -c -F -f words.txt", "c2", ["-c", "-F", "-f", WORDS]),
    ("c5  -c -o ERROR|WARN|\\d{3}-\\d{4}", "c5", ["-c", "-o", "-e", "ERROR|WARN", "-e", r"\d{3}-\d{4}"]),
    ("c1  -n 'Sherlock Holmes' (lines printed)", "c1", ["-n", "-F", "Sherlock Holmes"]),
]


def run(binary, args, path, env):
    t = time.time()
    r = subprocess.run([binary, *args, path], capture_output=True, env=env)
    return time.time() - t, r.returncode, r.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=1.0)
    ap.add_argument("--verbose", action="store_true", help="print the drop-in's own timing lines (UGREP_B200_VERBOSE)")
    a = ap.parse_args()
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "ugrep_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""),
               UGREP_B200_REQUIRE="1")
    made = {}
    for label, cname, args in CASES:
        if cname not in made:
            block = corpus.block(cname, 64 << 20)
            path = "/dev/shm/ugx_dropin_%s.txt" % cname
            with open(path, "wb") as f:
                for _ in range(max(1, int(a.gib * (1 << 30)) // block.size)):
                    f.write(block.tobytes())
            made[cname] = path
        path = made[cname]
        size = os.path.getsize(path)
        run(NEW, args, path, env)  # warm: CUDA context creation is part of every CLI start, page cache is not
        t_new, rc_new, out_new = run(NEW, args, path, env)
        if a.verbose:
            r = subprocess.run([NEW, *args, path], capture_output=True, env=dict(env, UGREP_B200_VERBOSE="1"))
            sys.stdout.write(r.stderr.decode("utf-8", "replace"))
        t_ref, rc_ref, out_ref = run(REF, args, path, env)
        same = rc_new == rc_ref and out_new == out_ref
        print("%-44s %5.2f GiB  reference %6.2f s (%5.2f GB/s)  drop-in %6.2f s (%5.2f GB/s)  x%.1f  output %s"
              % (label, size / (1 << 30), t_ref, size / t_ref / 1e9, t_new, size / t_new / 1e9, t_ref / t_new,
                 "identical" if same else "DIFFERENT"))
    for p in made.values():
        os.unlink(p)


if __name__ == "__main__":
    main()
