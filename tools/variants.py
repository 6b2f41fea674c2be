#!/usr/bin/env python3
"""Build A/B variants of one kernel source with different -D flags (tuning sweeps on the GPU box).

    python tools/variants.py stream_count.cu v_spans2:-DUGX_SC_SPANS=2 v_minb4:-DUGX_SC_MINB=4 ...

NOTE: only macros that are private to that one source may be varied.  A macro shared through a header (for example
UGX_SC_REGION, which capi.cu and stream_count.cu read too) gives an inconsistent library — that hung a kernel once.

Each variant becomes ugrep_b200/build/<name>.so (all other objects are shared with the main build);
select one at run time with UGX_LIB=ugrep_b200/build/<name>.so.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ugrep_b200 import build as B  # noqa: E402


def main():
    src = sys.argv[1]
    variants = [a.split(":", 1) for a in sys.argv[2:]]
    B.build()
    objdir = os.path.join(B.HERE, "build")
    others = [os.path.join(objdir, os.path.splitext(s)[0] + ".o") for s in B.SOURCES if s != src]

    def one(v):
        name, flags = v
        obj = os.path.join(objdir, name + ".o")
        lib = os.path.join(objdir, name + ".so")
        cmd = [B.nvcc(), "-O3", "-std=c++17", "-lineinfo", *B.ARCH, "-Xcompiler", "-fPIC,-O2,-Wall", "-Xptxas", "-v",
               *flags.split(), "-c", "-o", obj, os.path.join(B.CSRC, src)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            return name, r.stdout + r.stderr
        regs = [ln for ln in (r.stdout + r.stderr).splitlines() if "registers" in ln or "spill" in ln]
        r2 = subprocess.run([B.nvcc(), *B.ARCH, "-shared", "-o", lib, obj, *others, "-lcudart"], capture_output=True, text=True)
        return name, "\n".join(regs[:4]) + r2.stdout + r2.stderr

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        for name, log in ex.map(one, variants):
            print("==", name)
            print(log)


if __name__ == "__main__":
    main()
