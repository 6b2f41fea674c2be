import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import oracle_lib as O
from ugrep_b200 import api, corpus
sc = api.Scanner(0)
path = "/root/repo/ugrep_b200/patterns/c3b.ugxp"
pat = api.Pattern.load(path, 0); op = O.OraclePattern(path)
block = corpus.block("c3", 32 << 20)
reps = 150   # ~4.7 GiB: offsets beyond 2^32
dev = torch.from_numpy(block).cuda().repeat(reps)
want = op.find_all(block)
tot = sc.find_all_device(pat, dev, base_offset=7, base_line=3)
assert tot.matches == reps * len(want), (tot.matches, reps * len(want))
nl = int((block == 10).sum())
for r in (0, 129, reps - 1):
    got = sc.fetch(r * len(want), len(want))
    assert bool(np.all(got["offset"] == want["offset"] + r * len(block) + 7)), r
    assert bool(np.all(got["line"] == want["line"] + r * nl + 3)), r
print("big offsets ok: n = %.2f GiB, records %d, last offset %d, kernel %s, %.1f ms" % (dev.numel() / 2**30, tot.matches, int(got["offset"][-1]), tot.kernel, tot.kernel_ms))
path = "/root/repo/ugrep_b200/patterns/c5.ugxp"
pat = api.Pattern.load(path, 0); op = O.OraclePattern(path)
block = corpus.block("c5", 32 << 20)
dev = torch.from_numpy(block).cuda().repeat(reps)
assert sc.count_matches(pat, dev).matches == reps * op.count_matches(block)
assert sc.count_newlines(dev).newlines == reps * int((block == 10).sum())
print("c5 4.7 GiB counts ok")
