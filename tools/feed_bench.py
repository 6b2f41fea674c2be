import sys, time, os
sys.path.insert(0,'/root/repo')
import numpy as np, torch
from ugrep_b200 import api, corpus
blk=corpus.block('c2', 64<<20); host=np.tile(blk, 32)   # ~2.5 GiB pageable
pat=api.Pattern.load('/root/repo/ugrep_b200/patterns/c2.ugxp',0); sc=api.Scanner(0)
for opt in (0,1):
    sc.set_option('no_feeder',opt)
    sc.count_lines(pat,host)
    t0=time.perf_counter(); n=3
    for _ in range(n): r=sc.count_lines(pat,host)
    dt=(time.perf_counter()-t0)/n
    print('no_feeder=%d: %.1f GB/s e2e from pageable memory (%d matches)'%(opt, host.size/dt/1e9, r.matches))
