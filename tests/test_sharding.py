"""Multi-GPU host logic on the CPU (gloo, world size 2): line-aligned sharding of a corpus, the one exchange
step of the path (all-gather of per-shard {matches, newlines}) and the line-number / byte-offset bases it
yields.  The per-shard scan is done by the oracle here (test infrastructure); on the GPU box bench.py does the
same exchange over NCCL with the CUDA scan."""
import os
import socket
import sys

import numpy as np
import pytest

import oracle_lib as O
from ugrep_b200 import corpus, sharding

PAT_DIR = os.path.join(O.ROOT, "ugrep_b200", "patterns")


def test_cut_points_are_line_aligned_and_cover_the_buffer():
    data = corpus.block("c5", 300000)
    for world in (1, 2, 3, 8):
        cuts = sharding.line_aligned_cuts(data, world)
        assert cuts[0] == 0 and cuts[-1] == len(data) and len(cuts) == world + 1
        assert all(a <= b for a, b in zip(cuts, cuts[1:]))
        for c in cuts[1:-1]:
            assert c == 0 or data[c - 1] == 10
    # degenerate inputs: no newline at all, empty
    one = np.frombuffer(b"x" * 1000, dtype=np.uint8)
    assert sharding.line_aligned_cuts(one, 4) == [0, 1000, 1000, 1000, 1000]
    assert sharding.line_aligned_cuts(np.zeros(0, dtype=np.uint8), 2) == [0, 0, 0]


def test_tiled_cuts_equal_the_cuts_of_the_materialised_corpus():
    """bench.py shards a logical corpus (one block repeated R times) without building it on the host: the cuts
    computed from the block alone, and the bytes materialised per shard, must equal the plain ones"""
    import torch
    block = corpus.block("c5", 50000)
    tblock = torch.from_numpy(block)
    for reps, world in ((1, 1), (3, 2), (9, 8), (17, 4), (8, 8), (5, 3)):
        data = np.tile(block, reps)
        cuts = sharding.tiled_cuts(block, reps, world)
        assert cuts == sharding.line_aligned_cuts(data, world), (reps, world)
        for r in range(world):
            got = sharding.materialize_tiled(tblock, cuts[r], cuts[r + 1]).numpy()
            assert got.tobytes() == data[cuts[r]:cuts[r + 1]].tobytes(), (reps, world, r)
    with pytest.raises(ValueError):
        sharding.tiled_cuts(np.frombuffer(b"no newline", dtype=np.uint8), 2, 2)


def test_bases_from_gathered_counts():
    bases = sharding.bases_from_counts([(5, 10), (0, 3), (7, 0)], [0, 100, 250, 300])
    assert bases == [(0, 0, 0), (5, 10, 100), (5, 13, 250)]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = corpus.block("c3", 400000)
        op = O.OraclePattern(os.path.join(PAT_DIR, "c3b.ugxp"))
        cuts = sharding.line_aligned_cuts(data, world)
        shard = data[cuts[rank]:cuts[rank + 1]]
        rec = op.find_all(shard)
        counts = sharding.all_gather_counts(len(rec), int((shard == 10).sum()))
        assert sharding.CountExchange("cpu")(len(rec), int((shard == 10).sum())) == counts  # the loop form of the same exchange
        bases = sharding.bases_from_counts(counts, cuts)
        first_match, base_line, base_off = bases[rank]
        rec = rec.copy()
        rec["line"] += base_line
        rec["offset"] += base_off
        q.put((rank, first_match, rec.tobytes(), sum(c[0] for c in counts)))
    finally:
        dist.destroy_process_group()


def test_two_rank_scan_equals_single_scan():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    data = corpus.block("c3", 400000)
    op = O.OraclePattern(os.path.join(PAT_DIR, "c3b.ugxp"))
    want = op.find_all(data)
    merged = b"".join(g[2] for g in got)
    assert merged == want.tobytes()
    assert got[0][3] == len(want) and got[1][1] == len(np.frombuffer(got[0][2], dtype=want.dtype))
