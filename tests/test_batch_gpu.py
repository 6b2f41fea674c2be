"""ugx_count_batch: many files in one launch, each keeping its own end of buffer — per-file counts
must equal the oracle's scan of every file ALONE (what `ugrep -c` / `ugrep -c -o` print per file)."""
import os

import numpy as np
import pytest

import golden_lib as G
import oracle_lib as O
from ugrep_b200 import corpus

pytestmark = pytest.mark.gpu
PAT_DIR = os.path.join(O.ROOT, "ugrep_b200", "patterns")


def make_files(cname, rng, count):
    base = corpus.block(cname, 400000).tobytes()
    files = [b"", b"\n", base[:1], base[:15], base[:16], base[:17]]
    for _ in range(count):
        lo = int(rng.integers(0, len(base) - 1))
        n = int(rng.choice([3, 40, 300, 2000, 16383, 16384, 16385, 40000, 70000]))
        files.append(base[lo:lo + n])       # cut anywhere: unterminated last lines, matches cut by the end of a file
    return files


@pytest.mark.parametrize("name,cname", [("c5", "c5"), ("c1", "c1"), ("c2", "c2"), ("c3b", "c3"), ("c4", "c4"), ("wb", "c1"),
                                        ("bol", "c5"), ("w_the", "c1"), ("empty_line", "c1")])
def test_batch_counts_equal_the_files_scanned_alone(name, cname):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the -m gpu tests must run on the B200 box")
    from ugrep_b200 import api
    path = os.path.join(PAT_DIR, name + ".ugxp")
    if not os.path.exists(path):
        path = G.pattern_path(name)
    pat = api.Pattern.load(path, 0)
    op = O.OraclePattern(path)
    sc = api.Scanner(0)
    files = make_files(cname, np.random.default_rng(3), 60)
    got, tot = sc.count_batch(pat, files, "lines")
    want = [op.count_lines(f) for f in files]
    assert list(map(int, got)) == want, name
    assert tot.matches == sum(want) and tot.launches == 1 and tot.kernel == "scan_batch_kernel"
    got, tot = sc.count_batch(pat, files, "matches")
    assert list(map(int, got)) == [op.count_matches(f) for f in files], name


def test_batch_of_a_source_tree_worth_of_files():
    """20 000 small files: one launch"""
    from ugrep_b200 import api
    path = os.path.join(PAT_DIR, "c5.ugxp")
    pat = api.Pattern.load(path, 0)
    op = O.OraclePattern(path)
    sc = api.Scanner(0)
    base = corpus.block("c5", 2 << 20).tobytes()
    rng = np.random.default_rng(8)
    files = []
    for _ in range(20000):
        lo = int(rng.integers(0, len(base) - 4096))
        files.append(base[lo:lo + int(rng.integers(0, 4096))])
    got, tot = sc.count_batch(pat, files, "matches")
    assert tot.launches == 1
    idx = rng.integers(0, len(files), size=300)
    for i in idx:
        assert int(got[i]) == op.count_matches(files[i]), i
    assert tot.matches == int(got.sum())
    assert sc.count_batch(pat, [], "lines")[1].matches == 0
