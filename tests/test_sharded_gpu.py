"""ugx_sharded_*: one process, several devices (here: the box's device 0 named several times, and every visible
device when there are more) — line-aligned cuts, per-shard scans, host-summed bases — against the oracle's single
sequential scan of the whole buffer; through ctypes and through a plain C++ caller."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from ugrep_b200 import corpus

pytestmark = pytest.mark.gpu
PAT_DIR = os.path.join(O.ROOT, "ugrep_b200", "patterns")
LIB = os.path.join(O.ROOT, "ugrep_b200", "libugrep_b200.so")


def device_lists():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the -m gpu tests must run on the B200 box")
    n = torch.cuda.device_count()
    lists = [[0], [0, 0], [0, 0, 0, 0, 0]]
    if n > 1:
        lists.append(list(range(n)))
    return lists


@pytest.mark.parametrize("pname,cname", [("c5", "c5"), ("c3b", "c3"), ("c2", "c2"), ("c1", "c1")])
def test_sharded_scan_equals_the_single_scan(pname, cname):
    from ugrep_b200 import api
    path = os.path.join(PAT_DIR, pname + ".ugxp")
    op = O.OraclePattern(path)
    data = corpus.block(cname, 3 << 20)
    want_rec = op.find_all(data)
    want_lines = op.count_lines(data)
    nl = int((data == 10).sum())
    for devs in device_lists():
        sh = api.Sharded(path, devs)
        tot, _, info = sh.scan(data, "lines")
        assert tot.matches == want_lines and tot.newlines == nl, (pname, devs)
        assert [i["begin"] for i in info][0] == 0 and info[-1]["end"] == len(data)
        for a, b in zip(info, info[1:]):
            assert a["end"] == b["begin"] and (b["begin"] == 0 or data[b["begin"] - 1] == 10)
        tot, _, _ = sh.scan(data, "matches")
        assert tot.matches == len(want_rec), (pname, devs)
        tot, rec, info = sh.scan(data, "records", cap=len(want_rec) + 10)
        assert len(rec) == len(want_rec) and bool(np.all(rec == want_rec)), (pname, devs)
        assert sum(i["matches"] for i in info) == len(want_rec)
        assert [i["record_base"] for i in info] == list(np.cumsum([0] + [i["matches"] for i in info[:-1]]))
        with pytest.raises(api.UgxError):
            sh.scan(data, "records", cap=max(0, len(want_rec) - 1))
        sh.close()


def test_sharded_degenerate_inputs():
    from ugrep_b200 import api
    path = os.path.join(PAT_DIR, "c5.ugxp")
    op = O.OraclePattern(path)
    sh = api.Sharded(path, [0, 0, 0])
    for data in (b"", b"\n", b"ERROR", b"x" * 1000, b"WARN\n" * 3, b"a\nERROR 555-1234"):
        tot, rec, _ = sh.scan(data, "records", cap=64)
        want = op.find_all(data)
        assert len(rec) == len(want) and bool(np.all(rec == want)), data
        assert sh.scan(data, "lines")[0].matches == op.count_lines(data), data
    sh.set_option("pin", 1)
    big = corpus.block("c5", 1 << 20)
    assert sh.scan(big, "matches")[0].matches == op.count_matches(big)


def test_sharded_from_a_plain_cpp_caller(tmp_path):
    """no Python, no torch.distributed: a C++ program links the library and shards a file over the devices it names"""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available on this box")
    exe = str(tmp_path / "sharded_test")
    cmd = ["g++", "-std=c++17", "-O1", "-I" + os.path.join(O.ROOT, "include"), "-o", exe,
           os.path.join(O.ROOT, "tests", "cpp", "sharded_test.cpp"), "-L" + os.path.dirname(LIB), "-lugrep_b200",
           "-Wl,-rpath," + os.path.dirname(LIB)]
    subprocess.run(cmd, check=True)
    path = os.path.join(PAT_DIR, "c5.ugxp")
    op = O.OraclePattern(path)
    data = corpus.block("c5", 2 << 20)
    f = tmp_path / "in.log"
    f.write_bytes(data.tobytes())
    r = subprocess.run([exe, path, str(f), "0,0,0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = dict(kv.split("=") for kv in r.stdout.split())
    want = op.find_all(data)
    assert int(got["lines"]) == op.count_lines(data)
    assert int(got["matches"]) == len(want) == int(got["records"])
    assert int(got["offset_sum"]) == int(want["offset"].sum()) and int(got["line_sum"]) == int(want["line"].sum())
