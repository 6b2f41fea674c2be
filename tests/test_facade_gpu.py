"""The C++ mirror of reflex::Matcher (include/ugrep_b200/matcher.hpp) on the GPU: the reference's caller loops
(find / skip('\\n') / lineno() / first() / size()) replayed over device-produced records, against the golden
outputs of the unmodified reference."""
import base64
import os
import shutil
import subprocess

import pytest

import golden_lib as G
import oracle_lib as O

pytestmark = pytest.mark.gpu
ROOT = O.ROOT
LIB = os.path.join(ROOT, "ugrep_b200", "libugrep_b200.so")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available on this box")
    out = str(tmp_path_factory.mktemp("facade") / "facade_test")
    cmd = ["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), "-o", out,
           os.path.join(ROOT, "tests", "cpp", "facade_test.cpp"), "-L" + os.path.dirname(LIB), "-lugrep_b200",
           "-Wl,-rpath," + os.path.dirname(LIB)]
    for d in ("/usr/local/cuda/lib64",):
        if os.path.isdir(d):
            cmd += ["-L" + d, "-Wl,-rpath," + d]
    subprocess.run(cmd, check=True)
    return out


@pytest.mark.parametrize("name", ["c1", "c3b", "c3c", "c5", "hello_U", "wb", "alt3", "icase"])
def test_facade_loops_match_reference_goldens(exe, name, tmp_path):
    if name not in G.pattern_names():
        pytest.skip("no golden pattern %s" % name)
    op = O.OraclePattern(G.pattern_path(name))
    n = 0
    for case, data in G.cases(name):
        if len(data) > (1 << 20) or n >= 4:
            continue
        n += 1
        f = tmp_path / ("in%d.txt" % n)
        f.write_bytes(data)

        r = subprocess.run([exe, G.pattern_path(name), "all", str(f)], capture_output=True)
        assert r.returncode == 0, r.stderr
        head, text = r.stdout.split(b"\n", 3)[:3], r.stdout.split(b"\n", 3)[3]
        got = dict(h.split(b"=") for h in head)
        assert int(got[b"cl"]) == case["lines"], (name, case["input"])
        assert int(got[b"cm"]) == case["matches"], (name, case["input"])
        assert int(got[b"loop"]) == case["lines"], (name, case["input"], "find+skip loop")
        assert text == G.format_list(data, op.find_all(data))
        if "list" in case:
            assert text == base64.b64decode(case["list"])
