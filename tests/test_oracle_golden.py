"""Pins the oracle (oracle/oracle.c, the plain-C restatement of the reference's buffer-scan
path) on the CPU:

  1. against the committed golden fixtures made with the unmodified reference
     (tests/golden/golden.json, tools/make_golden.py): `ugrep -c`, `ugrep -c -o`,
     `ugrep -n -b -o` on 47 patterns x edge inputs + seeded corpus blocks;
  2. against the reference's OWN golden files for this path (tests/out/Hello_*-c.out, -co.out,
     -on.out, -onkbT.out), read from /root/reference when it is present;
  3. against the unmodified reference library run live (oracle/_ref/refscan) on random
     slices, when it is present.
"""
import os
import re

import numpy as np
import pytest

import golden_lib as G
import oracle_lib as O

REF_TESTS = "/root/reference/tests"


@pytest.mark.parametrize("name", G.pattern_names())
def test_oracle_vs_committed_golden(name):
    op = O.OraclePattern(G.pattern_path(name))
    for case, data in G.cases(name):
        assert op.count_lines(data) == case["lines"], (name, case["input"], "lines")
        assert op.count_matches(data) == case["matches"], (name, case["input"], "matches")
        rec = op.find_all(data)
        assert len(rec) == case["matches"]
        G.check_list(case, data, rec)


def test_golden_covers_every_prefilter_family():
    """the pattern suite exercises (nearly) every routine family of Matcher::init_advance"""
    seen = {O.OraclePattern(G.pattern_path(n)).advance for n in G.pattern_names()}
    names = {0: "none", 1: "pin1_one", 2: "pin1_pma", 3: "pin1_pmh", 4: "pin_one", 5: "pin_pma", 6: "pin_pmh", 7: "min1",
             8: "min2", 9: "min3", 10: "min4", 11: "pma", 12: "char", 13: "char_pma", 14: "char_pmh", 15: "string",
             16: "string_pma", 17: "string_pmh"}
    missing = sorted(names[k] for k in names if k not in seen)
    # char_pma / char_pmh (len_ == 1 with min_ > 0) cannot be produced by the reference's compiler in SIMD builds: a
    # one-byte literal prefix is kept only when the state after it accepts, and then nothing follows for the predictor
    # (min_ = 0, advance_char); otherwise len_ is reset to 0 (lib/pattern.cpp:4328-4334).  Every other family is covered.
    assert set(missing) == {"char_pma", "char_pmh"}, missing


SGR = re.compile(rb"\x1b\[[0-9;]*m")
HELLO_FILES = ["Hello.bat", "Hello.class", "Hello.java", "Hello.pdf", "Hello.sh", "Hello.txt", "empty.txt", "emptyline.txt"]


def _ref_golden(name):
    with open(os.path.join(REF_TESTS, "out", name), "rb") as f:
        return [SGR.sub(b"", ln) for ln in f.read().split(b"\n") if ln]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_TESTS, "out")), reason="/root/reference is not present")
@pytest.mark.parametrize("pat,stem", [("hello_U", "Hello_Hello"), ("nomatch_U", "Hello_nomatch")])
def test_oracle_vs_reference_own_goldens(pat, stem):
    """tests/verify.sh: `ugrep --color=always --sort -U OPS PAT Hello.* empty.txt emptyline.txt`"""
    op = O.OraclePattern(G.pattern_path(pat))
    data = {}
    for fn in HELLO_FILES:
        with open(os.path.join(REF_TESTS, fn), "rb") as f:
            data[fn] = f.read()
    # -c: file:count for every file
    want = dict(ln.rsplit(b":", 1) for ln in _ref_golden(stem + "-c.out"))
    assert {fn.encode(): b"%d" % op.count_lines(data[fn]) for fn in HELLO_FILES} == want
    # -co: file:matches
    want = dict(ln.rsplit(b":", 1) for ln in _ref_golden(stem + "-co.out"))
    assert {fn.encode(): b"%d" % op.count_matches(data[fn]) for fn in HELLO_FILES} == want
    # -on: file:line:text (text files; binary files print "Binary file ... matches")
    # -onkbT: file:line:column:offset:<TAB>text
    lines_on = _ref_golden(stem + "-on.out")
    lines_kb = _ref_golden(stem + "-onkbT.out")
    for fn in HELLO_FILES:
        rec = op.find_all(data[fn])
        mine_on, mine_kb = [], []
        for r in rec:
            o, n = int(r["offset"]), int(r["len"])
            bol = data[fn].rfind(b"\n", 0, o) + 1
            mine_on.append(b"%s:%d:%s" % (fn.encode(), r["line"], data[fn][o:o + n]))
            mine_kb.append(b"%s:%6d:%3d:%7d:\t%s" % (fn.encode(), r["line"], o - bol + 1, o, data[fn][o:o + n]))
        got_on = [ln for ln in lines_on if ln.startswith(fn.encode() + b":")]
        got_kb = [ln for ln in lines_kb if ln.startswith(fn.encode() + b":")]
        binary = any(b"Binary file " + fn.encode() in ln for ln in lines_on)
        if binary:
            assert len(rec) > 0
            continue
        assert mine_on == got_on, fn
        assert mine_kb == got_kb, fn


@pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref (the built reference) is not present")
@pytest.mark.parametrize("name", ["c1", "c2", "c3b", "c3c", "c4", "c5", "email", "two_caps", "bw", "eol", "space", "dotstar"])
def test_oracle_vs_live_reference_on_random_slices(name):
    """same pattern, random line-UNaligned slices of a corpus block: exercises the end-of-buffer rules"""
    entry = G.golden()["patterns"][name]
    popts = [os.path.join(O.ROOT, "ugrep_b200", "patterns", "words.txt") if p == "@WORDS@" else p for p in entry["popts"]]
    op = O.OraclePattern(G.pattern_path(name))
    block = G.input_bytes(entry["cases"][-1]["input"])
    rng = np.random.default_rng(5)
    for _ in range(6):
        a = int(rng.integers(0, len(block) - 5000))
        n = int(rng.integers(1, 5000))
        data = block[a:a + n]
        rc, out = O.ref_scan("list", popts, data)
        rec = op.find_all(data)
        assert G.format_list(data, rec) == out, (name, a, n)
        rc, out = O.ref_scan("cl", popts, data)
        assert b"%d\n" % op.count_lines(data) == out, (name, a, n)
