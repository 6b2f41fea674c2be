#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/ with the UNMODIFIED
reference (oracle/_ref/ugrep CLI and oracle/_ref/refscan over libreflex, both built
from /root/reference by oracle/Makefile).

    python tools/make_golden.py          (needs /root/reference; run in the build container)

Writes
    tests/golden/patterns/<name>.ugxp    compiled pattern (opcode words + prefilter fields)
    tests/golden/golden.json             per (pattern, input): what the reference prints for
                                         `ugrep -c`, `ugrep -c -o`, `ugrep -n -b -o`
Inputs are either literal byte strings (stored base64 in the JSON) or seeded corpus blocks
(ugrep_b200.corpus.block(name, nbytes); the JSON pins their sha256).

For every case the CLI (256 KiB streaming window) and the in-place library scan
(AbstractMatcher::buffer, the mode the GPU path mirrors) must agree, else generation fails.
"""
import base64
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ugrep_b200 import corpus  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
UGREP = os.path.join(REF, "ugrep")
REFSCAN = os.path.join(REF, "refscan")
OUT = os.path.join(ROOT, "tests", "golden")

# name -> pattern options as given to ugrep / refscan (pattern options only)
PATTERNS = {
    # the five BASELINE.json configs and the two config-3 companions
    "c1": ["-F", "-e", "Sherlock Holmes"],
    "c2": ["-F", "-f", "@WORDS@"],
    "c3": ["-e", "[A-Z][a-z]+ing\\b"],
    "c3b": ["-e", "[A-Z][a-z]+ing"],
    "c3c": ["-e", "[A-Z][a-z]{1,9}ing\\b"],
    "c4": ["-i", "-e", "\\p{Greek}+|naïve\\w*"],
    "c5": ["-e", "ERROR|WARN", "-e", "\\d{3}-\\d{4}"],
    # the reference's own test patterns (tests/verify.sh: -U 'Hello', 'nomatch')
    "hello_U": ["-U", "-e", "Hello"],
    "nomatch_U": ["-U", "-e", "nomatch"],
    "hello_Uw": ["-U", "-w", "-e", "Hello"],
    "hello_F": ["-F", "-e", "Hello"],
    # one per prefilter routine family / DFA feature
    "char1": ["-F", "-e", "e"],
    "char2": ["-F", "-e", "th"],
    "char3": ["-F", "-e", "the"],
    "str4": ["-F", "-e", "that"],
    "str_cls": ["-e", "the[a-z]"],
    "str_long": ["-e", "Sherlock Hol[a-z]+"],
    "alt3": ["-F", "-e", "ERROR", "-e", "WARN", "-e", "INFO"],
    "alt_words": ["-e", "morning|Running|Walking|Thinking"],
    "icase": ["-i", "-e", "error"],
    "opt": ["-e", "colou?r|th[ae]n"],
    "dotstar": ["-e", "a.*b"],
    "neg": ["-e", "[^ ]+ing"],
    "space": ["-e", "\\s+"],
    "email": ["-e", "\\w+@\\w+\\.com"],
    "digits_U": ["-U", "-e", "\\d{3}-\\d{4}"],
    "bol": ["-e", "^ERROR"],
    "bol2": ["-e", "^[0-9T:-]+ ERROR"],
    "eol": ["-e", "id=[0-9]+$"],
    "wb": ["-e", "\\bthe\\b"],
    "bw": ["-e", "\\<th[a-z]*\\>"],
    "w_the": ["-w", "-e", "the"],
    "w_greek": ["-w", "-e", "\\p{Greek}+"],
    "greek1": ["-e", "\\p{Greek}"],
    "upper_word": ["-e", "[A-Z][a-z]+"],
    "two_caps": ["-e", "[A-Z][a-z]+ [A-Z][a-z]+"],
    "hex": ["-U", "-e", "[0-9a-f]{4,}"],
    # remaining routine families (probe: tools/make_golden.py prints the selected fields)
    "min2": ["-e", "e[a-z]"],
    "min4": ["-e", "e[a-z]{4}"],
    "str_pmh": ["-e", "the[a-z]{4}"],
    "pin_pma": ["-e", "[a-z]{2}[0-9]"],
    "pin_pma_lb": ["-e", "s[a-z]*ing"],
    "pin1_pma": ["-e", "k[a-z]{2}"],
    "pin1_pmh": ["-e", "a[a-z]c[a-z]"],
    "pin1_one_lb": ["-e", "[a-z]+@"],
    "pin_one": ["-e", "ERROR|W"],
    "min1_uni": ["-e", "\\d"],
    "min3": ["-e", "e[a-z][a-z]"],
    # matcher option N (ugrep -Y, switched on by CNF::anchor for ^... / ...$ patterns, src/cnf.hpp:200-204):
    # empty matches are reported (lib/matcher.cpp:681-728); min_ == 0 then means no prefilter at all (:804)
    "empty_line": ["-e", "^$"],
    "bol_only": ["-e", "^"],
    "eol_only": ["-e", "$"],
    "xstar": ["-e", "x*"],
    "xstar_Y": ["-Y", "-e", "x*"],
    "opt_eol": ["-e", "[0-9]*$"],
    "bol_xstar": ["-e", "^x*"],
    "hello_Y": ["-Y", "-U", "-e", "Hello"],
    "the_Y": ["-Y", "-e", "th*e"],
    "cr_eol": ["-e", "e$"],
}

# Patterns pinned by the in-place LIBRARY scan only (refscan over reflex::Matcher with AbstractMatcher::buffer, the
# mode the GPU path mirrors).  For these the CLI's output is not the matcher's: `ugrep -o` prints the whole line for
# an empty match and the CLI's streaming window sees `$` at the end of an unterminated last line differently from
# the one-pass in-place buffer (measured: tools/make_golden.py refuses a CLI/library disagreement for every other
# pattern).  The caller loops are the same three loops of Grep::search (oracle/refscan.cpp scan()).
LIB_ONLY = {"empty_line", "bol_only", "eol_only", "xstar_Y"}

EDGE = [
    b"",
    b"\n",
    b"\n\n\n",
    b"Hello",
    b"Hello\n",
    b"Sherlock Holmes",
    b"xSherlock HolmesSherlock Holmes Sherlock Holme\nSherlock Holmes",
    b"Running\nWalking and Thinking\n\nmorning Singing",
    b"a" * 70000 + b" Running the " + b"b" * 70000 + b"\n" + b"Walking",
    b"2026-01-02T03:04:05 ERROR svc01 call 555-1234 ext 123-4567 id=1\n" * 3 + b"WARN 111-2222",
    "naïve Ωmega αβγ NAÏVE\nκόσμος the\n".encode("utf-8"),
    b"\r\nWalking\r\nthe\r\n",
    b"the the the thethe then than colour color a@b.com xx@yy.com\n" * 40,
    b"ERROR\nERROR x\n ERROR\nxERROR id=7\nid=12 \nid=3",
    b"Hello World\nhello Hello_ Hello9 (Hello) Hello\n\nHelloHello\n",
    b"abc\nxxx\n\nyxxy 12\n\n\n5\nthe\r\n\r\nthhe",
    b"x",
]

# seeded corpus blocks: (corpus name, bytes)
BLOCKS = [("c1", 96 << 10), ("c2", 96 << 10), ("c2s", 96 << 10), ("c3", 96 << 10), ("c4", 96 << 10), ("c5", 96 << 10)]
# which blocks a pattern is run on (every pattern runs on every edge input)
BLOCKS_FOR = {
    "c1": ["c1", "c3"], "c2": ["c2", "c2s"], "c3": ["c3", "c1"], "c3b": ["c3", "c1"], "c3c": ["c3", "c1"], "c4": ["c4"],
    "c5": ["c5"], "icase": ["c5"], "alt3": ["c5"], "digits_U": ["c5"], "bol": ["c5"], "bol2": ["c5"], "eol": ["c5"],
    "hex": ["c5"], "pin_pma": ["c5", "c1"], "pin_one": ["c5"], "min1_uni": ["c5"], "opt_eol": ["c5"], "empty_line": ["c1"], "bol_only": ["c1"], "eol_only": ["c1"], "xstar": ["c1"],
    "xstar_Y": ["c1"], "bol_xstar": ["c1"], "pin1_one_lb": ["c4", "c1"], "w_greek": ["c4"], "greek1": ["c4"], "email": ["c4", "c1"],
}
DEFAULT_BLOCKS = ["c1", "c3"]


def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, **kw)


def reference_outputs(popts, data: bytes, d: str, lib_only: bool = False):
    path = os.path.join(d, "input.txt")
    with open(path, "wb") as f:
        f.write(data)
    res = {}
    for mode, flags, smode in (("lines", ["-c"], "cl"), ("matches", ["-c", "-o"], "cm"), ("list", ["-n", "-b", "-o"], "list")):
        cli = run([UGREP, "--no-config", *flags, *popts, path])
        if cli.returncode not in (0, 1):
            raise SystemExit("ugrep failed: %r %s" % (popts, cli.stderr[:300]))
        lib = run([REFSCAN, "scan", smode, *popts, path])
        if lib.returncode not in (0, 1):
            raise SystemExit("refscan failed: %r %s" % (popts, lib.stderr[:300]))
        if lib_only:
            res[mode] = lib.stdout
            continue
        if cli.stdout != lib.stdout:
            raise SystemExit("CLI and in-place library scan disagree for %r mode %s (%d vs %d bytes)"
                             % (popts, mode, len(cli.stdout), len(lib.stdout)))
        res[mode] = cli.stdout
    return res


def main():
    os.makedirs(os.path.join(OUT, "patterns"), exist_ok=True)
    words = os.path.join(ROOT, "ugrep_b200", "patterns", "words.txt")
    blocks = {name: corpus.block(name, nbytes).tobytes() for name, nbytes in BLOCKS}
    golden = {"generator": "tools/make_golden.py", "reference": "GerHobbelt/ugrep 7.4.2 (oracle/_ref/ugrep, oracle/_ref/refscan)",
              "edge_inputs": [base64.b64encode(e).decode() for e in EDGE],
              "blocks": {name: {"nbytes": nbytes, "len": len(blocks[name]),
                                "sha256": hashlib.sha256(blocks[name]).hexdigest()} for name, nbytes in BLOCKS},
              "patterns": {}}
    with tempfile.TemporaryDirectory() as d:
        for name, popts in PATTERNS.items():
            popts = [words if p == "@WORDS@" else p for p in popts]
            out = os.path.join(OUT, "patterns", name + ".ugxp")
            r = run([REFSCAN, "dump", *popts, "-o", out], text=True)
            if r.returncode != 0:
                raise SystemExit("refscan dump failed for %s: %s" % (name, r.stderr))
            entry = {"popts": [("@WORDS@" if p == words else p) for p in popts],
                     "fields": r.stderr.strip().replace("refscan: ", ""), "cases": []}
            if name in LIB_ONLY:
                entry["pinned_by"] = "library scan (refscan) only"

            inputs = [("edge", i, e) for i, e in enumerate(EDGE)]
            inputs += [("block", b, blocks[b]) for b in BLOCKS_FOR.get(name, DEFAULT_BLOCKS)]
            for kind, key, data in inputs:
                res = reference_outputs(popts, data, d, name in LIB_ONLY)
                lst = res["list"]
                case = {"input": [kind, key],
                        "lines": int(res["lines"].strip() or 0),
                        "matches": int(res["matches"].strip() or 0),
                        "list_sha256": hashlib.sha256(lst).hexdigest(),
                        "list_bytes": len(lst)}
                if len(lst) <= 600:
                    case["list"] = base64.b64encode(lst).decode()
                entry["cases"].append(case)
            golden["patterns"][name] = entry
            print("%-10s %s  cases=%d" % (name, entry["fields"], len(entry["cases"])))
    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(OUT, "golden.json"))


if __name__ == "__main__":
    main()
