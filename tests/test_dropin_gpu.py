"""The drop-in proper: the REFERENCE's own CLI built with its matcher replaced by integration/b200matcher.hpp
(integration/_build/ugrep-b200: reference src/*.cpp + one expression of src/ugrep.cpp:8902 changed at build time),
run next to the unmodified reference CLI (oracle/_ref/ugrep) on the same files with the option matrix of the
reference's tests/verify.sh that reaches this path: every byte of stdout and the exit code must agree.
UGREP_B200_REQUIRE=1 makes a pattern the GPU library refuses a hard error, so no case passes on reflex::Matcher."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from ugrep_b200 import corpus

pytestmark = pytest.mark.gpu

B200 = os.path.join(O.ROOT, "integration", "_build", "ugrep-b200")
REF = O.REF_UGREP

needs_binaries = pytest.mark.skipif(not (os.access(B200, os.X_OK) and os.access(REF, os.X_OK)),
                                    reason="integration/_build/ugrep-b200 or oracle/_ref/ugrep did not travel")


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("dropin")
    out = {}
    texts = {
        "english.txt": corpus.block("c3", 150000).tobytes(),
        "logs.txt": corpus.block("c5", 150000).tobytes(),
        "greek.txt": corpus.block("c4", 100000).tobytes(),
        "hello.txt": b"Hello World\nhello Hello_ Hello9 (Hello) Hello\n\nHelloHello\nno match here\nHello",
        "crlf.txt": b"the\r\nWalking the\r\n\r\nthe end\r\n",
        "empty.txt": b"",
        "emptyline.txt": b"\n",
    }
    for name, data in texts.items():
        p = d / name
        p.write_bytes(data)
        out[name] = str(p)
    return out


def run(binary, args, env_extra=None):
    env = dict(os.environ)
    env.pop("GREP_COLORS", None)
    env.pop("GREP_COLOR", None)
    if env_extra:
        env.update(env_extra)
    r = subprocess.run([binary, "--no-config", "--color=never", *args], capture_output=True, env=env, timeout=300)
    return r.returncode, r.stdout, r.stderr


# (pattern options, files) — all in the GPU library's scope
PATTERNS = [
    (["Hello"], ["hello.txt", "english.txt", "empty.txt", "emptyline.txt"]),
    (["-F", "Hello"], ["hello.txt"]),
    (["-w", "Hello"], ["hello.txt"]),
    (["-i", "hello"], ["hello.txt"]),
    (["-w", "the"], ["english.txt", "crlf.txt"]),
    (["[A-Z][a-z]+ing"], ["english.txt"]),
    (["-e", "ERROR|WARN", "-e", "\\d{3}-\\d{4}"], ["logs.txt"]),
    (["^2026-0[1-3]"], ["logs.txt"]),
    (["id=[0-9]+$"], ["logs.txt"]),
    (["\\bthe\\b"], ["english.txt", "crlf.txt"]),
    (["-i", "\\p{Greek}+|naïve\\w*"], ["greek.txt"]),
    (["nomatch"], ["hello.txt", "english.txt"]),
]
# output modes of Grep::search (src/ugrep.cpp:10346-13276) that go through match(FIND)
MODES = [["-c"], ["-c", "-o"], ["-o"], ["-o", "-n"], ["-o", "-n", "-k", "-b", "-T"], ["-n"], [], ["-n", "-k", "-b", "-T"],
         ["-v", "-c"], ["-v", "-n"], ["-l"], ["-q"], ["-C2", "-n"], ["-A1"], ["-B1", "-n"], ["-m2", "-n"], ["-y", "-n"],
         ["--mmap", "-n", "-b", "-o"], ["--mmap", "-c"], ["-x", "-c"], ["-w", "-n"]]


def _cases():
    """every (pattern, file) pair with a rotating slice of the output modes: each process start costs a CUDA context,
    so the full cross product (it passes: 12 patterns x files x 19 modes, 15 minutes) is thinned to ~70 runs in which
    every mode still meets several patterns; UGX_DROPIN_FULL=1 runs the cross product"""
    full = os.environ.get("UGX_DROPIN_FULL", "") not in ("", "0")
    k = 0
    for popts, names in PATTERNS:
        for name in names:
            if full:
                modes = MODES
            else:
                modes = [MODES[(k + j * 7) % len(MODES)] for j in range(3)]
                k += 1
            for mode in modes:
                yield popts, name, mode


@needs_binaries
def test_dropin_equals_the_reference_cli(files):
    seen = set()
    for popts, name, mode in _cases():
        args = [*mode, *popts, files[name]]
        want = run(REF, args)
        got = run(B200, args, {"UGREP_B200_REQUIRE": "1"})
        assert got[0] == want[0], (args, got[2][:300])
        assert got[1] == want[1], (args, got[1][:200], want[1][:200])
        seen.add(tuple(mode))
    assert len(seen) == len(MODES)


@needs_binaries
def test_dropin_many_files_and_threads(files):
    """several files, worker threads with cloned matchers (src/ugrep.cpp:4204-4215), --sort for a fixed order"""
    paths = [files[n] for n in ("english.txt", "logs.txt", "hello.txt", "greek.txt", "crlf.txt")]
    for args, jobss in ((["-c", "the"], ("-J4",)), (["-n", "-w", "the"], ("-J1", "-J4")), (["-l", "Hello"], ("-J4",))):
        for jobs in jobss:
            full = ["--sort", jobs, *args, *paths]
            want = run(REF, full)
            got = run(B200, full, {"UGREP_B200_REQUIRE": "1"})
            assert got[0] == want[0] and got[1] == want[1], (full, got[2][:300])


@needs_binaries
def test_dropin_reports_its_engine_and_refuses_loudly(files):
    rc, out, err = run(B200, ["-c", "Hello", files["hello.txt"]], {"UGREP_B200_VERBOSE": "1"})
    assert rc == 0 and b"served by libugrep_b200" in err
    # a lookahead is outside the library's scope: with REQUIRE that is an error, without it reflex::Matcher serves it
    rc, out, err = run(B200, ["-c", "Hello(?=_)", files["hello.txt"]], {"UGREP_B200_REQUIRE": "1"})
    assert rc == 2 and b"outside the GPU path's scope" in err
    rc, out, err = run(B200, ["-c", "Hello(?=_)", files["hello.txt"]], {"UGREP_B200_VERBOSE": "1"})
    assert (rc, out) == run(REF, ["-c", "Hello(?=_)", files["hello.txt"]])[:2] and b"stays with reflex::Matcher" in err


@needs_binaries
def test_dropin_on_a_large_file(tmp_path):
    """64 MiB through the drop-in (one device scan, records replayed in batches) against the reference CLI"""
    data = corpus.block("c5", 64 << 20)
    p = tmp_path / "big.log"
    p.write_bytes(data.tobytes())
    for args in (["-c", "-o", "-e", "ERROR|WARN", "-e", "\\d{3}-\\d{4}"], ["-c", "ERROR"], ["-n", "-b", "-o", "WARN"]):
        want = run(REF, [*args, str(p)])
        got = run(B200, [*args, str(p)], {"UGREP_B200_REQUIRE": "1"})
        assert got[0] == want[0] and got[1] == want[1], args
