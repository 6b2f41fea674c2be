"""GPU parity tests proper: the CUDA path, called through the C ABI, against the
oracle (oracle/oracle.c) on the same seeded inputs; against the unmodified
reference (oracle/_ref/ugrep, refscan) where it travelled with the snapshot."""
import os

import numpy as np
import pytest

import oracle_lib as O
from ugrep_b200 import corpus

pytestmark = pytest.mark.gpu

PAT_DIR = os.path.join(O.ROOT, "ugrep_b200", "patterns")
# config -> (pattern file, corpus, mode)
CONFIGS = {
    "c1": ("c1", "c1", "lines"),
    "c2": ("c2", "c2", "lines"),
    "c2s": ("c2", "c2s", "lines"),
    "c3": ("c3", "c3", "list"),
    "c3b": ("c3b", "c3", "list"),
    "c3c": ("c3c", "c3", "list"),
    "c4": ("c4", "c4", "lines"),
    "c5": ("c5", "c5", "matches"),
}


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the -m gpu tests must run on the B200 box")
    from ugrep_b200 import api
    return api, api.Scanner(0)


def run_mode(api, sc, pat, data, mode):
    if mode == "lines":
        return sc.count_lines(pat, data).matches
    if mode == "matches":
        return sc.count_matches(pat, data).matches
    rec, tot = sc.find_all(pat, data)
    assert tot.matches == len(rec)
    return rec


def oracle_mode(op, data, mode):
    if mode == "lines":
        return op.count_lines(data)
    if mode == "matches":
        return op.count_matches(data)
    return op.find_all(data)


def same(a, b):
    if isinstance(a, np.ndarray):
        return len(a) == len(b) and bool(np.all(a == b))
    return a == b


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("mode_override", [None, "lines", "matches", "list"])
def test_config_vs_oracle(gpu, name, mode_override):
    api, sc = gpu
    pfile, cname, mode = CONFIGS[name]
    mode = mode_override or mode
    path = os.path.join(PAT_DIR, pfile + ".ugxp")
    pat = api.Pattern.load(path, 0)
    op = O.OraclePattern(path)
    data = corpus.block(cname, 3 << 20)
    got = run_mode(api, sc, pat, data, mode)
    want = oracle_mode(op, data, mode)
    assert same(got, want), "%s/%s: cuda %r oracle %r" % (name, mode, got if not isinstance(got, np.ndarray) else len(got),
                                                        want if not isinstance(want, np.ndarray) else len(want))
    # device-resident input gives the same answer as the host-staged one
    import torch
    dev = torch.from_numpy(data).cuda()
    got2 = run_mode(api, sc, pat, dev, mode)
    assert same(got2, want)
    if mode == "list":
        assert sc.count_matches(pat, dev).newlines == O.newlines(data)


EDGE_INPUTS = [
    b"",
    b"\n",
    b"\n\n\n",
    b"Sherlock Holmes",
    b"Sherlock Holmes\n",
    b"xSherlock HolmesSherlock Holmes Sherlock Holme\nSherlock Holmes",
    b"Running\nWalking and Thinking\n\nmorning Singing",
    b"a" * 70000 + b" Running " + b"b" * 70000 + b"\n" + b"Walking",
    b"2026-01-02T03:04:05 ERROR svc01 call 555-1234 ext 123-4567 id=1\n" * 3 + b"WARN 111-2222",
    "naïve Ωmega αβγ NAÏVE\nκόσμος\n".encode("utf-8"),
    b"\r\nWalking\r\n",
]


@pytest.mark.parametrize("name", list(CONFIGS))
def test_edge_inputs(gpu, name):
    api, sc = gpu
    pfile, _, _ = CONFIGS[name]
    path = os.path.join(PAT_DIR, pfile + ".ugxp")
    pat = api.Pattern.load(path, 0)
    op = O.OraclePattern(path)
    for data in EDGE_INPUTS:
        for mode in ("lines", "matches", "list"):
            got = run_mode(api, sc, pat, data, mode)
            want = oracle_mode(op, data, mode)
            assert same(got, want), "%s/%s on %r" % (name, mode, data[:60])


def test_ragged_sizes(gpu):
    """every buffer length around the strip / tile boundaries"""
    api, sc = gpu
    path = os.path.join(PAT_DIR, "c5.ugxp")
    pat = api.Pattern.load(path, 0)
    op = O.OraclePattern(path)
    base = corpus.block("c5", 40000)
    for n in list(range(0, 200)) + [16383, 16384, 16385, 32767, 32768, 32769, len(base)]:
        data = base[:n]
        assert sc.count_matches(pat, data).matches == op.count_matches(data), n
        assert sc.count_lines(pat, data).matches == op.count_lines(data), n
        if n % 7 == 0 or n > 1000:
            rec, _ = sc.find_all(pat, data, base_offset=1000, base_line=50)
            want = op.find_all(data)
            assert len(rec) == len(want), n
            assert bool(np.all(rec["offset"] == want["offset"] + 1000)) and bool(np.all(rec["line"] == want["line"] + 50)), n
            assert bool(np.all(rec["len"] == want["len"])) and bool(np.all(rec["cap"] == want["cap"])), n


def test_records_dense_strips_and_long_lines(gpu):
    """single-pass records: strips with more matches than the shared-memory slots hold (re-run path), lines far
    longer than a tile, and a first call whose staging guess is too small (second staging pass)"""
    api, _ = gpu
    sc = api.Scanner(0)
    path = os.path.join(PAT_DIR, "c5.ugxp")
    pat = api.Pattern.load(path, 0)
    op = O.OraclePattern(path)
    dense = (b"ERROR WARN 555-1234 WARN ERROR 123-4567 " * 40 + b"\n") * 30          # ~10 matches per 64 bytes
    long_line = b"x" * 50000 + b" ERROR " + b"y" * 40000 + b" 555-0000 WARN\n" + b"WARN\n" * 5
    huge = dense * 40                                                               # > n / 96 records: staging regrows
    for data in (dense, long_line, long_line + dense + long_line, huge):
        rec, tot = sc.find_all(pat, data)
        want = op.find_all(data)
        assert len(rec) == len(want) == tot.matches
        assert bool(np.all(rec == want))
    sc.set_option("two_pass_records", 1)
    rec2, _ = sc.find_all(pat, huge)
    assert bool(np.all(rec2 == op.find_all(huge)))


@pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref (the built reference) did not travel")
@pytest.mark.parametrize("name,cli", [
    ("c1", ["-c", "-F", "Sherlock Holmes"]),
    ("c3", ["-n", "-b", "-o", "[A-Z][a-z]+ing\\b"]),
    ("c3b", ["-n", "-b", "-o", "[A-Z][a-z]+ing"]),
    ("c3c", ["-n", "-b", "-o", "[A-Z][a-z]{1,9}ing\\b"]),
    ("c4", ["-i", "-c", "\\p{Greek}+|naïve\\w*"]),
    ("c5", ["-c", "-o", "-e", "ERROR|WARN", "-e", "\\d{3}-\\d{4}"]),
])
def test_config_vs_reference_cli(gpu, name, cli):
    """bit-exact against the reference's own output on the same input"""
    api, sc = gpu
    pfile, cname, mode = CONFIGS[name]
    pat = api.Pattern.load(os.path.join(PAT_DIR, pfile + ".ugxp"), 0)
    data = corpus.block(cname, 2 << 20)
    rc, ref = O.ref_cli(cli, data)
    got = run_mode(api, sc, pat, data, mode)
    mine = O.format_list(data, got) if mode == "list" else b"%d\n" % got
    assert mine == ref


# ---- every committed golden case (made with the unmodified reference), through every kernel route ----
import golden_lib as G  # noqa: E402

ROUTES = {"default": {}, "generic": {"force_generic": 1}, "no_span": {"no_span": 1}, "legacy_any": {"legacy_any": 1},
          "stream": {"stream_dfa": 1}, "stream_nl": {"stream_dfa": 1, "count_newlines": 1},
          "two_pass": {"two_pass_records": 1}, "match_lines": {"match_lines": 1}}


@pytest.mark.parametrize("route", list(ROUTES))
@pytest.mark.parametrize("name", G.pattern_names())
def test_golden_cases(gpu, name, route):
    api, _ = gpu
    sc = api.Scanner(0)
    for k, v in ROUTES[route].items():
        sc.set_option(k, v)
    try:
        pat = api.Pattern.load(G.pattern_path(name), 0)
    except api.UgxError as ex:
        assert ex.code == 2, ex  # UGX_E_UNSUPPORTED: outside the path's scope, rejected loudly
        pytest.skip("out of scope: %s" % ex)
    for case, data in G.cases(name):
        if route == "two_pass":
            rec, _ = sc.find_all(pat, data)
            assert len(rec) == case["matches"]
            G.check_list(case, data, rec)
            continue
        t = sc.count_lines(pat, data)
        assert t.matches == case["lines"], (name, route, case["input"], "lines")
        if route == "stream_nl" and t.newlines:
            assert t.newlines == data.count(b"\n"), (name, case["input"], "newlines")
        if route in ("default", "generic", "match_lines", "no_span"):
            assert sc.count_matches(pat, data).matches == case["matches"], (name, route, case["input"], "matches")
        if route in ("default", "no_span"):
            rec, _ = sc.find_all(pat, data)
            assert len(rec) == case["matches"]
            G.check_list(case, data, rec)


def test_stream_count_ragged_and_chained(gpu):
    """the streaming count at every length around chunk / span / block / region boundaries, with newlines placed
    so that lines straddle regions and whole regions hold no newline"""
    api, sc = gpu
    sc.set_option("stream_dfa", 1)
    sc2 = api.Scanner(0)
    sc2.set_option("stream_dfa", 1)
    sc2.set_option("count_newlines", 1)
    rng = np.random.default_rng(5)
    for pname in ("c1", "c2", "c4", "w_the"):
        path = os.path.join(PAT_DIR, pname + ".ugxp") if pname.startswith("c") else G.pattern_path(pname)
        pat = api.Pattern.load(path, 0)
        op = O.OraclePattern(path)
        base = corpus.block("c4" if pname == "c4" else "c1", 100000).copy()
        lit = b"Sherlock Holmes" if pname == "c1" else None
        # long lines: drop most newlines, then sprinkle a few so that some 16 KiB regions have none
        nl = np.flatnonzero(base == 10)
        drop = nl[rng.random(len(nl)) < 0.97]
        long_lines = base.copy()
        long_lines[drop] = ord(" ")
        if lit is not None:
            for at in (5, 16370, 16384 - 7, 32768 + 100, 49152 - 15, 65536 - 1, 70000):
                long_lines[at:at + len(lit)] = np.frombuffer(lit, dtype=np.uint8)
                base[at:at + len(lit)] = np.frombuffer(lit, dtype=np.uint8)
        for src in (base, long_lines):
            for n in list(range(0, 70)) + [511, 512, 513, 2047, 2048, 2049, 2055, 2056, 2057, 16383, 16384, 16385,
                                           16391, 16392, 16393, 32768, 49152 + 17, 65535, 65536, 65537, 99999, len(src)]:
                data = src[:n]
                want = op.count_lines(data)
                assert sc.count_lines(pat, data).matches == want, (pname, n)
                t = sc2.count_lines(pat, data)
                assert t.matches == want and t.newlines == int((data == 10).sum()), (pname, n, "nl")
    sc.set_option("stream_dfa", 0)


def test_pipelined_host_buffer_equals_device_scan(gpu):
    """host buffers >= 64 MiB on the streaming route are copied in chunks while earlier chunks are scanned; the
    region summaries chain the launches.  Counts must equal the one-launch device-resident scan and the oracle
    (additive over line-aligned repetitions of one block)."""
    import torch
    api, sc = gpu
    for pname, cname in (("c1", "c1"), ("c2", "c2")):
        path = os.path.join(PAT_DIR, pname + ".ugxp")
        pat = api.Pattern.load(path, 0)
        op = O.OraclePattern(path)
        block = corpus.block(cname, 5 << 20)
        assert block[-1] == 10
        reps = 15  # 75 MiB: three chunks of the pipeline, the last one partial
        # a line that straddles every block boundary would break additivity: blocks end with a newline
        host = np.tile(block, reps)
        # put a literal across a 32 MiB chunk boundary (c1) so that the lag of one region is exercised
        if pname == "c1":
            at = (32 << 20) - 7
            host = host.copy()
            host[at:at + 15] = np.frombuffer(b"Sherlock Holmes", dtype=np.uint8)
        want_dev = sc.count_lines(pat, torch.from_numpy(host).cuda())
        pinned = torch.from_numpy(host).pin_memory()
        got = sc.count_lines(pat, pinned.numpy())
        assert got.launches >= 3, got          # page-locked memory: chunked copy overlapped with the scan
        assert got.matches == want_dev.matches
        sc.set_option("no_pipeline", 1)
        assert sc.count_lines(pat, pinned.numpy()).matches == want_dev.matches
        sc.set_option("no_pipeline", 0)
        # pageable memory (what ugrep hands over): feeder threads through pinned slots, then one launch
        got = sc.count_lines(pat, host)
        assert got.launches == 1 and got.matches == want_dev.matches
        sc.set_option("no_feeder", 1)
        assert sc.count_lines(pat, host).matches == want_dev.matches
        sc.set_option("no_feeder", 0)
        assert sc.count_matches(pat, host).matches == sc.count_matches(pat, pinned.numpy()).matches
        if pname == "c2":
            assert got.matches == reps * op.count_lines(block)


def test_never_firing_prefilter_on_a_pipelined_host_buffer(gpu, tmp_path):
    """a pattern whose bitap table admits no candidate (config 3's situation) with lbk == 0 and a bounded DFA, on a
    host buffer big enough for the chunked H2D pipeline: the scan must still see the copied bytes (the newline count
    it reports comes from the text)"""
    api, sc = gpu
    raw = bytearray(open(G.pattern_path("min2"), "rb").read())
    tap_off = 24 + 12 * 4 + 256 + 256
    raw[tap_off:tap_off + 2048] = b"\xff" * 2048
    path = tmp_path / "never.ugxp"
    path.write_bytes(bytes(raw))
    pat = api.Pattern.load(str(path), 0)
    host = np.tile(corpus.block("c1", 5 << 20), 14)   # 70 MiB > 2 pipeline chunks
    t = sc.count_lines(pat, host)
    assert t.matches == 0 and t.kernel == "count_newlines_kernel"
    assert t.newlines == int((host == 10).sum())


def test_count_newlines(gpu):
    """ugx_count_newlines = reflex::nlcount: every length around the 16-byte vectors, device and host buffers"""
    import torch
    api, sc = gpu
    base = corpus.block("c5", 300000)
    for n in list(range(0, 70)) + [255, 256, 257, 4095, 4096, 4097, 65535, 65536, 65537, len(base)]:
        data = base[:n]
        assert sc.count_newlines(data).newlines == int((data == 10).sum()), n
    big = np.tile(corpus.block("c1", 1 << 20), 40)
    assert sc.count_newlines(torch.from_numpy(big).cuda()).newlines == int((big == 10).sum())
    tricky = np.frombuffer(b"\n\x0b\n\x0b\x0a\x8a\x0a\xff\n" * 1000, dtype=np.uint8)  # bytes one off '\n', high bits set
    assert sc.count_newlines(tricky).newlines == int((tricky == 10).sum())


def test_misaligned_device_pointers_and_random_slices(gpu):
    """device buffers that do not start on a 16-byte boundary (views into a larger tensor) and random slices that cut
    lines anywhere: every mode against the oracle"""
    import torch
    api, sc = gpu
    rng = np.random.default_rng(11)
    for pname, cname in (("c1", "c1"), ("c2", "c2"), ("c4", "c4"), ("c5", "c5"), ("c3b", "c3")):
        path = os.path.join(PAT_DIR, pname + ".ugxp")
        pat = api.Pattern.load(path, 0)
        op = O.OraclePattern(path)
        base = corpus.block(cname, 400000)
        dev = torch.from_numpy(base).cuda()
        for _ in range(6):
            lo = int(rng.integers(0, 50000))
            hi = int(rng.integers(lo + 1, len(base)))
            data = base[lo:hi]
            view = dev[lo:hi]
            assert sc.count_lines(pat, view).matches == op.count_lines(data), (pname, lo, hi)
            assert sc.count_matches(pat, view).matches == op.count_matches(data), (pname, lo, hi)
            rec, _ = sc.find_all(pat, view)
            assert same(rec, op.find_all(data)), (pname, lo, hi)
            assert sc.count_newlines(view).newlines == int((data == 10).sum())


def test_full_size_properties(gpu):
    """BASELINE-sized inputs, checked through size-independent properties: a corpus tiled from R line-aligned copies
    of one seeded block must give R times the block's counts (the block's counts come from the ORACLE), and its match
    records must be the block's records repeated with shifted offsets and line numbers — including offsets and
    record indices beyond 2^32 (c3b at 4.75 GiB)."""
    import torch
    api, sc = gpu
    block_bytes = 32 << 20
    for pname, cname, mode, gib in (("c1", "c1", "lines", 4), ("c2", "c2", "lines", 4), ("c2", "c2s", "lines", 2),
                                    ("c4", "c4", "lines", 8), ("c5", "c5", "matches", 8), ("c5", "c5", "lines", 2),
                                    ("c3", "c3", "list", 4), ("c3b", "c3", "list", 4.75)):
        path = os.path.join(PAT_DIR, pname + ".ugxp")
        pat = api.Pattern.load(path, 0)
        op = O.OraclePattern(path)
        block = corpus.block(cname, block_bytes)
        assert block[-1] == 10
        reps = int(gib * (1 << 30)) // len(block)
        dev = torch.from_numpy(block).cuda().repeat(reps)
        nl = int((block == 10).sum())
        if mode == "lines":
            assert sc.count_lines(pat, dev).matches == reps * op.count_lines(block), (pname, cname)
        elif mode == "matches":
            t = sc.count_matches(pat, dev)
            assert t.matches == reps * op.count_matches(block), pname
            assert t.newlines == reps * nl, pname
        else:
            want = op.find_all(block)
            tot = sc.find_all_device(pat, dev, base_offset=7, base_line=3)
            assert tot.matches == reps * len(want), pname
            assert tot.newlines == reps * nl, pname
            if len(want):
                if gib > 4:
                    assert int(want["offset"][-1]) + (reps - 1) * len(block) > 1 << 32
                for r in (0, reps // 2, reps - 1):
                    got = sc.fetch(r * len(want), len(want))
                    assert bool(np.all(got["offset"] == want["offset"] + r * len(block) + 7)), (pname, r)
                    assert bool(np.all(got["line"] == want["line"] + r * nl + 3)), (pname, r)
                    assert bool(np.all(got["len"] == want["len"])) and bool(np.all(got["cap"] == want["cap"])), (pname, r)
        del dev
        torch.cuda.empty_cache()


def test_long_lines_are_scanned_by_all_warps(gpu):
    """lines far longer than a region (here: one line of 8 MiB with records, one of 256 MiB with counts) go through
    the span kernels like any other text — regions take their chain state from the 512 bytes before them — and must
    equal the oracle's sequential find loop; a throughput floor guards against a one-thread fallback"""
    import torch
    api, sc = gpu
    path = os.path.join(PAT_DIR, "c5.ugxp")
    pat = api.Pattern.load(path, 0)
    op = O.OraclePattern(path)
    base = corpus.block("c5", 32 << 20).copy()
    base[base == 10] = 32
    small = base[:8 << 20].copy()
    small[-5:] = np.frombuffer(b" end\n", dtype=np.uint8)
    rec, tot = sc.find_all(pat, small)
    assert tot.kernel == "span_scan_kernel"
    want = op.find_all(small)
    assert len(want) > 100000 and same(rec, want)
    assert int(rec["line"].max()) == 1
    big = np.tile(base, 8)
    big[-5:] = np.frombuffer(b" end\n", dtype=np.uint8)
    dev = torch.from_numpy(big).cuda()
    t = sc.count_matches(pat, dev)
    assert t.kernel == "span_scan_kernel" and t.newlines == 1
    assert t.matches == op.count_matches(big) and t.matches > 1000000
    best = min(sc.count_matches(pat, dev).kernel_ms for _ in range(3))
    assert big.size / best / 1e6 >= 50.0, "%.1f GB/s on a single 256 MiB line" % (big.size / best / 1e6)


def test_span_kernels_hand_over_what_they_cannot_vouch_for(gpu):
    """inputs outside the span kernels' guarantees must come back exact through the line-at-a-time kernels: a match
    longer than a window across a region start inside a long line, a look-back run longer than the look-ahead bound,
    a match of 64 KiB or more, an attempt that fails at the very end of a long last line"""
    api, sc = gpu
    cases = [
        ("dotstar", b"x" * 20000 + b"a" + b"y" * 40000 + b"b zz a b\n" + b"a b\n" * 10),
        ("pin_pma_lb", b"the s" + b"a" * 1100000 + b"ing sing\nsing song\n"),
        ("dotstar", b"a" + b"q" * 70000 + b"b\n" + b"ab\n" * 5),
        ("c5", b"x" * 100000 + b" ERROR 555-12"),
    ]
    for name, data in cases:
        path = G.pattern_path(name) if name != "c5" else os.path.join(PAT_DIR, "c5.ugxp")
        pat = api.Pattern.load(path, 0)
        op = O.OraclePattern(path)
        rec, tot = sc.find_all(pat, data)
        assert same(rec, op.find_all(data)), name
        assert tot.flags & 2 and tot.kernel != "span_scan_kernel", name   # UGX_TOT_SPAN_HANDOVER
        t = sc.count_matches(pat, data)
        assert t.matches == op.count_matches(data) and t.flags & 1 and t.newlines == data.count(b"\n"), name
    # ... and the ordinary case does stay with the spans
    pat = api.Pattern.load(os.path.join(PAT_DIR, "c5.ugxp"), 0)
    t = sc.count_matches(pat, corpus.block("c5", 1 << 20))
    assert t.kernel == "span_scan_kernel" and t.flags == 1
    # ugx_count_lines on its streaming kernels reports newlines only when asked to count them (flag UGX_TOT_NEWLINES)
    pat2 = api.Pattern.load(os.path.join(PAT_DIR, "c2.ugxp"), 0)
    blk = corpus.block("c2", 1 << 20)
    assert sc.count_lines(pat2, blk).flags & 1 == 0
    sc.set_option("count_newlines", 1)
    t = sc.count_lines(pat2, blk)
    assert t.flags & 1 and t.newlines == int((blk == 10).sum())
    sc.set_option("count_newlines", 0)


def test_literal_compiled_by_the_library_itself(gpu):
    """config 1 without any reference binary: ugx_compile_literal + ugx_pattern_create, against the oracle running
    the reference-compiled pattern"""
    api, sc = gpu
    data = corpus.block("c1", 4 << 20)
    op = O.OraclePattern(os.path.join(PAT_DIR, "c1.ugxp"))
    pat = api.Pattern.literal(b"Sherlock Holmes", 0)
    t = sc.count_lines(pat, data)
    assert t.matches == op.count_lines(data) and t.kernel == "count_lines_literal_kernel"
    rec, _ = sc.find_all(pat, data)
    assert same(rec, op.find_all(data))
    for lit in (b"the", b"e", b"water long little", "naïve".encode()):
        pat = api.Pattern.literal(lit, 0)
        blk = corpus.block("c4" if lit[0] > 127 or b"\xc3" in lit else "c1", 1 << 20)
        raw = blk.tobytes()
        assert sc.count_matches(pat, blk).matches == raw.count(lit), lit   # non-overlapping occurrences
        assert sc.count_lines(pat, blk).matches == sum(1 for ln in raw.split(b"\n") if lit in ln), lit


def test_word_list_compiled_by_the_library_itself(gpu):
    """config 2 without any reference binary: ugx_compile_words + ugx_pattern_create"""
    api, sc = gpu
    op = O.OraclePattern(os.path.join(PAT_DIR, "c2.ugxp"))
    pat = api.Pattern.words(corpus.words_list(), 0)
    for cname in ("c2", "c2s"):
        data = corpus.block(cname, 4 << 20)
        assert sc.count_lines(pat, data).matches == op.count_lines(data)
        assert sc.count_matches(pat, data).matches == op.count_matches(data)


def test_word_list_too_big_for_shared_memory(gpu, tmp_path):
    """3 000 words: a 400 KB transition table, read from global memory (Stepper's unstaged form) by the streaming count,
    the span kernels and the records path; oracle on the same compiled pattern, plus a plain substring count"""
    api, sc = gpu
    words = corpus.words_list(11, 3000)
    opc, pf = api.compile_words(words)
    path = str(tmp_path / "w3000.ugxp")
    api.write_ugxp(path, opc, pf)
    op = O.OraclePattern(path)
    pat = api.Pattern.words(words, 0)
    assert pat.info["table_in_smem"] == 0 and pat.info["table_bytes"] > 300000
    data = corpus.block("c2", 2 << 20)
    want = op.count_lines(data)
    assert want == sum(1 for ln in data.tobytes().split(b"\n") if any(w in ln for w in words))
    assert sc.count_lines(pat, data).matches == want
    assert sc.count_matches(pat, data).matches == op.count_matches(data)
    rec, _ = sc.find_all(pat, data)
    assert same(rec, op.find_all(data))


def test_scanners_on_two_host_threads_share_a_pattern(gpu):
    """a ugx_pattern is shareable, a ugx_scanner belongs to one host thread: two threads scanning at once with
    different patterns (different table sizes, hence different shared-memory needs of the same kernels)"""
    import threading
    api, _ = gpu
    jobs = []
    for pname, cname in (("c2", "c2"), ("c5", "c5"), ("c4", "c4"), ("c3b", "c3")):
        path = os.path.join(PAT_DIR, pname + ".ugxp")
        data = corpus.block(cname, 2 << 20)
        op = O.OraclePattern(path)
        jobs.append((api.Pattern.load(path, 0), data, op.count_lines(data), op.count_matches(data)))
    errors = []

    def worker(seed):
        try:
            sc = api.Scanner(0)
            for i in range(12):
                pat, data, want_l, want_m = jobs[(seed + i) % len(jobs)]
                if sc.count_lines(pat, data).matches != want_l or sc.count_matches(pat, data).matches != want_m:
                    errors.append((seed, i))
        except Exception as ex:  # noqa: BLE001
            errors.append((seed, repr(ex)))

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
