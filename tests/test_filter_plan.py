"""CPU test of the host-side export (DFA flattening + first-stage filter plan, csrc/pattern_host.cpp) through the
host-only entry point ugx_plan_describe: the plan must be a SUPERSET of the reference's candidate predicate on
interior positions — the property that makes the two-stage evaluation of the position-parallel kernels exact.
The reference predicate comes from the oracle (ora_candidates); stage 1 is simulated here with numpy from the
tables the library reports, exactly as the kernels evaluate it (stream_count.cu / stream_literal.cu)."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

import golden_lib as G
import oracle_lib as O
from ugrep_b200 import corpus

LIB = os.path.join(O.ROOT, "ugrep_b200", "libugrep_b200.so")


class PlanInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("states", "classes", "table_bytes", "first_acc", "first_leaf", "max_match_len",
                                          "advance", "has_meta", "newline_live", "kind", "nterms")] + \
               [("t_off", C.c_uint32 * 3), ("a_off", C.c_uint32 * 2), ("a_chr", C.c_uint32 * 2),
                ("h4_terms", C.c_uint32), ("h4_shift", C.c_uint32), ("pm2", C.c_uint32), ("pm2_shift", C.c_uint32),
                ("lut", C.c_uint32 * 256), ("covers", C.c_uint32)]


def load_ugxp(path):
    b = open(path, "rb").read()
    nop, regex_len, pf_size, flags = struct.unpack_from("<4I", b, 8)
    pf = b[24:24 + pf_size]
    opc = b[24 + pf_size:24 + pf_size + 4 * nop]
    return opc, nop, pf, flags


def describe(path):
    lib = C.CDLL(LIB)
    lib.ugx_last_error.restype = C.c_char_p
    opc, nop, pf, flags = load_ugxp(path)
    info = PlanInfo()
    rc = lib.ugx_plan_describe(opc, nop, pf, flags, C.byref(info))
    return rc, info, pf


def stage1(info, pf, data):
    """pass mask of the first stage for positions 0 .. n-25 (interior), as the kernels evaluate it"""
    n = len(data)
    m = n - 24
    d = data.astype(np.uint32)
    ok = np.ones(m, dtype=bool)
    pma = np.frombuffer(pf, dtype=np.uint8, count=4096, offset=48 + 256 + 256 + 2048).astype(np.uint32)
    pmh = np.frombuffer(pf, dtype=np.uint8, count=4096, offset=48 + 256 + 256 + 2048 + 4096).astype(np.uint32)
    if info.kind == 2:
        ok &= d[0:m] == info.a_chr[0]
        ok &= d[info.a_off[1]:info.a_off[1] + m] == info.a_chr[1]
        return ok, "anchor2"
    used = []
    if info.h4_terms:
        g = np.zeros(n, dtype=np.uint32)
        for back in range(4):  # g(p) = hash of bytes p-3 .. p
            g[3:] ^= (d[3 - back:n - back] << (3 * back)) & 4095
        for t in range(info.h4_terms):
            p = np.arange(m) + info.h4_shift + 3 + t
            ok &= ((pmh[g[p]] >> (3 + t)) & 1) == 0
        used.append("h4x%d" % info.h4_terms)
    elif info.pm2:
        c0 = d[info.pm2_shift:info.pm2_shift + m]
        c1 = d[info.pm2_shift + 1:info.pm2_shift + 1 + m]
        a, b = pma[c0], pma[((c0 << 3) ^ c1) & 4095]
        q7, q6, q5, q4 = (a >> 7) & 1, (a >> 6) & 1, (b >> 5) & 1, (b >> 4) & 1
        ok &= (q7 & q5 & (q4 | q6)) == 0
        used.append("pm2")
    elif info.kind == 3:
        lut = np.array(list(info.lut), dtype=np.uint32)
        for t in range(info.nterms):
            ok &= ((lut[d[info.t_off[t]:info.t_off[t] + m]] >> (8 * t)) & 1) == 0
        used.append("lut%d" % info.nterms)
    return ok, "+".join(used) or "all"


@pytest.mark.parametrize("name", G.pattern_names())
def test_stage1_is_a_superset_of_the_reference_candidates(name):
    path = G.pattern_path(name)
    rc, info, pf = describe(path)
    if rc == 2:
        pytest.skip("out of scope (rejected at upload)")
    assert rc == 0
    op = O.OraclePattern(path)
    assert info.advance == op.advance
    total_c = total_p = 0
    for cname in ("c1", "c2", "c3", "c4", "c5"):
        data = corpus.block(cname, 60000)
        cand = op.candidates(data)
        ok, how = stage1(info, pf, data)
        m = len(ok)
        missed = np.flatnonzero(cand[:m] & ~ok)
        assert missed.size == 0, "%s (%s): stage 1 drops reference candidates at %s of %s" % (name, how, missed[:5], cname)
        total_c += int(cand[:m].sum())
        total_p += int(ok.sum())
    assert total_p >= total_c


@pytest.mark.parametrize("name", G.pattern_names())
def test_covers_flag_agrees_with_the_oracle(name):
    """`covers` (prefilter_covers_matches: proven by enumeration over the DFA) says that every position starting a
    non-empty match passes the reference's candidate predicate; check its conclusion against the oracle, position by
    position, on the golden inputs and corpus blocks (interior positions: the proof leaves the buffer's end out)"""
    path = G.pattern_path(name)
    rc, info, pf = describe(path)
    if rc == 2:
        pytest.skip("out of scope (rejected at upload)")
    if not info.covers or info.advance == 0:  # advance_none: every position is attempted, nothing to prove
        return
    op = O.OraclePattern(path)
    inputs = [np.frombuffer(d[:20000], dtype=np.uint8) for _, d in G.cases(name)]
    inputs += [corpus.block(c, 20000) for c in ("c2", "c4", "c5")]
    checked = 0
    for a in inputs:
        if len(a) < 40:
            continue
        cand = op.candidates(a)
        for p in np.flatnonzero(~cand[:len(a) - 32])[:6000]:
            cap, ln = op.match_at(a, int(p))
            assert not (cap and ln), (name, int(p))
            checked += 1
    assert checked >= 0


def test_covers_is_proven_for_the_word_list_and_refused_for_config3():
    pat = os.path.join(O.ROOT, "ugrep_b200", "patterns")
    assert describe(os.path.join(pat, "c2.ugxp"))[1].covers == 1   # PMH over min = 4 bytes of a tree DFA
    assert describe(os.path.join(pat, "c4.ugxp"))[1].covers == 1   # PM4, min = 2: "whatever follows" from the table bits
    assert describe(os.path.join(pat, "c1.ugxp"))[1].covers == 0   # `one`: the predicate is the match itself
    assert describe(os.path.join(pat, "c3.ugxp"))[1].covers == 0   # the prefilter has false negatives
    assert describe(os.path.join(pat, "c5.ugxp"))[1].covers == 0   # look-back: the attempt set is more than cand()


def test_dfa_export_shapes_of_the_configs():
    pat = os.path.join(O.ROOT, "ugrep_b200", "patterns")
    rc, c2, _ = describe(os.path.join(pat, "c2.ugxp"))
    assert rc == 0 and c2.h4_terms == 1 and c2.max_match_len < 64          # tree DFA of the word list: bounded
    words = open(os.path.join(pat, "words.txt")).read().split()
    assert c2.max_match_len == max(len(w) for w in words)
    rc, c4, _ = describe(os.path.join(pat, "c4.ugxp"))
    assert rc == 0 and c4.max_match_len == 0xFFFFFFFF and c4.pm2 == 1      # \p{Greek}+ : unbounded
    rc, c1, _ = describe(os.path.join(pat, "c1.ugxp"))
    assert rc == 0 and c1.kind == 2 and c1.a_off[0] == 0 and c1.a_chr[0] == ord("S") and c1.max_match_len == 15
    for info in (c2, c4, c1):
        assert 1 <= info.first_acc <= info.first_leaf <= info.states
        assert info.table_bytes == info.states * info.classes * 2


def test_covers_is_sound_on_random_word_lists(tmp_path):
    """word lists compiled by the library itself (min lengths 1..8: the PM4, bitap and hashed-predictor routines):
    wherever the proof says `covers`, no position outside the oracle's candidate set starts a match — on text made of
    the words, their prefixes and noise"""
    from ugrep_b200 import api
    rng = np.random.default_rng(2024)
    proven = 0
    for trial in range(40):
        alpha = [b"ab", b"abcdefgh", b"etaoinshr", b"0123456789-"][trial % 4]
        lo = int(rng.integers(1, 7))
        words = set()
        while len(words) < int(rng.choice([1, 2, 5, 20, 100])):
            n = int(rng.integers(lo, lo + 5))
            words.add(bytes(int(alpha[int(i)]) for i in rng.integers(0, len(alpha), size=n)))
        words = sorted(words)
        icase = trial % 5 == 4   # (-i: an uppercase twin for every lowercase edge; the text below mixes the cases)
        try:
            opc, pf = api.compile_words(words, icase=icase)
        except api.UgxError:
            continue
        path = str(tmp_path / ("w%d.ugxp" % trial))
        api.write_ugxp(path, opc, pf)
        rc, info, _ = describe(path)
        assert rc == 0
        if not info.covers or info.advance == 0:
            continue
        proven += 1
        op = O.OraclePattern(path)
        parts = []
        for _ in range(400):
            k = int(rng.integers(0, 4))
            w = words[int(rng.integers(0, len(words)))]
            parts.append(w if k == 0 else w[:int(rng.integers(0, len(w) + 1))] if k == 1 else
                         bytes(int(alpha[int(i)]) for i in rng.integers(0, len(alpha), size=int(rng.integers(1, 6)))) if k == 2
                         else rng.choice([b" ", b"\n", b"x"]))
        text = b"".join(parts)
        if icase:
            text = bytes((c - 32 if (97 <= c <= 122 and rng.random() < 0.4) else c) for c in text)
        a = np.frombuffer(text + b"\n" + b" " * 40, dtype=np.uint8)
        cand = op.candidates(a)
        for p in np.flatnonzero(~cand[:len(a) - 32]):
            cap, ln = op.match_at(a, int(p))
            assert not (cap and ln), (trial, words[:5], int(p))
    assert proven >= 10
