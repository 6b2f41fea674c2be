"""ugx_compile_literal (the part for `-F 'literal'`): byte-identical to what the unmodified reference's
pattern compiler produces — against the committed .ugxp files everywhere, against `refscan dump` on random literals
where the built reference is present."""
import os
import struct

import numpy as np
import pytest

import golden_lib as G
import oracle_lib as O
from ugrep_b200 import api

PRE = 24
PF = api.PREFILTER_BYTES


def parts(path):
    raw = open(path, "rb").read()
    nop = struct.unpack_from("<I", raw, 8)[0]
    return raw[PRE:PRE + PF], np.frombuffer(raw, dtype="<u4", count=nop, offset=PRE + PF)


@pytest.mark.parametrize("path,lit", [
    (os.path.join(O.ROOT, "ugrep_b200", "patterns", "c1.ugxp"), b"Sherlock Holmes"),
    (G.pattern_path("hello_F"), b"Hello"), (G.pattern_path("char1"), b"e"), (G.pattern_path("char2"), b"th"),
    (G.pattern_path("char3"), b"the"), (G.pattern_path("str4"), b"that"),
])
def test_literal_compile_equals_committed_reference_output(path, lit):
    pf, opc = parts(path)
    got_opc, got_pf = api.compile_literal(lit)
    assert got_opc.tolist() == opc.tolist()
    assert got_pf == pf


def test_literal_compile_scope():
    for bad in (b"", b"a\nb", b"a\rb", b"x" * 255):
        with pytest.raises(api.UgxError) as e:
            api.compile_literal(bad)
        assert e.value.code == 2
    api.compile_literal(b"x" * 254)


@pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref (the built reference) is not present")
def test_literal_compile_equals_the_live_reference_on_random_literals(tmp_path):
    rng = np.random.default_rng(4)
    alphabet = (b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789 _-.,:;!?@#$%&*()[]{}<>|/+=~^'\""
                + "éüßñαβγ日本".encode("utf-8"))
    lits = [b"Sherlock Holmes", b"aa", b"aaa", b"abab", b"zq", b"\\E", b"a\\Eb\\", b"\\Q\\E", b"\xff\xfe", b"-e", b"x" * 254]
    for _ in range(1000):
        n = int(rng.integers(1, 40))
        lits.append(bytes(alphabet[int(i)] for i in rng.integers(0, len(alphabet), size=n)))
    out = str(tmp_path / "p.ugxp")
    bad = 0
    for lit in lits:
        try:
            arg = lit.decode("utf-8")
        except UnicodeDecodeError:
            arg = None
        if arg is None or arg.startswith("-") and False:
            # bytes that are not valid UTF-8 cannot be passed through argv as str: hand them over as bytes
            import subprocess
            r = subprocess.run([O.REF_SCAN.encode(), b"dump", b"-F", b"-e", lit, b"-o", out.encode()], capture_output=True)
            assert r.returncode == 0, r.stderr
        else:
            O.ref_dump(["-F", "-e", arg], out)
        pf, opc = parts(out)
        got_opc, got_pf = api.compile_literal(lit)
        if got_opc.tolist() != opc.tolist() or got_pf != pf:
            bad += 1
            f = struct.unpack_from("<12I", pf, 0)
            g = struct.unpack_from("<12I", got_pf, 0)
            print("MISMATCH", lit, f, g)
    assert bad == 0


# ---- ugx_compile_words: `ugrep -F -f words.txt` (config 2)

def test_wordlist_compile_equals_committed_reference_output():
    """the config-2 pattern (1 000 words) and the three-literal golden, as the unmodified reference compiled them"""
    from ugrep_b200 import corpus
    pf, opc = parts(os.path.join(O.ROOT, "ugrep_b200", "patterns", "c2.ugxp"))
    got_opc, got_pf = api.compile_words(corpus.words_list())
    assert got_opc.tolist() == opc.tolist() and got_pf == pf
    pf, opc = parts(G.pattern_path("alt3"))
    got_opc, got_pf = api.compile_words([b"ERROR", b"WARN", b"INFO"])
    assert got_opc.tolist() == opc.tolist() and got_pf == pf
    # one word is one literal
    a = api.compile_words([b"Sherlock Holmes"])
    b = api.compile_literal(b"Sherlock Holmes")
    assert a[0].tolist() == b[0].tolist() and a[1] == b[1]


def test_wordlist_compile_icase_equals_committed_reference_output():
    """`-F -i -f LIST`: strings lowered into the tree, an uppercase twin for every lowercase edge (tools/make_icase_golden.py)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_icase_golden", os.path.join(O.ROOT, "tools", "make_icase_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    pf, opc = parts(os.path.join(O.ROOT, "tests", "golden", "words_icase.ugxp"))
    got_opc, got_pf = api.compile_words(mod.WORDS, icase=True)
    assert got_opc.tolist() == opc.tolist()
    assert got_pf == pf
    # and the compiled pattern matches case-insensitively in the oracle
    op = O.OraclePattern(os.path.join(O.ROOT, "tests", "golden", "words_icase.ugxp"))
    text = b"an ERROR here\nwarn: Fatal TIMEOUT2\nnothing\nSHERLOCK holmes and NA\xc3\xafVE\n"
    assert op.count_lines(text) == 3 and op.count_matches(text) == 6


@pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref (the built reference) is not present")
def test_single_literal_icase_equals_the_live_reference(tmp_path):
    """`-F -i 'literal'`: the word-list compiler on a list of one"""
    out = str(tmp_path / "p.ugxp")
    rng = np.random.default_rng(4)
    n = 0
    for _ in range(120):
        lit = bytes(int(x) for x in rng.choice(list(b"abcXYZ eE0-_.:"), size=int(rng.integers(1, 20))))
        if lit.strip() != lit:
            continue
        O.ref_dump(["-F", "-i", "-e", lit.decode()], out)
        pf, opc = parts(out)
        got_opc, got_pf = api.compile_words([lit], icase=True)
        assert got_opc.tolist() == opc.tolist() and got_pf == pf, lit
        n += 1
    assert n >= 60


@pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref (the built reference) is not present")
def test_plain_regex_alternations_compile_like_fixed_strings(tmp_path):
    """`ugrep [-i] -e A -e B` without -F, no regex operator in A, B: same compiled form (api.compile_plain)"""
    from ugrep_b200 import corpus
    out = str(tmp_path / "p.ugxp")
    rng = np.random.default_rng(6)
    eng = [w.encode() for w in corpus._ENGLISH]
    punct = [b"id=7", b"a-b", b"x:y", b"/usr/lib", b"a,b;c", b"u@h", b"#tag", b"50%", b"R&D", b"~x", b"it's", b"<b>", b"say \"hi\"", b"a b"]
    for k in range(40):
        words = [bytes(w) for w in rng.choice(eng, size=int(rng.integers(1, 8)), replace=False)]
        if k % 3 == 0:
            words.append(punct[k % len(punct)])
        icase = k % 2 == 1
        args = ["-i"] if icase else []
        for w in words:
            args += ["-e", w.decode()]
        O.ref_dump(args, out)
        pf, opc = parts(out)
        got_opc, got_pf = api.compile_plain(words, icase=icase)
        assert got_opc.tolist() == opc.tolist() and got_pf == pf, words
    # escapes, \\Q..\\E and top-level alternation inside ONE pattern
    fixed = [r"a\.b", r"x\*y", r"\(a\)", r"a\|b", r"c:\\dir", r"\[x\]", r"a\+b", r"wh\?", r"\$9", r"\^x", r"a\{2\}", r"tab\there",
             r"a\-b", r"\Qa.b*\E", r"foo\.bar|baz\.qux", r"one|two|three", r"\Q(x)\E|y\/z", r"na\xc3\xafve".replace(r"\xc3\xaf", "ï")]
    punct_all = "!\"#%&',-/:;@`"      # escapable; \\~ \\< \\> \\= \\_ \\e are rewritten or rejected by the reference: refused here
    bare_only = "<=>_~"
    ops = ".[](){}*+?|^$\\"
    for k in range(40):
        alts = []
        for _ in range(int(rng.integers(1, 5))):
            w = ""
            for _ in range(int(rng.integers(1, 9))):
                r = rng.random()
                if r < 0.6:
                    w += chr(int(rng.choice(list(b"abcdexyzAB019"))))
                elif r < 0.8:
                    w += "\\" + ops[int(rng.integers(0, len(ops)))]
                elif r < 0.9:
                    both = punct_all + bare_only
                    w += both[int(rng.integers(0, len(both)))]
                else:
                    w += "\\" + punct_all[int(rng.integers(0, len(punct_all)))]
            alts.append(w)
        fixed.append("|".join(alts))
    for k, rx in enumerate(fixed):
        icase = k % 3 == 2 and rx.isascii()   # (-i with non-ASCII letters outside \\Q..\\E: refused, see below)
        O.ref_dump((["-i"] if icase else []) + ["-e", rx], out)
        pf, opc = parts(out)
        got_opc, got_pf = api.compile_plain([rx.encode()], icase=icase)
        assert got_opc.tolist() == opc.tolist() and got_pf == pf, rx
    for bad in (b"a.b", b"x*", b"(a)", b"^a", b"a$", b"a\\b", b"a\\d", b"[ab]", b"a{2}", b"a+", b"a?", b"a||b", b"|a", b"\\Qabc",
                b"a\\x41", b"\\xe9", b"a]", b"a}", b"a\\", b"a\\~b", b"a\\<b", b"a\\eb", b"a\\_b"):
        with pytest.raises(api.UgxError):
            api.compile_plain([bad])
    # a leading (?i) is the inline form of -i
    O.ref_dump(["-e", "(?i)Bzaq|bxcd|Hello"], out)
    pf, opc = parts(out)
    got_opc, got_pf = api.compile_plain([b"(?i)Bzaq|bxcd|Hello"])
    assert got_opc.tolist() == opc.tolist() and got_pf == pf
    with pytest.raises(api.UgxError):
        api.compile_plain(["naïve".encode()], icase=True)
    api.compile_plain(["\\Qnaïve\\E".encode()], icase=True)


def test_wordlist_compile_scope():
    for bad in ([b""], [b"ok", b""], [b"a\nb"], [b"a\x00"]):
        with pytest.raises(api.UgxError) as e:
            api.compile_words(bad)
        assert e.value.code == 2


def _random_lists():
    from ugrep_b200 import corpus
    rng = np.random.default_rng(5)
    eng = [w.encode() for w in corpus._ENGLISH]
    al = b"abcdefghijklmnopqrstuvwxyz0123456789"
    for i in range(25):
        yield [bytes(w) for w in rng.choice(eng, size=int(rng.integers(1, 60)), replace=False)]
    for i in range(25):
        yield corpus._syllable_words(np.random.default_rng(100 + i), int(rng.integers(2, 400)), int(rng.integers(1, 3)),
                                     int(rng.integers(3, 7)), set())
    for i in range(15):
        pre = bytes(int(x) for x in rng.choice(list(b"abcdexyz"), size=int(rng.integers(1, 6))))
        yield [pre + bytes(w) for w in rng.choice(eng, size=int(rng.integers(1, 30)), replace=False)]
    for i in range(15):
        yield sorted({bytes(int(x) for x in rng.choice(list(b"ab01"), size=int(rng.integers(1, 7)))) for _ in range(int(rng.integers(1, 40)))})
    for i in range(25):
        n, k, L = int(rng.integers(20, 500)), int(rng.integers(4, 36)), int(rng.integers(6, 14))
        yield sorted({bytes(int(x) for x in rng.choice(list(al[:k]), size=int(rng.integers(max(1, L - 2), L + 1)))) for _ in range(n)})
    yield corpus._syllable_words(np.random.default_rng(500), 9000, 3, 8, set())   # beyond 64K opcode words: LONG jumps
    yield [w.encode() for w in ("naïve", "café", "über", "日本語", "Привет", "中文", "señor", "αβγ")]


@pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref (the built reference) is not present")
@pytest.mark.parametrize("icase", [False, True])
def test_wordlist_compile_equals_the_live_reference_on_random_lists(tmp_path, icase):
    """byte equality of opcode words and prefilter block with `refscan dump -F [-i] -f`; a list the library refuses must
    be one for which the reference's analysis cut the DFA (cut_ != 0), and only those"""
    wf = tmp_path / "w.txt"
    out = str(tmp_path / "p.ugxp")
    n_ok = n_cut = 0
    rng = np.random.default_rng(9)
    for words in _random_lists():
        if icase:  # mixed case in the list itself
            words = [bytes((c - 32 if (97 <= c <= 122 and rng.random() < 0.3) else c) for c in w) for w in words]
        wf.write_bytes(b"\n".join(words) + b"\n")
        O.ref_dump(["-F", "-i", "-f", str(wf)] if icase else ["-F", "-f", str(wf)], out)
        pf, opc = parts(out)
        cut = struct.unpack_from("<12I", pf, 0)[11]
        try:
            got_opc, got_pf = api.compile_words(words, icase=icase)
        except api.UgxError as e:
            assert e.code == 2 and cut != 0, words[:3]
            n_cut += 1
            continue
        assert cut == 0, words[:3]
        assert got_opc.tolist() == opc.tolist() and got_pf == pf, words[:3]
        n_ok += 1
    assert n_ok >= 90
