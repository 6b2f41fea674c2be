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
