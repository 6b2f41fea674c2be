"""The selection rule the span kernels (csrc/span_scan.cu) are built on, checked on the CPU against the oracle's
find loop for every golden pattern in the kernels' scope (no META edges, no option W, start state not accepting):

    A        = { p : some prefilter candidate k >= p has only look-back bytes in [p, k) and k - p <= lbk }
    D(p)     = length of the anchored longest match at p (0 = none)
    matches  = the chain  c -> first p >= c with p in A and D(p) > 0,  c := p + D(p)

i.e. Matcher::match(FIND) with its look-back / retry rules (lib/matcher.cpp:54-70, 627-658) attempts exactly the
positions of A at or after the cursor, in increasing order, until one accepts.  Attempts that fail after reading up
to the end of the buffer take the reference's end-of-buffer path (:623 `if (!at_end())`), which the kernels hand to
the line-at-a-time form; the model below does the same by stopping the comparison at the first such attempt."""
import struct

import numpy as np
import pytest

import golden_lib as G
import oracle_lib as O


def _fields(path):
    raw = open(path, "rb").read()
    nop, _, _, flags = struct.unpack_from("<IIII", raw, 8)
    f = struct.unpack_from("<12I", raw, 24)
    names = ("len", "min", "pin", "lcp", "lcs", "bmd", "npy", "one", "bol", "lbk", "lbm", "cut")
    d = dict(zip(names, f))
    off = 24 + 48 + 256 + 256 + 2048 + 4096 + 4096
    d["cbk"] = np.unpackbits(np.frombuffer(raw, dtype=np.uint8, count=32, offset=off), bitorder="little").astype(bool)
    d["flags"] = flags
    pre = 24 + 48 + 256 + 256 + 2048 + 4096 + 4096 + 32 + 32 + 256
    d["opc"] = np.frombuffer(raw, dtype="<u4", count=nop, offset=pre)
    return d


def _in_scope(f):
    opc = f["opc"]
    goto = (opc << np.uint32(8)) >= (opc & np.uint32(0xff000000))
    code = opc >> np.uint32(24)
    has_meta = bool(np.any(~goto & (code >= 1) & (code <= 0x0c)))
    acc0 = (not goto[0]) and code[0] == 0xfe
    nonadv = f["len"] == 0 and f["min"] == 0 and (f["flags"] & 1)
    return not has_meta and not (f["flags"] & 2) and not acc0 and not nonadv and f["lbk"] in (0, 0xffff)


def _model(op, f, data):
    a = np.frombuffer(data, dtype=np.uint8)
    n = len(a)
    cand = op.candidates(a)
    A = cand.copy()
    if f["lbk"]:
        R = f["cbk"][a]
        hot = False
        for p in range(n - 1, -1, -1):
            hot = bool(cand[p]) or (bool(R[p]) and hot)
            A[p] = hot
    out = []
    c = 0
    for p in np.flatnonzero(A):
        p = int(p)
        if p < c:
            continue
        cap, ln = op.match_at(a, p)
        if cap and ln:
            out.append((p, ln, cap))
            c = p + ln
    return out


NAMES = [n for n in G.pattern_names() if _in_scope(_fields(G.pattern_path(n)))]


def test_scope_is_not_empty():
    assert {"c5", "c3b", "c2", "c4", "pin_pma_lb", "pin1_one_lb", "two_caps", "dotstar", "digits_U"} <= set(NAMES)


@pytest.mark.parametrize("name", NAMES)
def test_chain_over_attempt_set_equals_find(name):
    path = G.pattern_path(name)
    op = O.OraclePattern(path)
    f = _fields(path)
    for case, data in G.cases(name):
        if len(data) > 20000:
            data = data[:20000]
            data = data[:data.rfind(b"\n") + 1]
        if not data.endswith(b"\n"):
            data = data + b"\n"   # the end-of-buffer path is the line-at-a-time form's job (see the module docstring)
        want = op.find_all(data)
        got = _model(op, f, data)
        assert [(int(r["offset"]), int(r["len"]), int(r["cap"])) for r in want] == got, (name, case["input"])


# ---- the k-gram viability table (csrc/pattern_host.cpp build_viability): a position it rejects never starts a match

def _viability(path):
    import ctypes as C
    from ugrep_b200 import api
    L = api.lib()
    raw = open(path, "rb").read()
    nop = struct.unpack_from("<I", raw, 8)[0]
    pre = 24 + 48 + 256 + 256 + 2048 + 4096 + 4096 + 32 + 32 + 256
    opc = (C.c_uint32 * nop).from_buffer_copy(raw[pre:pre + 4 * nop])
    k, stride = C.c_uint32(), C.c_uint32()
    t01 = (C.c_uint32 * 256)()
    t23 = (C.c_uint32 * 256)()
    pair = (C.c_uint8 * 16384)()
    bits = (C.c_uint32 * 8192)()
    npair, words = C.c_uint32(), C.c_uint32()
    L.ugx_viability_describe.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    rc = L.ugx_viability_describe(opc, nop, C.byref(k), C.byref(stride), t01, t23, pair, 16384, C.byref(npair), bits, 8192,
                                  C.byref(words))
    assert rc == 0
    return (k.value, stride.value, np.frombuffer(t01, dtype=np.uint32).copy(), np.frombuffer(t23, dtype=np.uint32).copy(),
            np.frombuffer(pair, dtype=np.uint8)[:npair.value].copy(), np.frombuffer(bits, dtype=np.uint32)[:words.value].copy())


def _viable_mask(v, a):
    """the device's test (csrc/viability.cuh viable16) for positions 0 .. len(a) - 4"""
    k, stride, t01, t23, pair, bits = v
    m = len(a) - 3
    code = pair[(t01[a[0:m]] & 0xffff).astype(np.int64) + (t01[a[1:m + 1]] >> 16)].astype(np.int64)
    idx = code * stride + (t23[a[2:m + 2]] & 0xffff) + (t23[a[3:m + 3]] >> 16)
    return ((bits[idx >> 5] >> (idx & 31).astype(np.uint32)) & 1).astype(bool)


@pytest.mark.parametrize("name", NAMES)
def test_viability_never_rejects_a_match_start(name):
    path = G.pattern_path(name)
    op = O.OraclePattern(path)
    v = _viability(path)
    assert v[0] in (0, 2, 3, 4)
    if v[0] == 0:
        return
    for case, data in G.cases(name):
        a = np.frombuffer(data[:30000], dtype=np.uint8)
        if len(a) < 8:
            continue
        viable = _viable_mask(v, a)
        for p in np.flatnonzero(~viable)[:4000]:
            cap, ln = op.match_at(a, int(p))
            assert not (cap and ln), (name, case["input"], int(p))


def test_viability_is_selective_on_the_configs():
    """the table is worth its lookups: on the c5 corpus it spares most of the attempt set (a third of all positions)"""
    import os
    from ugrep_b200 import corpus
    v = _viability(os.path.join(O.ROOT, "ugrep_b200", "patterns", "c5.ugxp"))
    assert v[0] == 4
    assert _viable_mask(v, corpus.block("c5", 100000)).mean() < 0.08
