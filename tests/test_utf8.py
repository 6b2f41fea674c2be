"""reflex::isutf8 (ugrep's binary-file test): the oracle's restatement against the golden vectors made
with the unmodified reference (tools/make_utf8_golden.py), and — on the GPU — utf8_check_kernel against both."""
import base64
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from ugrep_b200 import corpus

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "utf8.json")


def cases():
    with open(GOLDEN) as f:
        return [(base64.b64decode(d), bool(v)) for d, v in json.load(f)["cases"]]


def test_oracle_isutf8_equals_the_reference_vectors():
    cs = cases()
    assert len(cs) > 1500 and 0.2 < sum(v for _, v in cs) / len(cs) < 0.8
    for data, want in cs:
        assert O.isutf8(data) == want, data[:60]


@pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref (the built reference) is not present")
def test_oracle_isutf8_equals_the_live_reference(tmp_path):
    import subprocess
    rng = np.random.default_rng(9)
    text = corpus.block("c4", 4000).tobytes()
    paths, datas = [], []
    for i in range(120):
        a = bytearray(text[:int(rng.integers(1, 4000))])
        if i % 3:
            a[int(rng.integers(0, len(a)))] = int(rng.integers(0, 256))
        p = tmp_path / ("f%d" % i)
        p.write_bytes(bytes(a))
        paths.append(str(p))
        datas.append(bytes(a))
    r = subprocess.run([O.REF_SCAN, "isutf8", *paths], capture_output=True, text=True, check=True)
    assert [bool(int(x)) for x in r.stdout.split()] == [O.isutf8(d) for d in datas]


@pytest.mark.gpu
def test_gpu_check_text_equals_reference_vectors_and_oracle():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the -m gpu tests must run on the B200 box")
    from ugrep_b200 import api
    sc = api.Scanner(0)
    for data, want in cases():
        got = sc.check_text(data)
        assert got["is_utf8"] == want, data[:60]
        assert got["has_nul"] == (0 in data), data[:60]
    # sequences cut by every span / chunk boundary, device-resident and misaligned views
    text = corpus.block("c4", 3 << 20)
    dev = torch.from_numpy(text).cuda()
    rng = np.random.default_rng(2)
    for _ in range(40):
        lo = int(rng.integers(0, 5000))
        hi = int(rng.integers(lo + 1, len(text)))
        assert sc.check_text(dev[lo:hi])["is_utf8"] == O.isutf8(text[lo:hi]), (lo, hi)
    big = np.tile(text, 40)
    assert sc.check_text(big)["is_utf8"] and not sc.check_text(big)["has_nul"]
    for at, val in ((len(big) - 1, 0xE2), (len(big) // 2, 0x00), (777777, 0xC0), (511, 0xFF), (16 * 31 + 15, 0xF5)):
        b2 = big.copy()
        b2[at] = val
        got = sc.check_text(torch.from_numpy(b2).cuda())
        assert got["is_utf8"] == O.isutf8(b2) and got["has_nul"] == O.has_nul(b2), (at, val)
