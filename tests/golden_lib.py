"""Access to the committed golden fixtures (tests/golden/, written by tools/make_golden.py
with the unmodified reference).  Shared by the CPU oracle tests and the GPU parity tests."""
from __future__ import annotations

import base64
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(HERE, "golden")
PATTERN_DIR = os.path.join(GOLDEN_DIR, "patterns")

_golden = None
_blocks = {}


def golden() -> dict:
    global _golden
    if _golden is None:
        with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
            _golden = json.load(f)
    return _golden


def pattern_names() -> list[str]:
    return sorted(golden()["patterns"])


def pattern_path(name: str) -> str:
    return os.path.join(PATTERN_DIR, name + ".ugxp")


def input_bytes(ref) -> bytes:
    """ref = ["edge", index] or ["block", corpus name]"""
    kind, key = ref
    g = golden()
    if kind == "edge":
        return base64.b64decode(g["edge_inputs"][key])
    if key not in _blocks:
        from ugrep_b200 import corpus
        meta = g["blocks"][key]
        data = corpus.block(key, meta["nbytes"]).tobytes()
        if hashlib.sha256(data).hexdigest() != meta["sha256"]:
            raise RuntimeError("corpus block %s drifted from the one the goldens were made on" % key)
        _blocks[key] = data
    return _blocks[key]


def cases(name: str):
    for c in golden()["patterns"][name]["cases"]:
        yield c, input_bytes(c["input"])


def format_list(data: bytes, rec: np.ndarray) -> bytes:
    """What `ugrep -n -b -o` prints for these records (src/output.cpp:339-402: line:offset:text)."""
    out = []
    for r in rec:
        o = int(r["offset"])
        out.append(b"%d:%d:%s\n" % (int(r["line"]), o, data[o:o + int(r["len"])]))
    return b"".join(out)


def check_list(case: dict, data: bytes, rec: np.ndarray) -> None:
    text = format_list(data, rec)
    assert len(text) == case["list_bytes"], (case["input"], len(text), case["list_bytes"])
    assert hashlib.sha256(text).hexdigest() == case["list_sha256"], case["input"]
    if "list" in case:
        assert text == base64.b64decode(case["list"])
