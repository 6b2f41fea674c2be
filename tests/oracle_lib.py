"""ctypes access to the test oracle (oracle/_ref/liboracle.so) and to the
unmodified reference built into oracle/_ref (ugrep CLI, refscan harness).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
REF_UGREP = os.path.join(REF_DIR, "ugrep")
REF_SCAN = os.path.join(REF_DIR, "refscan")
LIB_PATH = os.path.join(REF_DIR, "liboracle.so")


class Match(C.Structure):
    _fields_ = [("line", C.c_uint64), ("offset", C.c_uint64), ("len", C.c_uint32), ("cap", C.c_uint32)]


MATCH_DTYPE = np.dtype([("line", "<u8"), ("offset", "<u8"), ("len", "<u4"), ("cap", "<u4")])


def build_oracle() -> None:
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "port"], check=True)


def have_reference() -> bool:
    return os.access(REF_UGREP, os.X_OK) and os.access(REF_SCAN, os.X_OK)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build_oracle()
        L = C.CDLL(LIB_PATH)
        L.ora_pattern_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.ora_pattern_destroy.argtypes = [C.c_void_p]
        L.ora_advance_kind.argtypes = [C.c_void_p]
        for name in ("ora_count_lines", "ora_count_matches"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.ora_find_all.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                   C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.ora_count_newlines.argtypes = [C.c_void_p, C.c_uint64]
        L.ora_count_newlines.restype = C.c_uint64
        L.ora_isutf8.argtypes = [C.c_void_p, C.c_uint64]
        L.ora_has_nul.argtypes = [C.c_void_p, C.c_uint64]
        L.ora_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ora_match_at.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def _as_u8(data) -> np.ndarray:
    if isinstance(data, (bytes, bytearray)):
        return np.frombuffer(bytes(data), dtype=np.uint8)
    return np.ascontiguousarray(data, dtype=np.uint8)


class OraclePattern:
    def __init__(self, path: str):
        self.handle = C.c_void_p()
        rc = lib().ora_pattern_load(path.encode(), C.byref(self.handle))
        if rc != 0:
            raise RuntimeError("ora_pattern_load(%s) -> %d" % (path, rc))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                lib().ora_pattern_destroy(self.handle)
                self.handle = None
        except Exception:  # interpreter shutdown: module globals may be gone already
            pass

    @property
    def advance(self) -> int:
        return lib().ora_advance_kind(self.handle)

    def count_lines(self, data) -> int:
        a = _as_u8(data)
        n = C.c_uint64()
        lib().ora_count_lines(self.handle, a.ctypes.data, a.size, C.byref(n))
        return n.value

    def count_matches(self, data) -> int:
        a = _as_u8(data)
        n = C.c_uint64()
        lib().ora_count_matches(self.handle, a.ctypes.data, a.size, C.byref(n))
        return n.value

    def find_all(self, data, base_offset: int = 0, base_line: int = 0) -> np.ndarray:
        a = _as_u8(data)
        n = C.c_uint64()
        cap = max(1024, a.size + 16)
        out = np.zeros(cap, dtype=MATCH_DTYPE)
        rc = lib().ora_find_all(self.handle, a.ctypes.data, a.size, base_offset, base_line,
                                out.ctypes.data, cap, C.byref(n))
        if rc != 0:
            raise RuntimeError("ora_find_all -> %d" % rc)
        return out[: n.value].copy()

    def candidates(self, data) -> np.ndarray:
        a = _as_u8(data)
        bm = np.zeros((a.size + 7) // 8, dtype=np.uint8)
        lib().ora_candidates(self.handle, a.ctypes.data, a.size, bm.ctypes.data)
        return np.unpackbits(bm, bitorder="little")[: a.size].astype(bool)

    def match_at(self, data, k: int):
        a = _as_u8(data)
        ln = C.c_uint64()
        cap = lib().ora_match_at(self.handle, a.ctypes.data, a.size, k, C.byref(ln))
        return cap, ln.value


def newlines(data) -> int:
    a = _as_u8(data)
    return lib().ora_count_newlines(a.ctypes.data, a.size)


def isutf8(data) -> bool:
    a = _as_u8(data)
    return bool(lib().ora_isutf8(a.ctypes.data, a.size))


def has_nul(data) -> bool:
    a = _as_u8(data)
    return bool(lib().ora_has_nul(a.ctypes.data, a.size))


# ---- the unmodified reference (oracle/_ref) ----

def ref_dump(popts: list[str], out_path: str) -> str:
    """Compile a pattern with the reference and write it as UGXP."""
    r = subprocess.run([REF_SCAN, "dump", *popts, "-o", out_path], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("refscan dump %r failed: %s" % (popts, r.stderr))
    return r.stderr.strip()


def ref_cli(args: list[str], data, filename: str | None = None) -> tuple[int, bytes]:
    """Run the reference ugrep CLI on one file holding `data`; returns (exit code, stdout)."""
    a = _as_u8(data)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, filename or "input.txt")
        a.tofile(path)
        r = subprocess.run([REF_UGREP, "--no-config", *args, path], capture_output=True)
        return r.returncode, r.stdout


def ref_scan(mode: str, popts: list[str], data) -> tuple[int, bytes]:
    """Run the reference library in-place (refscan scan)."""
    a = _as_u8(data)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "input.txt")
        a.tofile(path)
        r = subprocess.run([REF_SCAN, "scan", mode, *popts, path], capture_output=True)
        return r.returncode, r.stdout


def format_list(data, rec: np.ndarray) -> bytes:
    """What `ugrep -n -b -o` prints for these records."""
    a = _as_u8(data)
    raw = a.tobytes()
    out = []
    for r in rec:
        o = int(r["offset"])
        out.append(b"%d:%d:%s\n" % (int(r["line"]), o, raw[o:o + int(r["len"])]))
    return b"".join(out)
