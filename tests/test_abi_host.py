"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/ugrep_b200.h
declares, the product fails loudly without a CUDA device (no CPU fallback), the UGXP container parses, and the
C++ mirror of reflex::Matcher (include/ugrep_b200/matcher.hpp) compiles against the library."""
import ctypes as C
import os
import re
import shutil
import subprocess

import pytest

import golden_lib as G
import oracle_lib as O

ROOT = O.ROOT
HEADER = os.path.join(ROOT, "include", "ugrep_b200.h")
LIB = os.path.join(ROOT, "ugrep_b200", "libugrep_b200.so")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"\b(ugx_[a-z_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        from ugrep_b200 import build
        build.build()
    return C.CDLL(LIB)


def test_library_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libugrep_b200.so does not export %s" % n
    from ugrep_b200 import api
    assert set(api.EXPORTS) == set(names)
    assert lib.ugx_abi_version() == 2


def test_struct_layouts_match_the_header():
    """the ctypes mirrors in api.py and the oracle's pattern loader agree with the header's sizes"""
    text = open(HEADER).read()
    assert "UGX_BTAP 2048" in text and "UGX_HASH 4096" in text
    # ugx_prefilter: 12 u32 + 256 + 256 + 2048 + 4096 + 4096 + 32 + 32 + 256
    want = 12 * 4 + 256 + 256 + 2048 + 4096 + 4096 + 32 + 32 + 256
    hdr = open(G.pattern_path("c1"), "rb").read(24)
    assert hdr[:8] == b"UGXP\x01\x00\x00\x00"
    nop, regex_len, pf_size, flags = (int.from_bytes(hdr[8 + 4 * i:12 + 4 * i], "little") for i in range(4))
    assert pf_size == want
    assert os.path.getsize(G.pattern_path("c1")) == 24 + pf_size + 4 * nop + regex_len


def test_no_cpu_fallback_without_a_device(lib):
    """on a box without a usable GPU every entry point fails with UGX_E_CUDA; nothing scans on the CPU"""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    lib.ugx_last_error.restype = C.c_char_p
    h = C.c_void_p()
    rc = lib.ugx_pattern_load(os.fsencode(G.pattern_path("c1")), 0, C.byref(h))
    assert rc == 3, rc  # UGX_E_CUDA
    assert b"no CPU fallback" in lib.ugx_last_error()
    s = C.c_void_p()
    assert lib.ugx_scanner_create(0, None, C.byref(s)) == 3


def test_out_of_scope_patterns_are_rejected_before_any_cuda_call(lib):
    """REDO/HEAD/TAIL opcodes, a live newline transition ...: UGX_E_UNSUPPORTED from the host-side export"""
    lib.ugx_last_error.restype = C.c_char_p
    pf = (C.c_uint8 * 11128)()
    # one state: TAKE 1, then a GOTO on '\n' to itself -> newline transition is live
    opc = (C.c_uint32 * 3)(0xFE000001, (0x0A << 24) | (0x0A << 16) | 0, 0x00FFFFFF)
    h = C.c_void_p()
    rc = lib.ugx_pattern_create(opc, 3, pf, 0, 0, C.byref(h))
    assert rc == 2, (rc, lib.ugx_last_error())
    # HEAD opcode (lookahead)
    opc = (C.c_uint32 * 2)(0xFB000000, 0x00FFFFFF)
    assert lib.ugx_pattern_create(opc, 2, pf, 0, 0, C.byref(h)) == 2


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_cpp_facade_compiles_and_fails_loudly_without_gpu(tmp_path, lib):
    exe = str(tmp_path / "facade_test")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-o", exe,
           os.path.join(ROOT, "tests", "cpp", "facade_test.cpp"), "-L" + os.path.dirname(LIB), "-lugrep_b200",
           "-Wl,-rpath," + os.path.dirname(LIB)]
    for d in ("/usr/local/cuda/lib64",):
        if os.path.isdir(d):
            cmd += ["-L" + d, "-Wl,-rpath," + d]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except ImportError:
        has_gpu = False
    if not has_gpu:
        f = tmp_path / "in.txt"
        f.write_bytes(b"Sherlock Holmes\n")
        r = subprocess.run([exe, G.pattern_path("c1"), "cl", str(f)], capture_output=True, text=True)
        assert r.returncode == 4 and "no CPU fallback" in r.stderr
