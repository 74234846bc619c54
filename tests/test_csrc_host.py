"""CPU check of the host+device logic in bzip3_b200/csrc (CRC algebra, single-lane LZP and CM coder):
the same functions the CUDA kernels call are compiled with g++ and compared with the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from bzip3_b200 import synth
from tests import refs

ROOT = refs.ROOT
SO = os.path.join(ROOT, "tests", "_build", "libhostcheck.so")
SRC = os.path.join(ROOT, "tests", "native", "host_check.cpp")


@pytest.fixture(scope="module")
def H():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    deps = [SRC] + [os.path.join(ROOT, "bzip3_b200", "csrc", f) for f in ("common.cuh", "crc.cuh", "lzp.cuh", "cm.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-fPIC", "-shared", "-fvisibility=hidden",
                               "-o", SO, SRC])
    L = C.CDLL(SO)
    L.hc_crc_chunked.restype = C.c_uint32
    L.hc_crc_chunked.argtypes = [refs.u8p, C.c_uint32, C.c_uint32, C.c_uint32]
    L.hc_lzp_encode.restype = C.c_int32
    L.hc_lzp_encode.argtypes = [refs.u8p, C.c_int32, refs.u8p]
    L.hc_lzp_decode.restype = C.c_int32
    L.hc_lzp_decode.argtypes = [refs.u8p, C.c_int32, refs.u8p, C.c_int32]
    L.hc_cm_encode.restype = C.c_int32
    L.hc_cm_encode.argtypes = [refs.u8p, C.c_int32, refs.u8p]
    L.hc_cm_decode.restype = None
    L.hc_cm_decode.argtypes = [refs.u8p, C.c_int32, refs.u8p, C.c_int32]
    return L


CASES = synth.edge_cases()
IDS = [c[0] for c in CASES]


def arr(b):
    return np.frombuffer(bytes(b), dtype=np.uint8).copy()


@pytest.mark.parametrize("name,data", CASES, ids=IDS)
def test_crc_chunk_algebra(H, name, data):
    O = refs.oracle()
    a = arr(data)
    want = O.orc_crc32(1, refs.ptr(a), len(a))
    for chunk in (2048, 7, 1 << 20):
        assert H.hc_crc_chunked(refs.ptr(a), len(a), 1, chunk) == want


@pytest.mark.parametrize("name,data", CASES, ids=IDS)
def test_lzp_serial(H, name, data):
    O = refs.oracle()
    a = arr(data)
    n = len(a)
    pad = np.zeros(n + 64, np.uint8)
    pad[:n] = a
    o1 = np.zeros(n + 64, np.uint8)
    o2 = np.zeros(n + 64, np.uint8)
    lut = np.zeros(1 << 18, np.int32)
    r1 = H.hc_lzp_encode(refs.ptr(pad), n, refs.ptr(o1))
    r2 = O.orc_lzp_encode(refs.ptr(pad), n, refs.ptr(o2), lut.ctypes.data_as(refs.i32p))
    assert r1 == r2
    if r1 > 0:
        assert bytes(o1[:r1]) == bytes(o2[:r2])
        for cut in (r1, r1 - 1, r1 // 2, 4, 3):
            d1 = np.zeros(refs.bound(n) + 64, np.uint8)
            d2 = np.zeros(refs.bound(n) + 64, np.uint8)
            s1 = H.hc_lzp_decode(refs.ptr(o1), cut, refs.ptr(d1), refs.bound(n))
            s2 = O.orc_lzp_decode(refs.ptr(o2), cut, refs.ptr(d2), refs.bound(n), lut.ctypes.data_as(refs.i32p))
            assert s1 == s2
            if s1 > 0:
                assert bytes(d1[:s1]) == bytes(d2[:s2])


@pytest.mark.parametrize("name,data", CASES, ids=IDS)
def test_cm_serial(H, name, data):
    O = refs.oracle()
    a = arr(data)
    n = len(a)
    o1 = np.zeros(2 * n + 64, np.uint8)
    o2 = np.zeros(2 * n + 64, np.uint8)
    r1 = H.hc_cm_encode(refs.ptr(a), n, refs.ptr(o1))
    r2 = O.orc_cm_encode(refs.ptr(a), n, refs.ptr(o2))
    assert r1 == r2 and bytes(o1[:r1]) == bytes(o2[:r2])
    for insize in (r1, max(r1 - 2, 0), 0):
        d1 = np.zeros(n + 8, np.uint8)
        d2 = np.zeros(n + 8, np.uint8)
        H.hc_cm_decode(refs.ptr(o1), insize, refs.ptr(d1), n)
        O.orc_cm_decode(refs.ptr(o2), insize, refs.ptr(d2), n)
        assert bytes(d1) == bytes(d2)
