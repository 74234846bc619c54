"""The multi-kernel stages -- CRC, mRLE, suffix-array BWT (radix sort, scans, prefix doubling), inverse BWT -- run on the
CPU thread-block emulator with their REAL host-side launch sequences (bzip3_b200/csrc/*.cuh compiled by g++, see
tests/native/emu_stages.cpp) and compared bit-for-bit with the oracle.  Same purpose and same limits as
tests/test_emu_kernels.py: it proves the algorithms and their launch logic in a container without a GPU, not race
freedom and nothing about speed; the GPU parity tests remain the gate."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from bzip3_b200 import synth
from tests import refs

ROOT = refs.ROOT
SO = os.path.join(ROOT, "tests", "_build", "libemustages.so")
SRCS = [os.path.join(ROOT, "tests", "native", f) for f in ("emu_stages.cpp", "cta_emu.cpp")]
DEPS = SRCS + [os.path.join(ROOT, "tests", "native", "cta_emu.h")] + [
    os.path.join(ROOT, "bzip3_b200", "csrc", f) for f in ("common.cuh", "scan.cuh", "radix_sort.cuh", "crc.cuh", "mrle.cuh",
                                                          "sufsort.cuh", "unbwt.cuh", "lzp.cuh", "lzp_scan.cuh")]
_lib = None


def emu():
    global _lib
    if _lib is None:
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-x", "c++",
                                   "-o", SO] + SRCS)
        L = C.CDLL(SO)
        L.emu_stage_crc.restype = C.c_uint32
        L.emu_stage_crc.argtypes = [refs.u8p, C.c_uint32, C.c_uint32]
        L.emu_stage_rle_encode.restype = C.c_int32
        L.emu_stage_rle_encode.argtypes = [refs.u8p, C.c_uint32, refs.u8p]
        L.emu_stage_rle_decode.restype = C.c_int
        L.emu_stage_rle_decode.argtypes = [refs.u8p, C.c_uint32, refs.u8p, C.c_uint32]
        L.emu_stage_bwt.restype = C.c_int32
        L.emu_stage_bwt.argtypes = [refs.u8p, C.c_uint32, refs.u8p]
        L.emu_stage_unbwt.restype = C.c_int
        L.emu_stage_unbwt.argtypes = [refs.u8p, C.c_uint32, C.c_int32, refs.u8p]
        L.emu_stage_lzp_encode.restype = C.c_int32
        L.emu_stage_lzp_encode.argtypes = [refs.u8p, C.c_int32, refs.u8p]
        _lib = L
    return _lib


def arr(b):
    return np.frombuffer(bytes(b), dtype=np.uint8).copy()


CASES = [(name, arr(d)) for name, d in synth.edge_cases()]
IDS = [c[0] for c in CASES]


@pytest.mark.parametrize("name,data", CASES, ids=IDS)
def test_crc_kernel(name, data):
    E, O = emu(), refs.oracle()
    a = data[:20000]
    pad = np.zeros(len(a) + 16, np.uint8)
    pad[:len(a)] = a
    for init in (1, 0xDEADBEEF):
        assert E.emu_stage_crc(refs.ptr(pad), len(a), init) == O.orc_crc32(init, refs.ptr(pad), len(a))


@pytest.mark.parametrize("name,data", CASES, ids=IDS)
def test_mrle_kernels(name, data):
    E, O = emu(), refs.oracle()
    a = data[:6000]
    n = len(a)
    pad = np.zeros(n + 16, np.uint8)
    pad[:n] = a
    want = np.zeros(2 * n + 64, np.uint8)
    got = np.zeros(2 * n + 64, np.uint8)
    rw = O.orc_mrle_encode(refs.ptr(pad), n, refs.ptr(want))
    rg = E.emu_stage_rle_encode(refs.ptr(pad), n, refs.ptr(got))
    assert rg == rw
    assert bytes(got[:rg]) == bytes(want[:rw])
    for cut in (rw, rw - 1, rw // 2, 33, 32, 31):   # truncated input: the staleness quirk of the reference included
        if cut < 0:
            continue
        dw = np.zeros(n + 8, np.uint8)
        dg = np.zeros(n + 8, np.uint8)
        ew = O.orc_mrle_decode(refs.ptr(want), refs.ptr(dw), n, cut)
        eg = E.emu_stage_rle_decode(refs.ptr(want), cut, refs.ptr(dg), n)
        assert eg == ew, (cut, eg, ew)
        assert bytes(dg[:n]) == bytes(dw[:n]), cut


@pytest.mark.parametrize("name,data", CASES, ids=IDS)
def test_bwt_and_inverse(name, data):
    E, O = emu(), refs.oracle()
    a = data[:1500]
    n = len(a)
    pad = np.zeros(n + 64, np.uint8)
    pad[:n] = a
    want = np.zeros(n + 64, np.uint8)
    got = np.zeros(n + 64, np.uint8)
    iw = O.orc_bwt(refs.ptr(pad), refs.ptr(want), n)
    ig = E.emu_stage_bwt(refs.ptr(pad), n, refs.ptr(got))
    assert ig == iw
    assert bytes(got[:n]) == bytes(want[:n])
    if n == 0:
        return
    back = np.zeros(n + 64, np.uint8)
    assert E.emu_stage_unbwt(refs.ptr(want), n, iw, refs.ptr(back)) == 0
    assert bytes(back[:n]) == bytes(a)


def test_unbwt_on_corrupt_input_matches_oracle():
    """A wrong primary index or a damaged last column still yields the reference's deterministic bytes."""
    E, O = emu(), refs.oracle()
    rng = np.random.default_rng(77)
    base = synth.zipf_text(1500, seed=5)
    n = len(base)
    L0 = np.zeros(n + 64, np.uint8)
    idx = O.orc_bwt(refs.ptr(base), refs.ptr(L0), n)
    for trial in range(12):
        L = L0.copy()
        k = idx
        if trial % 3 != 0:
            for _ in range(int(rng.integers(1, 6))):
                L[int(rng.integers(0, n))] = int(rng.integers(0, 256))
        if trial % 2 == 1:
            k = int(rng.integers(1, n + 1))
        want = np.zeros(n + 64, np.uint8)
        got = np.zeros(n + 64, np.uint8)
        sw = O.orc_unbwt(refs.ptr(L), refs.ptr(want), n, k)
        sg = E.emu_stage_unbwt(refs.ptr(L), n, k, refs.ptr(got))
        assert (sg == 0) == (sw == 0), (trial, sg, sw)
        if sw == 0:
            assert bytes(got[:n]) == bytes(want[:n]), trial
    for bad in (0, -3, n + 1):
        got = np.zeros(n + 64, np.uint8)
        assert E.emu_stage_unbwt(refs.ptr(L0), n, bad, refs.ptr(got)) != 0


def _lzp_cases():
    rng = np.random.default_rng(5)
    rep = np.tile(rng.integers(0, 256, 700, dtype=np.uint8), 60)          # long matches, period 700
    runs = np.repeat(rng.integers(0, 3, 300, dtype=np.uint8), 150)        # every context repeats, matches back to back
    esc = rng.choice(np.array([0xF2, 0x41, 0x42], np.uint8), 30000)       # escape bytes with live slots
    mix = np.concatenate([rep[:9000], rng.integers(0, 256, 5000, dtype=np.uint8), rep[:9000], runs[:6000], esc[:4000]])
    near = synth.source_corpus(40 << 10, seed=31)
    dup = np.concatenate([near, rng.integers(0, 256, 3000, dtype=np.uint8), near[5000:30000], near[:9000]])   # far repeats
    extra = [("periodic_42k", rep), ("runs_45k", runs), ("escapes_30k", esc), ("mix_33k", mix), ("far_repeats_77k", dup),
             ("source_96k", synth.source_corpus(96 << 10, seed=21)), ("log_64k", synth.log_stream(64 << 10, seed=22)),
             ("zipf_64k", synth.zipf_text(64 << 10, seed=23))]
    return CASES + [(n, np.ascontiguousarray(d)) for n, d in extra]


LZP_CASES = _lzp_cases()


@pytest.mark.parametrize("name,data", LZP_CASES, ids=[c[0] for c in LZP_CASES])
def test_lzp_scan_encoder(name, data):
    """hash keys -> sort -> links -> candidate codes -> commit engine, against the oracle's lzp_encode_block"""
    E, O = emu(), refs.oracle()
    n = len(data)
    pad = np.zeros(n + 64, np.uint8)
    pad[:n] = data
    want = np.zeros(n + 64, np.uint8)
    got = np.zeros(n + 64, np.uint8)
    lut = np.zeros(1 << 18, np.int32)
    rw = O.orc_lzp_encode(refs.ptr(pad), n, refs.ptr(want), lut.ctypes.data_as(refs.i32p))
    rg = E.emu_stage_lzp_encode(refs.ptr(pad), n, refs.ptr(got))
    assert rg == rw, (rg, rw)
    if rw > 0:
        d = np.nonzero(got[:rw] != want[:rw])[0]
        assert len(d) == 0, ("first difference at", int(d[0]), "of", rw)
