"""Pins the CPU oracle (oracle/bz3_oracle.c) against
  (1) the reference's golden vector  examples/shakespeare.txt(.bz3)  (Makefile.am:81-83),
  (2) known answers computed by the compiled reference (SURVEY.md 8c / BASELINE.md section 2),
  (3) the unmodified reference compiled from /root/reference into oracle/_ref (when present),
stage by stage and block by block.  CPU only."""
import ctypes as C
import hashlib
import os
import struct

import numpy as np
import pytest

from bzip3_b200 import synth
from tests import refs

O = refs.oracle()
needs_ref = pytest.mark.skipif(not refs.have_ref(), reason="oracle/_ref not built (no /root/reference here)")
CASES = synth.edge_cases()
IDS = [c[0] for c in CASES]


def arr(b):
    return np.frombuffer(bytes(b), dtype=np.uint8).copy() if len(b) else np.zeros(1, np.uint8)[:0].copy()


def golden(name):
    with open(os.path.join(refs.GOLDEN, name), "rb") as f:
        return f.read()


# ------------------------------------------------------------------ golden vectors / KATs
def parse_cli_stream(blob):
    """bzip3 CLI container (src/main.c:174-179, :249-253): 'BZ3v1' u32 block_size, then [csize][osize][payload]*"""
    assert blob[:5] == b"BZ3v1"
    bs = struct.unpack("<I", blob[5:9])[0]
    at, blocks = 9, []
    while at < len(blob):
        cs, osz = struct.unpack("<ii", blob[at:at + 8])
        blocks.append((blob[at + 8:at + 8 + cs], osz))
        at += 8 + cs
    return bs, blocks


def test_golden_decode_shakespeare():
    bs, blocks = parse_cli_stream(golden("shakespeare.txt.bz3"))
    assert bs == 4 << 20 and len(blocks) == 2
    plain = b""
    for enc, osz in blocks:
        out, r, e = refs.oracle_decode_block(enc, osz, bs)
        assert r == osz and e == 0
        plain += out
    assert plain == golden("shakespeare.txt")


def test_kat_encode_shakespeare_b8():
    # computed by the compiled reference: bzip3 -e -b 8 examples/shakespeare.txt
    data = golden("shakespeare.txt")
    enc, r, e = refs.oracle_encode_block(data, 8 << 20)
    assert e == 0 and r == 1229797
    crc, idx, model, lzp = struct.unpack("<IiBi", enc[:13])
    assert (crc, idx, model, lzp) == (0x18A1405E, 1980452, 2, 5314513)
    stream = b"BZ3v1" + struct.pack("<I", 8 << 20) + struct.pack("<ii", r, len(data)) + enc
    assert len(stream) == 1229814
    assert hashlib.sha256(stream).hexdigest() == "6ed262b586d6e58aa00429ac1776b3ca29ca59283c2008f43378fef87755cee6"


def test_crc_known_values():
    assert O.orc_crc32(1, refs.ptr(arr(b"")), 0) == 1
    d = arr(golden("shakespeare.txt"))
    assert O.orc_crc32(1, refs.ptr(d), len(d)) == 0x18A1405E


@pytest.mark.parametrize("name", ["63_byte_file.bin", "65_byte_file.bin"])
def test_seed_files_roundtrip(name):
    data = golden(name)
    enc, r, e = refs.oracle_encode_block(data, 65 * 1024)
    assert r > 0
    if len(data) < 64:
        assert r == len(data) + 8 and struct.unpack("<i", enc[4:8])[0] == -1
    out, r2, e2 = refs.oracle_decode_block(enc, len(data), 65 * 1024)
    assert out == data


# ------------------------------------------------------------------ stage-level differential vs the reference
@needs_ref
@pytest.mark.parametrize("name,data", CASES, ids=IDS)
def test_stage_crc(name, data):
    R = refs.ref_stages()
    a = arr(data)
    assert O.orc_crc32(1, refs.ptr(a), len(a)) == R.ref_crc32(1, refs.ptr(a), len(a))


@needs_ref
@pytest.mark.parametrize("name,data", CASES, ids=IDS)
def test_stage_mrle(name, data):
    R = refs.ref_stages()
    a = arr(data)
    n = len(a)
    o1 = np.zeros(2 * n + 64, np.uint8)
    o2 = np.zeros(2 * n + 64, np.uint8)
    r1 = O.orc_mrle_encode(refs.ptr(a), n, refs.ptr(o1))
    r2 = R.ref_mrlec(refs.ptr(a.copy()), n, refs.ptr(o2))
    assert r1 == r2 and bytes(o1[:r1]) == bytes(o2[:r2])
    d1 = np.zeros(n + 8, np.uint8)
    d2 = np.zeros(n + 8, np.uint8)
    assert O.orc_mrle_decode(refs.ptr(o1), refs.ptr(d1), n, r1) == R.ref_mrled(refs.ptr(o2), refs.ptr(d2), n, r2) == 0
    assert bytes(d1[:n]) == bytes(d2[:n]) == bytes(a)
    # truncated / wrong-length inputs must agree on the failure flag and on the bytes produced
    for cut in (r1 - 1, r1 // 2, 33, 32, 31):
        if cut < 0:
            continue
        d1[:] = 0
        d2[:] = 0
        e1 = O.orc_mrle_decode(refs.ptr(o1), refs.ptr(d1), n, cut)
        e2 = R.ref_mrled(refs.ptr(o2), refs.ptr(d2), n, cut)
        assert e1 == e2 and bytes(d1) == bytes(d2)


@needs_ref
@pytest.mark.parametrize("name,data", CASES, ids=IDS)
def test_stage_lzp(name, data):
    R = refs.ref_stages()
    a = arr(data)
    n = len(a)
    pad = np.zeros(n + 64, np.uint8)
    pad[:n] = a
    o1 = np.zeros(n + 64, np.uint8)
    o2 = np.zeros(n + 64, np.uint8)
    lut1 = np.zeros(1 << 18, np.int32)
    lut2 = np.zeros(1 << 18, np.int32)
    l1 = lut1.ctypes.data_as(refs.i32p)
    l2 = lut2.ctypes.data_as(refs.i32p)
    r1 = O.orc_lzp_encode(refs.ptr(pad), n, refs.ptr(o1), l1)
    r2 = R.ref_lzp_compress(refs.ptr(pad), refs.ptr(o2), n, l2)
    assert r1 == r2
    if r1 > 0:
        assert bytes(o1[:r1]) == bytes(o2[:r2])
        cap = refs.bound(n) + 64
        d1 = np.zeros(cap, np.uint8)
        d2 = np.zeros(cap, np.uint8)
        s1 = O.orc_lzp_decode(refs.ptr(o1), r1, refs.ptr(d1), refs.bound(n), l1)
        s2 = R.ref_lzp_decompress(refs.ptr(o2), refs.ptr(d2), r2, refs.bound(n), l2)
        assert s1 == s2 == n and bytes(d1[:n]) == bytes(a)
        for cut in (r1 - 1, r1 - 2, r1 // 2, 5, 4, 3):
            if cut < 0:
                continue
            d1[:] = 0
            d2[:] = 0
            s1 = O.orc_lzp_decode(refs.ptr(o1), cut, refs.ptr(d1), refs.bound(n), l1)
            s2 = R.ref_lzp_decompress(refs.ptr(o2), refs.ptr(d2), cut, refs.bound(n), l2)
            assert s1 == s2
            if s1 > 0:
                assert bytes(d1[:s1]) == bytes(d2[:s2])


@needs_ref
@pytest.mark.parametrize("name,data", CASES, ids=IDS)
def test_stage_bwt(name, data):
    R = refs.ref_stages()
    a = arr(data)
    n = len(a)
    u1 = np.zeros(n + 8, np.uint8)
    u2 = np.zeros(n + 8, np.uint8)
    A = np.zeros(n + 256, np.int32)
    i1 = O.orc_bwt(refs.ptr(a), refs.ptr(u1), n)
    i2 = R.ref_bwt(refs.ptr(a), refs.ptr(u2), A.ctypes.data_as(refs.i32p), n)
    assert i1 == i2 and bytes(u1[:n]) == bytes(u2[:n])
    t1 = np.zeros(n + 8, np.uint8)
    t2 = np.zeros(n + 8, np.uint8)
    A[:] = 0
    assert O.orc_unbwt(refs.ptr(u1), refs.ptr(t1), n, i1) == R.ref_unbwt(refs.ptr(u2), refs.ptr(t2),
                                                                         A.ctypes.data_as(refs.i32p), n, i2) == 0
    assert bytes(t1[:n]) == bytes(t2[:n]) == bytes(a)
    for bad in (0, -3, n + 1):
        if n >= 2:
            A[:] = 0
            assert O.orc_unbwt(refs.ptr(u1), refs.ptr(t1), n, bad) == R.ref_unbwt(
                refs.ptr(u2), refs.ptr(t2), A.ctypes.data_as(refs.i32p), n, bad) == -1


@needs_ref
@pytest.mark.parametrize("name,data", CASES, ids=IDS)
def test_stage_cm(name, data):
    R = refs.ref_stages()
    a = arr(data)
    n = len(a)
    o1 = np.zeros(2 * n + 64, np.uint8)
    o2 = np.zeros(2 * n + 64, np.uint8)
    r1 = O.orc_cm_encode(refs.ptr(a), n, refs.ptr(o1))
    r2 = R.ref_cm_encode(refs.ptr(a.copy()), n, refs.ptr(o2))
    assert r1 == r2 and bytes(o1[:r1]) == bytes(o2[:r2])
    for insize in (r1, max(r1 - 3, 0), r1 // 2, 0):  # truncated payloads read as 0xFF.. like read_in
        d1 = np.zeros(n + 8, np.uint8)
        d2 = np.zeros(n + 8, np.uint8)
        O.orc_cm_decode(refs.ptr(o1), insize, refs.ptr(d1), n)
        R.ref_cm_decode(refs.ptr(o2), insize, refs.ptr(d2), n)
        assert bytes(d1) == bytes(d2)
        if insize == r1:
            assert bytes(d1[:n]) == bytes(a)


# ------------------------------------------------------------------ block-level differential
@needs_ref
@pytest.mark.parametrize("name,data", CASES, ids=IDS)
def test_block_encode_decode_vs_reference(name, data):
    L = refs.ref()
    bs = max(65 * 1024, len(data))
    enc_r, r_r, e_r = refs.api_encode_block(L, data, bs)
    enc_o, r_o, e_o = refs.oracle_encode_block(data, bs)
    assert (r_r, e_r) == (r_o, e_o) and enc_r == enc_o
    dec_r, d_r, de_r = refs.api_decode_block(L, enc_r, len(data), bs)
    dec_o, d_o, de_o = refs.oracle_decode_block(enc_o, len(data), bs)
    assert (d_r, de_r) == (d_o, de_o) == (len(data), 0) and dec_r == dec_o == data


@needs_ref
def test_block_too_big():
    L = refs.ref()
    data = bytes(70000)
    assert refs.api_encode_block(L, data, 65 * 1024)[1:] == refs.oracle_encode_block(data, 65 * 1024)[1:] == (-1, -6)


def hostile_variants(enc, osz, bs, rng):
    """(enc, orig_size, buffer_size, compressed_size) tuples in the spirit of examples/fuzz-decode-block.c:173-207."""
    out = [(enc, osz, None, None), (enc, osz + 1, None, None), (enc, max(osz - 1, 0), None, None),
           (enc, osz, 8, None), (enc, osz, len(enc) - 1, None), (enc, osz, None, -5), (enc, osz, None, 4),
           (enc, -1, None, None), (enc, refs.bound(bs) + 1, None, None), (enc, osz, osz, None),
           (enc[:len(enc) // 2], osz, None, None), (enc[:9], osz, None, None), (enc[:12], osz, None, None)]
    for _ in range(12):
        b = bytearray(enc)
        k = int(rng.integers(0, len(b)))
        b[k] ^= 1 << int(rng.integers(0, 8))
        out.append((bytes(b), osz, None, None))
    for field in (4, 8, 9, 13):  # bwt index, model byte, size fields
        if len(enc) > field + 4:
            for v in (0, 1, 0x7FFFFFFF, 0xFFFFFFFE, osz + 7):
                b = bytearray(enc)
                if field == 8:
                    b[8] = v & 0xFF
                else:
                    b[field:field + 4] = struct.pack("<I", v & 0xFFFFFFFF)
                out.append((bytes(b), osz, None, None))
    return out


@needs_ref
@pytest.mark.parametrize("name", ["raw63", "coded65_text", "zeros_4k", "random_10k", "repeat_block_5000x20",
                                  "long_runs", "zipf_200k"])
def test_hostile_decode_error_parity(name):
    L = refs.ref()
    data = dict(CASES)[name]
    bs = max(65 * 1024, len(data))
    enc, r, e = refs.oracle_encode_block(data, bs)
    rng = np.random.default_rng(len(data))
    for k, (venc, osz, bsz, csz) in enumerate(hostile_variants(enc, len(data), bs, rng)):
        got_r = refs.api_decode_block(L, venc, osz, bs, buffer_size=bsz, compressed_size=csz)
        got_o = refs.oracle_decode_block(venc, osz, bs, buffer_size=bsz, compressed_size=csz)
        assert got_r[1:] == got_o[1:], (name, k, got_r[1:], got_o[1:])
        if got_r[1] >= 0:
            assert got_r[0] == got_o[0]


@needs_ref
def test_medium_corpora_block_parity():
    L = refs.ref()
    for gen, n in ((synth.zipf_text, 1_500_000), (synth.source_corpus, 1_500_000), (synth.mixed, 1_200_000),
                   (synth.log_stream, 800_000)):
        data = gen(n).tobytes() if gen is not synth.mixed else gen(n, segment=300_000).tobytes()
        bs = 2 << 20
        a = refs.api_encode_block(L, data, bs)
        b = refs.oracle_encode_block(data, bs)
        assert a == b
        assert refs.oracle_decode_block(b[0], len(data), bs)[0] == data
