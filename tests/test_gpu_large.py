"""Large blocks on the GPU against the compiled reference (oracle/_ref): the encoded block must equal the reference's
bz3_encode_block output byte for byte and decode both ways, up to the metric's block size (256 MiB, BASELINE.json
configs[2]) and the format's maximum (511 MiB, configs[4]; src/libbz3.c:536); plus the batch API at the benchmark's
block size.  The reference's side of the two big cases runs on a host thread while the GPU works (about a minute each)."""
import os
import struct

import numpy as np
import pytest

import bzip3_b200
from bzip3_b200 import synth
from tests import refs

pytestmark = pytest.mark.gpu


def roundtrip(block_mib, gen, check_reference):
    n = block_mib << 20
    data = gen(n, seed=1234 + block_mib)
    with bzip3_b200.Bz3State(n) as s:
        enc, r = s.encode_block(data.tobytes())
        assert r > 0 and s.last_error == 0
        crc, idx, model = struct.unpack("<IiB", enc[:9])
        assert crc == refs.oracle().orc_crc32(1, refs.ptr(data), n)
        assert 1 <= idx <= n
        if check_reference and refs.have_ref():
            want = refs.api_encode_block(refs.ref(), data.tobytes(), n)[0]
            assert want == enc, "differs from the reference encoder"
        dec, r2 = s.decode_block(enc, n)
        assert r2 == n and s.last_error == 0 and dec == data.tobytes()


def test_roundtrip_64mib_source_block():
    roundtrip(64, synth.source_corpus, check_reference=True)


def test_roundtrip_32mib_mixed_block_with_incompressible_segments():
    roundtrip(32, lambda n, seed: synth.mixed(n, seed=seed, segment=4 << 20), check_reference=True)


def test_batch_of_16mib_blocks_matches_reference():
    bs = 16 << 20
    datas = [synth.zipf_text(bs, seed=77).tobytes(), synth.log_stream(bs // 2, seed=78).tobytes()]
    states = [bzip3_b200.Bz3State(bs) for _ in datas]
    try:
        bufs = []
        for d in datas:
            b = np.zeros(bzip3_b200.bound(bs) + 64, np.uint8)
            b[:len(d)] = np.frombuffer(d, np.uint8)
            bufs.append(b)
        sizes = bzip3_b200.encode_blocks(states, bufs, [len(d) for d in datas])
        assert all(s.last_error == 0 for s in states)
        if refs.have_ref():
            for d, b, sz in zip(datas, bufs, sizes):
                assert refs.api_encode_block(refs.ref(), d, bs)[0] == bytes(b[:sz])
        bzip3_b200.decode_blocks(states, bufs, [len(b) for b in bufs], sizes, [len(d) for d in datas])
        for d, b, s in zip(datas, bufs, states):
            assert s.last_error == 0 and bytes(b[:len(d)]) == d
    finally:
        for s in states:
            s.close()


def test_many_blocks_of_mixed_sizes_at_once():
    """40 states with different block sizes coded by 40 host threads at once, twice: sorts with different pass counts
    (8 for the first round of a suffix sort, 2 * ceil(log2(n + 1)) / 8 afterwards, 3 for LZP, 1 for the inverse BWT) overlap
    in time, so every per-function launch attribute must be the same for all of them.  (A histogram kernel whose shared
    memory limit followed the launching thread's pass count failed here with "too many resources requested for launch".)"""
    rng = np.random.default_rng(4040)
    sizes = [int(x) for x in rng.integers(70 << 10, 3 << 20, 40)]
    gens = [synth.zipf_text, synth.source_corpus, synth.log_stream]
    datas = [gens[i % 3](n, seed=500 + i).tobytes()[:n] for i, n in enumerate(sizes)]
    states = [bzip3_b200.Bz3State(max(len(d), 65 << 10)) for d in datas]
    try:
        for rep in range(2):
            bufs = []
            for d in datas:
                b = np.zeros(bzip3_b200.bound(len(d)) + 64, np.uint8)
                b[:len(d)] = np.frombuffer(d, np.uint8)
                bufs.append(b)
            enc = bzip3_b200.encode_blocks(states, bufs, [len(d) for d in datas])
            assert all(s.last_error == 0 for s in states), [s.last_error for s in states]
            assert all(e > 0 for e in enc), enc
            if rep == 0:
                for i in (0, 1, 2, 17, 39):
                    want, r, _ = refs.oracle_encode_block(datas[i], max(len(datas[i]), 65 << 10))
                    assert r == enc[i] and want == bytes(bufs[i][:enc[i]]), i
            bzip3_b200.decode_blocks(states, bufs, [len(b) for b in bufs], enc, [len(d) for d in datas])
            for d, b, s in zip(datas, bufs, states):
                assert s.last_error == 0 and bytes(b[:len(d)]) == d
    finally:
        for s in states:
            s.close()


def _against_reference_big(n, data):
    """encode on the GPU and with the reference at the same time, compare, decode both ways"""
    import threading
    assert refs.have_ref(), "oracle/_ref is missing: run `make -C oracle ref` where /root/reference exists"
    R = refs.ref()
    cap = refs.bound(n) + 64
    rbuf = np.zeros(cap, np.uint8)
    rbuf[:n] = data
    rst = R.bz3_new(n)
    assert rst
    rres = {}

    def ref_side():
        rres["size"] = R.bz3_encode_block(rst, refs.ptr(rbuf), n)
        rres["err"] = R.bz3_last_error(rst)

    th = threading.Thread(target=ref_side)
    th.start()
    gbuf = np.zeros(cap, np.uint8)
    gbuf[:n] = data
    L = bzip3_b200.lib()
    try:
        with bzip3_b200.Bz3State(n) as s:
            r = L.bz3_encode_block(s.handle, refs.ptr(gbuf), n)
            assert r > 0 and s.last_error == 0, (r, s.last_error)
            th.join()
            assert rres["err"] == 0 and rres["size"] == r, (rres, r)
            assert np.array_equal(gbuf[:r], rbuf[:r]), "block differs from the reference's bz3_encode_block output"
            # the reference decodes the GPU's block (host thread) while the GPU decodes the reference's
            def ref_decode():
                rres["dec"] = R.bz3_decode_block(rst, refs.ptr(rbuf), cap, r, n)
                rres["derr"] = R.bz3_last_error(rst)
            rbuf[:r] = gbuf[:r]
            th2 = threading.Thread(target=ref_decode)
            th2.start()
            r2 = L.bz3_decode_block(s.handle, refs.ptr(gbuf), cap, r, n)
            assert r2 == n and s.last_error == 0, (r2, s.last_error)
            assert np.array_equal(gbuf[:n], data), "GPU decode of the block differs from the input"
            th2.join()
            assert rres["dec"] == n and rres["derr"] == 0, rres
            assert np.array_equal(rbuf[:n], data), "the reference decodes the GPU's block to something else"
    finally:
        if th.is_alive():
            th.join()
        R.bz3_free(rst)


def test_256mib_block_equals_the_reference():
    """the metric's block size: one 256 MiB block of the synthetic source corpus (BASELINE.json configs[2])"""
    n = 256 << 20
    _against_reference_big(n, synth.source_corpus(n, seed=synth.SEED_SOURCE))


def test_511mib_block_equals_the_reference():
    """the largest block the format allows (src/libbz3.c:536): 511 MiB of the synthetic log stream (configs[4])"""
    n = 511 << 20
    _against_reference_big(n, synth.log_stream(n, seed=synth.SEED_LOG))
