"""Size-independent properties at large block sizes (GPU): encode -> decode round trip, CRC of the decoded block,
agreement of the header fields with the reference where the CPU reference is cheap enough to run, and the
block-sharded batch API at the benchmark's block size.  The 256 MiB case (BASELINE.json configs[2], minutes of GPU time
at the current coder rate) runs only with BZ3_TEST_HUGE=1."""
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


@pytest.mark.skipif(os.environ.get("BZ3_TEST_HUGE") != "1", reason="set BZ3_TEST_HUGE=1 (several minutes of GPU time)")
def test_roundtrip_256mib_block():
    roundtrip(256, synth.source_corpus, check_reference=False)


def test_no_promoted_kernel_was_retired():
    """Runs last in this module: the round trips above used the kernels the self-test chose; none of them may have needed
    the round-1 kernels' second opinion on a good block (decode_checked in bz3_api.cu)."""
    assert bzip3_b200.lib().bz3_b200_demotions() == 0
