"""compute-sanitizer target (memcheck / racecheck / initcheck): hostile blocks and every shipped kernel on hardware.

For a handful of small inputs (plain text, a run-heavy one that takes the mRLE path, a repetitive one that takes the
LZP path, incompressible bytes, a raw < 64 byte block) the block is encoded on the GPU, then every hostile variant of
tests/test_oracle.py::hostile_variants (wrong sizes, truncated payloads, flipped bits, forged header fields -- in the
spirit of the reference's examples/fuzz-decode-block.c:173-207) is decoded on the GPU and compared with the oracle's
verdict (return value, error number, bytes).  Run as
    compute-sanitizer --tool memcheck  python tools/sanitize_hostile.py
    compute-sanitizer --tool racecheck python tools/sanitize_hostile.py
The point is the sanitizer's report ("0 errors"): no kernel may read or write out of bounds, or race, on corrupt input."""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import bzip3_b200  # noqa: E402
from bzip3_b200 import synth  # noqa: E402
from tests import refs  # noqa: E402
from tests.test_oracle import hostile_variants  # noqa: E402

BS = 65 * 1024 + 4096
rng = np.random.default_rng(11)
per = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
cases = {
    "text": synth.zipf_text(per, seed=3).tobytes(),
    "runs_mrle": bytes(np.repeat(rng.integers(0, 5, per // 40 + 1, dtype=np.uint8), 40)[:per]),
    "repeats_lzp": (synth.log_stream(per // 3, seed=4).tobytes()) * 3,
    "random": bytes(rng.integers(0, 256, per, dtype=np.uint8)),
    "raw_under_64": b"x" * 50,
}
checked = 0
with bzip3_b200.Bz3State(BS) as s:
    for name, data in cases.items():
        enc, r = s.encode_block(data)
        want = refs.oracle_encode_block(data, BS)
        assert r == want[1] and enc == want[0], name
        dec, r2 = s.decode_block(enc, len(data))
        assert dec == data, name
        vr = np.random.default_rng(len(data))
        for k, (venc, osz, bsz, csz) in enumerate(hostile_variants(enc, len(data), BS, vr)):
            got = s.decode_block(venc, osz, buffer_size=bsz, compressed_size=csz)
            err = s.last_error
            ow = refs.oracle_decode_block(venc, osz, BS, buffer_size=bsz, compressed_size=csz)
            assert got[1] == ow[1], (name, k, got[1], ow[1])
            if got[1] >= 0:
                assert got[0] == ow[0], (name, k)
            else:
                assert err == ow[2], (name, k, err, ow[2])
            checked += 1
        print("%-14s model byte %s: %d hostile variants decoded like the oracle" % (name, enc[8] if len(enc) > 8 else "-", k + 1), flush=True)
print("sanitize_hostile ok: %d decodes" % checked)
