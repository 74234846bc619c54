"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every stage kernel on tiny inputs."""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import bzip3_b200
from bzip3_b200 import synth

rng = np.random.default_rng(5)
cases = [synth.zipf_text(40_000, seed=3).tobytes(), bytes(3000) + bytes(rng.integers(0, 256, 5000, dtype=np.uint8)),
         (synth.log_stream(30_000, seed=4).tobytes()) * 2, b"x" * 70]
with bzip3_b200.Bz3State(1 << 17) as s:
    if True:
        for d in cases:
            enc, r = s.encode_block(d)
            assert r > 0, r
            dec, r2 = s.decode_block(enc, len(d))
            assert dec == d
print("sanitize smoke ok")
