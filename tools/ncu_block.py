"""One block through bz3_encode_block + bz3_decode_block, for ncu:
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file gpurun_out/ncu_block.csv python tools/ncu_block.py [mib] [generator]
(then tools/ncu_summary.py turns the csv into the per-kernel table under profiles/).  ctypes + numpy only."""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import bzip3_b200  # noqa: E402
from bzip3_b200 import synth  # noqa: E402

mib = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0
gen = getattr(synth, sys.argv[2]) if len(sys.argv) > 2 else synth.zipf_text
n = int(mib * (1 << 20))
data = gen(n, seed=synth.SEED_ZIPF_TEXT).tobytes()
with bzip3_b200.Bz3State(n) as s:
    enc, r = s.encode_block(data)
    assert r > 0, r
    dec, r2 = s.decode_block(enc, n)
    assert dec == data
print("block of %d bytes -> %d, model byte %d" % (n, r, enc[8]))
