"""clock64 phase breakdown of the entropy-stage kernels (needs tools/variants/lib_cmprof.so = the library built
with -DBZ_CM_PROFILE; see tools/README in DESIGN.md section 6).  Prints cycles per input byte for every
phase of the encoder (two model stages, coder lane) and of the decoder's chain warp.
ctypes + numpy only."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["BZ3_B200_LIB"] = os.path.join(ROOT, "tools", "variants", "lib_cmprof.so")
import bzip3_b200  # noqa: E402
from bzip3_b200 import synth  # noqa: E402

CM = 5
u8p = C.POINTER(C.c_uint8)


def main():
    mib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    n = int(mib * (1 << 20))
    L = bzip3_b200.lib()
    out = {}
    sets = [("zipf_text", synth.zipf_text(n, seed=4242)), ("mixed", synth.mixed(n, seed=4244, segment=max(n // 4, 1 << 16)))]

    def prof():
        p = (C.c_ulonglong * 48)()
        L.bz3_b200_debug_cm_profile(p)
        return [int(x) for x in p]

    with bzip3_b200.Bz3State(max(n, 1 << 20)) as st:
        for name, data in sets:
            data = np.ascontiguousarray(data[:n])
            bwt = np.zeros(n + 64, np.uint8)
            L.bz3_b200_stage_bwt(st.handle, data.ctypes.data_as(u8p), n, bwt.ctypes.data_as(u8p))
            enc = np.zeros(2 * n + 64, np.uint8)
            rec = {}
            r = L.bz3_b200_stage_cm_encode(st.handle, bwt.ctypes.data_as(u8p), n, enc.ctypes.data_as(u8p))
            p = prof()
            rec["encoder"] = {"stage1_busy": p[13] / n, "stage2_busy": p[14] / n, "coder_busy": p[15] / n}
            back = np.zeros(n + 8, np.uint8)
            L.bz3_b200_stage_cm_decode(st.handle, enc.ctypes.data_as(u8p), r, back.ctypes.data_as(u8p), n)
            assert bytes(back[:n]) == bytes(bwt[:n]), name
            p = prof()
            rec["decoder_chain_warp"] = dict(zip(("wait_ptab", "fast_tier", "exact_tier", "publish_wait_byte", "redo_frac"),
                                                 [x / n for x in p[8:13]]))
            out[name] = rec
            print(name)
            for k, d in rec.items():
                print("  %-16s" % k, "  ".join("%s=%.1f" % (a, b) if b >= 1.5 else "%s=%.3f" % (a, b) for a, b in d.items()), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "cm_prof2.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
