"""clock64 phase breakdown of the entropy-stage kernels (needs tools/variants/lib_cmprof.so = the library built
with -DBZ_CM_PROFILE; see tools/README in DESIGN.md section 6).  Prints cycles per input byte for every
phase of: encoder v0 / v4, decoder v0 (tree), v4 (lane-parallel chain warp), v5 (all paths v2).
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
            for v in [int(x) for x in os.environ.get('BZ3_PROF_ENC', '0,4,6').split(',')]:
                L.bz3_b200_set_variant(st.handle, CM + 100, v)
                r = L.bz3_b200_stage_cm_encode(st.handle, bwt.ctypes.data_as(u8p), n, enc.ctypes.data_as(u8p))
                p = prof()
                rec["enc_v%d" % v] = {"stage1_busy": p[13] / n, "stage2_busy": p[14] / n, "coder_busy": p[15] / n,
                                      "exact_tier_cycles": p[32] / n, "exact_tier_bytes_frac": p[33] / n}
            for v in [int(x) for x in os.environ.get('BZ3_PROF_DEC', '0,4,5,6,7,8,9,10').split(',')]:
                L.bz3_b200_set_variant(st.handle, CM + 200, v)
                back = np.zeros(n + 8, np.uint8)
                L.bz3_b200_stage_cm_decode(st.handle, enc.ctypes.data_as(u8p), r, back.ctypes.data_as(u8p), n)
                assert bytes(back[:n]) == bytes(bwt[:n]), (name, v)
                p = prof()
                if v == 0:
                    rec["dec_v0_chain"] = dict(zip(("wait_ptab", "fast_tier", "exact_tier", "publish_wait_byte", "redo_frac"),
                                                   [x / n for x in p[8:13]]))
                elif v == 10:
                    rec["dec_v10_walker"] = dict(zip(("first_table", "walk", "publish_wait_spec", "wait_real", "miss_frac"),
                                                     [x / n for x in p[8:13]]))
                elif v == 4:
                    rec["dec_v4_chain"] = dict(zip(("wait_ptab", "round1_fast", "round1_exact", "round2_fast", "round2_exact",
                                                    "publish_wait_byte", "fb1_frac", "fb2_frac"), [x / n for x in p[16:24]]))
                elif v == 5:
                    rec["dec_v5_thread0"] = dict(zip(("predict", "wait_S1", "walk", "wait_S2", "fallback", "learn", "fallback_frac"),
                                                     [x / n for x in p[24:31]]))
                else:
                    rec["dec_v%d_walker0" % v] = dict(zip(("wait_B1", "walk", "wait_B2", "slow_path", "tail", "slow_frac",
                                                           "serial_frac"), [x / n for x in p[36:43]]))
            L.bz3_b200_set_variant(st.handle, CM, 0)
            out[name] = rec
            print(name)
            for k, d in rec.items():
                print("  %-16s" % k, "  ".join("%s=%.1f" % (a, b) if b >= 1.5 else "%s=%.3f" % (a, b) for a, b in d.items()), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "cm_prof2.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
