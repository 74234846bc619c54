"""Runs single stages of the CUDA codec on synthetic data (used under ncu / for stage timings).
usage: python tools/stage_driver.py <stage> <MiB> [repeat]     stage in {bwt, unbwt, cm_enc, cm_dec, lzp, rle, block}"""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import bzip3_b200  # noqa: E402
from bzip3_b200 import synth  # noqa: E402


def main():
    stage = sys.argv[1]
    mib = float(sys.argv[2])
    rep = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    n = int(mib * (1 << 20))
    gen = synth.zipf_text if len(sys.argv) < 5 else getattr(synth, sys.argv[4])
    data = gen(n, seed=4242)
    bs = max(n, 65 * 1024)
    L = bzip3_b200.lib()
    u8p = C.POINTER(C.c_uint8)
    with bzip3_b200.Bz3State(bs) as s:
        import os
        out = np.zeros(bzip3_b200.bound(n) + 64, np.uint8)
        back = np.zeros(bzip3_b200.bound(n) + 64, np.uint8)
        pi, po, pb = (a.ctypes.data_as(u8p) for a in (data, out, back))
        if stage.endswith("_bwt"):  # entropy stage on BWT output (what it sees inside a block)
            tmp = np.zeros(n + 64, np.uint8)
            L.bz3_b200_stage_bwt(s.handle, pi, n, tmp.ctypes.data_as(u8p))
            data = tmp[:n].copy()
            pi = data.ctypes.data_as(u8p)
            stage = stage[:-4]
        for r in range(rep):
            t0 = time.perf_counter()
            if stage == "bwt":
                idx = L.bz3_b200_stage_bwt(s.handle, pi, n, po)
                info = f"idx={idx}"
            elif stage == "unbwt":
                idx = L.bz3_b200_stage_bwt(s.handle, pi, n, po)
                t0 = time.perf_counter()
                st = L.bz3_b200_stage_unbwt(s.handle, po, n, idx, pb)
                info = f"status={st} ok={bytes(back[:n]) == data.tobytes()}"
            elif stage == "cm_enc":
                r2 = L.bz3_b200_stage_cm_encode(s.handle, pi, n, po)
                info = f"out={r2}"
            elif stage == "cm_dec":
                r2 = L.bz3_b200_stage_cm_encode(s.handle, pi, n, po)
                t0 = time.perf_counter()
                L.bz3_b200_stage_cm_decode(s.handle, po, r2, pb, n)
                info = f"ok={bytes(back[:n]) == data.tobytes()}"
            elif stage == "lzp":
                r2 = L.bz3_b200_stage_lzp_encode(s.handle, pi, n, po)
                t1 = time.perf_counter()
                r3 = L.bz3_b200_stage_lzp_decode(s.handle, po, r2, pb, bzip3_b200.bound(n)) if r2 > 0 else -1
                info = f"enc={r2} ({(t1 - t0) * 1e3:.1f} ms) dec={r3} ok={r2 <= 0 or bytes(back[:n]) == data.tobytes()}"
            elif stage == "rle":
                r2 = L.bz3_b200_stage_rle_encode(s.handle, pi, n, po)
                t1 = time.perf_counter()
                e = L.bz3_b200_stage_rle_decode(s.handle, po, r2, pb, n)
                info = f"enc={r2} ({(t1 - t0) * 1e3:.1f} ms) err={e} ok={bytes(back[:n]) == data.tobytes()}"
            elif stage == "block":
                enc, r2 = s.encode_block(data.tobytes())
                t1 = time.perf_counter()
                dec, r3 = s.decode_block(enc, n)
                info = f"enc={r2} ({(t1 - t0) * 1e3:.1f} ms) ok={dec == data.tobytes()} enc_ms={s.stage_ms(False)} dec_ms={s.stage_ms(True)}"
            else:
                raise SystemExit("unknown stage")
            dt = time.perf_counter() - t0
            print(f"{stage} {mib} MiB rep {r}: {dt * 1e3:.2f} ms  {n / (1 << 20) / dt:.1f} MiB/s  {info}", flush=True)


if __name__ == "__main__":
    main()
