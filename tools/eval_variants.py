"""One-shot GPU evaluation of the entropy-stage kernel variants: parity against the oracle and device time.

Usage (GPU box):  python tools/eval_variants.py [--mib 2] [--out gpurun_out/eval_variants.json]
Imports only ctypes/numpy (no torch) so that it starts in seconds.  For every data set the BWT of the data is
made on the GPU, coded by the oracle on the CPU (the known answer), then every encoder variant must
reproduce the oracle's bytes and every decoder variant must reproduce the input -- also on a truncated
stream -- while the stage call is timed (wall clock around a synchronous stage call: one H2D + kernel + D2H;
the copies are microseconds against a kernel of hundreds of milliseconds)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bzip3_b200  # noqa: E402
from bzip3_b200 import synth  # noqa: E402
from tests import refs  # noqa: E402

CM = 5  # BZ3_STAGE_CM
ENC_VARIANTS = {0: "chunked, select+mul.hi coder lane (round-1 default)", 4: "chunked, one-multiply coder lane, two-tier",
                6: "chunked, one-multiply coder lane, branch-free byte + resume at first event",
                10: "round 2: coder lane with in-place shifts (straight-line byte, second copy after a shift)"}
DEC_VARIANTS = {0: "tree, serial chain warp (round-1 default)", 3: "all paths, first edition",
                4: "tree, lane-parallel chain warp", 5: "all paths, one multiply per level",
                6: "walker warps (all-paths walk) + model threads", 7: "walker warps + slim model threads",
                8: "walker warps (stop after 3 levels when outside their eighth) + slim model threads",
                9: "as 8, branch-light parity-unrolled model threads",
                10: "round 2: serial walker with 3-level-deep table prefetch, in-place shifts, named-barrier hand-offs"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=float, default=2.0)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "eval_variants.json"))
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--sets", default="zipf_text,source,mixed")
    ap.add_argument("--enc", default="0,4,6")
    ap.add_argument("--dec", default="0,4,6,7,8,9,5")
    args = ap.parse_args()
    n = int(args.mib * (1 << 20))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    L = bzip3_b200.lib()
    O = refs.oracle()
    u8p = refs.u8p
    sets = [("zipf_text", synth.zipf_text(n, seed=4242)), ("source", synth.source_corpus(n, seed=4243)),
            ("mixed", synth.mixed(n, seed=4244, segment=max(n // 4, 1 << 16)))]
    sets = [x for x in sets if x[0] in args.sets.split(",")]
    result = {"n": n, "sets": {}, "ok": True}
    with bzip3_b200.Bz3State(max(n, 1 << 20)) as st:
        try:   # what the self-test of the library chose for this process
            result["defaults_in_effect"] = {"cm_enc": L.bz3_b200_get_variant(st.handle, CM + 100),
                                            "cm_dec": L.bz3_b200_get_variant(st.handle, CM + 200),
                                            "lzp": L.bz3_b200_get_variant(st.handle, 3)}
            print("defaults in effect:", result["defaults_in_effect"], flush=True)
        except Exception:
            pass
        prep = {}
        for name, data in sets:
            data = np.ascontiguousarray(data[:n])
            bwt = np.zeros(n + 64, np.uint8)
            idx = L.bz3_b200_stage_bwt(st.handle, data.ctypes.data_as(u8p), n, bwt.ctypes.data_as(u8p))
            assert idx > 0, idx
            want = np.zeros(2 * n + 64, np.uint8)
            t0 = time.perf_counter()
            rw = O.orc_cm_encode(bwt.ctypes.data_as(u8p), n, want.ctypes.data_as(u8p))
            cpu_enc = time.perf_counter() - t0
            back = np.zeros(n + 8, np.uint8)
            t0 = time.perf_counter()
            O.orc_cm_decode(want.ctypes.data_as(u8p), rw, back.ctypes.data_as(u8p), n)
            cpu_dec = time.perf_counter() - t0
            cut = max(rw - 5, 0)
            dw = np.zeros(n + 8, np.uint8)
            O.orc_cm_decode(want.ctypes.data_as(u8p), cut, dw.ctypes.data_as(u8p), n)
            prep[name] = (bwt, want, rw, cut, dw)
            result["sets"][name] = {"ratio": rw / n, "oracle_cpu_enc_MBps": n / cpu_enc / 1e6,
                                    "oracle_cpu_dec_MBps": n / cpu_dec / 1e6, "enc": {}, "dec": {}}
            print("%s: ratio %.3f, oracle on one host core: enc %.1f MB/s, dec %.1f MB/s" % (
                name, rw / n, n / cpu_enc / 1e6, n / cpu_dec / 1e6), flush=True)
        # known-good kernels first, so that a fault in a new one cannot hide the baseline
        order = [("enc", int(v)) for v in args.enc.split(",") if v] + [("dec", int(v)) for v in args.dec.split(",") if v]
        for kind, v in order:
            for name, _ in sets:
                bwt, want, rw, cut, dw = prep[name]
                best, ok = 1e9, True
                if kind == "enc":
                    L.bz3_b200_set_variant(st.handle, CM + 100, v)
                    for _ in range(args.reps):
                        got = np.zeros(2 * n + 64, np.uint8)
                        t0 = time.perf_counter()
                        rg = L.bz3_b200_stage_cm_encode(st.handle, bwt.ctypes.data_as(u8p), n, got.ctypes.data_as(u8p))
                        best = min(best, time.perf_counter() - t0)
                        ok = ok and rg == rw and bytes(got[:rw]) == bytes(want[:rw])
                else:
                    L.bz3_b200_set_variant(st.handle, CM + 200, v)
                    for _ in range(args.reps):
                        got = np.zeros(n + 8, np.uint8)
                        t0 = time.perf_counter()
                        rc = L.bz3_b200_stage_cm_decode(st.handle, want.ctypes.data_as(u8p), rw, got.ctypes.data_as(u8p), n)
                        best = min(best, time.perf_counter() - t0)
                        ok = ok and rc == 0 and bytes(got[:n]) == bytes(bwt[:n])
                    dg = np.zeros(n + 8, np.uint8)   # truncated stream: same bytes as the oracle produces from it
                    L.bz3_b200_stage_cm_decode(st.handle, want.ctypes.data_as(u8p), cut, dg.ctypes.data_as(u8p), n)
                    ok = ok and bytes(dg[:n]) == bytes(dw[:n])
                L.bz3_b200_set_variant(st.handle, CM, 0)
                d = {"ok": bool(ok), "ms": best * 1e3, "MBps": n / best / 1e6, "cycles_per_byte_at_1.965GHz": best * 1.965e9 / n}
                result["sets"][name][kind][v] = d
                result["ok"] = result["ok"] and bool(ok)
                print("  %s v%d %-10s %-5s %8.1f ms  %6.2f MB/s  %6.0f cyc/B   %s" % (
                    kind, v, name, "OK" if ok else "FAIL", d["ms"], d["MBps"], d["cycles_per_byte_at_1.965GHz"],
                    (ENC_VARIANTS if kind == "enc" else DEC_VARIANTS)[v]), flush=True)
                with open(args.out, "w") as f:
                    json.dump(result, f, indent=1)
        # LZP pre-pass: one-window kernels (variant 3) against variant 2 (windows in flight / bulk decoder)
        LZP = 3  # BZ3_STAGE_LZP
        lzp_sets = sets + [("log", synth.log_stream(n, seed=4245))]
        result["lzp"] = {}
        for name, data in lzp_sets:
            data = np.ascontiguousarray(data[:n])
            pad = np.zeros(n + 64, np.uint8)
            pad[:n] = data
            lut = np.zeros(1 << 18, np.int32)
            lp = lut.ctypes.data_as(refs.i32p)
            want = np.zeros(n + 64, np.uint8)
            rw = O.orc_lzp_encode(pad.ctypes.data_as(u8p), n, want.ctypes.data_as(u8p), lp)
            rec = {"lzp_size": int(rw)}
            for v in (3, 2):   # 3 = one window per step (round-1 kernels), 2 = windows in flight / bulk decoder
                L.bz3_b200_set_variant(st.handle, LZP, v)
                got = np.zeros(n + 64, np.uint8)
                t0 = time.perf_counter()
                rg = L.bz3_b200_stage_lzp_encode(st.handle, pad.ctypes.data_as(u8p), n, got.ctypes.data_as(u8p))
                te = time.perf_counter() - t0
                ok = rg == rw and (rw <= 0 or bytes(got[:rw]) == bytes(want[:rw]))
                td, okd = 0.0, True
                if rw > 0:
                    cap = refs.bound(n)
                    back = np.zeros(cap + 64, np.uint8)
                    t0 = time.perf_counter()
                    rd = L.bz3_b200_stage_lzp_decode(st.handle, want.ctypes.data_as(u8p), rw, back.ctypes.data_as(u8p), cap)
                    td = time.perf_counter() - t0
                    okd = rd == n and bytes(back[:n]) == bytes(data)
                    dw = np.zeros(cap + 64, np.uint8)   # truncated input: same verdict and bytes as the oracle
                    dg = np.zeros(cap + 64, np.uint8)
                    sw = O.orc_lzp_decode(want.ctypes.data_as(u8p), rw // 2, dw.ctypes.data_as(u8p), cap, lp)
                    sg = L.bz3_b200_stage_lzp_decode(st.handle, want.ctypes.data_as(u8p), rw // 2, dg.ctypes.data_as(u8p), cap)
                    okd = okd and sg == sw and (sw <= 0 or bytes(dg[:sw]) == bytes(dw[:sw]))
                rec["v%d" % v] = {"enc_ok": bool(ok), "enc_ms": te * 1e3, "dec_ok": bool(okd), "dec_ms": td * 1e3}
                result["ok"] = result["ok"] and bool(ok) and bool(okd)
                print("  lzp v%d %-10s enc %-5s %8.2f ms (%7.1f MB/s)   dec %-5s %8.2f ms   lzp_size %d" % (
                    v, name, "OK" if ok else "FAIL", te * 1e3, n / te / 1e6, "OK" if okd else "FAIL", td * 1e3, rw), flush=True)
            L.bz3_b200_set_variant(st.handle, LZP, 0)
            result["lzp"][name] = rec
        # whole-block round trip with the new kernels selected through the public block API
        blk = np.ascontiguousarray(sets[0][1][: min(n, 1 << 20)])
        enc_o, r_o, e_o = refs.oracle_encode_block(bytes(blk), max(n, 1 << 20))
        combos = {}
        for ve, vd in ((4, 4), (6, 8), (6, 9)):
            L.bz3_b200_set_variant(st.handle, CM + 100, ve)
            L.bz3_b200_set_variant(st.handle, CM + 200, vd)
            buf = np.zeros(refs.bound(max(n, 1 << 20)) + 64, np.uint8)
            buf[: len(blk)] = blk
            r = L.bz3_encode_block(st.handle, buf.ctypes.data_as(u8p), len(blk))
            same = r == r_o and bytes(buf[:r]) == enc_o
            r2 = L.bz3_decode_block(st.handle, buf.ctypes.data_as(u8p), len(buf), r, len(blk))
            rt = r2 == len(blk) and bytes(buf[: len(blk)]) == bytes(blk)
            combos["enc%d_dec%d" % (ve, vd)] = {"encode_equals_oracle": bool(same), "roundtrip": bool(rt)}
            result["ok"] = result["ok"] and same and rt
            print("block enc v%d dec v%d: encode==oracle %s, round trip %s" % (ve, vd, same, rt), flush=True)
        L.bz3_b200_set_variant(st.handle, CM, 0)
        result["block"] = combos
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(result, f, indent=1)
    print("ALL OK" if result["ok"] else "SOME FAILED")
    return 0 if result["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
