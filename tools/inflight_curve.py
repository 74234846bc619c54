"""Blocks in flight on ONE GPU: round-trip MiB/s of k device-resident blocks coded at once (one state, one stream and
one host thread per block: bz3_b200_encode_resident_many / decode_resident_many = what bz3_encode_blocks runs minus
the PCIe copies), for growing k.  Answers "what do 148 (296) blocks in flight buy on a B200" with a measurement.

Usage (GPU box):  python tools/inflight_curve.py [--mib 4] [--ks 1,6,37,74,148,296] [--out gpurun_out/inflight.json]
ctypes + numpy only.  Every block is a different slice of the same Zipf text; the first block of every k is checked
against its input after the round trip."""
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

u8p = C.POINTER(C.c_uint8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=float, default=4.0)
    ap.add_argument("--ks", default="1,6,37,74,148,296")
    ap.add_argument("--kind", default="zipf")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "inflight.json"))
    a = ap.parse_args()
    n = int(a.mib * (1 << 20))
    ks = [int(x) for x in a.ks.split(",")]
    L = bzip3_b200.lib()
    kmax = max(ks)
    gen = {"zipf": synth.zipf_text, "source": synth.source_corpus}[a.kind]
    pool = np.ascontiguousarray(gen(n + 4096 * kmax, seed=77))
    states = []
    for i in range(kmax):
        s = L.bz3_new(max(n, 1 << 20))
        assert s, "bz3_new failed at state %d" % i
        states.append(s)
    rows = []
    for k in ks:
        hs = (C.c_void_p * k)(*states[:k])
        sizes = (C.c_int32 * k)(*([n] * k))
        res = (C.c_int32 * k)()
        for i in range(k):
            blk = pool[i * 4096: i * 4096 + n]
            assert L.bz3_b200_upload(states[i], blk.ctypes.data_as(u8p), n) == 0
        t0 = time.perf_counter()
        L.bz3_b200_encode_resident_many(hs, sizes, res, k)
        te = time.perf_counter() - t0
        cs = [int(x) for x in res]
        assert min(cs) > 0, cs
        csz = (C.c_int32 * k)(*cs)
        osz = (C.c_int32 * k)(*([n] * k))
        res2 = (C.c_int32 * k)()
        t0 = time.perf_counter()
        L.bz3_b200_decode_resident_many(hs, csz, osz, res2, k)
        td = time.perf_counter() - t0
        assert all(int(x) == n for x in res2), list(res2)
        back = np.zeros(n, np.uint8)
        assert L.bz3_b200_download(states[0], back.ctypes.data_as(u8p), n) == 0
        assert bytes(back) == bytes(pool[:n]), "round trip of block 0 differs"
        mib = k * n / (1 << 20)
        row = {"blocks": k, "block_mib": a.mib, "enc_s": te, "dec_s": td, "enc_MiBps": mib / te, "dec_MiBps": mib / td,
               "roundtrip_MiBps": mib / (te + td)}
        rows.append(row)
        print("k=%4d  enc %7.2f s (%8.1f MiB/s)  dec %7.2f s (%8.1f MiB/s)  round trip %8.1f MiB/s" % (
            k, te, row["enc_MiBps"], td, row["dec_MiBps"], row["roundtrip_MiBps"]), flush=True)
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        with open(a.out, "w") as f:
            json.dump({"kind": a.kind, "n": n, "rows": rows}, f, indent=1)
    for s in states:
        L.bz3_free(s)


if __name__ == "__main__":
    main()
