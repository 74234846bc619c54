"""Does a long-running tiny kernel on one stream slow down multi-CTA kernels on another stream?
Thread A: serial (1-thread) LZP encode of 16 MiB (seconds).  Thread B: BWT stage of 4 MiB in a loop."""
import ctypes as C
import sys
import threading
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import bzip3_b200  # noqa: E402
from bzip3_b200 import synth  # noqa: E402

u8p = C.POINTER(C.c_uint8)
L = bzip3_b200.lib()
big = synth.zipf_text(16 << 20, seed=1)
small = synth.zipf_text(4 << 20, seed=2)
sa = bzip3_b200.Bz3State(16 << 20)
sb = bzip3_b200.Bz3State(4 << 20)
outa = np.zeros(bzip3_b200.bound(len(big)) + 64, np.uint8)
outb = np.zeros(bzip3_b200.bound(len(small)) + 64, np.uint8)


def bwt_times(k):
    ts = []
    for _ in range(k):
        t0 = time.perf_counter()
        L.bz3_b200_stage_bwt(sb.handle, small.ctypes.data_as(u8p), len(small), outb.ctypes.data_as(u8p))
        ts.append((time.perf_counter() - t0) * 1e3)
    return ts


print("bwt alone ms:", [round(x, 2) for x in bwt_times(6)])
mode = sys.argv[1] if len(sys.argv) > 1 else "lzp"
L.bz3_b200_set_variant(sa.handle, 3, 1)  # serial LZP kernel
L.bz3_b200_set_variant(sa.handle, 5, 1)  # single-lane CM kernel


def long_job():
    t0 = time.perf_counter()
    if mode == "lzp":
        L.bz3_b200_stage_lzp_encode(sa.handle, big.ctypes.data_as(u8p), len(big), outa.ctypes.data_as(u8p))
    else:
        L.bz3_b200_stage_cm_encode(sa.handle, big.ctypes.data_as(u8p), 2 << 20, outa.ctypes.data_as(u8p))
    print(f"long {mode} job took {(time.perf_counter() - t0) * 1e3:.0f} ms")


th = threading.Thread(target=long_job)
th.start()
time.sleep(0.3)
print(f"bwt while {mode} kernel runs ms:", [round(x, 2) for x in bwt_times(6)])
th.join()
print("bwt alone again ms:", [round(x, 2) for x in bwt_times(3)])
