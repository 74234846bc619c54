import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
os.environ["BZ3_B200_LIB"] = os.path.join(os.path.dirname(__file__), "variants", "lib_cmprof.so")
import bzip3_b200
from bzip3_b200 import synth
L = bzip3_b200.lib()
n = 2 << 20
data = synth.zipf_text(n, seed=4242)
u8p = C.POINTER(C.c_uint8)
with bzip3_b200.Bz3State(n) as s:
    t = np.zeros(n + 64, np.uint8); L.bz3_b200_stage_bwt(s.handle, data.ctypes.data_as(u8p), n, t.ctypes.data_as(u8p))
    enc = np.zeros(2 * n, np.uint8); r = L.bz3_b200_stage_cm_encode(s.handle, t.ctypes.data_as(u8p), n, enc.ctypes.data_as(u8p))
    out = np.zeros(n + 64, np.uint8); L.bz3_b200_stage_cm_decode(s.handle, enc.ctypes.data_as(u8p), r, out.ctypes.data_as(u8p), n)
    assert bytes(out[:n]) == bytes(t[:n])
    prof = (C.c_ulonglong * 16)(); L.bz3_b200_debug_cm_profile(prof)
    p = [x / n for x in prof]
    print("cycles/byte  model(node1): A=%.0f wait_ptab=%.0f precompute=%.0f wait_byte=%.0f" % tuple(p[0:4]))
    print("cycles/byte  model(node200): A=%.0f wait_ptab=%.0f precompute=%.0f wait_byte=%.0f" % tuple(p[4:8]))
    print("cycles/byte  encoder busy: stage1=%.0f stage2=%.0f coder=%.0f" % tuple(p[13:16]))
    print("cycles/byte  chain: loop=%.0f wait_ptab=%.0f B=%.0f publish+refill=%.0f wait_byte=%.0f" % tuple(p[8:13]))
