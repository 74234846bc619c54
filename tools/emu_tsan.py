"""Race check without a GPU: the kernels on the CPU thread-block emulator under ThreadSanitizer.

tests/native/cta_emu.h announces every CUDA thread to TSan as a fiber (switches carry no implied synchronisation) and
every barrier / barrier-OR / warp collective as a release-acquire edge of its own generation; atomics are real atomics.
Two CUDA threads touching the same shared- or global-memory location with no barrier, collective or atomic between
them are then reported as a data race -- e.g. the round-1 bug of the tree decoder (chain warp overwrites the
published byte after a speculation hit) shows up as "read in cm_dec_model_thread / previous write in
cm_decode_tree_kernel" when it is put back.  Out of reach: races BETWEEN thread blocks of one grid (they run one after
another here), memory-ordering subtleties below the barrier level.

    python tools/emu_tsan.py          (re-executes itself with libtsan preloaded; takes a few minutes)

Expected output: one line per job, "no race reported" -- except cm_decode_kernel, whose chain warp stores the decoded byte
from all 32 lanes at once (the same value to the same address, twice per byte: "2 x bz3::cm_decode_kernel", benign by
design; every job that decodes shows it)."""
import ctypes as C
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KSO, SSO, LSO = "/tmp/libemucheck_tsan.so", "/tmp/libemustages_tsan.so", "/tmp/libbzip3_emu_tsan.so"
N = 1500


def child(kind, variant):
    sys.path.insert(0, ROOT)
    import numpy as np
    from bzip3_b200 import synth
    from tests import refs
    O = refs.oracle()
    u8p = refs.u8p

    def bwt(d):
        o = np.zeros(len(d) + 16, np.uint8)
        O.orc_bwt(refs.ptr(d), refs.ptr(o), len(d))
        return o[:len(d)].copy()

    if kind == "lib":   # the whole library: host threads of the batch API queueing for `variant` stage workspaces
        import bzip3_b200
        bs = 66 * 1024
        datas = [synth.zipf_text(700 + 60 * k, seed=k).tobytes() for k in range(4)]
        states = [bzip3_b200.Bz3State(bs) for _ in datas]
        bufs = []
        for d in datas:
            b = np.zeros(bzip3_b200.bound(bs) + 64, np.uint8)
            b[:len(d)] = np.frombuffer(d, np.uint8)
            bufs.append(b)
        sizes = [len(d) for d in datas]
        out = bzip3_b200.encode_blocks(states, bufs, sizes)
        ok = all(bytes(b[:r]) == refs.oracle_encode_block(d, bs)[0] for d, b, r in zip(datas, bufs, out))
        errs = bzip3_b200.decode_blocks(states, bufs, [len(b) for b in bufs], out, sizes)
        ok = ok and all(e == 0 and bytes(b[:len(d)]) == d for d, b, e in zip(datas, bufs, errs))
        for s in states:
            s.close()
    elif kind == "stream":   # the container front end: reader / `variant` workers / writer threads (csrc/stream.h)
        import tempfile
        import bzip3_b200
        bs = 66 * 1024
        L = bzip3_b200.lib()
        line = synth.log_stream(300, seed=5).tobytes()
        data = (line * (2 * bs // len(line) + 40))[: 2 * bs + 5000]
        with tempfile.TemporaryDirectory() as tmp:
            src, dst, back = (os.path.join(tmp, n) for n in ("a", "b", "c"))
            open(src, "wb").write(data)

            def run(fn, a, b, *args):
                fi, fo = os.open(a, os.O_RDONLY), os.open(b, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o600)
                try:
                    return fn(fi, fo, *args, None, None)
                finally:
                    os.close(fi)
                    os.close(fo)
            ok = run(L.bz3_b200_encode_fd, src, dst, bs, variant) == 0
            ok = ok and run(L.bz3_b200_decode_fd, dst, back, variant) == 0 and open(back, "rb").read() == data
            blob = open(dst, "rb").read()
            open(dst, "wb").write(blob[: len(blob) - 40])
            ok = ok and run(L.bz3_b200_decode_fd, dst, back, variant) == -23
    elif kind in ("enc", "dec"):
        L = C.CDLL(KSO)
        L.emu_cm_encode.restype = C.c_int32
        L.emu_cm_encode.argtypes = [u8p, C.c_int32, u8p]
        L.emu_cm_decode.argtypes = [u8p, C.c_int32, u8p, C.c_int32]
        data = bwt(synth.zipf_text(N, seed=7))
        n = len(data)
        want = np.zeros(2 * n + 64, np.uint8)
        rw = O.orc_cm_encode(refs.ptr(data), n, refs.ptr(want))
        if kind == "enc":
            got = np.zeros(2 * n + 64, np.uint8)
            ok = L.emu_cm_encode(refs.ptr(data), n, refs.ptr(got)) == rw and bytes(got[:rw]) == bytes(want[:rw])
        else:
            back = np.zeros(n + 8, np.uint8)
            L.emu_cm_decode(refs.ptr(want), rw, refs.ptr(back), n)
            ok = bytes(back[:n]) == bytes(data)
    elif kind == "lzp":   # encoder: scan kernels + in-order commit engine (emu_stages); decoder: one thread block (emu_check)
        L, S2 = C.CDLL(KSO), C.CDLL(SSO)
        src = synth.log_stream(5000, seed=3)
        m = len(src)
        pad = np.zeros(m + 64, np.uint8)
        pad[:m] = src
        lut = np.zeros(1 << 18, np.int32)
        lp = lut.ctypes.data_as(refs.i32p)
        lw = np.zeros(m + 64, np.uint8)
        r0 = O.orc_lzp_encode(refs.ptr(pad), m, refs.ptr(lw), lp)
        enc, dec = S2.emu_stage_lzp_encode, L.emu_lzp_decode
        for f in (enc, dec):
            f.restype = C.c_int32
        enc.argtypes = [u8p, C.c_int32, u8p]
        dec.argtypes = [u8p, C.c_int32, u8p, C.c_int32, refs.i32p]
        lg = np.zeros(m + 64, np.uint8)
        ok = enc(refs.ptr(pad), m, refs.ptr(lg)) == r0 and bytes(lg[:r0]) == bytes(lw[:r0])
        d = np.zeros(refs.bound(m) + 64, np.uint8)
        lut[:] = 0
        ok = ok and dec(refs.ptr(lw), r0, refs.ptr(d), refs.bound(m), lp) == m and bytes(d[:m]) == bytes(src)
    else:  # the multi-kernel stages
        L = C.CDLL(SSO)
        L.emu_stage_crc.restype = C.c_uint32
        L.emu_stage_crc.argtypes = [u8p, C.c_uint32, C.c_uint32]
        L.emu_stage_bwt.argtypes = [u8p, C.c_uint32, u8p]
        L.emu_stage_unbwt.argtypes = [u8p, C.c_uint32, C.c_int32, u8p]
        L.emu_stage_rle_encode.argtypes = [u8p, C.c_uint32, u8p]
        L.emu_stage_rle_decode.argtypes = [u8p, C.c_uint32, u8p, C.c_uint32]
        d = synth.zipf_text(900, seed=1)
        n = len(d)
        ok = L.emu_stage_crc(refs.ptr(d), n, 1) == O.orc_crc32(1, refs.ptr(d), n)
        o1, o2, b = (np.zeros(n + 64, np.uint8) for _ in range(3))
        i1 = L.emu_stage_bwt(refs.ptr(d), n, refs.ptr(o1))
        i2 = O.orc_bwt(refs.ptr(d), refs.ptr(o2), n)
        ok = ok and i1 == i2 and bytes(o1[:n]) == bytes(o2[:n])
        ok = ok and L.emu_stage_unbwt(refs.ptr(o2), n, i2, refs.ptr(b)) == 0 and bytes(b[:n]) == bytes(d)
        r = np.repeat(np.arange(30, dtype=np.uint8), 40)
        e1, e2, bk = np.zeros(3 * len(r) + 64, np.uint8), np.zeros(3 * len(r) + 64, np.uint8), np.zeros(len(r) + 8, np.uint8)
        rg = L.emu_stage_rle_encode(refs.ptr(r), len(r), refs.ptr(e1))
        rw = O.orc_mrle_encode(refs.ptr(r), len(r), refs.ptr(e2))
        ok = ok and rg == rw and bytes(e1[:rg]) == bytes(e2[:rw])
        ok = ok and L.emu_stage_rle_decode(refs.ptr(e2), rw, refs.ptr(bk), len(r)) == 0 and bytes(bk[:len(r)]) == bytes(r)
    print("RESULT", "bit-exact" if ok else "WRONG OUTPUT", flush=True)


def main():
    if len(sys.argv) == 3:
        return child(sys.argv[1], int(sys.argv[2]))
    native = os.path.join(ROOT, "tests", "native")
    flags = ["g++", "-O1", "-g", "-fsanitize=thread", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-x", "c++"]
    subprocess.check_call(flags + ["-o", KSO, os.path.join(native, "emu_check.cpp"), os.path.join(native, "cta_emu.cpp")])
    subprocess.check_call(flags + ["-o", SSO, os.path.join(native, "emu_stages.cpp"), os.path.join(native, "cta_emu.cpp")])
    subprocess.check_call(flags[:-2] + ["-DBZ_EMU", "-include", os.path.join(native, "cta_emu.h"), "-x", "c++", "-o", LSO,
                                        os.path.join(ROOT, "bzip3_b200", "csrc", "bz3_api.cu"), os.path.join(native, "cta_emu.cpp"),
                                        "-lpthread"])
    tsan = subprocess.check_output(["g++", "-print-file-name=libtsan.so"], text=True).strip()
    env = dict(os.environ, LD_PRELOAD=tsan, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0")
    jobs = [("enc", 0), ("dec", 0), ("lzp", 0), ("stages", 0), ("lib", 1), ("lib", 2), ("lib", 4), ("stream", 3)]
    only = os.environ.get("EMU_TSAN_ONLY")   # e.g. EMU_TSAN_ONLY=lib,stages
    jobs = [j for j in jobs if not only or j[0] in only.split(",")]
    bad = 0
    for kind, v in jobs:
        if kind == "lib":
            env = dict(env, BZ3_B200_LIB=LSO, BZ3_B200_ARENAS=str(v))
        if kind == "stream":
            env = dict(env, BZ3_B200_LIB=LSO)
        out = subprocess.run([sys.executable, os.path.abspath(__file__), kind, str(v)], env=env, capture_output=True, text=True)
        text = out.stdout + out.stderr
        result = "bit-exact" if "RESULT bit-exact" in text else "NO RESULT / WRONG OUTPUT"
        races = {}
        for blk in text.split("WARNING: ThreadSanitizer: data race")[1:]:
            fr = re.findall(r"#0 (?:void )?([\w:<>, ]+?)\(", blk)
            key = " <-> ".join(sorted(set(f.strip() for f in fr[:2])))
            races[key] = races.get(key, 0) + 1
        name = {"enc": "CM encoder", "dec": "CM decoder", "lzp": "LZP encoder + decoder", "stages": "CRC / mRLE / BWT / inverse BWT",
                "lib": "library, 4-block batch, workspaces", "stream": "container front end, 3 blocks, workers"}[kind]
        line = "%-30s %d: %s, " % (name, v, result)
        line += "no race reported" if not races else "; ".join("%d x %s" % (c, k) for k, c in races.items())
        print(line, flush=True)
        bad += result != "bit-exact"
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
