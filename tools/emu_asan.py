"""Memory check of the whole library without a GPU: builds bzip3_b200/csrc/bz3_api.cu for the CPU thread-block
emulator (tests/native/cta_emu.h) with -fsanitize=address and round-trips every edge case of the test corpus through
the block API with the default and the opt-in kernels.  "Device" buffers are heap blocks there and shared memory is
ordinary memory, so an out-of-bounds access of any kernel is an AddressSanitizer report.

    python tools/emu_asan.py        (re-executes itself with libasan preloaded)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = "/tmp/libbzip3_emu_asan.so"


def main():
    if os.environ.get("BZ3_EMU_ASAN_CHILD") != "1":
        emu_h = os.path.join(ROOT, "tests", "native", "cta_emu.h")
        subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=address", "-fno-omit-frame-pointer", "-std=c++17", "-fPIC",
                               "-shared", "-fvisibility=hidden", "-DBZ_EMU", "-include", emu_h, "-x", "c++", "-o", SO,
                               os.path.join(ROOT, "bzip3_b200", "csrc", "bz3_api.cu"),
                               os.path.join(ROOT, "tests", "native", "cta_emu.cpp"), "-lpthread"])
        asan = subprocess.check_output(["g++", "-print-file-name=libasan.so"], text=True).strip()
        env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0",
                   BZ3_EMU_ASAN_CHILD="1", BZ3_B200_LIB=SO)
        return subprocess.call([sys.executable, os.path.abspath(__file__)], env=env)
    sys.path.insert(0, ROOT)
    import bzip3_b200
    from bzip3_b200 import synth
    from tests import refs
    bs = 65 * 1024 + 1024
    cases = [(n, bytes(d[:1200])) for n, d in synth.edge_cases()]
    with bzip3_b200.Bz3State(bs) as s:
        if True:
            for name, data in cases:
                enc, r = s.encode_block(data)
                want = refs.oracle_encode_block(data, bs)
                assert r == want[1] and enc == want[0], name
                dec, r2 = s.decode_block(enc, len(data))
                assert dec == data, name
                if len(enc) > 40:   # a damaged block must not make any kernel read or write out of bounds either
                    bad = bytearray(enc)
                    bad[len(bad) // 2] ^= 0x10
                    s.decode_block(bytes(bad), len(data))
                    s.decode_block(enc[: len(enc) // 2], len(data))
            print("%d cases round-tripped, damaged and truncated blocks decoded, no AddressSanitizer report" % len(cases), flush=True)
    # the file container front end (csrc/stream.h): reader, workers and writer threads on a file of three blocks, whole
    # and cut short inside a block
    import ctypes as C
    import tempfile
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
        assert run(L.bz3_b200_encode_fd, src, dst, bs, 3) == 0
        assert run(L.bz3_b200_decode_fd, dst, back, 2) == 0 and open(back, "rb").read() == data
        blob = open(dst, "rb").read()
        open(dst, "wb").write(blob[: len(blob) - 40])
        assert run(L.bz3_b200_decode_fd, dst, back, 3) == -23
    print("container front end: 3 blocks through reader / 3 workers / writer, whole and truncated, no AddressSanitizer report", flush=True)
    print("done")
    return 0


if __name__ == "__main__":
    sys.exit(main())
