"""bzip3_b200 -- host-side mirror of the libbz3 block API on top of the CUDA library.

The product is `libbzip3_b200.so` (C ABI of include/libbz3.h + include/bz3_b200.h, built from
bzip3_b200/csrc by bzip3_b200/build.py).  This package is the thin Python binding used by the tests
and the benchmark; names follow the reference API (kspalaiologos/bzip3 include/libbz3.h).  There is no
CPU implementation here: if the library is missing or no GPU is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BZ3_B200_LIB") or os.path.join(_HERE, "libbzip3_b200.so")  # env override: tuning builds

BZ3_OK = 0
BZ3_ERR_OUT_OF_BOUNDS = -1
BZ3_ERR_BWT = -2
BZ3_ERR_CRC = -3
BZ3_ERR_MALFORMED_HEADER = -4
BZ3_ERR_TRUNCATED_DATA = -5
BZ3_ERR_DATA_TOO_BIG = -6
BZ3_ERR_INIT = -7
BZ3_ERR_DATA_SIZE_TOO_SMALL = -8

STAGES = ("h2d", "crc", "rle", "lzp", "bwt", "cm", "d2h")

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_lib = None


class Bz3Error(RuntimeError):
    pass


def bound(n: int) -> int:
    """bz3_bound (reference src/libbz3.c:510)."""
    return n + n // 50 + 32


def lib():
    """Loads libbzip3_b200.so and declares its prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Bz3Error(f"{LIB_PATH} is missing: run `python -m bzip3_b200.build` (nvcc, sm_100a). "
                       "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.bz3_version.restype = C.c_char_p
    L.bz3_new.restype = C.c_void_p
    L.bz3_new.argtypes = [C.c_int32]
    L.bz3_free.argtypes = [C.c_void_p]
    L.bz3_last_error.restype = C.c_int8
    L.bz3_last_error.argtypes = [C.c_void_p]
    L.bz3_strerror.restype = C.c_char_p
    L.bz3_strerror.argtypes = [C.c_void_p]
    L.bz3_bound.restype = C.c_size_t
    L.bz3_bound.argtypes = [C.c_size_t]
    L.bz3_min_memory_needed.restype = C.c_size_t
    L.bz3_min_memory_needed.argtypes = [C.c_int32]
    L.bz3_encode_block.restype = C.c_int32
    L.bz3_encode_block.argtypes = [C.c_void_p, _u8p, C.c_int32]
    L.bz3_decode_block.restype = C.c_int32
    L.bz3_decode_block.argtypes = [C.c_void_p, _u8p, C.c_size_t, C.c_int32, C.c_int32]
    L.bz3_encode_blocks.restype = None
    L.bz3_encode_blocks.argtypes = [C.POINTER(C.c_void_p), C.POINTER(_u8p), _i32p, C.c_int32]
    L.bz3_decode_blocks.restype = None
    L.bz3_decode_blocks.argtypes = [C.POINTER(C.c_void_p), C.POINTER(_u8p), C.POINTER(C.c_size_t), _i32p, _i32p,
                                    C.c_int32]
    L.bz3_compress.restype = C.c_int
    L.bz3_compress.argtypes = [C.c_uint32, _u8p, _u8p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.bz3_decompress.restype = C.c_int
    L.bz3_decompress.argtypes = [_u8p, _u8p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.bz3_orig_size_sufficient_for_decode.restype = C.c_int
    L.bz3_orig_size_sufficient_for_decode.argtypes = [_u8p, C.c_size_t, C.c_int32]
    # extensions (include/bz3_b200.h)
    L.bz3_b200_device_count.restype = C.c_int
    L.bz3_b200_state_device.restype = C.c_int
    L.bz3_b200_state_device.argtypes = [C.c_void_p]
    L.bz3_b200_set_devices.restype = C.c_int
    L.bz3_b200_set_devices.argtypes = [C.c_int]
    L.bz3_b200_device_bytes.restype = C.c_size_t
    L.bz3_b200_device_bytes.argtypes = [C.c_void_p]
    u64p = C.POINTER(C.c_uint64)
    L.bz3_b200_encode_fd.restype = C.c_int
    L.bz3_b200_encode_fd.argtypes = [C.c_int, C.c_int, C.c_int32, C.c_int, u64p, u64p]
    L.bz3_b200_decode_fd.restype = C.c_int
    L.bz3_b200_decode_fd.argtypes = [C.c_int, C.c_int, C.c_int, u64p, u64p]
    L.bz3_b200_encode_fd2.restype = C.c_int
    L.bz3_b200_encode_fd2.argtypes = [C.c_int, C.c_int, C.c_int32, C.c_int, C.c_int, u64p, u64p]
    L.bz3_b200_decode_fd2.restype = C.c_int
    L.bz3_b200_decode_fd2.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, u64p, u64p]
    L.bz3_b200_workspace_bytes.restype = C.c_size_t
    L.bz3_b200_workspace_bytes.argtypes = [C.c_void_p]
    L.bz3_b200_upload.restype = C.c_int
    L.bz3_b200_upload.argtypes = [C.c_void_p, _u8p, C.c_int32]
    L.bz3_b200_download.restype = C.c_int
    L.bz3_b200_download.argtypes = [C.c_void_p, _u8p, C.c_int32]
    L.bz3_b200_encode_resident.restype = C.c_int32
    L.bz3_b200_encode_resident.argtypes = [C.c_void_p, C.c_int32]
    L.bz3_b200_decode_resident.restype = C.c_int32
    L.bz3_b200_decode_resident.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    L.bz3_b200_encode_resident_many.restype = None
    L.bz3_b200_encode_resident_many.argtypes = [C.POINTER(C.c_void_p), _i32p, _i32p, C.c_int32]
    L.bz3_b200_decode_resident_many.restype = None
    L.bz3_b200_decode_resident_many.argtypes = [C.POINTER(C.c_void_p), _i32p, _i32p, _i32p, C.c_int32]
    L.bz3_b200_stats_reset.argtypes = [C.c_void_p]
    L.bz3_b200_stage_ms.restype = C.c_double
    L.bz3_b200_stage_ms.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.bz3_b200_kernel_launches.restype = C.c_uint64
    L.bz3_b200_kernel_launches.argtypes = [C.c_void_p]
    L.bz3_b200_last_sort_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), _i32p, C.POINTER(C.c_double)]
    L.bz3_b200_stage_crc.restype = C.c_uint32
    L.bz3_b200_stage_crc.argtypes = [C.c_void_p, _u8p, C.c_int32]
    L.bz3_b200_stage_rle_encode.restype = C.c_int32
    L.bz3_b200_stage_rle_encode.argtypes = [C.c_void_p, _u8p, C.c_int32, _u8p]
    L.bz3_b200_stage_rle_decode.restype = C.c_int
    L.bz3_b200_stage_rle_decode.argtypes = [C.c_void_p, _u8p, C.c_int32, _u8p, C.c_int32]
    L.bz3_b200_stage_lzp_encode.restype = C.c_int32
    L.bz3_b200_stage_lzp_encode.argtypes = [C.c_void_p, _u8p, C.c_int32, _u8p]
    L.bz3_b200_stage_lzp_decode.restype = C.c_int32
    L.bz3_b200_stage_lzp_decode.argtypes = [C.c_void_p, _u8p, C.c_int32, _u8p, C.c_int32]
    L.bz3_b200_stage_bwt.restype = C.c_int32
    L.bz3_b200_stage_bwt.argtypes = [C.c_void_p, _u8p, C.c_int32, _u8p]
    L.bz3_b200_stage_unbwt.restype = C.c_int32
    L.bz3_b200_stage_unbwt.argtypes = [C.c_void_p, _u8p, C.c_int32, C.c_int32, _u8p]
    L.bz3_b200_stage_cm_encode.restype = C.c_int32
    L.bz3_b200_stage_cm_encode.argtypes = [C.c_void_p, _u8p, C.c_int32, _u8p]
    L.bz3_b200_stage_cm_decode.restype = C.c_int
    L.bz3_b200_stage_cm_decode.argtypes = [C.c_void_p, _u8p, C.c_int32, _u8p, C.c_int32]
    _lib = L
    return L


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_u8p)


class Bz3State:
    """Owns one `struct bz3_state` (device block buffers + stream; stage workspaces are shared per device).  Mirrors bz3_new / bz3_free."""

    def __init__(self, block_size: int):
        self.L = lib()
        self.block_size = block_size
        self.handle = self.L.bz3_new(block_size)
        if not self.handle:
            raise Bz3Error(f"bz3_new({block_size}) failed (block size out of range, no GPU, or out of device memory)")

    def close(self):
        if self.handle:
            self.L.bz3_free(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def last_error(self) -> int:
        return int(self.L.bz3_last_error(self.handle))

    def strerror(self) -> str:
        return self.L.bz3_strerror(self.handle).decode()

    def encode_block(self, data) -> tuple[bytes | None, int]:
        """bz3_encode_block on a host buffer.  Returns (block bytes or None, return value)."""
        a = np.frombuffer(bytes(data), dtype=np.uint8)
        buf = np.zeros(bound(max(len(a), self.block_size)) + 64, dtype=np.uint8)
        buf[:len(a)] = a
        r = self.L.bz3_encode_block(self.handle, _ptr(buf), len(a))
        return (bytes(buf[:r]) if r >= 0 else None), r

    def decode_block(self, block, orig_size: int, buffer_size: int | None = None,
                     compressed_size: int | None = None) -> tuple[bytes | None, int]:
        cap = bound(self.block_size) + 64
        buf = np.zeros(max(cap, len(block)), dtype=np.uint8)
        buf[:len(block)] = np.frombuffer(bytes(block), dtype=np.uint8)
        bs = cap if buffer_size is None else buffer_size
        cs = len(block) if compressed_size is None else compressed_size
        r = self.L.bz3_decode_block(self.handle, _ptr(buf), bs, cs, orig_size)
        return (bytes(buf[:r]) if r >= 0 else None), r

    def stage_ms(self, decode: bool = False) -> dict:
        return {name: self.L.bz3_b200_stage_ms(self.handle, i, 1 if decode else 0) for i, name in enumerate(STAGES)}

    def launches(self) -> int:
        return int(self.L.bz3_b200_kernel_launches(self.handle))

    def stats_reset(self):
        self.L.bz3_b200_stats_reset(self.handle)


def encode_blocks(states, buffers, sizes):
    """bz3_encode_blocks: buffers are numpy uint8 arrays of capacity bound(size); sizes updated in place."""
    L = lib()
    n = len(states)
    hs = (C.c_void_p * n)(*[s.handle for s in states])
    bp = (_u8p * n)(*[_ptr(b) for b in buffers])
    sz = (C.c_int32 * n)(*sizes)
    L.bz3_encode_blocks(hs, bp, sz, n)
    return [int(x) for x in sz]


def decode_blocks(states, buffers, buffer_sizes, sizes, orig_sizes):
    L = lib()
    n = len(states)
    hs = (C.c_void_p * n)(*[s.handle for s in states])
    bp = (_u8p * n)(*[_ptr(b) for b in buffers])
    bs = (C.c_size_t * n)(*buffer_sizes)
    sz = (C.c_int32 * n)(*sizes)
    osz = (C.c_int32 * n)(*orig_sizes)
    L.bz3_decode_blocks(hs, bp, bs, sz, osz, n)
    return [int(s.last_error) for s in states]
