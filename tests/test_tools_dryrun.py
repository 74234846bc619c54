"""Dry run of tools/eval_variants.py on the CPU: the GPU library is replaced by a stand-in whose stage calls go
to the kernel emulator (tests/native) and, for BWT and whole blocks, to the oracle.  This exercises the script's
own logic (variant order, parity checks, JSON) so that a scarce GPU call is not lost to a typo; it says nothing
about the GPU library."""
import ctypes as C
import importlib.util
import json
import os
import sys

import numpy as np

import bzip3_b200
from tests import refs
from tests.test_emu_kernels import emu

ROOT = refs.ROOT


def _arr(p, n):
    return np.ctypeslib.as_array(p, shape=(max(int(n), 1),))


class FakeLib:
    """Mimics the ctypes functions of libbzip3_b200.so that the evaluation script uses."""

    def __init__(self):
        self.E = emu()
        self.O = refs.oracle()
        self.enc, self.dec, self.lzp = 0, 0, 0
        self.lut = np.zeros(1 << 18, np.int32)
        self.calls = []

    def bz3_b200_set_variant(self, h, stage, v):
        if stage == 5:
            self.enc = self.dec = v
        elif stage == 105:
            self.enc = v
        elif stage == 205:
            self.dec = v
        elif stage == 3:
            self.lzp = v

    def bz3_b200_get_variant(self, h, stage):
        return {105: self.enc, 205: self.dec, 3: self.lzp}.get(stage, 0)

    def bz3_b200_stage_bwt(self, h, pin, n, pout):
        return self.O.orc_bwt(pin, pout, n)

    def bz3_b200_stage_cm_encode(self, h, pin, n, pout):
        self.calls.append(("enc", self.enc))
        return self.E.emu_cm_encode(self.enc, pin, n, pout)

    def bz3_b200_stage_cm_decode(self, h, pin, insize, pout, n):
        self.calls.append(("dec", self.dec))
        return self.E.emu_cm_decode(self.dec, pin, insize, pout, n)

    def bz3_b200_stage_lzp_encode(self, h, pin, n, pout):
        lp = self.lut.ctypes.data_as(refs.i32p)
        self.calls.append(("lzp_enc", self.lzp))
        return (self.E.emu_lzp_encode_pf if self.lzp == 2 else self.E.emu_lzp_encode)(pin, n, pout, lp)

    def bz3_b200_stage_lzp_decode(self, h, pin, n, pout, cap):
        lp = self.lut.ctypes.data_as(refs.i32p)
        self.calls.append(("lzp_dec", self.lzp))
        return (self.E.emu_lzp_decode_bulk if self.lzp == 2 else self.E.emu_lzp_decode)(pin, n, pout, cap, lp)

    def bz3_encode_block(self, h, pbuf, n):
        err = C.c_int8(0)
        return self.O.orc_encode_block(1 << 20, pbuf, n, C.byref(err))

    def bz3_decode_block(self, h, pbuf, bufsize, csize, osize):
        err = C.c_int8(0)
        return self.O.orc_decode_block(1 << 20, pbuf, bufsize, csize, osize, C.byref(err))


class FakeState:
    def __init__(self, block_size):
        self.handle = 1
        self.block_size = block_size

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def test_eval_variants_script_logic(tmp_path, monkeypatch, capsys):
    fake = FakeLib()
    monkeypatch.setattr(bzip3_b200, "lib", lambda: fake)
    monkeypatch.setattr(bzip3_b200, "Bz3State", FakeState)
    spec = importlib.util.spec_from_file_location("eval_variants", os.path.join(ROOT, "tools", "eval_variants.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = tmp_path / "eval.json"
    monkeypatch.setattr(sys, "argv", ["eval_variants.py", "--mib", "0.004", "--reps", "1", "--out", str(out)])
    rc = mod.main()
    text = capsys.readouterr().out
    assert rc == 0, text
    assert "ALL OK" in text
    res = json.load(open(out))
    assert res["ok"] is True
    # every variant the script announces was really selected for its calls
    assert {v for k, v in fake.calls if k == "enc"} == set(mod.ENC_VARIANTS)
    assert {v for k, v in fake.calls if k == "dec"} >= set(mod.DEC_VARIANTS) - {3}
    assert {v for k, v in fake.calls if k.startswith("lzp")} == {3, 2}
    for name, rec in res["sets"].items():
        assert all(d["ok"] for d in rec["enc"].values()) and all(d["ok"] for d in rec["dec"].values()), name
