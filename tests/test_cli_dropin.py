"""Drop-in check of the C ABI with the reference's own, unmodified command-line front end (src/main.c) linked
against libbzip3_b200.so (oracle/_ref/bzip3_cli_on_b200, built by oracle/Makefile where /root/reference exists):
its .bz3 output must be byte-identical to the reference binary's and both must decode each other's files."""
import os
import subprocess

import pytest

from bzip3_b200 import synth
from tests import refs

pytestmark = pytest.mark.gpu
CLI_B200 = os.path.join(refs.ROOT, "oracle", "_ref", "bzip3_cli_on_b200")


@pytest.mark.skipif(not (os.path.exists(CLI_B200) and os.path.exists(refs.REF_CLI)),
                    reason="oracle/_ref CLI binaries not present")
@pytest.mark.parametrize("jobs", [1, 3])
def test_reference_cli_on_cuda_library(tmp_path, jobs):
    data = (synth.zipf_text(700_000, seed=3).tobytes() + bytes(50_000) + synth.log_stream(400_000, seed=4).tobytes()
            + b"tail")
    src = tmp_path / "in.bin"
    src.write_bytes(data)
    outs = {}
    for name, exe in (("ref", refs.REF_CLI), ("b200", CLI_B200)):
        out = tmp_path / f"{name}.bz3"
        with open(src, "rb") as fi, open(out, "wb") as fo:
            subprocess.run([exe, "-e", "-b", "1", "-j", str(jobs)], stdin=fi, stdout=fo, check=True, timeout=600)
        outs[name] = out.read_bytes()
    assert outs["ref"] == outs["b200"], "CLI output differs from the reference binary"
    for enc_by, dec_exe in (("ref", CLI_B200), ("b200", refs.REF_CLI)):
        r = subprocess.run([dec_exe, "-d", "-j", str(jobs)], input=outs[enc_by], capture_output=True, check=True,
                           timeout=600)
        assert r.stdout == data
    # -t (test mode) through the CUDA library
    subprocess.run([CLI_B200, "-t", "-j", str(jobs)], input=outs["b200"], check=True, timeout=600)
