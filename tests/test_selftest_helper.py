"""The self-test behind the default kernels runs in a helper process (bzip3_b200/bz3_selftest, spawned by the library at
the first bz3_new; kernel_autoselect / selftest_in_child in bz3_api.cu), so that a candidate kernel that hangs or
faults can never take the caller's CUDA context down.  The parent/child protocol is exercised here on the CPU: the
library is built for the thread-block emulator WITH the spawn path (-DBZ_EMU_SPAWN_TEST), next to the real helper
binary -- and next to helpers that hang, crash or answer nonsense."""
import os
import shutil
import subprocess
import sys

import pytest

from tests import refs

ROOT = refs.ROOT
DIR = os.path.join(ROOT, "tests", "_build", "spawn")
SO = os.path.join(DIR, "libbzip3_b200.so")
HELPER = os.path.join(DIR, "bz3_selftest")
CSRC = os.path.join(ROOT, "bzip3_b200", "csrc")
EMU_H = os.path.join(ROOT, "tests", "native", "cta_emu.h")

SCRIPT = (
    "import sys\n"
    "sys.path.insert(0, %r)\n"
    "import bzip3_b200\n"
    "from bzip3_b200 import synth\n"
    "from tests import refs\n"
    "bs = 66 * 1024\n"
    "with bzip3_b200.Bz3State(bs) as s:\n"
    "    L = s.L\n"
    "    print('CHOICE', L.bz3_b200_get_variant(s.handle, 105), L.bz3_b200_get_variant(s.handle, 205), L.bz3_b200_get_variant(s.handle, 3))\n"
    "    data = synth.zipf_text(1200, seed=1).tobytes()\n"
    "    enc, r = s.encode_block(data)\n"
    "    want = refs.oracle_encode_block(data, bs)\n"
    "    dec, r2 = s.decode_block(enc, len(data))\n"
    "    print('EXACT', r == want[1] and enc == want[0] and dec == data)\n" % ROOT)


@pytest.fixture(scope="module")
def spawn_dir():
    os.makedirs(DIR, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [EMU_H, os.path.join(ROOT, "tests", "native", "cta_emu.cpp")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-DBZ_EMU",
                               "-DBZ_EMU_SPAWN_TEST", "-include", EMU_H, "-x", "c++", "-o", SO,
                               os.path.join(CSRC, "bz3_api.cu"), os.path.join(ROOT, "tests", "native", "cta_emu.cpp"),
                               "-lpthread", "-ldl"])
    subprocess.check_call(["g++", "-O2", "-o", os.path.join(DIR, "bz3_selftest.real"),
                           os.path.join(CSRC, "selftest_helper.cpp"), "-ldl"])
    return DIR


MARKER = os.path.join(DIR, "hung.marker")   # the "a helper had to be killed" note, kept out of /tmp for the tests


def run(extra_env, keep_marker=False):
    if not keep_marker and os.path.exists(MARKER):
        os.remove(MARKER)
    env = dict(os.environ, BZ3_B200_LIB=SO, BZ3_B200_VERBOSE="1", BZ3_B200_SELFTEST_MARKER=MARKER)
    env.update(extra_env)
    out = subprocess.run([sys.executable, "-c", SCRIPT], env=env, capture_output=True, text=True, timeout=600)
    return out.stdout + out.stderr


def install(helper_text=None):
    if os.path.exists(HELPER):
        os.remove(HELPER)
    if helper_text is None:
        shutil.copy(os.path.join(DIR, "bz3_selftest.real"), HELPER)
    elif helper_text:
        with open(HELPER, "w") as f:
            f.write(helper_text)
        os.chmod(HELPER, 0o755)


def test_choice_comes_from_the_helper_process(spawn_dir):
    install()
    text = run({})
    assert "self-test in the helper process" in text, text
    assert "EXACT True" in text, text
    # the helper's own verdict (it runs the comparison in ITS process) is what the parent reports
    direct = subprocess.run([HELPER, SO, "0"], capture_output=True, text=True, timeout=600)
    assert direct.returncode == 0 and direct.stdout.startswith("BZ3SELFTEST ")
    e, d, l = direct.stdout.split()[1:4]
    assert e in ("0", "6") and d in ("0", "8", "9") and l in ("3", "2")


@pytest.mark.parametrize("name,helper,env", [
    ("hangs", "#!/bin/sh\nexec sleep 100\n", {"BZ3_B200_SELFTEST_TIMEOUT": "2"}),
    ("nonsense", "#!/bin/sh\necho BZ3SELFTEST 6 5 2\n", {}),
    ("crashes", "#!/bin/sh\nkill -SEGV $$\n", {}),
    ("silent", "#!/bin/sh\nexit 0\n", {}),
    ("missing", "", {}),
])
def test_a_misbehaving_helper_leaves_the_proven_kernels(spawn_dir, name, helper, env):
    install(helper)
    text = run(env)
    assert "CHOICE 0 0 3" in text, text
    assert "EXACT True" in text, text
    assert "did not finish" in text, text


def test_pinned_stages_and_switch_off(spawn_dir):
    install()
    text = run({"BZ3_B200_AUTOSELECT": "0"})
    assert "CHOICE 0 0 3" in text and "self-test off" in text, text
    text = run({"BZ3_B200_CM_ENC": "4", "BZ3_B200_CM_DEC": "5", "BZ3_B200_LZP": "2"})
    assert "CHOICE 4 5 2" in text and "EXACT True" in text, text


def test_a_killed_helper_is_remembered(spawn_dir):
    """A helper that hung costs the deadline once: the next process on the machine sees the marker, skips the self-test and
    keeps the round-1 kernels; a stale marker (older than an hour) is ignored."""
    import time
    install("#!/bin/sh\nexec sleep 100\n")
    t0 = time.time()
    text = run({"BZ3_B200_SELFTEST_TIMEOUT": "2"})
    assert "CHOICE 0 0 3" in text and "did not finish" in text, text
    assert os.path.exists(MARKER)
    install()   # even a working helper is not asked while the marker is fresh
    t0 = time.time()
    text = run({"BZ3_B200_SELFTEST_TIMEOUT": "60"}, keep_marker=True)
    assert "CHOICE 0 0 3" in text and "had to be killed on this machine" in text and "EXACT True" in text, text
    old = time.time() - 7200
    os.utime(MARKER, (old, old))
    text = run({}, keep_marker=True)
    assert "self-test in the helper process" in text, text
    os.remove(MARKER)
