"""Drop-in acceptance: the reference's own test programs (test/iterative/*.c, public API only),
compiled UNMODIFIED against include/ and linked against libsavgol_b200.so instead of the reference
library (oracle/Makefile target `dropin`, built where /root/reference exists; the binaries travel in
oracle/_ref/).  Every one of the reference's 71 checks must pass with the CUDA library underneath."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref")


@pytest.mark.parametrize("name,min_pass", [("test_savgol_b200", 25), ("test_savgol_stream_b200", 19), ("test_savgol2d_b200", 27)])
def test_reference_test_program_passes_on_our_library(name, min_pass):
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        pytest.skip("drop-in binaries not built (make -C oracle dropin)")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    out = p.stdout
    assert p.returncode == 0, out[-2000:] + p.stderr[-2000:]
    assert "[FAIL]" not in out, out[-2000:]
    assert out.count("[PASS]") >= min_pass, out[-2000:]


def test_reference_demo_program_runs():
    exe = os.path.join(BIN, "test_savgol_main_b200")
    if not os.path.exists(exe):
        pytest.skip("drop-in binaries not built")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0
    assert re.search(r"Verification: PASS \(0 mismatches\)", p.stdout), p.stdout[-1500:]


def test_extension_api_from_plain_c():
    # examples/c_extensions.c: batch / multichannel stream / checkpoint / 2D batch called from C99 with host
    # buffers, each checked against the reference-shaped single-call API of the same library
    exe = os.path.join(BIN, "c_extensions_b200")
    if not os.path.exists(exe):
        pytest.skip("drop-in binaries not built")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
    assert p.stdout.count("[PASS]") == 4 and "[FAIL]" not in p.stdout, p.stdout


def test_multi_gpu_api_from_plain_c():
    # examples/c_multi_gpu.c: savgol_apply_batch_multi (host batch sharded / one long signal partitioned) and
    # savgol_apply_slices (device-resident slices, halos read from the neighbours' memory) from C99, each compared bit
    # for bit with the single-GPU call.  With one visible GPU the device list repeats it.
    exe = os.path.join(BIN, "c_multi_gpu_b200")
    if not os.path.exists(exe):
        pytest.skip("drop-in binaries not built")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
    assert p.stdout.count("[PASS]") == 5 and "[FAIL]" not in p.stdout, p.stdout
