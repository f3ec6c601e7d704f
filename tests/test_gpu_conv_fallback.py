"""conv_tc.cu is the fallback for shapes the window-kernel planner rejects (3-channel stride-1 stems,
channel counts that are not multiples of 16, ...).  AICAM_NO_WIN=1 forces EVERY layer onto it; the switch is
read once per process, hence the subprocess."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUBSET = "1x1_16_32 or 3x3_16_16_res1 or 3x3s2_32_64 or cout80 or cout256_res2_relu or f32_out_1x1 or 1x1s2_down or stem_3_16"


def test_fallback_kernel_forced_on_representative_shapes():
    env = dict(os.environ, AICAM_NO_WIN="1", AICAM_NO_S2D="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_conv.py"), "-q", "-x", "-m", "gpu",
                        "-k", SUBSET, "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-1500:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "8 passed" in r.stdout, tail
