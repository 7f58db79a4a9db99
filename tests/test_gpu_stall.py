"""The sort's look-back watchdog must surface as GSR_ERR_SORT_STALLED (include/gsrast_b200.h), not as a silently wrong
frame.  gsrast_b200/libgsrast_b200_stall.so is the same library with radix_sort.cu built -DGSR_FORCE_STALL (tile 1 of
every onesweep pass reports a stall); it is loaded in a subprocess through GSRAST_B200_LIB so the default library of
this test process stays untouched."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STALL_LIB = os.path.join(ROOT, "gsrast_b200", "libgsrast_b200_stall.so")

SCRIPT = r"""
import sys
sys.path.insert(0, %(root)r)
sys.path.insert(0, %(root)r + "/tests")
import numpy as np, torch
from gsrast_b200 import _lib, camera, scene
from gsrast_b200.views import ViewRenderer
from helpers import run_cuda
assert _lib.LIB_PATH.endswith("_stall.so"), _lib.LIB_PATH
sc = scene.make_config_scene("C1", P=40000)[0]
cam = camera.default_camera(640, 360)

def code(fn):
    try:
        fn()
        return 0
    except RuntimeError as e:
        return int(str(e).split()[2].rstrip(":"))

# 1. a call that synchronises (timings) reports the stall itself
c1 = code(lambda: run_cuda(sc, cam, timings=True))
# 2. an asynchronous call reports it at the latest on the next call of the same thread
c2a = code(lambda: run_cuda(sc, cam))
c2b = code(lambda: run_cuda(sc, cam))
# 3. the host-delivering renderer call synchronises at its end and reports it
vr = ViewRenderer.from_scene(sc, 640, 360, device="cuda")
packed = np.stack([cam.packed()] * 3).astype(np.float32)
c3 = code(lambda: vr.render_host(packed, cam.tan_fovx, cam.tan_fovy))
vr.close()
print("CODES", c1, c2a, c2b, c3)
"""


def test_forced_stall_is_reported():
    if not os.path.exists(STALL_LIB):
        pytest.skip("libgsrast_b200_stall.so not built")
    env = dict(os.environ, GSRAST_B200_LIB=STALL_LIB)
    out = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("CODES")][-1]
    c1, c2a, c2b, c3 = [int(x) for x in line.split()[1:]]
    from gsrast_b200 import _lib

    assert c1 == _lib.ERR_SORT_STALLED
    assert _lib.ERR_SORT_STALLED in (c2a, c2b)
    assert c3 == _lib.ERR_SORT_STALLED


def test_default_library_reports_no_stall(oracle):
    """Same sequence on the shipped library: no error, and the frame is the oracle's."""
    import numpy as np

    from gsrast_b200 import camera, scene
    from helpers import assert_parity, run_cuda, run_oracle

    sc = scene.make_config_scene("C1", P=40000)[0]
    cam = camera.default_camera(640, 360)
    cu = run_cuda(sc, cam, timings=True)
    assert_parity(cu, run_oracle(oracle, sc, cam))
    assert cu["times"]["kernel_launches"] > 0
    assert np.isfinite(cu["out_color"]).all()
