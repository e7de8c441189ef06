"""Boundary proof: the reference's unmodified `main_flow.main()` runs end to end with `unimatch_b200.UniMatch` bound in
place of its own class and writes the same flow files (tests/dropin_main_flow.py; build container only -- the reference
checkout does not exist on the GPU box)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/demo/flow-davis"), reason="reference checkout not present")
def test_unmodified_main_flow_with_the_dropin_class(tmp_path):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_main_flow.py"), str(tmp_path)],
                       capture_output=True, text=True, timeout=1500, cwd=str(tmp_path))
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("DROPIN ")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads(line[-1][7:])
    assert len(res["files"]) >= 2 and res["shape"][2] == 2
    # same weights, same frames, fp32 both sides (CPU kernels restate the CUDA ops exactly up to the fp16 hi/lo operand
    # split): the written flows agree to ~1e-4 px; stated tolerance 1e-2 px mean as everywhere (tests/stage_checks.py)
    assert res["mean_epe"] <= 1e-2 and res["max_epe"] <= 1e-1, res
