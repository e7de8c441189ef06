"""The teacher-forced stage harness (tests/stage_checks.py) on the CPU at a small shape, with the oracle-backed CPU kernels
of tests/refops.py standing in for the CUDA ops: checks the harness itself and the host orchestration of every
`UniMatch._stage_*` method.  The CUDA kernels take the same harness at 480x832 in tests/test_stages_gpu.py."""
import torch

import refops
import stage_checks


def test_stage_harness_small_shape_on_cpu():
    refops.register_cpu_kernels()
    res = stage_checks.run(torch.device("cpu"), H=128, W=192, report=lambda *_: None)
    assert "e2e" in res and res["e2e"][0] <= stage_checks.E2E_MEAN
