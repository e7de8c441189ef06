"""The oracle restatement vs the golden vectors produced by the reference itself
(tests/golden/make_golden.py ran `/root/reference` in the build container; see cases.py).
Runs anywhere (CPU).  Tolerance: the vectors were bit-exact at generation time; a different host may
pick different BLAS blocking, so allow fp32 rounding (1e-5 relative to the tensor's max magnitude)
on ops and the measured self-noise of the network end to end (cases.E2E_NOISE)."""
import os

import pytest
import torch

import cases

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "golden.pt"))["vectors"]


@pytest.mark.parametrize("name", sorted(cases.OP_CASES))
def test_op_matches_reference_vector(name):
    make_in, run = cases.OP_CASES[name]
    sd = cases.synthetic_state_dict(seed=326, **cases.OP_CASE_WEIGHTS[name])
    with torch.no_grad():
        got = run(make_in(), sd)
    ref = GOLD[name]
    assert got.shape == ref.shape
    tol = 2e-5 * max(ref.abs().max().item(), 1.0)
    assert (got - ref).abs().max().item() <= tol


@pytest.mark.parametrize("name", sorted(cases.E2E_CASES))
def test_e2e_matches_reference_vector(name):
    got = cases.e2e_oracle(name)
    ref = GOLD[name]
    assert got.shape == ref.shape
    mean, mx = cases.epe(got, ref)
    assert mean <= cases.e2e_tolerance(name), (mean, mx)
