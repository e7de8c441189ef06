"""Generate tests/golden/golden_post.pt by running the REFERENCE's post-processing helpers (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_post.py

`forward_backward_consistency_check` (unimatch/geometry.py:75-96) and `InputPadder` (utils/utils.py:6-24) are imported from
`/root/reference` read-only, run on the seeded inputs of `cases.py`, compared with the oracle restatements, and their outputs
stored for the box where the reference does not exist.
"""
import os
import sys
import warnings

import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")

import cases  # noqa: E402
from cases import O  # noqa: E402

from unimatch.geometry import forward_backward_consistency_check  # noqa: E402
from utils.utils import InputPadder  # noqa: E402


def main():
    out = {}
    fwd, bwd = cases.fb_inputs()
    ref = forward_backward_consistency_check(fwd, bwd)
    got = O.fb_consistency(fwd, bwd)
    assert all(torch.equal(r, g) for r, g in zip(ref, got)), "oracle fb_consistency != reference"
    print("fb_consistency: occluded fraction fwd %.3f bwd %.3f (bit-exact vs oracle)" % (ref[0].mean(), ref[1].mean()))
    out["fb_consistency"] = torch.stack(ref).to(torch.uint8)
    pads = []
    gen = torch.Generator().manual_seed(5)
    for dims, mode, factor in cases.PADDER_CASES:
        p = InputPadder(dims, mode=mode, padding_factor=factor)
        assert list(p._pad) == O.pad_amounts(dims[-2], dims[-1], mode, factor), (dims, mode, factor)
        x = torch.randn(dims, generator=gen)
        (y,) = p.pad(x)
        (yo,) = O.pad_inputs(list(p._pad), x)
        assert torch.equal(y, yo) and torch.equal(p.unpad(y), x) and torch.equal(O.unpad_output(list(p._pad), y), x)
        assert y.shape[-2] % factor == 0 and y.shape[-1] % factor == 0
        pads.append(list(p._pad))
    out["padder"] = pads
    path = os.path.join(HERE, "golden_post.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes; pads", pads)


if __name__ == "__main__":
    main()
