"""Post-processing around the matching path: InputPadder, the batched inference wrapper and
the fused forward-backward consistency check, against the reference's golden outputs and the oracle."""
import os

import pytest
import torch

import cases
import refops
from cases import O
from unimatch_b200 import UniMatch
from unimatch_b200.inference import InputPadder, forward_backward_consistency_check, infer_flow

GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_post.pt"))


def test_oracle_matches_reference_golden():
    fwd, bwd = cases.fb_inputs()
    got = torch.stack(O.fb_consistency(fwd, bwd)).to(torch.uint8)
    assert torch.equal(got, GOLD["fb_consistency"])
    for (dims, mode, factor), pad in zip(cases.PADDER_CASES, GOLD["padder"]):
        assert O.pad_amounts(dims[-2], dims[-1], mode, factor) == pad


def test_input_padder_matches_reference():
    gen = torch.Generator().manual_seed(5)
    for (dims, mode, factor), pad in zip(cases.PADDER_CASES, GOLD["padder"]):
        p = InputPadder(dims, mode=mode, padding_factor=factor)
        assert list(p._pad) == pad
        x = torch.randn(dims, generator=gen)
        a, b = p.pad(x, 2 * x)
        assert a.shape[-2] % factor == 0 and a.shape[-1] % factor == 0 and torch.equal(b, 2 * a)
        assert torch.equal(a, O.pad_inputs(pad, x)[0])
        assert torch.equal(p.unpad(a), x)


def _tiny_model():
    cfg, sd, batch, call = cases.e2e_setup("e2e_gmflow_s1_bidir")
    m = UniMatch(**cfg["model"]).eval()
    m.load_state_dict(sd)
    return m, sd, cfg, call


@pytest.mark.parametrize("hw,size", [((60, 90), None), ((90, 60), None), ((64, 96), (64, 128))])
def test_infer_flow_host_logic_cpu(hw, size):
    """Resize-to-multiple, portrait transpose, flow rescale, fwd|bwd split and occlusion masks: the product wrapper (CPU
    kernels installed for this test only) against the oracle's restatement of evaluate_flow.py:711-792."""
    refops.register_cpu_kernels()
    m, sd, cfg, call = _tiny_model()
    kw = {k: v for k, v in call.items() if k != "pred_bidir_flow"}
    gen = torch.Generator().manual_seed(9)
    img0 = torch.rand((1, 3, *hw), generator=gen) * 255
    img1 = torch.rand((1, 3, *hw), generator=gen) * 255
    got = infer_flow(m, img0, img1, padding_factor=16, inference_size=size, pred_bidir_flow=True,
                     fwd_bwd_consistency_check=True, **kw)

    mk = {k: cfg["model"][k] for k in ("num_scales", "upsample_factor", "reg_refine")}

    def fwd_fn(a, b, bidir):
        return O.forward(sd, a, b, pred_bidir_flow=bidir, **mk, **kw)["flow_preds"][-1]

    ref = O.infer_flow(fwd_fn, img0, img1, 16, inference_size=size, pred_bidir_flow=True, fwd_bwd_consistency_check=True)
    assert set(got) == {"flow", "flow_bwd", "fwd_occ", "bwd_occ"}
    for k in ("flow", "flow_bwd"):
        assert got[k].shape == ref[k].shape == (1, 2, *hw)
        assert (got[k] - ref[k]).abs().max().item() <= 2e-3 * max(ref[k].abs().max().item(), 1.0), k
    for k in ("fwd_occ", "bwd_occ"):
        assert got[k].shape == (1, *hw)
        assert (got[k] != ref[k]).float().mean().item() < 0.02, k          # the flows differ by ~1e-3 px: few pixels flip


def test_infer_flow_argument_errors():
    m, *_ = _tiny_model()
    x = torch.zeros((1, 3, 32, 48))
    with pytest.raises(ValueError):
        infer_flow(m, x, x, padding_factor=16, fwd_bwd_consistency_check=True)
    with pytest.raises(ValueError):
        forward_backward_consistency_check(torch.zeros(1, 3, 4, 4), torch.zeros(1, 3, 4, 4))


@pytest.mark.gpu
def test_fb_consistency_kernel_matches_reference():
    fwd, bwd = cases.fb_inputs()
    occ = forward_backward_consistency_check(fwd.cuda(), bwd.cuda())
    got = torch.stack([o.cpu() for o in occ])
    ref = GOLD["fb_consistency"].float()
    assert got.shape == ref.shape and set(got.unique().tolist()) <= {0.0, 1.0}
    flips = got != ref
    # a pixel may only differ where the test statistic sits within rounding distance of the threshold
    margin = torch.stack(O.fb_consistency_margin(fwd, bwd))
    assert flips.float().mean().item() < 1e-3
    assert not flips.any() or margin[flips].max().item() < 1e-4


@pytest.mark.gpu
def test_fb_consistency_kernel_full_size_properties():
    """480x832: swapping the arguments swaps the outputs; identical opposite constant flows are consistent everywhere
    except where the warp leaves the image (zero padding)."""
    gen = torch.Generator().manual_seed(3)
    f = (torch.randn((2, 2, 480, 832), generator=gen) * 3).cuda()
    b = (torch.randn((2, 2, 480, 832), generator=gen) * 3).cuda()
    o1 = forward_backward_consistency_check(f, b)
    o2 = forward_backward_consistency_check(b, f)
    assert torch.equal(o1[0], o2[1]) and torch.equal(o1[1], o2[0])
    c = torch.zeros((1, 2, 480, 832), device="cuda")
    c[:, 0] = 5.0
    fo, bo = forward_backward_consistency_check(c, -c)
    assert fo[..., :, :-5].sum().item() == 0 and bo[..., :, 5:].sum().item() == 0
    assert fo[..., :, -4:].min().item() == 1.0 and bo[..., :, :4].min().item() == 1.0
