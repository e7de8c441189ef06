"""Batched inference driver around the drop-in module.

Mirrors what the reference's own drivers do around `model(...)` -- `InputPadder` (utils/utils.py:6-24), the
resize-to-multiple / resize-back / flow-rescale logic of `inference_flow` (evaluate_flow.py:711-755) and the
forward-backward consistency check (geometry.py:75-96, evaluate_flow.py:774-792) -- but on BATCHES of pairs that are already
device tensors, with the occlusion test as one fused kernel (`um_fb_consistency`) instead of two warps and six
elementwise passes.  File IO, visualisation and video handling stay out of scope.
"""
import math

import torch
import torch.nn.functional as F

from . import ops  # noqa: F401  (registers torch.ops.unimatch_sm100.*)

_OPS = torch.ops.unimatch_sm100


class InputPadder:
    """Replicate-pads [..., H, W] tensors so that H and W are divisible by `padding_factor`; same constructor, `pad` and
    `unpad` as the reference class (utils/utils.py:6-24): centred padding for mode 'sintel', bottom-only in height otherwise."""

    def __init__(self, dims, mode="sintel", padding_factor=8):
        self.ht, self.wd = dims[-2:]
        ph, pw = (-self.ht) % padding_factor, (-self.wd) % padding_factor
        top = ph // 2 if mode == "sintel" else 0
        self._pad = [pw // 2, pw - pw // 2, top, ph - top]

    def pad(self, *inputs):
        return [F.pad(x, self._pad, mode="replicate") for x in inputs]

    def unpad(self, x):
        h, w = x.shape[-2:]
        left, right, top, bottom = self._pad
        return x[..., top:h - bottom, left:w - right]


def forward_backward_consistency_check(fwd_flow, bwd_flow, alpha=0.01, beta=0.5):
    """geometry.py:75-96: (fwd_occ, bwd_occ), float [B,H,W], 1 = occluded.  One fused kernel over the planar flows."""
    if fwd_flow.dim() != 4 or bwd_flow.dim() != 4 or fwd_flow.size(1) != 2 or bwd_flow.size(1) != 2:
        raise ValueError("forward_backward_consistency_check expects [B,2,H,W] flows")
    return _OPS.fb_consistency(fwd_flow.contiguous(), bwd_flow.contiguous(), float(alpha), float(beta))


@torch.no_grad()
def infer_flow(model, image1, image2, *, padding_factor, inference_size=None, pred_bidir_flow=False,
               fwd_bwd_consistency_check=False, **model_kwargs):
    """`inference_flow` (evaluate_flow.py:711-755, :774-792) for a batch of pairs `[B,3,H,W]` in [0,255].

    Portrait inputs are transposed (the model is trained with width > height), images are resized to the nearest
    multiple of `padding_factor` (or to `inference_size`), the flow is resized back and its components rescaled.
    Returns {'flow': [B,2,H,W]} plus 'flow_bwd' when `pred_bidir_flow` and 'fwd_occ' / 'bwd_occ' ([B,H,W], 1 = occluded)
    when `fwd_bwd_consistency_check`.  `model_kwargs` are forwarded to `model(...)` (attn_type, attn_splits_list, ...)."""
    if fwd_bwd_consistency_check and not pred_bidir_flow:
        raise ValueError("fwd_bwd_consistency_check needs pred_bidir_flow=True (evaluate_flow.py:774-792)")
    transposed = image1.size(-2) > image1.size(-1)
    if transposed:
        image1, image2 = image1.transpose(-2, -1), image2.transpose(-2, -1)
    ori = tuple(image1.shape[-2:])
    if inference_size is None:
        size = (int(math.ceil(ori[0] / padding_factor)) * padding_factor, int(math.ceil(ori[1] / padding_factor)) * padding_factor)
    else:
        size = (int(inference_size[0]), int(inference_size[1]))
    resized = size != ori
    if resized:
        image1 = F.interpolate(image1, size=size, mode="bilinear", align_corners=True)
        image2 = F.interpolate(image2, size=size, mode="bilinear", align_corners=True)
    if model_kwargs.setdefault("task", "flow") != "flow":
        raise ValueError("infer_flow drives the flow task only")
    flow = model(image1.contiguous(), image2.contiguous(), pred_bidir_flow=pred_bidir_flow, **model_kwargs)["flow_preds"][-1]
    if resized:
        flow = F.interpolate(flow, size=ori, mode="bilinear", align_corners=True)
        flow[:, 0] = flow[:, 0] * ori[1] / size[1]
        flow[:, 1] = flow[:, 1] * ori[0] / size[0]
    if transposed:
        flow = flow.transpose(-2, -1)       # axes only -- the reference leaves the (u, v) components in place (evaluate_flow.py:757-758)
    out = {"flow": flow}
    if pred_bidir_flow:
        half = flow.shape[0] // 2
        out["flow"], out["flow_bwd"] = flow[:half], flow[half:]
        if fwd_bwd_consistency_check:
            out["fwd_occ"], out["bwd_occ"] = forward_backward_consistency_check(out["flow"], out["flow_bwd"])
    return out
