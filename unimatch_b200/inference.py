"""Batched inference drivers around the drop-in module.

Mirrors what the reference's own drivers do around `model(...)` -- `InputPadder` (utils/utils.py:6-24), the
resize-to-multiple / resize-back / rescale logic of `inference_flow` (evaluate_flow.py:711-755), `inference_stereo`
(evaluate_stereo.py:712-843, including the horizontal-flip trick for right / bidirectional disparity, :789-796, :829-836) and
`inference_depth` (evaluate_depth.py:297-419), and the forward-backward consistency check (geometry.py:75-96,
evaluate_flow.py:774-792) -- but on BATCHES of pairs that are already device tensors, with every resize (flow-component /
disparity rescale and flips folded in) as one `um_resize_bilinear` launch and the occlusion test as one fused kernel
(`um_fb_consistency`).  File IO, visualisation and video handling stay out of scope.

`BatchedFlowRunner` is the serving-side driver (evaluate_flow.py:686-760 runs one frame pair at a time): frames are padded to a
fixed bucket, host batches are staged through two pinned buffers and copied on a side stream while the previous batch
computes, and the fixed-shape forward is replayed as a CUDA graph.
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


def _inference_size(ori, padding_factor, inference_size):
    if inference_size is None:
        return (int(math.ceil(ori[0] / padding_factor)) * padding_factor, int(math.ceil(ori[1] / padding_factor)) * padding_factor)
    return (int(inference_size[0]), int(inference_size[1]))


def _resize(x, size, scale=None, flip=False):
    """[B, C<=3, H, W] -> [B, C, *size] bilinear, align_corners=True, optional per-channel scale and horizontal flip."""
    return _OPS.resize_bilinear(x.float().contiguous(), int(size[0]), int(size[1]), scale, bool(flip))


def _hflip(x):
    """torchvision hflip of a [B, C, H, W] tensor as the same kernel at unchanged size."""
    return _resize(x, x.shape[-2:], None, True)


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
    size = _inference_size(ori, padding_factor, inference_size)
    resized = size != ori
    if resized:
        image1, image2 = _resize(image1, size), _resize(image2, size)
    if model_kwargs.setdefault("task", "flow") != "flow":
        raise ValueError("infer_flow drives the flow task only")
    flow = model(image1.contiguous(), image2.contiguous(), pred_bidir_flow=pred_bidir_flow, **model_kwargs)["flow_preds"][-1]
    if resized:                                          # resize back + rescale (u, v) by the size ratios, one launch
        flow = _resize(flow, ori, [ori[1] / size[1], ori[0] / size[0]])
    if transposed:
        flow = flow.transpose(-2, -1)       # axes only -- the reference leaves the (u, v) components in place (evaluate_flow.py:757-758)
    out = {"flow": flow}
    if pred_bidir_flow:
        half = flow.shape[0] // 2
        out["flow"], out["flow_bwd"] = flow[:half], flow[half:]
        if fwd_bwd_consistency_check:
            out["fwd_occ"], out["bwd_occ"] = forward_backward_consistency_check(out["flow"], out["flow_bwd"])
    return out


@torch.no_grad()
def infer_stereo(model, left, right, *, padding_factor=16, inference_size=None, pred_bidir_disp=False, pred_right_disp=False,
                 **model_kwargs):
    """`inference_stereo` (evaluate_stereo.py:776-836) for a batch of ImageNet-normalised pairs `[B,3,H,W]`.

    `pred_bidir_disp`: the right-view disparity comes from the SAME network on the mirrored, swapped pair, batched with the
    original pair (hflip trick, :789-792) and mirrored back (:829-836); `pred_right_disp`: only that.  Returns
    {'disp': [B,H,W]} (+ 'disp_right' for the bidirectional case); disparities are rescaled by the width ratio."""
    if pred_bidir_disp and pred_right_disp:
        raise ValueError("choose one of pred_bidir_disp / pred_right_disp")
    ori = tuple(left.shape[-2:])
    size = _inference_size(ori, padding_factor, inference_size)
    resized = size != ori
    if resized:
        left, right = _resize(left, size), _resize(right, size)
    b = left.shape[0]
    if pred_bidir_disp:
        left, right = torch.cat((left, _hflip(right)), dim=0), torch.cat((right, _hflip(left)), dim=0)
    elif pred_right_disp:
        left, right = _hflip(right), _hflip(left)
    model_kwargs["task"] = "stereo"
    disp = model(left.float().contiguous(), right.float().contiguous(), **model_kwargs)["flow_preds"][-1]     # [B or 2B, H, W]
    disp = disp.unsqueeze(1)
    sc = [ori[1] / float(size[1])] if resized else None
    if pred_bidir_disp:
        main = _resize(disp[:b], ori, sc) if resized else disp[:b]
        other = _resize(disp[b:], ori, sc, flip=True)              # resize back, rescale and mirror back in one pass
        return {"disp": main.squeeze(1), "disp_right": other.squeeze(1)}
    if resized or pred_right_disp:
        disp = _resize(disp, ori, sc, flip=pred_right_disp)
    return {"disp": disp.squeeze(1)}


@torch.no_grad()
def infer_depth(model, img_ref, img_tgt, intrinsics, pose, *, padding_factor=16, inference_size=None, min_depth=0.5,
                max_depth=10.0, num_depth_candidates=64, depth_from_argmax=False, pred_bidir_depth=False, **model_kwargs):
    """`inference_depth` (evaluate_depth.py:360-400) for a batch of ImageNet-normalised view pairs `[B,3,H,W]`,
    intrinsics `[B,3,3]` and relative poses `[B,4,4]` (target <- reference).  `min_depth` / `max_depth` are metric depths;
    the model receives their inverses, as in the reference (:389-390).  Returns {'depth': [B,H,W]} (+ 'depth_bwd')."""
    ori = tuple(img_ref.shape[-2:])
    size = _inference_size(ori, padding_factor, inference_size)
    resized = size != ori
    if resized:
        img_ref, img_tgt = _resize(img_ref, size), _resize(img_tgt, size)
    model_kwargs["task"] = "depth"
    depth = model(img_ref.float().contiguous(), img_tgt.float().contiguous(), intrinsics=intrinsics, pose=pose,
                  min_depth=1.0 / max_depth, max_depth=1.0 / min_depth, num_depth_candidates=num_depth_candidates,
                  depth_from_argmax=depth_from_argmax, pred_bidir_depth=pred_bidir_depth, **model_kwargs)["flow_preds"][-1]
    if resized:
        depth = _resize(depth.unsqueeze(1), ori).squeeze(1)
    if pred_bidir_depth:
        half = depth.shape[0] // 2
        return {"depth": depth[:half], "depth_bwd": depth[half:]}
    return {"depth": depth}


class BatchedFlowRunner:
    """Fixed-bucket, double-buffered, graph-replayed flow inference for a stream of host frame pairs.

    * bucket: every pair is replicate-padded (`InputPadder`, mode 'sintel') from `frame_size` up to a multiple of
      `padding_factor`; batches are always `batch` pairs (a short last batch is padded with copies of its last pair), so the
      device sees ONE shape and the forward can be captured once;
    * prefetch: two pinned host staging buffers and two device input buffers; the host->device copy of batch i+1 runs on
      a side stream while batch i computes, the device->host copy of result i overlaps batch i+1;
    * replay: the forward on the static input buffer is captured in a CUDA graph after two eager warm-up runs (which also
      build the module's cached operand planes outside the capture).

    `run(pairs)` takes an iterable of (img1, img2) CPU tensors `[3,H,W]` in [0,255] and yields unpadded flows `[2,H,W]`
    (CPU, pinned staging reused -- copy them if you keep them)."""

    def __init__(self, model, frame_size, batch, device, padding_factor=32, use_graph=True, **model_kwargs):
        self.model, self.kw, self.batch, self.dev = model, dict(model_kwargs), int(batch), torch.device(device)
        self.kw.setdefault("task", "flow")
        self.padder = InputPadder(frame_size, mode="sintel", padding_factor=padding_factor)
        h, w = frame_size
        left, right, top, bottom = self.padder._pad
        self.hp, self.wp = h + top + bottom, w + left + right
        shape = (self.batch, 3, self.hp, self.wp)
        self.pin = [[torch.empty(shape).pin_memory() for _ in range(2)] for _ in range(2)]       # [slot][view]
        self.dev_in = [[torch.empty(shape, device=self.dev) for _ in range(2)] for _ in range(2)]
        self.out_pin = [torch.empty((self.batch, 2, self.hp, self.wp)).pin_memory() for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.graphs = [None, None]
        self.static_out = [None, None]
        self.use_graph = use_graph

    def _forward(self, slot):
        a, b = self.dev_in[slot]
        return self.model(a, b, **self.kw)["flow_preds"][-1]

    def _capture(self):
        with torch.cuda.device(self.dev):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    for slot in range(2):
                        self.dev_in[slot][0].zero_(); self.dev_in[slot][1].zero_()
                        self._forward(slot)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for slot in range(2):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.static_out[slot] = self._forward(slot)
                self.graphs[slot] = g
            torch.cuda.synchronize()

    def _stage(self, slot, chunk):
        """host side of one batch: pad into the pinned buffers, then enqueue the H2D copy on the side stream"""
        for i in range(self.batch):
            img1, img2 = chunk[min(i, len(chunk) - 1)]
            p1, p2 = self.padder.pad(img1[None].float(), img2[None].float())
            self.pin[slot][0][i].copy_(p1[0]); self.pin[slot][1][i].copy_(p2[0])
        with torch.cuda.stream(self.copy_stream):
            for v in range(2):
                self.dev_in[slot][v].copy_(self.pin[slot][v], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return ev

    @torch.no_grad()
    def run(self, pairs):
        with torch.cuda.device(self.dev):
            if self.use_graph and self.graphs[0] is None:
                self._capture()
            it = iter(pairs)

            def next_chunk():
                chunk = []
                for p in it:
                    chunk.append(p)
                    if len(chunk) == self.batch:
                        break
                return chunk

            cur = next_chunk()
            if not cur:
                return
            main = torch.cuda.current_stream()
            self.copy_stream.wait_stream(main)
            ready = self._stage(0, cur)
            slot = 0
            done_ev, done_n = None, 0
            while cur:
                main.wait_event(ready)
                if self.use_graph:
                    self.graphs[slot].replay()
                    flow = self.static_out[slot]
                else:
                    flow = self._forward(slot)
                fwd_done = torch.cuda.Event()
                fwd_done.record(main)
                nxt = next_chunk()
                if nxt:                                        # stage the next batch while this one computes
                    ready = self._stage(slot ^ 1, nxt)
                if done_ev is not None:                        # hand out the previous batch's flows
                    done_ev.synchronize()
                    for i in range(done_n):
                        yield self.padder.unpad(self.out_pin[slot ^ 1][i])
                with torch.cuda.stream(self.copy_stream):      # D2H of this batch's flows behind its forward
                    self.copy_stream.wait_event(fwd_done)
                    self.out_pin[slot].copy_(flow, non_blocking=True)
                    done_ev = torch.cuda.Event()
                    done_ev.record(self.copy_stream)
                done_n = len(cur)
                cur, slot = nxt, slot ^ 1
            done_ev.synchronize()
            for i in range(done_n):
                yield self.padder.unpad(self.out_pin[slot ^ 1][i])
