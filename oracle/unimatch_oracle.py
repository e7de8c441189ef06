"""CPU oracle for the UniMatch matching inference path  --  TEST INFRASTRUCTURE ONLY.

This file is a CPU (torch fp32, ATen ops) restatement of the reference algorithm for the hot
path named in BASELINE.json.  It is *not* product code: only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s CPU-baseline / `--impl reference` legs may import it.  The product path
(`unimatch_b200`) never imports anything from `oracle/` and fails loudly without its CUDA library.

Pinning: the reference ships no tests / golden vectors, so parity is pinned by
running the reference itself in the build container: `tests/golden/make_golden.py` imports
`/root/reference/unimatch`, loads the same synthetic weights, and (a) asserts this oracle equals the
reference on every stage and end to end (bit-exact at equal thread count; the same ATen ops are
issued in the same order), and (b) writes small golden vectors to `tests/golden/*.pt` that travel
to the GPU box, where `tests/test_oracle_golden.py` re-checks this oracle against them.

Every function cites the reference file:line it restates (paths relative to /root/reference).
Weights arrive as a flat mapping with the reference's own state_dict keys.
Only eval-mode semantics are restated (the callers use model.eval() + no_grad, evaluate_flow.py:19,33).
"""
import math

import torch
import torch.nn.functional as F

C_FEAT = 128


# ----------------------------------------------------------------------------------------------
# geometry helpers                                                     (unimatch/geometry.py)
# ----------------------------------------------------------------------------------------------
def pixel_grid(b, h, w, homogeneous=False, device=None):
    """geometry.py:5-21 coords_grid: channel 0 = x, channel 1 = y (, channel 2 = 1)."""
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    planes = [xs, ys] + ([torch.ones_like(xs)] if homogeneous else [])
    g = torch.stack(planes, dim=0).float()[None].repeat(b, 1, 1, 1)
    return g if device is None else g.to(device)


def window_offsets(r_y, r_x, device):
    """geometry.py:24-32 generate_window_grid: [2ry+1, 2rx+1, 2] of (dx, dy), k = iy*(2rx+1)+ix."""
    dx = torch.linspace(-r_x, r_x, 2 * r_x + 1, device=device)
    dy = torch.linspace(-r_y, r_y, 2 * r_y + 1, device=device)
    gx, gy = torch.meshgrid([dx, dy], indexing="ij")
    return torch.stack((gx, gy), -1).transpose(0, 1).float()


def to_unit_coords(coords, h, w):
    """geometry.py:35-38 normalize_coords (pixel -> [-1,1], align_corners convention)."""
    c = torch.tensor([(w - 1) / 2.0, (h - 1) / 2.0], dtype=torch.float32, device=coords.device)
    return (coords - c) / c


def sample_bilinear(img, xy):
    """geometry.py:41-62 bilinear_sample: xy [B,2,H,W] in pixels; zeros padding, align_corners."""
    _, _, h, w = xy.shape
    gx = 2 * xy[:, 0] / (w - 1) - 1
    gy = 2 * xy[:, 1] / (h - 1) - 1
    grid = torch.stack([gx, gy], dim=-1)
    return F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=True)


def warp_by_flow(feature, flow):
    """geometry.py:65-72 flow_warp."""
    b, _, h, w = feature.shape
    return sample_bilinear(feature, pixel_grid(b, h, w).to(flow.device) + flow)


def fb_consistency(fwd_flow, bwd_flow, alpha=0.01, beta=0.5):
    """geometry.py:75-96 forward_backward_consistency_check: occlusion masks (1 = occluded) of a flow pair [B,2,H,W]."""
    mag = torch.norm(fwd_flow, dim=1) + torch.norm(bwd_flow, dim=1)
    diff_fwd = torch.norm(fwd_flow + warp_by_flow(bwd_flow, fwd_flow), dim=1)
    diff_bwd = torch.norm(bwd_flow + warp_by_flow(fwd_flow, bwd_flow), dim=1)
    thr = alpha * mag + beta
    return (diff_fwd > thr).float(), (diff_bwd > thr).float()


def fb_consistency_margin(fwd_flow, bwd_flow, alpha=0.01, beta=0.5):
    """|diff - threshold| per pixel for both directions: pixels whose margin is ~1e-5 may legitimately flip between
    implementations that round differently (test helper, not part of the reference)."""
    mag = torch.norm(fwd_flow, dim=1) + torch.norm(bwd_flow, dim=1)
    thr = alpha * mag + beta
    d0 = torch.norm(fwd_flow + warp_by_flow(bwd_flow, fwd_flow), dim=1)
    d1 = torch.norm(bwd_flow + warp_by_flow(fwd_flow, bwd_flow), dim=1)
    return (d0 - thr).abs(), (d1 - thr).abs()


def pad_amounts(h, w, mode="sintel", padding_factor=8):
    """utils/utils.py:9-16 InputPadder.__init__: [left, right, top, bottom] replicate padding that makes H, W divisible by
    `padding_factor` (centred for 'sintel', bottom-only in height otherwise)."""
    ph, pw = (-h) % padding_factor, (-w) % padding_factor
    top = ph // 2 if mode == "sintel" else 0
    return [pw // 2, pw - pw // 2, top, ph - top]


def pad_inputs(pad, *tensors):
    """utils/utils.py:18-19 InputPadder.pad."""
    return [F.pad(t, pad, mode="replicate") for t in tensors]


def unpad_output(pad, x):
    """utils/utils.py:21-24 InputPadder.unpad."""
    h, w = x.shape[-2:]
    return x[..., pad[2]:h - pad[3], pad[0]:w - pad[1]]


def infer_flow(forward_fn, image1, image2, padding_factor, inference_size=None, pred_bidir_flow=False,
               fwd_bwd_consistency_check=False):
    """evaluate_flow.py:711-755, :774-792 (inference_flow) on batched tensors: transpose portrait inputs, resize to the
    nearest multiple of `padding_factor` (or a fixed size), run, resize the flow back and rescale its components,
    optionally split fwd | bwd and derive occlusion masks.  `forward_fn(img1, img2, pred_bidir_flow)` -> [B or 2B, 2, H, W]."""
    transposed = image1.size(-2) > image1.size(-1)
    if transposed:
        image1, image2 = torch.transpose(image1, -2, -1), torch.transpose(image2, -2, -1)
    ori = image1.shape[-2:]
    size = list(inference_size) if inference_size is not None else [
        int(math.ceil(ori[0] / padding_factor)) * padding_factor, int(math.ceil(ori[1] / padding_factor)) * padding_factor]
    resized = size[0] != ori[0] or size[1] != ori[1]
    if resized:
        image1 = F.interpolate(image1, size=size, mode="bilinear", align_corners=True)
        image2 = F.interpolate(image2, size=size, mode="bilinear", align_corners=True)
    flow = forward_fn(image1, image2, pred_bidir_flow)
    if resized:
        flow = F.interpolate(flow, size=ori, mode="bilinear", align_corners=True)
        flow[:, 0] = flow[:, 0] * ori[-1] / size[-1]
        flow[:, 1] = flow[:, 1] * ori[-2] / size[-2]
    if transposed:
        flow = torch.transpose(flow, -2, -1)          # as the reference does: axes swapped, components left in place
    out = {"flow": flow}
    if pred_bidir_flow:
        half = flow.shape[0] // 2
        out["flow"], out["flow_bwd"] = flow[:half], flow[half:]
        if fwd_bwd_consistency_check:
            out["fwd_occ"], out["bwd_occ"] = fb_consistency(out["flow"], out["flow_bwd"])
    return out


def infer_stereo(forward_fn, left, right, padding_factor=16, inference_size=None, pred_bidir_disp=False, pred_right_disp=False):
    """evaluate_stereo.py:776-836 (inference_stereo) on batched, already normalised tensors: resize, optional hflip trick
    (right / bidirectional disparity), run, resize back and rescale by the width ratio, flip back.
    `forward_fn(left, right)` -> [B or 2B, H, W]."""
    ori = left.shape[-2:]
    size = list(inference_size) if inference_size is not None else [
        int(math.ceil(ori[0] / padding_factor)) * padding_factor, int(math.ceil(ori[1] / padding_factor)) * padding_factor]
    resized = size[0] != ori[0] or size[1] != ori[1]
    if resized:
        left = F.interpolate(left, size=size, mode="bilinear", align_corners=True)
        right = F.interpolate(right, size=size, mode="bilinear", align_corners=True)
    hflip = lambda t: torch.flip(t, dims=[-1])
    if pred_bidir_disp:
        new_left, new_right = hflip(right), hflip(left)
        left, right = torch.cat((left, new_left), dim=0), torch.cat((right, new_right), dim=0)
    if pred_right_disp:
        left, right = hflip(right), hflip(left)
    disp = forward_fn(left, right)
    if resized:
        disp = F.interpolate(disp.unsqueeze(1), size=ori, mode="bilinear", align_corners=True).squeeze(1)
        disp = disp * ori[-1] / float(size[-1])
    if pred_right_disp:
        disp = hflip(disp)
    if pred_bidir_disp:
        half = disp.shape[0] // 2
        return {"disp": disp[:half], "disp_right": hflip(disp[half:])}
    return {"disp": disp}


def infer_depth(forward_fn, img_ref, img_tgt, padding_factor=16, inference_size=None, pred_bidir_depth=False):
    """evaluate_depth.py:360-400 (inference_depth) on batched tensors: resize, run, resize the depth back.
    `forward_fn(img_ref, img_tgt)` -> [B or 2B, H, W] depth."""
    ori = img_ref.shape[-2:]
    size = list(inference_size) if inference_size is not None else [
        int(math.ceil(ori[0] / padding_factor)) * padding_factor, int(math.ceil(ori[1] / padding_factor)) * padding_factor]
    resized = size[0] != ori[0] or size[1] != ori[1]
    if resized:
        img_ref = F.interpolate(img_ref, size=size, mode="bilinear", align_corners=True)
        img_tgt = F.interpolate(img_tgt, size=size, mode="bilinear", align_corners=True)
    depth = forward_fn(img_ref, img_tgt)
    if resized:
        depth = F.interpolate(depth.unsqueeze(1), size=ori, mode="bilinear", align_corners=True).squeeze(1)
    if pred_bidir_depth:
        half = depth.shape[0] // 2
        return {"depth": depth[:half], "depth_bwd": depth[half:]}
    return {"depth": depth}


def rigid_flow_from_depth(depth, K, pose):
    """geometry.py:99-195 compute_flow_with_depth_pose (back_project -> camera_transform -> reproject)."""
    b, h, w = depth.shape
    grid = pixel_grid(b, h, w, homogeneous=True, device=depth.device)
    pts = torch.inverse(K).bmm(grid.view(b, 3, -1)).view(b, 3, h, w) * depth.unsqueeze(1)      # :99-111
    pts = torch.bmm(pose[:, :3, :3], pts.view(b, 3, -1)) + pose[:, :3, -1:]                      # :114-131
    proj = torch.bmm(K, pts.view(b, 3, -1)).view(b, 3, h, w)                                     # :134-157
    z = proj[:, 2].clamp(min=1e-3)
    uv = torch.stack([proj[:, 0] / z, proj[:, 1] / z], dim=1).view(b, 2, h, w)
    return uv - pixel_grid(b, h, w, device=depth.device)


# ----------------------------------------------------------------------------------------------
# window split / merge, masks, position encoding                        (unimatch/utils.py, position.py)
# ----------------------------------------------------------------------------------------------
def split_windows(x, k, channel_last=False):
    """utils.py:34-58 split_feature."""
    if channel_last:
        b, h, w, c = x.shape
        return x.view(b, k, h // k, k, w // k, c).permute(0, 1, 3, 2, 4, 5).reshape(b * k * k, h // k, w // k, c)
    b, c, h, w = x.shape
    return x.view(b, c, k, h // k, k, w // k).permute(0, 2, 4, 1, 3, 5).reshape(b * k * k, c, h // k, w // k)


def merge_windows(x, k, channel_last=False):
    """utils.py:61-81 merge_splits."""
    if channel_last:
        b, h, w, c = x.shape
        nb = b // k // k
        return x.view(nb, k, k, h, w, c).permute(0, 1, 3, 2, 4, 5).contiguous().view(nb, k * h, k * w, c)
    b, c, h, w = x.shape
    nb = b // k // k
    return x.view(nb, k, k, c, h, w).permute(0, 3, 1, 4, 2, 5).contiguous().view(nb, c, k * h, k * w)


def shift_mask_2d(h, w, wh, ww, sh, sw, device):
    """utils.py:84-108 generate_shift_window_attn_mask: additive 0 / -100 per window [K*K, Lw, Lw]."""
    ids = torch.zeros((1, h, w, 1), device=device)
    cnt = 0
    for ys in (slice(0, -wh), slice(-wh, -sh), slice(-sh, None)):
        for xs in (slice(0, -ww), slice(-ww, -sw), slice(-sw, None)):
            ids[:, ys, xs, :] = cnt
            cnt += 1
    win = split_windows(ids, w // ww, channel_last=True).view(-1, wh * ww)
    diff = win.unsqueeze(1) - win.unsqueeze(2)
    return diff.masked_fill(diff != 0, float(-100.0)).masked_fill(diff == 0, float(0.0))


def shift_mask_1d(w, ww, sw, device):
    """utils.py:199-216 generate_shift_window_attn_mask_1d: [K, ww, ww]."""
    ids = torch.zeros((1, w, 1), device=device)
    cnt = 0
    for xs in (slice(0, -ww), slice(-ww, -sw), slice(-sw, None)):
        ids[:, xs, :] = cnt
        cnt += 1
    win = ids.view(1, w // ww, ww, 1).view(-1, ww)
    diff = win.unsqueeze(1) - win.unsqueeze(2)
    return diff.masked_fill(diff != 0, float(-100.0)).masked_fill(diff == 0, float(0.0))


def sine_position(x, num_pos_feats=64, temperature=10000):
    """position.py:26-45 PositionEmbeddingSine.forward (normalize=True, scale=2*pi) -> [B,128,H,W]."""
    b, _, h, w = x.shape
    ones = torch.ones((b, h, w), device=x.device)
    y_e = ones.cumsum(1, dtype=torch.float32)
    x_e = ones.cumsum(2, dtype=torch.float32)
    eps = 1e-6
    y_e = y_e / (y_e[:, -1:, :] + eps) * (2 * math.pi)
    x_e = x_e / (x_e[:, :, -1:] + eps) * (2 * math.pi)
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32, device=x.device)
    dim_t = temperature ** (2 * (dim_t // 2) / num_pos_feats)
    px = x_e[:, :, :, None] / dim_t
    py = y_e[:, :, :, None] / dim_t
    px = torch.stack((px[:, :, :, 0::2].sin(), px[:, :, :, 1::2].cos()), dim=4).flatten(3)
    py = torch.stack((py[:, :, :, 0::2].sin(), py[:, :, :, 1::2].cos()), dim=4).flatten(3)
    return torch.cat((py, px), dim=3).permute(0, 3, 1, 2)


def add_position(f0, f1, splits):
    """utils.py:111-131 feature_add_position: per-window sine encoding added to both views."""
    if splits > 1:
        a, b_ = split_windows(f0, splits), split_windows(f1, splits)
        pos = sine_position(a)
        return merge_windows(a + pos, splits), merge_windows(b_ + pos, splits)
    pos = sine_position(f0)
    return f0 + pos, f1 + pos


def convex_upsample(flow, mask, factor, is_depth=False):
    """utils.py:134-152 upsample_flow_with_mask."""
    b, fc, h, w = flow.shape
    m = torch.softmax(mask.view(b, 1, 9, factor, factor, h, w), dim=2)
    mult = 1 if is_depth else factor
    nb = F.unfold(mult * flow, [3, 3], padding=1).view(b, fc, 9, 1, 1, h, w)
    up = torch.sum(m * nb, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(b, fc, factor * h, factor * w)


# ----------------------------------------------------------------------------------------------
# attention                                                             (unimatch/attention.py)
# ----------------------------------------------------------------------------------------------
def attn_full(q, k, v):
    """attention.py:8-16 single_head_full_attention."""
    s = torch.matmul(q, k.permute(0, 2, 1)) / (q.size(2) ** 0.5)
    return torch.matmul(torch.softmax(s, dim=2), v)


def attn_full_1d(q, k, v, h, w):
    """attention.py:19-42 single_head_full_attention_1d (attention along each image row)."""
    b, _, c = q.shape
    q, k, v = q.view(b, h, w, c), k.view(b, h, w, c), v.view(b, h, w, c)
    s = torch.matmul(q, k.permute(0, 1, 3, 2)) / (c ** 0.5)
    return torch.matmul(torch.softmax(s, dim=-1), v).view(b, -1, c)


def attn_window_2d(q, k, v, splits, shift, h, w, mask):
    """attention.py:45-104 single_head_split_window_attention."""
    b, _, c = q.shape
    nb = b * splits * splits
    wh, ww = h // splits, w // splits
    q, k, v = q.view(b, h, w, c), k.view(b, h, w, c), v.view(b, h, w, c)
    if shift:
        sh, sw = wh // 2, ww // 2
        q = torch.roll(q, shifts=(-sh, -sw), dims=(1, 2))
        k = torch.roll(k, shifts=(-sh, -sw), dims=(1, 2))
        v = torch.roll(v, shifts=(-sh, -sw), dims=(1, 2))
    q = split_windows(q, splits, channel_last=True)
    k = split_windows(k, splits, channel_last=True)
    v = split_windows(v, splits, channel_last=True)
    s = torch.matmul(q.view(nb, -1, c), k.view(nb, -1, c).permute(0, 2, 1)) / (c ** 0.5)
    if shift:
        s += mask.repeat(b, 1, 1)
    out = torch.matmul(torch.softmax(s, dim=-1), v.view(nb, -1, c))
    out = merge_windows(out.view(nb, wh, ww, c), splits, channel_last=True)
    if shift:
        out = torch.roll(out, shifts=(sh, sw), dims=(1, 2))
    return out.view(b, -1, c)


def attn_window_1d(q, k, v, splits, shift, h, w, mask):
    """attention.py:107-163 single_head_split_window_attention_1d."""
    b, _, c = q.shape
    nb = b * splits * h
    ww = w // splits
    q, k, v = q.view(b * h, w, c), k.view(b * h, w, c), v.view(b * h, w, c)
    if shift:
        sw = ww // 2
        q, k, v = (torch.roll(t, shifts=-sw, dims=1) for t in (q, k, v))
    q, k, v = (t.view(b * h, splits, ww, c).view(nb, ww, c) for t in (q, k, v))   # utils.py:155-169
    s = torch.matmul(q, k.permute(0, 2, 1)) / (c ** 0.5)
    if shift:
        s += mask.repeat(b * h, 1, 1)
    out = torch.matmul(torch.softmax(s, dim=-1), v)
    out = out.view(b, h, splits, ww, c).view(b, h, w, c)                           # utils.py:172-182
    if shift:
        out = torch.roll(out, shifts=sw, dims=2)
    return out.view(b, -1, c)


# ----------------------------------------------------------------------------------------------
# transformer                                                           (unimatch/transformer.py)
# ----------------------------------------------------------------------------------------------
def transformer_layer(sd, pfx, source, target, h, w, mask2d, mask1d, attn_type, shift, splits, has_ffn):
    """transformer.py:42-144 TransformerLayer.forward."""
    is_self = (source - target).abs().max() < 1e-6                                    # :55
    q = F.linear(source, sd[pfx + "q_proj.weight"])
    k = F.linear(target, sd[pfx + "k_proj.weight"])
    v = F.linear(target, sd[pfx + "v_proj.weight"])
    if attn_type == "swin" and splits > 1:                                            # :62-74
        msg = attn_window_2d(q, k, v, splits, shift, h, w, mask2d)
    elif attn_type == "self_swin2d_cross_1d":                                          # :76-98
        if is_self:
            msg = attn_window_2d(q, k, v, splits, shift, h, w, mask2d) if splits > 1 else attn_full(q, k, v)
        else:
            msg = attn_full_1d(q, k, v, h, w)
    elif attn_type == "self_swin2d_cross_swin1d":                                      # :100-132
        if is_self:
            msg = attn_window_2d(q, k, v, splits, shift, h, w, mask2d) if splits > 1 else attn_full(q, k, v)
        elif splits > 1:
            msg = attn_window_1d(q, k, v, splits, shift, h, w, mask1d)
        else:
            msg = attn_full_1d(q, k, v, h, w)
    else:
        msg = attn_full(q, k, v)                                                       # :134-135
    msg = F.linear(msg, sd[pfx + "merge.weight"])
    msg = F.layer_norm(msg, (msg.shape[-1],), sd[pfx + "norm1.weight"], sd[pfx + "norm1.bias"])
    if has_ffn:                                                                        # :140-142
        x = F.linear(torch.cat([source, msg], dim=-1), sd[pfx + "mlp.0.weight"])
        x = F.linear(F.gelu(x), sd[pfx + "mlp.2.weight"])
        msg = F.layer_norm(x, (x.shape[-1],), sd[pfx + "norm2.weight"], sd[pfx + "norm2.bias"])
    return source + msg


def feature_transformer(sd, f0, f1, attn_type, splits, num_layers=6):
    """transformer.py:226-294 FeatureTransformer.forward."""
    b, c, h, w = f0.shape
    f0 = f0.flatten(-2).permute(0, 2, 1)
    f1 = f1.flatten(-2).permute(0, 2, 1)
    mask2d = mask1d = None
    if "swin" in attn_type and splits > 1:
        wh, ww = h // splits, w // splits
        mask2d = shift_mask_2d(h, w, wh, ww, wh // 2, ww // 2, f0.device)
    if "swin1d" in attn_type and splits > 1:
        ww = w // splits
        mask1d = shift_mask_1d(w, ww, ww // 2, f0.device)
    c0 = torch.cat((f0, f1), dim=0)
    c1 = torch.cat((f1, f0), dim=0)
    for i in range(num_layers):
        shift = "swin" in attn_type and splits > 1 and i % 2 == 1
        pfx = "transformer.layers.%d." % i
        c0 = transformer_layer(sd, pfx + "self_attn.", c0, c0, h, w, mask2d, mask1d, attn_type, shift, splits, False)
        c0 = transformer_layer(sd, pfx + "cross_attn_ffn.", c0, c1, h, w, mask2d, mask1d, attn_type, shift, splits, True)
        c1 = torch.cat(c0.chunk(chunks=2, dim=0)[::-1], dim=0)
    f0, f1 = c0.chunk(chunks=2, dim=0)
    f0 = f0.view(b, h, w, c).permute(0, 3, 1, 2).contiguous()
    f1 = f1.view(b, h, w, c).permute(0, 3, 1, 2).contiguous()
    return f0, f1


# ----------------------------------------------------------------------------------------------
# matching                                                              (unimatch/matching.py)
# ----------------------------------------------------------------------------------------------
def global_corr_flow(f0, f1, bidir=False):
    """matching.py:7-36 global_correlation_softmax -> flow [B(,x2),2,H,W]."""
    b, c, h, w = f0.shape
    a = f0.view(b, c, -1).permute(0, 2, 1)
    corr = torch.matmul(a, f1.view(b, c, -1)).view(b, h, w, h, w) / (c ** 0.5)
    init = pixel_grid(b, h, w).to(corr.device)
    grid = init.view(b, 2, -1).permute(0, 2, 1)
    corr = corr.view(b, h * w, h * w)
    if bidir:
        corr = torch.cat((corr, corr.permute(0, 2, 1)), dim=0)
        init = init.repeat(2, 1, 1, 1)
        grid = grid.repeat(2, 1, 1)
        b = b * 2
    prob = F.softmax(corr, dim=-1)
    match = torch.matmul(prob, grid).view(b, h, w, 2).permute(0, 3, 1, 2)
    return match - init


def _window_corr(f0, f1, r_y, r_x, flow=None):
    """Shared body of matching.py:39-73 / :86-121 / :154-187: grid_sample a (2ry+1)x(2rx+1) window."""
    b, c, h, w = f0.shape
    init = pixel_grid(b, h, w).to(f0.device)
    coords = init.view(b, 2, -1).permute(0, 2, 1)
    offs = window_offsets(r_y, r_x, f0.device).reshape(-1, 2).repeat(b, 1, 1, 1)
    pts = coords.unsqueeze(-2) + offs
    if flow is not None:
        pts = pts + flow.view(b, 2, -1).permute(0, 2, 1).unsqueeze(-2)
    win = F.grid_sample(f1, to_unit_coords(pts, h, w), padding_mode="zeros", align_corners=True).permute(0, 2, 1, 3)
    a = f0.permute(0, 2, 3, 1).contiguous().view(b, h * w, 1, c)
    corr = torch.matmul(a, win).view(b, h * w, -1) / (c ** 0.5)
    return corr, pts, init


def local_corr_flow(f0, f1, radius):
    """matching.py:39-83 local_correlation_softmax -> flow [B,2,H,W]."""
    b, _, h, w = f0.shape
    corr, pts, init = _window_corr(f0, f1, radius, radius)
    ok = (pts[..., 0] >= 0) & (pts[..., 0] < w) & (pts[..., 1] >= 0) & (pts[..., 1] < h)
    corr[~ok] = -1e9
    prob = F.softmax(corr, -1)
    match = torch.matmul(prob.unsqueeze(-2), pts).squeeze(-2).view(b, h, w, 2).permute(0, 3, 1, 2)
    return match - init


def local_corr_volume(f0, f1, flow, radius):
    """matching.py:86-123 local_correlation_with_flow -> [B,(2r+1)^2,H,W] (no softmax)."""
    b, _, h, w = f0.shape
    corr, _, _ = _window_corr(f0, f1, radius, radius, flow=flow)
    return corr.view(b, h, w, -1).permute(0, 3, 1, 2).contiguous()


def global_corr_disp(f0, f1):
    """matching.py:126-151 global_correlation_softmax_stereo -> disparity [B,1,H,W]."""
    b, c, h, w = f0.shape
    xs = torch.linspace(0, w - 1, w, device=f0.device)
    corr = torch.matmul(f0.permute(0, 2, 3, 1), f1.permute(0, 2, 1, 3)) / (c ** 0.5)
    upper = torch.triu(torch.ones((w, w)), diagonal=1).type_as(f0)
    keep = (upper == 0).unsqueeze(0).unsqueeze(0).repeat(b, h, 1, 1)
    corr[~keep] = -1e9
    prob = F.softmax(corr, dim=-1)
    match = (xs.view(1, 1, 1, w) * prob).sum(-1)
    return (xs.view(1, 1, w).repeat(b, h, 1) - match).unsqueeze(1)


def local_corr_disp(f0, f1, radius):
    """matching.py:154-200 local_correlation_softmax_stereo -> disparity residual [B,1,H,W]."""
    b, _, h, w = f0.shape
    corr, pts, init = _window_corr(f0, f1, 0, radius)
    ok = (pts[..., 0] >= 0) & (pts[..., 0] < w) & (pts[..., 1] >= 0) & (pts[..., 1] < h)
    corr[~ok] = -1e9
    prob = F.softmax(corr, -1)
    match = torch.matmul(prob.unsqueeze(-2), pts).squeeze(-2).view(b, h, w, 2).permute(0, 3, 1, 2).contiguous()
    return -(match - init)[:, :1]


def plane_sweep_warp(f1, K, pose, depth, clamp_min_depth=1e-3):
    """matching.py:239-282 warp_with_pose_depth_candidates -> [B,C,D,H,W]."""
    b, d, h, w = depth.shape
    c = f1.size(1)
    grid = pixel_grid(b, h, w, homogeneous=True, device=depth.device)
    pts = torch.inverse(K).bmm(grid.view(b, 3, -1))
    pts = torch.bmm(pose[:, :3, :3], pts).unsqueeze(2).repeat(1, 1, d, 1) * depth.view(b, 1, d, h * w)
    pts = pts + pose[:, :3, -1:].unsqueeze(-1)
    pts = torch.bmm(K, pts.view(b, 3, -1)).view(b, 3, d, h * w)
    uv = pts[:, :2] / pts[:, -1:].clamp(min=clamp_min_depth)
    gx = 2 * uv[:, 0] / (w - 1) - 1
    gy = 2 * uv[:, 1] / (h - 1) - 1
    g = torch.stack([gx, gy], dim=-1)
    return F.grid_sample(f1, g.view(b, d * h, w, 2), mode="bilinear", padding_mode="zeros",
                         align_corners=True).view(b, c, d, h, w)


def depth_corr(f0, f1, K, pose, cand, from_argmax=False, bidir=False):
    """matching.py:203-236 correlation_softmax_depth -> inverse depth [B(,x2),1,H,W]."""
    c = f0.size(1)
    if bidir:
        f0, f1 = torch.cat((f0, f1), dim=0), torch.cat((f1, f0), dim=0)
        K = K.repeat(2, 1, 1)
        pose = torch.cat((pose, torch.inverse(pose)), dim=0)
        cand = cand.repeat(2, 1, 1, 1)
    vol = plane_sweep_warp(f1, K, pose, 1.0 / cand)
    corr = (f0.unsqueeze(2) * vol).sum(1) / (c ** 0.5)
    prob = F.softmax(corr, dim=1)
    if from_argmax:
        return torch.gather(cand, dim=1, index=torch.argmax(prob, dim=1, keepdim=True))
    return (prob * cand).sum(dim=1, keepdim=True)


# ----------------------------------------------------------------------------------------------
# self-attention propagation                                            (unimatch/attention.py:166-253)
# ----------------------------------------------------------------------------------------------
def propagate_global(sd, feat, flow):
    """attention.py:194-215: q = q_proj(x); k = k_proj(q)  [sic]; out = softmax(q k^T/sqrt(C)) flow."""
    b, c, h, w = feat.shape
    x = feat.view(b, c, h * w).permute(0, 2, 1)
    q = F.linear(x, sd["feature_flow_attn.q_proj.weight"], sd["feature_flow_attn.q_proj.bias"])
    k = F.linear(q, sd["feature_flow_attn.k_proj.weight"], sd["feature_flow_attn.k_proj.bias"])
    v = flow.view(b, flow.size(1), h * w).permute(0, 2, 1)
    p = torch.softmax(torch.matmul(q, k.permute(0, 2, 1)) / (c ** 0.5), dim=-1)
    return torch.matmul(p, v).view(b, h, w, v.size(-1)).permute(0, 3, 1, 2)


def propagate_local(sd, feat, flow, radius):
    """attention.py:217-253: k = k_proj(x) (not k_proj(q)); zero-padded unfold of keys and flow."""
    b, c, h, w = feat.shape
    vc = flow.size(1)
    x = feat.view(b, c, -1).permute(0, 2, 1)
    q = F.linear(x, sd["feature_flow_attn.q_proj.weight"], sd["feature_flow_attn.q_proj.bias"]).reshape(b * h * w, 1, c)
    ks = 2 * radius + 1
    kp = F.linear(x, sd["feature_flow_attn.k_proj.weight"], sd["feature_flow_attn.k_proj.bias"])
    kp = kp.permute(0, 2, 1).reshape(b, c, h, w)
    kw = F.unfold(kp, kernel_size=ks, padding=radius).view(b, c, ks ** 2, h, w)
    kw = kw.permute(0, 3, 4, 1, 2).reshape(b * h * w, c, ks ** 2)
    fw = F.unfold(flow, kernel_size=ks, padding=radius).view(b, vc, ks ** 2, h, w)
    fw = fw.permute(0, 3, 4, 2, 1).reshape(b * h * w, ks ** 2, vc)
    p = torch.softmax(torch.matmul(q, kw) / (c ** 0.5), dim=-1)
    return torch.matmul(p, fw).view(b, h, w, vc).permute(0, 3, 1, 2).contiguous()


# ----------------------------------------------------------------------------------------------
# regression refinement                                                 (unimatch/reg_refine.py)
# ----------------------------------------------------------------------------------------------
def _conv(sd, key, x, padding):
    return F.conv2d(x, sd[key + ".weight"], sd.get(key + ".bias"), padding=padding)


def update_block(sd, net, inp, corr, flow):
    """reg_refine.py:106-119 BasicUpdateBlock.forward -> (net, mask|None, delta)."""
    e = "refine.encoder."
    cor = F.relu(_conv(sd, e + "convc1", corr, 0))                                    # :67-75
    cor = F.relu(_conv(sd, e + "convc2", cor, 1))
    flo = F.relu(_conv(sd, e + "convf1", flow, 3))
    flo = F.relu(_conv(sd, e + "convf2", flo, 1))
    mf = F.relu(_conv(sd, e + "conv", torch.cat([cor, flo], dim=1), 1))
    x = torch.cat([inp, torch.cat([mf, flow], dim=1)], dim=1)
    h = net
    for sfx, pad in (("1", (0, 2)), ("2", (2, 0))):                                   # :37-52
        hx = torch.cat([h, x], dim=1)
        z = torch.sigmoid(_conv(sd, "refine.gru.convz" + sfx, hx, pad))
        r = torch.sigmoid(_conv(sd, "refine.gru.convr" + sfx, hx, pad))
        q = torch.tanh(_conv(sd, "refine.gru.convq" + sfx, torch.cat([r * h, x], dim=1), pad))
        h = (1 - z) * h + z * q
    delta = _conv(sd, "refine.flow_head.conv2", F.relu(_conv(sd, "refine.flow_head.conv1", h, 1)), 1)   # :16-19
    mask = None
    if "refine.mask.0.weight" in sd:                                                  # :98-104
        mask = _conv(sd, "refine.mask.2", F.relu(_conv(sd, "refine.mask.0", h, 1)), 0)
    return h, mask, delta


# ----------------------------------------------------------------------------------------------
# backbone                                                              (unimatch/backbone.py, trident_conv.py)
# ----------------------------------------------------------------------------------------------
def _res_block(sd, pfx, x, stride):
    """backbone.py:6-36 ResidualBlock (InstanceNorm2d: eps 1e-5, no affine, no running stats)."""
    y = F.relu(F.instance_norm(F.conv2d(x, sd[pfx + "conv1.weight"], None, stride=stride, padding=1)))
    y = F.relu(F.instance_norm(F.conv2d(y, sd[pfx + "conv2.weight"], None, padding=1)))
    if pfx + "downsample.0.weight" in sd:
        x = F.instance_norm(F.conv2d(x, sd[pfx + "downsample.0.weight"], sd[pfx + "downsample.0.bias"], stride=stride))
    return F.relu(x + y)


def backbone(sd, x, num_scales):
    """backbone.py:104-133 CNNEncoder.forward; returns features high->low resolution."""
    x = F.relu(F.instance_norm(F.conv2d(x, sd["backbone.conv1.weight"], None, stride=2, padding=3)))
    x = _res_block(sd, "backbone.layer1.0.", x, 1)
    x = _res_block(sd, "backbone.layer1.1.", x, 1)
    x = _res_block(sd, "backbone.layer2.0.", x, 2)
    x = _res_block(sd, "backbone.layer2.1.", x, 1)
    x = _res_block(sd, "backbone.layer3.0.", x, 2 if num_scales == 1 else 1)
    x = _res_block(sd, "backbone.layer3.1.", x, 1)
    x = F.conv2d(x, sd["backbone.conv2.weight"], sd["backbone.conv2.bias"])
    if num_scales == 1:
        return [x]
    wt = sd["backbone.trident_conv.weight"]                                           # trident_conv.py:64-70
    return [F.conv2d(x, wt, None, stride=s, padding=1) for s in ((1, 2) if num_scales == 2 else (1, 2, 4, 8)[:num_scales])]


# ----------------------------------------------------------------------------------------------
# the boundary                                                          (unimatch/unimatch.py:95-367)
# ----------------------------------------------------------------------------------------------
def _upsampler(sd, flow, feat, factor, is_depth=False):
    """unimatch.py:81-93 upsample_flow (learned convex branch)."""
    m = F.conv2d(torch.cat((flow, feat), dim=1), sd["upsampler.0.weight"], sd["upsampler.0.bias"], padding=1)
    m = F.conv2d(F.relu(m), sd["upsampler.2.weight"], sd["upsampler.2.bias"])
    return convex_upsample(flow, m, factor, is_depth=is_depth)


@torch.no_grad()
def forward(sd, img0, img1, *, num_scales=1, upsample_factor=8, reg_refine=False,
            attn_type=None, attn_splits_list=None, corr_radius_list=None, prop_radius_list=None,
            num_reg_refine=1, pred_bidir_flow=False, task="flow", intrinsics=None, pose=None,
            min_depth=1.0 / 0.5, max_depth=1.0 / 10, num_depth_candidates=64, depth_from_argmax=False,
            pred_bidir_depth=False, taps=None):
    """Eval-mode restatement of UniMatch.forward (unimatch.py:95-367).  `taps`, if a dict, receives
    intermediate tensors (stage-level teacher-forcing points for the parity tests)."""
    if task == "flow":                                                                 # :122-124, utils.py:23-31
        mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1).to(img1.device)
        std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1).to(img1.device)
        img0, img1 = (img0 / 255.0 - mean) / std, (img1 / 255.0 - mean) / std
    feats = backbone(sd, torch.cat((img0, img1), dim=0), num_scales)[::-1]            # :64-79
    f0_list = [torch.chunk(f, 2, 0)[0] for f in feats]
    f1_list = [torch.chunk(f, 2, 0)[1] for f in feats]

    def tap(name, t):
        if taps is not None:
            taps[name] = t.clone()

    flow = None
    preds = []
    for s in range(num_scales):
        f0, f1 = f0_list[s], f1_list[s]
        if pred_bidir_flow and s > 0:
            f0, f1 = torch.cat((f0, f1), dim=0), torch.cat((f1, f0), dim=0)
        f0_ori, f1_ori = f0, f1
        tap("s%d.f0_ori" % s, f0)
        tap("s%d.f1_ori" % s, f1)
        up = upsample_factor * (2 ** (num_scales - 1 - s))
        if task == "depth":
            Ks = intrinsics.clone()
            Ks[:, :2] = Ks[:, :2] / up
        if s > 0:
            flow = F.interpolate(flow, scale_factor=2, mode="bilinear", align_corners=True) * 2
        if flow is not None:
            if task == "stereo":
                f1 = warp_by_flow(f1, torch.cat((-flow, torch.zeros_like(flow)), dim=1))
            else:
                f1 = warp_by_flow(f1, flow)
        splits = attn_splits_list[s]
        prop_r = prop_radius_list[s]
        tap("s%d.f0_in" % s, f0)
        tap("s%d.f1_in" % s, f1)
        f0, f1 = add_position(f0, f1, splits)
        f0, f1 = feature_transformer(sd, f0, f1, attn_type, splits)
        tap("s%d.f0_tr" % s, f0)
        tap("s%d.f1_tr" % s, f1)
        if task == "depth":
            b, _, h, w = f0.shape
            cand = torch.linspace(min_depth, max_depth, num_depth_candidates).type_as(f0)
            cand = cand.view(1, num_depth_candidates, 1, 1).repeat(b, 1, h, w)
            pred = depth_corr(f0, f1, Ks, pose, cand, depth_from_argmax, pred_bidir_depth)
        elif corr_radius_list[s] == -1:
            pred = global_corr_flow(f0, f1, pred_bidir_flow) if task == "flow" else global_corr_disp(f0, f1)
        else:
            r = corr_radius_list[s]
            pred = local_corr_flow(f0, f1, r) if task == "flow" else local_corr_disp(f0, f1, r)
        flow = flow + pred if flow is not None else pred
        if task == "stereo":
            flow = flow.clamp(min=0)
        tap("s%d.flow_corr" % s, flow)
        if (pred_bidir_flow or pred_bidir_depth) and s == 0:
            f0 = torch.cat((f0, f1), dim=0)
        flow = propagate_local(sd, f0, flow, prop_r) if prop_r > 0 else propagate_global(sd, f0, flow)
        tap("s%d.flow_prop" % s, flow)
        if s != num_scales - 1:
            continue
        if not reg_refine:                                                             # :246-264
            if task == "stereo":
                pad = torch.cat((-flow, torch.zeros_like(flow)), dim=1)
                out = -_upsampler(sd, pad, f0, upsample_factor)[:, :1]
            elif task == "depth":
                pad = torch.cat((flow, torch.zeros_like(flow)), dim=1)
                out = _upsampler(sd, pad, f0, upsample_factor, True).clamp(min=min_depth, max=max_depth)[:, :1]
            else:
                out = _upsampler(sd, flow, f0, upsample_factor)
            preds.append(out)
            continue
        for it in range(num_reg_refine):                                              # :272-354
            if task == "stereo":
                corr = local_corr_volume(f0_ori, f1_ori, torch.cat((-flow, torch.zeros_like(flow)), dim=1), 4)
            elif task == "depth":
                if pred_bidir_depth and it == 0:
                    Ks = Ks.repeat(2, 1, 1)
                    pose = torch.cat((pose, torch.inverse(pose)), dim=0)
                    f0_ori, f1_ori = torch.cat((f0_ori, f1_ori), dim=0), torch.cat((f1_ori, f0_ori), dim=0)
                corr = local_corr_volume(f0_ori, f1_ori, rigid_flow_from_depth(1.0 / flow.squeeze(1), Ks, pose), 4)
            else:
                corr = local_corr_volume(f0_ori, f1_ori, flow, 4)
            proj = F.conv2d(f0, sd["refine_proj.weight"], sd["refine_proj.bias"])
            net, inp = torch.chunk(proj, chunks=2, dim=1)
            net, mask, delta = update_block(sd, torch.tanh(net), torch.relu(inp), corr, flow.clone())
            if task == "depth":
                flow = (flow - delta).clamp(min=min_depth, max=max_depth)
            else:
                flow = flow + delta
            if task == "stereo":
                flow = flow.clamp(min=0)
            tap("refine%d.flow" % it, flow)
            if it == num_reg_refine - 1:
                if mask is not None:
                    tap("refine%d.mask" % it, mask)
                if task == "depth":
                    pad = torch.cat((flow, torch.zeros_like(flow)), dim=1)
                    out = _upsampler(sd, pad, f0, upsample_factor, True).clamp(min=min_depth, max=max_depth)[:, :1]
                else:
                    out = convex_upsample(flow, mask, upsample_factor)
                preds.append(out)
    if task == "stereo":
        preds = [p.squeeze(1) for p in preds]
    if task == "depth":
        preds = [1.0 / p.squeeze(1) for p in preds]
    return {"flow_preds": preds}
