"""The CPU statement of the fused FFN op (tests/refops.py::ffn_tc, what the GPU kernel is tested against) is itself the
reference's TransformerLayer FFN (transformer.py:137-144): message = mlp(cat[source, message]); message = norm2(message);
return source + message -- with mlp = Linear(256, 1024, bias=False), GELU (exact erf), Linear(1024, 128, bias=False)."""
import torch

import refops
from unimatch_b200 import ops


def test_refops_ffn_is_the_reference_ffn():
    gen = torch.Generator().manual_seed(11)
    rows, hidden = 512, 1024
    source, message = torch.randn((rows, 128), generator=gen), torch.randn((rows, 128), generator=gen)
    w1 = torch.randn((hidden, 256), generator=gen) * (2.0 / 256) ** 0.5
    w2 = torch.randn((128, hidden), generator=gen) * (1.0 / hidden) ** 0.5
    gamma, beta = torch.randn(128, generator=gen), torch.randn(128, generator=gen)
    # the reference's op sequence in plain fp32 torch
    h = torch.nn.functional.gelu(torch.cat([source, message], -1) @ w1.t())
    ref = source + torch.nn.functional.layer_norm(h @ w2.t(), (128,), gamma, beta, eps=1e-5)
    # the op's CPU statement on (hi, lo) planes
    planes = []
    for x in (source, message):
        buf = torch.zeros((2, rows, 128), dtype=torch.float16)
        refops.split_planes(x, buf, 0)
        planes.append(buf)
    out_f = torch.zeros((rows, 128))
    out_s = torch.zeros((2, rows, 128), dtype=torch.float16)
    refops.ffn_tc(planes[0], planes[1], ops.prep_conv_weight(w1[:, :, None, None], [128, 128], hidden),
                  ops.prep_conv_weight(w2[:, :, None, None], [hidden], 128), source, gamma, beta, out_f, out_s, rows)
    scale = ref.abs().max().item()
    assert (out_f - ref).abs().max().item() <= 2e-5 * scale
    assert ((out_s[0].float() + out_s[1].float()) - ref).abs().max().item() <= 2e-5 * scale
    assert ops.ffn_tc_supported(rows) and not ops.ffn_tc_supported(rows + 16)
