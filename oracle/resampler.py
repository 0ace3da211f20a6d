"""Restatement of src/ip_adapter/resampler.py:9-120 (perceiver Resampler of InstantID: face embedding (b,1,512) ->
16 image tokens of width 2048) as a function of its state dict.  TEST INFRASTRUCTURE.  Pinned by
tests/golden/resampler.pt (output of the unmodified reference module)."""
import math

import torch
import torch.nn.functional as F


def _ln(sd, name, x):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def _heads(x, heads):
    b, l, _ = x.shape
    return x.view(b, l, heads, -1).transpose(1, 2)


def resampler_forward(sd, x, heads, dim_head=64):
    latents = sd["latents"].repeat(x.size(0), 1, 1)
    x = F.linear(x, sd["proj_in.weight"], sd["proj_in.bias"])
    depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("layers."))
    for i in range(depth):
        a, f = f"layers.{i}.0", f"layers.{i}.1"
        xn, ln = _ln(sd, a + ".norm1", x), _ln(sd, a + ".norm2", latents)
        b, l, _ = ln.shape
        q = F.linear(ln, sd[a + ".to_q.weight"])
        k, v = F.linear(torch.cat((xn, ln), dim=-2), sd[a + ".to_kv.weight"]).chunk(2, dim=-1)
        q, k, v = _heads(q, heads), _heads(k, heads), _heads(v, heads)
        scale = 1 / math.sqrt(math.sqrt(dim_head))
        w = (q * scale) @ (k * scale).transpose(-2, -1)
        w = torch.softmax(w.float(), dim=-1).type(w.dtype)
        out = (w @ v).permute(0, 2, 1, 3).reshape(b, l, -1)
        latents = F.linear(out, sd[a + ".to_out.weight"]) + latents
        h = _ln(sd, f + ".0", latents)
        h = F.linear(F.gelu(F.linear(h, sd[f + ".1.weight"])), sd[f + ".3.weight"])
        latents = h + latents
    latents = F.linear(latents, sd["proj_out.weight"], sd["proj_out.bias"])
    return _ln(sd, "norm_out", latents)
