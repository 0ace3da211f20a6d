"""InstantID image projection (perceiver Resampler, src/ip_adapter/resampler.py:77-120; called from
instantid_single_pieline.py:221-243).  It runs once per call per identity on a (2, 1, 512) input, i.e. outside the
per-step hot path (SURVEY section 8 row A13 keeps it in torch): plain torch ops over the checkpoint's state dict
(keys of the reference module: latents, proj_in, layers.{i}.0.{norm1,norm2,to_q,to_kv,to_out}, layers.{i}.1.{0,1,3},
proj_out, norm_out)."""
import torch
import torch.nn.functional as F


def resampler_forward(sd, x: torch.Tensor, heads: int, dim_head: int = 64) -> torch.Tensor:
    def ln(prefix, t):
        return F.layer_norm(t, (t.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"])

    def split(t):
        b, n, _ = t.shape
        return t.reshape(b, n, heads, -1).transpose(1, 2)

    x = x.to(sd["latents"].dtype)
    lat = sd["latents"].expand(x.shape[0], -1, -1)
    x = F.linear(x, sd["proj_in.weight"], sd["proj_in.bias"])
    n_layers = len({k.split(".")[1] for k in sd if k.startswith("layers.")})
    s = dim_head ** -0.25  # q and k are each scaled by d^-1/4 before the product (resampler.py:66-67)
    for i in range(n_layers):
        att, ff = f"layers.{i}.0", f"layers.{i}.1"
        xn, ln_lat = ln(att + ".norm1", x), ln(att + ".norm2", lat)
        q = split(F.linear(ln_lat, sd[att + ".to_q.weight"]))
        k, v = F.linear(torch.cat([xn, ln_lat], dim=1), sd[att + ".to_kv.weight"]).chunk(2, dim=-1)
        w = torch.softmax(((q * s) @ (split(k) * s).transpose(-1, -2)).float(), dim=-1).to(q.dtype)
        o = (w @ split(v)).transpose(1, 2).reshape(lat.shape[0], lat.shape[1], -1)
        lat = lat + F.linear(o, sd[att + ".to_out.weight"])
        hid = F.gelu(F.linear(ln(ff + ".0", lat), sd[ff + ".1.weight"]))
        lat = lat + F.linear(hid, sd[ff + ".3.weight"])
    return ln("norm_out", F.linear(lat, sd["proj_out.weight"], sd["proj_out.bias"]))
