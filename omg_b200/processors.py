"""diffusers attention-processor plug-ins over the fused attention kernel.

The reference's extension point for this path is the processor API (`layer.set_processor(proc)`,
src/pipelines/lora_pipeline.py:140; `unet.set_attn_processor(dict)`, instantid_single_pieline.py:207) with the call
signature `proc(attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0, **kw)`
(lora_pipeline.py:67-76; the IP variants take no `scale`, attention_processor.py:324-331).  These classes keep that
signature and the attributes they read from `attn` (`to_q`, `to_k`, `to_v`, `to_out`, `heads`, `scale`,
`residual_connection`, `rescale_output_factor`), so they work on a real diffusers `Attention` module or on any
module that quacks like one; the q/k/v/out projections stay the host module's, only the attention core
(scores + softmax + controller edit + P.V, or the decoupled text + scale * ip sum) runs in `omg_attention`.

They are the operator-level drop-in for a maintainer who keeps the diffusers module tree; the whole-path pipelines in
omg_b200.pipelines do not go through them (they also fuse the projections, LayerNorm and LoRA).
"""
from typing import Optional

import torch

from . import ops
from .prompt_attention import AttentionReplace


def _core(q, k, v, heads, items, scale, out=None, out_weight=1.0, accumulate=False):
    q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    if out is None:
        out = torch.empty_like(q)
    return ops.attention(q, k, v, out, heads, q.shape[1], k.shape[1], items, scale=scale, out_weight=out_weight,
                         accumulate=accumulate)


def _finish(attn, h, residual, input_ndim, shape4):
    h = attn.to_out[0](h)
    h = attn.to_out[1](h)
    if input_ndim == 4:
        b, c, hh, ww = shape4
        h = h.transpose(-1, -2).reshape(b, c, hh, ww)
    if getattr(attn, "residual_connection", False):
        h = h + residual
    return h / getattr(attn, "rescale_output_factor", 1.0)


def _tokens(attn, hidden_states, temb):
    if getattr(attn, "spatial_norm", None) is not None:
        hidden_states = attn.spatial_norm(hidden_states, temb)
    shape4 = None
    if hidden_states.ndim == 4:
        shape4 = hidden_states.shape
        b, c, hh, ww = shape4
        hidden_states = hidden_states.view(b, c, hh * ww).transpose(1, 2)
    if getattr(attn, "group_norm", None) is not None:
        hidden_states = attn.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
    return hidden_states, shape4


class FusedRegionAttnProcessor:
    """RegionControlNet_AttnProcessor (src/pipelines/lora_pipeline.py:61-133) without the probability tensor: the
    prompt-to-prompt edit of the controller is applied through the kernel's batch-row remap.  Batch rows are
    (uncond0, uncond1, cond0, cond1) as in the pipelines (lora_pipeline.py:409,491)."""

    def __init__(self, attention_op=None, controller: Optional[AttentionReplace] = None, place_in_unet=None):
        self.attention_op, self.controller, self.place_in_unet = attention_op, controller, place_in_unet

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0,
                 **cross_attention_kwargs):
        if attention_mask is not None:
            raise NotImplementedError("attention masks are never used on the OMG path (prepare_attention_mask(None))")
        residual = hidden_states
        input_ndim = hidden_states.ndim
        x, shape4 = _tokens(attn, hidden_states, temb)
        is_cross = encoder_hidden_states is not None
        ctx = encoder_hidden_states if is_cross else x
        if is_cross and getattr(attn, "norm_cross", False):
            ctx = attn.norm_encoder_hidden_states(ctx)
        q, k, v = attn.to_q(x), attn.to_k(ctx), attn.to_v(ctx)
        B = q.shape[0]
        items = [(b, b, b, b) for b in range(B)]
        c = self.controller
        out = None
        if c is not None:
            if B != 4:
                raise ValueError("the prompt-to-prompt path expects the batch (uncond0, uncond1, cond0, cond1)")
            if is_cross:
                coef_base, coef_keep = c.cross_edit()
                # V rows mixed in value space: (M diag(alpha)) V_1 and diag(1-alpha) V_1 (V = ctx W_v is linear)
                v1 = v[3:4].contiguous()
                va = ops.ctx_mix(v1, coef_base.to(v.device).contiguous())
                vx = torch.cat([v, va], dim=0)
                items[3] = (3, 2, 2, 4)
                out = _core(q, k, vx, attn.heads, items, attn.scale)
                if coef_keep is not None:
                    vb = ops.ctx_mix(v1, coef_keep.to(v.device).contiguous())
                    _core(q, k, torch.cat([v, vb], dim=0), attn.heads, [(3, 3, 3, 4)], attn.scale, out=out,
                          accumulate=True)
            elif c.self_replace_active(q.shape[1]):
                items[3] = (3, 2, 2, 3)
            c.advance(1)
        if out is None:
            out = _core(q, k, v, attn.heads, items, attn.scale)
        return _finish(attn, out, residual, input_ndim, shape4)


class FusedAttnProcessor:
    """AttnProcessor / AttnProcessor2_0 (src/ip_adapter/attention_processor.py:15-85,207-293): plain attention."""

    def __init__(self, hidden_size=None, cross_attention_dim=None):
        pass

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        return FusedRegionAttnProcessor()(attn, hidden_states, encoder_hidden_states, attention_mask, temb)


class FusedIPAttnProcessor(torch.nn.Module):
    """IPAttnProcessor / IPAttnProcessor2_0 (src/ip_adapter/attention_processor.py:88-204,296-424): text attention
    over the first L-num_tokens context rows plus scale * image attention over the last num_tokens rows projected by
    this processor's own to_k_ip / to_v_ip; the two terms are separately normalised and summed inside the kernel
    (accumulate).  The reference's dead-store `attn_map` (:402-403) is not reproduced."""

    def __init__(self, hidden_size, cross_attention_dim=None, scale=1.0, num_tokens=4):
        super().__init__()
        self.hidden_size, self.cross_attention_dim = hidden_size, cross_attention_dim
        self.scale, self.num_tokens = scale, num_tokens
        self.to_k_ip = torch.nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)
        self.to_v_ip = torch.nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)

    def forward(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        if attention_mask is not None:
            raise NotImplementedError("attention masks are never used on the OMG path")
        residual = hidden_states
        input_ndim = hidden_states.ndim
        x, shape4 = _tokens(attn, hidden_states, temb)
        q = attn.to_q(x)
        B = q.shape[0]
        items = [(b, b, b, b) for b in range(B)]
        if encoder_hidden_states is None:
            out = _core(q, attn.to_k(x), attn.to_v(x), attn.heads, items, attn.scale)
            return _finish(attn, out, residual, input_ndim, shape4)
        end = encoder_hidden_states.shape[1] - self.num_tokens
        txt, ip = encoder_hidden_states[:, :end, :], encoder_hidden_states[:, end:, :]
        if getattr(attn, "norm_cross", False):
            txt = attn.norm_encoder_hidden_states(txt)
        out = _core(q, attn.to_k(txt), attn.to_v(txt), attn.heads, items, attn.scale)
        _core(q, self.to_k_ip(ip), self.to_v_ip(ip), attn.heads, items, attn.scale, out=out, out_weight=self.scale,
              accumulate=True)
        return _finish(attn, out, residual, input_ndim, shape4)
