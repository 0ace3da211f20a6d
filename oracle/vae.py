"""TEST INFRASTRUCTURE ONLY - fp32 CPU restatement of the SDXL VAE decoder.

What the reference runs after the loop (src/pipelines/lora_pipeline.py:634-661): the VAE is forced to fp32
(`needs_upcasting`, :642-646), `image = vae.decode(latents / vae.config.scaling_factor)`, then
`image_processor.postprocess`.  The module itself is diffusers 0.25 `AutoencoderKL` [3P: not under /root/reference,
not installed]; its published decoder (post_quant_conv 1x1 -> conv_in -> mid block (ResnetBlock2D, single-head
attention with a GroupNorm and residual, ResnetBlock2D) -> four UpDecoderBlock2D of three ResnetBlock2D each with a
nearest-2x + conv upsampler on all but the last -> GroupNorm + SiLU -> conv_out; all GroupNorms 32 groups, eps 1e-6,
no time embedding) is restated here on the diffusers key layout.  PARITY UNPINNED: no golden vectors exist for this
module offline; it is anchored on the published architecture and parameter shapes only.
"""
from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class VaeConfig:
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    latent_channels: int = 4
    out_channels: int = 3
    norm_num_groups: int = 32
    scaling_factor: float = 0.13025

    @staticmethod
    def sdxl() -> "VaeConfig":
        return VaeConfig()

    @staticmethod
    def tiny() -> "VaeConfig":
        return VaeConfig(block_out_channels=(64, 64, 128, 128))


def decoder_param_shapes(cfg: VaeConfig) -> Dict[str, Tuple[int, ...]]:
    S: Dict[str, Tuple[int, ...]] = {}

    def conv(n, i, o, k):
        S[n + ".weight"], S[n + ".bias"] = (o, i, k, k), (o,)

    def norm(n, c):
        S[n + ".weight"], S[n + ".bias"] = (c,), (c,)

    def res(n, i, o):
        norm(n + ".norm1", i)
        conv(n + ".conv1", i, o, 3)
        norm(n + ".norm2", o)
        conv(n + ".conv2", o, o, 3)
        if i != o:
            conv(n + ".conv_shortcut", i, o, 1)

    boc = cfg.block_out_channels
    top = boc[-1]
    conv("post_quant_conv", cfg.latent_channels, cfg.latent_channels, 1)
    conv("decoder.conv_in", cfg.latent_channels, top, 3)
    res("decoder.mid_block.resnets.0", top, top)
    a = "decoder.mid_block.attentions.0"
    norm(a + ".group_norm", top)
    for p in ("to_q", "to_k", "to_v", "to_out.0"):
        S[f"{a}.{p}.weight"], S[f"{a}.{p}.bias"] = (top, top), (top,)
    res("decoder.mid_block.resnets.1", top, top)
    prev = top
    rev = tuple(reversed(boc))
    for i, ch in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            res(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else ch, ch)
        if i < len(rev) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", ch, ch, 3)
        prev = ch
    norm("decoder.conv_norm_out", boc[0])
    conv("decoder.conv_out", boc[0], cfg.out_channels, 3)
    return S


def _gn(x, sd, n, groups, silu):
    y = F.group_norm(x, groups, sd[n + ".weight"], sd[n + ".bias"], eps=1e-6)
    return F.silu(y) if silu else y


def _res(x, sd, n, groups):
    h = F.conv2d(_gn(x, sd, n + ".norm1", groups, True), sd[n + ".conv1.weight"], sd[n + ".conv1.bias"], padding=1)
    h = F.conv2d(_gn(h, sd, n + ".norm2", groups, True), sd[n + ".conv2.weight"], sd[n + ".conv2.bias"], padding=1)
    if n + ".conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[n + ".conv_shortcut.weight"], sd[n + ".conv_shortcut.bias"])
    return x + h


def _attn(x, sd, n, groups):
    """diffusers `Attention(heads=1, dim_head=C, norm_num_groups, residual_connection=True)` on (B, C, H, W)."""
    B, C, H, W = x.shape
    t = _gn(x, sd, n + ".group_norm", groups, False).view(B, C, H * W).transpose(1, 2)
    q = F.linear(t, sd[n + ".to_q.weight"], sd[n + ".to_q.bias"])
    k = F.linear(t, sd[n + ".to_k.weight"], sd[n + ".to_k.bias"])
    v = F.linear(t, sd[n + ".to_v.weight"], sd[n + ".to_v.bias"])
    p = torch.softmax(q @ k.transpose(1, 2) * C ** -0.5, dim=-1)
    o = F.linear(p @ v, sd[n + ".to_out.0.weight"], sd[n + ".to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(B, C, H, W)


def decode(sd: Dict[str, torch.Tensor], latents: torch.Tensor, cfg: VaeConfig = VaeConfig()) -> torch.Tensor:
    """`vae.decode(latents / scaling_factor).sample`: (B, 4, h, w) -> (B, 3, 8h, 8w), nominally in [-1, 1]."""
    g = cfg.norm_num_groups
    z = latents.float() / cfg.scaling_factor
    z = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    x = F.conv2d(z, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    x = _res(x, sd, "decoder.mid_block.resnets.0", g)
    x = _attn(x, sd, "decoder.mid_block.attentions.0", g)
    x = _res(x, sd, "decoder.mid_block.resnets.1", g)
    n_up = len(cfg.block_out_channels)
    for i in range(n_up):
        for j in range(cfg.layers_per_block + 1):
            x = _res(x, sd, f"decoder.up_blocks.{i}.resnets.{j}", g)
        if i < n_up - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = F.conv2d(x, sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"],
                         sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"], padding=1)
    x = _gn(x, sd, "decoder.conv_norm_out", g, True)
    return F.conv2d(x, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)


def postprocess(image: torch.Tensor) -> torch.Tensor:
    """VaeImageProcessor.postprocess(output_type='pt') [3P]: denormalise to [0, 1]."""
    return (image / 2 + 0.5).clamp(0, 1)


def decoder_flops(cfg: VaeConfig, h: int, w: int) -> float:
    """2*MAC of every conv / linear + 4*N*N*C of the mid attention for one image at latent h x w."""
    S = decoder_param_shapes(cfg)
    total = 0.0
    lvl_of = {}
    rev = len(cfg.block_out_channels)
    for name, shape in S.items():
        if not name.endswith(".weight") or len(shape) < 2:
            continue
        if name.startswith("decoder.up_blocks."):
            i = int(name.split(".")[2])
            scale = 2 ** i * (2 if ".upsamplers." in name else 1)
        elif name.startswith(("decoder.conv_norm_out", "decoder.conv_out")):
            scale = 2 ** (rev - 1)
        else:
            scale = 1
        pix = (h * scale) * (w * scale)
        k = shape[2] * shape[3] if len(shape) == 4 else 1
        total += 2.0 * pix * shape[0] * shape[1] * k
    n = h * w
    total += 4.0 * n * n * cfg.block_out_channels[-1]
    return total
