"""Shared builders for parity tests: the same seeded synthetic weights feed the CUDA path (fp16) and the fp32
oracle (which receives the fp16-ROUNDED values, so weight quantisation is not counted as kernel error)."""
import torch

from omg_b200 import synthetic
from omg_b200.config import UNetConfig
from oracle import unet as ou


def r16(t):
    return t.half().float()


def ocfg(cfg: UNetConfig) -> ou.UNetConfig:
    return ou.UNetConfig(in_channels=cfg.in_channels, out_channels=cfg.out_channels,
                         block_out_channels=cfg.block_out_channels, layers_per_block=cfg.layers_per_block,
                         transformer_layers=cfg.transformer_layers, head_dim=cfg.head_dim,
                         cross_attention_dim=cfg.cross_attention_dim,
                         addition_time_embed_dim=cfg.addition_time_embed_dim, pooled_dim=cfg.pooled_dim,
                         norm_groups=cfg.norm_groups, cond_embed_channels=cfg.cond_embed_channels)


def weights(cfg, seed, controlnet=False, bias_std=0.05, jitter=0.1):
    sd = synthetic.make_state_dict(cfg, seed=seed, controlnet=controlnet, bias_std=bias_std, affine_jitter=jitter)
    return {k: r16(v) for k, v in sd.items()}


def lora(cfg, seed, rank=8):
    lo = synthetic.make_lora(cfg, seed, rank=rank)
    return {k: (r16(a), r16(b), s) for k, (a, b, s) in lo.items()}


def oracle_lora(adapters, global_scale):
    """[(lora, adapter_weight)] -> oracle format name -> [(A, B, scale)]."""
    out = {}
    for lo, w in adapters:
        for k, (a, b, s) in lo.items():
            out.setdefault(k, []).append((a, b, s * w * global_scale))
    return out


def to_nhwc8(x):
    """(B,4,H,W) fp32 -> (B,H,W,8) fp16 cuda with zero padding channels."""
    B, C, H, W = x.shape
    o = torch.zeros(B, H, W, 8, dtype=torch.float16, device="cuda")
    o[..., :C] = x.permute(0, 2, 3, 1).half().cuda()
    return o


def from_nhwc(y, c=4):
    return y[..., :c].permute(0, 3, 1, 2).float().cpu()


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()
