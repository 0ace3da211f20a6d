"""fp32 restatement of the SDXL UNet / ControlNet forward pass (diffusers==0.25.0 semantics [third-party]).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Functional style over a flat state dict that uses the diffusers
parameter names, NCHW activations, one un-fused torch op per reference op.  Reference call sites that fix the
interface: src/pipelines/lora_pipeline.py:520-529 (controlnet), :546-566 (main unet), :592-599 (concept unet).
"""
import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class UNetConfig:
    """SDXL-base-1.0 unet/config.json restated; `tiny()` keeps the topology at toy widths for CPU tests."""
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280)
    layers_per_block: int = 2
    transformer_layers: Tuple[int, ...] = (0, 2, 10)   # 0 = DownBlock2D / UpBlock2D (no attention)
    head_dim: int = 64
    cross_attention_dim: int = 2048
    addition_time_embed_dim: int = 256
    pooled_dim: int = 1280                              # text_embeds width; 1280 + 6*256 = 2816
    norm_groups: int = 32
    cond_embed_channels: Tuple[int, ...] = (16, 32, 96, 256)  # ControlNet conditioning embedding

    @property
    def time_embed_dim(self):
        return self.block_out_channels[0] * 4

    @property
    def add_in_dim(self):
        return self.pooled_dim + 6 * self.addition_time_embed_dim

    @staticmethod
    def sdxl():
        return UNetConfig()

    @staticmethod
    def tiny():
        return UNetConfig(block_out_channels=(64, 128, 256), transformer_layers=(0, 1, 2), cross_attention_dim=256,
                          addition_time_embed_dim=64, pooled_dim=128, cond_embed_channels=(16, 32, 64, 128))


# ------------------------------------------------------------------------------------------------ context
@dataclass
class Ctx:
    """Per-call state: weights, LoRA adapters, attention hook."""
    sd: Dict[str, torch.Tensor]
    cfg: UNetConfig
    # name -> list of (A [r,in], B [out,r], scale); scale = adapter_weight * alpha/r * cross_attention_kwargs.scale
    lora: Dict[str, List[Tuple[torch.Tensor, torch.Tensor, float]]] = field(default_factory=dict)
    # attention core: (name, q, k, v, is_cross, heads) -> (B, N, C)
    attn_core: Optional[Callable] = None
    # decoupled IP-adapter: dict name -> (to_k_ip [c, 2048], to_v_ip), number of image tokens, scale
    ip_weights: Optional[Dict[str, Tuple[torch.Tensor, torch.Tensor]]] = None
    ip_tokens: int = 16
    ip_scale: float = 1.0
    prefix: str = ""

    def w(self, name):
        return self.sd[self.prefix + name]

    def has(self, name):
        return (self.prefix + name) in self.sd


def linear(c: Ctx, name: str, x: torch.Tensor) -> torch.Tensor:
    y = F.linear(x, c.w(name + ".weight"), c.w(name + ".bias") if c.has(name + ".bias") else None)
    for A, Bm, s in c.lora.get(name, ()):  # peft un-merged LoRA: y += s * B(A x)
        y = y + s * F.linear(F.linear(x, A), Bm)
    return y


def conv(c: Ctx, name: str, x, stride=1, padding=1):
    return F.conv2d(x, c.w(name + ".weight"), c.w(name + ".bias") if c.has(name + ".bias") else None, stride=stride,
                    padding=padding)


def group_norm(c: Ctx, name: str, x, eps):
    return F.group_norm(x, c.cfg.norm_groups, c.w(name + ".weight"), c.w(name + ".bias"), eps)


def layer_norm(c: Ctx, name: str, x):
    return F.layer_norm(x, (x.shape[-1],), c.w(name + ".weight"), c.w(name + ".bias"), 1e-5)


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0) -> [cos | sin]."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t.float()[:, None] * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


# ------------------------------------------------------------------------------------------------ attention
def head_to_batch(x, heads):
    B, N, Cc = x.shape
    return x.reshape(B, N, heads, Cc // heads).permute(0, 2, 1, 3).reshape(B * heads, N, Cc // heads)


def batch_to_head(x, heads):
    Bh, N, d = x.shape
    return x.reshape(Bh // heads, heads, N, d).permute(0, 2, 1, 3).reshape(Bh // heads, N, heads * d)


def attention_scores(q, k, scale):
    """diffusers Attention.get_attention_scores: baddbmm(alpha=scale) then softmax over keys."""
    return torch.softmax(torch.bmm(q, k.transpose(1, 2)) * scale, dim=-1)


def plain_attn_core(name, q, k, v, is_cross, heads):
    d = q.shape[-1] // heads
    p = attention_scores(head_to_batch(q, heads), head_to_batch(k, heads), d ** -0.5)
    return batch_to_head(torch.bmm(p, head_to_batch(v, heads)), heads)


def make_p2p_attn_core(controller, place="mid"):
    """RegionControlNet_AttnProcessor (src/pipelines/lora_pipeline.py:67-133): materialise the probabilities,
    hand them to the controller (in place on the conditional half), then bmm."""
    def core(name, q, k, v, is_cross, heads):
        d = q.shape[-1] // heads
        p = attention_scores(head_to_batch(q, heads), head_to_batch(k, heads), d ** -0.5)
        p = controller(p, is_cross, place)
        return batch_to_head(torch.bmm(p, head_to_batch(v, heads)), heads)
    return core


def attention(c: Ctx, name: str, x, ctx=None):
    heads = x.shape[-1] // c.cfg.head_dim
    core = c.attn_core or plain_attn_core
    is_cross = ctx is not None
    q = linear(c, name + ".to_q", x)
    if is_cross and c.ip_weights is not None and name in c.ip_weights:
        # IPAttnProcessor (src/ip_adapter/attention_processor.py:113-195): split text / image tokens
        end = ctx.shape[1] - c.ip_tokens
        txt, ip = ctx[:, :end], ctx[:, end:]
        h = core(name, q, linear(c, name + ".to_k", txt), linear(c, name + ".to_v", txt), True, heads)
        wk, wv = c.ip_weights[name]
        h_ip = plain_attn_core(name, q, F.linear(ip, wk), F.linear(ip, wv), True, heads)
        h = h + c.ip_scale * h_ip
    else:
        src = ctx if is_cross else x
        h = core(name, q, linear(c, name + ".to_k", src), linear(c, name + ".to_v", src), is_cross, heads)
    return linear(c, name + ".to_out.0", h)


def transformer_block(c: Ctx, name: str, x, ctx):
    x = attention(c, name + ".attn1", layer_norm(c, name + ".norm1", x)) + x
    x = attention(c, name + ".attn2", layer_norm(c, name + ".norm2", x), ctx) + x
    h = linear(c, name + ".ff.net.0.proj", layer_norm(c, name + ".norm3", x))
    val, gate = h.chunk(2, dim=-1)
    x = linear(c, name + ".ff.net.2", val * F.gelu(gate)) + x
    return x


def transformer2d(c: Ctx, name: str, x, ctx, n_layers):
    B, Cc, H, W = x.shape
    res = x
    h = group_norm(c, name + ".norm", x, 1e-6)
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, Cc)
    h = linear(c, name + ".proj_in", h)
    for i in range(n_layers):
        h = transformer_block(c, f"{name}.transformer_blocks.{i}", h, ctx)
    h = linear(c, name + ".proj_out", h)
    return h.reshape(B, H, W, Cc).permute(0, 3, 1, 2) + res


def resnet(c: Ctx, name: str, x, emb):
    h = conv(c, name + ".conv1", F.silu(group_norm(c, name + ".norm1", x, 1e-5)))
    h = h + linear(c, name + ".time_emb_proj", F.silu(emb))[:, :, None, None]
    h = conv(c, name + ".conv2", F.silu(group_norm(c, name + ".norm2", h, 1e-5)))
    if c.has(name + ".conv_shortcut.weight"):
        x = conv(c, name + ".conv_shortcut", x, padding=0)
    return x + h


def embeddings(c: Ctx, timestep, text_embeds, time_ids, batch):
    cfg = c.cfg
    t = torch.as_tensor(timestep, dtype=torch.float32).reshape(-1)
    t = t.expand(batch) if t.numel() == 1 else t
    wdt = c.w("time_embedding.linear_1.weight").dtype  # diffusers: t_emb.to(dtype=sample.dtype); fp32 here unless the
    dev = c.w("time_embedding.linear_1.weight").device  # oracle is run as the fp16 library path (tests only)
    t_emb = timestep_embedding(t.to(dev), cfg.block_out_channels[0]).to(wdt)
    emb = linear(c, "time_embedding.linear_2", F.silu(linear(c, "time_embedding.linear_1", t_emb)))
    tid = timestep_embedding(time_ids.reshape(-1), cfg.addition_time_embed_dim).reshape(batch, -1).to(wdt)
    add = torch.cat([text_embeds, tid], dim=-1)
    aug = linear(c, "add_embedding.linear_2", F.silu(linear(c, "add_embedding.linear_1", add)))
    return emb + aug


def encoder(c: Ctx, h, emb, ctx):
    """conv_in output -> down blocks -> mid block.  Returns (mid output, 9 skip tensors)."""
    cfg = c.cfg
    skips = [h]
    nb = len(cfg.block_out_channels)
    for i in range(nb):
        for j in range(cfg.layers_per_block):
            h = resnet(c, f"down_blocks.{i}.resnets.{j}", h, emb)
            if cfg.transformer_layers[i] > 0:
                h = transformer2d(c, f"down_blocks.{i}.attentions.{j}", h, ctx, cfg.transformer_layers[i])
            skips.append(h)
        if i < nb - 1:
            h = conv(c, f"down_blocks.{i}.downsamplers.0.conv", h, stride=2, padding=1)
            skips.append(h)
    h = resnet(c, "mid_block.resnets.0", h, emb)
    h = transformer2d(c, "mid_block.attentions.0", h, ctx, cfg.transformer_layers[-1])
    h = resnet(c, "mid_block.resnets.1", h, emb)
    return h, skips


def unet_forward(c: Ctx, sample, timestep, ctx, text_embeds, time_ids, down_residuals=None, mid_residual=None):
    """UNet2DConditionModel.forward(sample, t, encoder_hidden_states, added_cond_kwargs, down_block_additional_
    residuals, mid_block_additional_residual)[0]."""
    cfg = c.cfg
    B = sample.shape[0]
    emb = embeddings(c, timestep, text_embeds, time_ids, B)
    h = conv(c, "conv_in", sample)
    h, skips = encoder(c, h, emb, ctx)
    if down_residuals is not None:
        skips = [s + r for s, r in zip(skips, down_residuals)]
        h = h + mid_residual
    nb = len(cfg.block_out_channels)
    for i in range(nb):
        layers = cfg.transformer_layers[nb - 1 - i]
        for j in range(cfg.layers_per_block + 1):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet(c, f"up_blocks.{i}.resnets.{j}", h, emb)
            if layers > 0:
                h = transformer2d(c, f"up_blocks.{i}.attentions.{j}", h, ctx, layers)
        if i < nb - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = conv(c, f"up_blocks.{i}.upsamplers.0.conv", h)
    h = F.silu(group_norm(c, "conv_norm_out", h, 1e-5))
    return conv(c, "conv_out", h)


def controlnet_forward(c: Ctx, sample, timestep, ctx, cond, conditioning_scale, text_embeds, time_ids):
    """ControlNetModel.forward(...) -> (9 down residuals, mid residual), each * conditioning_scale."""
    B = sample.shape[0]
    emb = embeddings(c, timestep, text_embeds, time_ids, B)
    h = conv(c, "conv_in", sample)
    e = F.silu(conv(c, "controlnet_cond_embedding.conv_in", cond))
    n = 2 * (len(c.cfg.cond_embed_channels) - 1)
    for i in range(n):
        e = F.silu(conv(c, f"controlnet_cond_embedding.blocks.{i}", e, stride=2 if i % 2 == 1 else 1))
    h = h + conv(c, "controlnet_cond_embedding.conv_out", e)
    h, skips = encoder(c, h, emb, ctx)
    down = [conv(c, f"controlnet_down_blocks.{i}", s, padding=0) * conditioning_scale for i, s in enumerate(skips)]
    mid = conv(c, "controlnet_mid_block", h, padding=0) * conditioning_scale
    return down, mid


# ------------------------------------------------------------------------------------------------ shapes
def param_shapes(cfg: UNetConfig, controlnet: bool = False) -> Dict[str, Tuple[int, ...]]:
    """Every parameter name -> shape of the restated model (diffusers naming)."""
    S: Dict[str, Tuple[int, ...]] = {}

    def lin(n, i, o, bias=True):
        S[n + ".weight"] = (o, i)
        if bias:
            S[n + ".bias"] = (o,)

    def cv(n, i, o, k=3):
        S[n + ".weight"] = (o, i, k, k)
        S[n + ".bias"] = (o,)

    def norm(n, ch):
        S[n + ".weight"] = (ch,)
        S[n + ".bias"] = (ch,)

    def res(n, i, o):
        norm(n + ".norm1", i)
        cv(n + ".conv1", i, o)
        lin(n + ".time_emb_proj", cfg.time_embed_dim, o)
        norm(n + ".norm2", o)
        cv(n + ".conv2", o, o)
        if i != o:
            cv(n + ".conv_shortcut", i, o, 1)

    def tr(n, ch, layers):
        norm(n + ".norm", ch)
        lin(n + ".proj_in", ch, ch)
        for k in range(layers):
            b = f"{n}.transformer_blocks.{k}"
            for a, kd in (("attn1", ch), ("attn2", cfg.cross_attention_dim)):
                lin(f"{b}.{a}.to_q", ch, ch, False)
                lin(f"{b}.{a}.to_k", kd, ch, False)
                lin(f"{b}.{a}.to_v", kd, ch, False)
                lin(f"{b}.{a}.to_out.0", ch, ch)
            for m in ("norm1", "norm2", "norm3"):
                norm(f"{b}.{m}", ch)
            lin(f"{b}.ff.net.0.proj", ch, 8 * ch)
            lin(f"{b}.ff.net.2", 4 * ch, ch)
        lin(n + ".proj_out", ch, ch)

    boc = cfg.block_out_channels
    nb = len(boc)
    cv("conv_in", cfg.in_channels, boc[0])
    lin("time_embedding.linear_1", boc[0], cfg.time_embed_dim)
    lin("time_embedding.linear_2", cfg.time_embed_dim, cfg.time_embed_dim)
    lin("add_embedding.linear_1", cfg.add_in_dim, cfg.time_embed_dim)
    lin("add_embedding.linear_2", cfg.time_embed_dim, cfg.time_embed_dim)
    skip_ch = [boc[0]]
    ch = boc[0]
    for i in range(nb):
        for j in range(cfg.layers_per_block):
            res(f"down_blocks.{i}.resnets.{j}", ch, boc[i])
            ch = boc[i]
            if cfg.transformer_layers[i] > 0:
                tr(f"down_blocks.{i}.attentions.{j}", ch, cfg.transformer_layers[i])
            skip_ch.append(ch)
        if i < nb - 1:
            cv(f"down_blocks.{i}.downsamplers.0.conv", ch, ch)
            skip_ch.append(ch)
    res("mid_block.resnets.0", ch, ch)
    tr("mid_block.attentions.0", ch, cfg.transformer_layers[-1])
    res("mid_block.resnets.1", ch, ch)
    if controlnet:
        cec = cfg.cond_embed_channels
        cv("controlnet_cond_embedding.conv_in", 3, cec[0])
        for i in range(len(cec) - 1):
            cv(f"controlnet_cond_embedding.blocks.{2 * i}", cec[i], cec[i])
            cv(f"controlnet_cond_embedding.blocks.{2 * i + 1}", cec[i], cec[i + 1])
        cv("controlnet_cond_embedding.conv_out", cec[-1], boc[0])
        for i, sc in enumerate(skip_ch):
            cv(f"controlnet_down_blocks.{i}", sc, sc, 1)
        cv("controlnet_mid_block", ch, ch, 1)
        return S
    skips = list(skip_ch)
    for i in range(nb):
        out = boc[nb - 1 - i]
        for j in range(cfg.layers_per_block + 1):
            res(f"up_blocks.{i}.resnets.{j}", ch + skips.pop(), out)
            ch = out
            if cfg.transformer_layers[nb - 1 - i] > 0:
                tr(f"up_blocks.{i}.attentions.{j}", ch, cfg.transformer_layers[nb - 1 - i])
        if i < nb - 1:
            cv(f"up_blocks.{i}.upsamplers.0.conv", ch, ch)
    norm("conv_norm_out", ch)
    cv("conv_out", ch, cfg.out_channels)
    return S


def attention_names(cfg: UNetConfig, controlnet: bool = False) -> List[str]:
    """attn1/attn2 module paths in forward order (the order the P2P controller counts them in)."""
    names = []

    def tr(n, layers):
        for k in range(layers):
            names.append(f"{n}.transformer_blocks.{k}.attn1")
            names.append(f"{n}.transformer_blocks.{k}.attn2")

    nb = len(cfg.block_out_channels)
    for i in range(nb):
        for j in range(cfg.layers_per_block):
            if cfg.transformer_layers[i] > 0:
                tr(f"down_blocks.{i}.attentions.{j}", cfg.transformer_layers[i])
    tr("mid_block.attentions.0", cfg.transformer_layers[-1])
    if not controlnet:
        for i in range(nb):
            for j in range(cfg.layers_per_block + 1):
                if cfg.transformer_layers[nb - 1 - i] > 0:
                    tr(f"up_blocks.{i}.attentions.{j}", cfg.transformer_layers[nb - 1 - i])
    return names


def unet_flops(cfg: UNetConfig, H: int, W: int, ctx_len: int = 77) -> float:
    """Analytic FLOPs of one sample-forward (2*MAC of every conv/linear + 4*N*L*c per attention)."""
    S = param_shapes(cfg)
    boc = cfg.block_out_channels
    nb = len(boc)
    total = 0.0

    def res_at(name):
        # spatial size of a module from its block index
        p = name.split(".")
        if p[0] == "down_blocks":
            lvl = int(p[1])
            if p[2] == "downsamplers":
                lvl += 1
        elif p[0] == "mid_block":
            lvl = nb - 1
        elif p[0] == "up_blocks":
            lvl = nb - 1 - int(p[1])
            if p[2] == "upsamplers":
                lvl -= 1
        else:
            lvl = 0
        return (H >> lvl) * (W >> lvl)

    for n, shp in S.items():
        if not n.endswith(".weight") or len(shp) == 1:
            continue
        hw = res_at(n)
        macs = 1
        for s in shp:
            macs *= s
        if "time_emb" in n or "time_embedding" in n or "add_embedding" in n:
            total += 2 * macs
        elif ".attn2.to_k" in n or ".attn2.to_v" in n:
            total += 2 * macs * ctx_len
        else:
            total += 2 * macs * hw
    for n in attention_names(cfg):
        hw = res_at(n)
        ch = S[n + ".to_q.weight"][0]
        L = hw if n.endswith("attn1") else ctx_len
        total += 4.0 * hw * L * ch
    return total
