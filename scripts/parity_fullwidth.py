"""Full-width parity of BASELINE configs 2, 3, 4 over ALL steps of both stages.

The checker is the fp32 oracle run ON THE GPU (pure torch, TF32 off): seconds per stage instead of an hour of host
cores.  For every config and stage this records, per step and for the final latents, the relative L2 distance to the
fp32 oracle of
  * ours        - the CUDA path through the public pipeline call (fp16 storage, fp32 accumulation), and
  * fp16_eager  - the SAME restatement run as fp16 torch eager (fp16 weights / activations / materialised
                  probabilities, un-merged LoRA): the arithmetic of the reference's diffusers + peft library path,
all three fed the same fp16-rounded weights, embeddings and initial noise.  One JSON line per (config, stage).

  python scripts/parity_fullwidth.py [2] [3] [4] [--steps 30] [--out profiles/r02_parity_fullwidth.jsonl]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omg_b200 import factory, synthetic  # noqa: E402
from omg_b200.config import UNetConfig  # noqa: E402
from omg_b200.pipelines import (ConceptModels, InstantidMultiConceptPipeline, LoraMultiConceptPipeline,  # noqa: E402
                                revise_regionally_controlnet_forward)
from omg_b200.prompt_attention import AttentionReplace  # noqa: E402
from omg_b200.unet import PackedUNet  # noqa: E402
from oracle import p2p as op2p  # noqa: E402
from oracle import unet as ou  # noqa: E402
from oracle.pipeline import Concept, denoise  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = "cuda"
SIZE = 1024
OCFG = ou.UNetConfig()


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


def cast_tree(x, dt):
    if torch.is_tensor(x):
        return x.to(dev, dt) if x.is_floating_point() else x.to(dev)
    if isinstance(x, dict):
        return {k: cast_tree(v, dt) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(cast_tree(v, dt) if not isinstance(v, float) else v for v in x)
    return x


def oracle_controller(prompts, dt):
    c = op2p.AttentionReplaceOracle(prompts, 50, {"default_": 1.0}, 0.4, SIZE // 32, SIZE // 32)
    c.num_att_layers = 140
    c.mapper = c.mapper.to(dev)
    c.cross_replace_alpha = c.cross_replace_alpha.to(dev, dt)
    return c


def run_oracle(dt, prompts, sd, lat0, ctx4, pooled4, tid, concept_specs, stage, steps, g, **kw):
    """concept_specs: list of dicts(ctx (2,77,D), pooled, mask, lora (oracle format) | None, ip (weights, scale) | None,
    tokens | None).  Returns (per-step latents list, final latents), fp32 on the device."""
    trace = []
    sdd = cast_tree(sd, dt)
    main = ou.Ctx(sdd, OCFG, attn_core=ou.make_p2p_attn_core(oracle_controller(prompts, dt)),
                  lora=cast_tree(kw.pop("main_lora", {}), dt))
    concepts = []
    for c in concept_specs:
        ck = dict(lora=cast_tree(c.get("lora") or {}, dt))
        if c.get("ip") is not None:
            ck.update(ip_weights=cast_tree(c["ip"][0], dt), ip_tokens=16, ip_scale=c["ip"][1])
        concepts.append(Concept(c["ctx"].to(dev, dt), c["pooled"].to(dev, dt), tid.repeat(2, 1).to(dev), c["mask"],
                                unet=ou.Ctx(sdd, OCFG, **ck),
                                image_tokens=None if c.get("tokens") is None else c["tokens"].to(dev, dt)))
    extra = {}
    for name in ("controlnet", "identitynet"):
        if kw.get(name) is not None:
            extra[name] = ou.Ctx(cast_tree(kw[name], dt), OCFG)
    for name in ("controlnet_cond", "identity_cond"):
        if kw.get(name) is not None:
            extra[name] = kw[name].to(dev, dt)
    for name in ("controlnet_scale", "identity_scale"):
        if name in kw:
            extra[name] = kw[name]
    with torch.no_grad():
        out = denoise(main, lat0.to(dev, dt), ctx4.to(dev, dt), pooled4.to(dev, dt), tid.repeat(4, 1).to(dev), concepts,
                      stage, steps, g, trace=lambda i, n, l: trace.append(l.float().clone()), **extra)
    return trace, out.float()


def compare(tag, cfg_id, stage, ours_trace, ours_final, o32, oh, t_oracle):
    tr32, f32 = o32
    trh, fh = oh
    rec = {"config": cfg_id, "stage": stage, "what": tag, "steps": len(tr32),
           "ours_final_rel_l2": [rel(ours_final[k], f32[k]) for k in range(2)],
           "fp16_eager_final_rel_l2": [rel(fh[k], f32[k]) for k in range(2)],
           "ours_vs_fp16_eager_final": [rel(ours_final[k], fh[k]) for k in range(2)],
           "ours_per_step": [round(rel(a, b), 6) for a, b in zip(ours_trace, tr32)],
           "fp16_eager_per_step": [round(rel(a, b), 6) for a, b in zip(trh, tr32)],
           "oracle_gpu_s": round(t_oracle, 1), "finite": bool(torch.isfinite(ours_final).all())}
    rec["ours_max_per_step"] = max(rec["ours_per_step"])
    rec["fp16_eager_max_per_step"] = max(rec["fp16_eager_per_step"])
    return rec


def tracing():
    tr = []
    return tr, (lambda pipe, i, t, kw: tr.append(kw["latents"].float().clone()) or {})


def config2(steps, emit, stages=(1, 2)):
    cfg = UNetConfig.sdxl()
    wl = factory.build_lora_workload(cfg, SIZE, 2, 32, steps, 7.5, device=dev, keep_state_dict=True)
    pipe, cm, kw = wl.pipe, wl.concept_models, dict(wl.call_kwargs)
    prompts, regions = kw["prompt"]
    lat0 = torch.randn(1, 4, SIZE // 8, SIZE // 8, generator=torch.Generator().manual_seed(14)).half()
    pe, ne, pp, np_ = pipe.encode_prompt(prompts, kw["negative_prompt"], 0.8)
    ctx4, pooled4 = torch.cat([ne, pe]).half(), torch.cat([np_, pp]).half()
    tid = torch.tensor([[SIZE, SIZE, 0, 0, SIZE, SIZE]], dtype=torch.float32)
    specs = []
    for k, (rp, rn) in enumerate(regions):
        e, n_, p_, np2 = cm.encode_prompt(rp, negative_prompt=rn)
        lo = {name: [(a, b, s * 0.8)] for name, (a, b, s) in cm._loras[kw["lora_list"][k]].items()}
        specs.append(dict(ctx=torch.cat([n_, e]).half(), pooled=torch.cat([np2, p_]).half(), mask=wl.masks[k], lora=lo))
    for stage in stages:
        tr, cb = tracing()
        extra = dict(region_masks=wl.masks) if stage == 2 else {}
        out = pipe(stage=stage, latents=lat0, callback_on_step_end=cb, **extra, **kw).images.float()
        wl.controller.reset()
        t0 = time.perf_counter()
        o32 = run_oracle(torch.float32, prompts, wl.state_dict, lat0, ctx4, pooled4, tid, specs, stage, steps, 7.5)
        torch.cuda.synchronize()
        t_or = time.perf_counter() - t0
        oh = run_oracle(torch.float16, prompts, wl.state_dict, lat0, ctx4, pooled4, tid, specs, stage, steps, 7.5)
        emit(compare("BASELINE config 2 (2 LoRA concepts, P2P), 128x128 latents", 2, stage, tr, out, o32, oh, t_or))


def resampler_sd(dim=1280, depth=4):
    g = torch.Generator().manual_seed(7)

    def rn(*s):
        return torch.randn(*s, generator=g) * (s[-1] ** -0.5)

    rs = {"latents": rn(1, 16, dim), "proj_in.weight": rn(dim, 512), "proj_in.bias": torch.zeros(dim),
          "proj_out.weight": rn(2048, dim), "proj_out.bias": torch.zeros(2048), "norm_out.weight": torch.ones(2048),
          "norm_out.bias": torch.zeros(2048)}
    for li in range(depth):
        a, f = f"layers.{li}.0", f"layers.{li}.1"
        for nm in ("norm1", "norm2"):
            rs[f"{a}.{nm}.weight"], rs[f"{a}.{nm}.bias"] = torch.ones(dim), torch.zeros(dim)
        rs[f"{a}.to_q.weight"], rs[f"{a}.to_kv.weight"], rs[f"{a}.to_out.weight"] = rn(dim, dim), rn(2 * dim, dim), rn(dim, dim)
        rs[f"{f}.0.weight"], rs[f"{f}.0.bias"] = torch.ones(dim), torch.zeros(dim)
        rs[f"{f}.1.weight"], rs[f"{f}.3.weight"] = rn(4 * dim, dim), rn(dim, 4 * dim)
    return rs


def config3(steps, emit):
    """OMG + InstantID, 2 identities: IP-adapter concept streams + IdentityNet, guidance 3.0, seed 53."""
    cfg = UNetConfig.sdxl()
    sd = synthetic.make_state_dict(cfg, seed=0, device=dev, dtype=torch.float16)
    idsd = synthetic.make_state_dict(cfg, seed=1, controlnet=True, device=dev, dtype=torch.float16)
    unet = PackedUNet(cfg, sd, device=dev)
    pipe = InstantidMultiConceptPipeline(unet, controlnet=PackedUNet(cfg, idsd, device=dev, controlnet=True))
    prompt = "two people on a beach"
    prompts = [prompt] * 2
    ctrl = AttentionReplace(prompts, 50, {"default_": 1.0}, 0.4, SIZE // 32, SIZE // 32)
    revise_regionally_controlnet_forward(pipe, ctrl)
    cm = ConceptModels(unet)
    rs = resampler_sd()
    ipw = synthetic.make_ip_adapter(cfg, 31, device=dev, dtype=torch.float16)
    cm.load_ip_adapter_instantid({k: v.to(dev) for k, v in rs.items()}, ipw, heads=20, dim_head=64, num_tokens=16)
    cm.set_ip_adapter_scale(0.8)
    faces = [torch.nn.functional.normalize(torch.randn(512, generator=torch.Generator().manual_seed(s)), dim=0) for s in (1, 2)]
    cond = torch.rand(3, SIZE, SIZE, generator=torch.Generator().manual_seed(7)).half().float()
    masks = synthetic.rect_masks(2, (SIZE, SIZE))
    regions = [("a man", "bad", None), ("a woman", "bad", None)]
    lat0 = torch.randn(1, 4, SIZE // 8, SIZE // 8, generator=torch.Generator().manual_seed(53)).half()
    common = dict(prompt=[prompts, regions], negative_prompt=["noisy"] * 2, guidance_scale=3.0,
                  num_inference_steps=steps, concept_models=cm, image=cond, controlnet_conditioning_scale=0.8,
                  face_embeds=faces, height=SIZE, width=SIZE, output_type="latent", latents=lat0)
    pe, ne, pp, np_ = pipe.encode_prompt(prompts, ["noisy"] * 2)
    ctx4, pooled4 = torch.cat([ne, pe]).half(), torch.cat([np_, pp]).half()
    tid = torch.tensor([[SIZE, SIZE, 0, 0, SIZE, SIZE]], dtype=torch.float32)
    specs = []
    for k, reg in enumerate(regions):
        e, n_, p_, np2 = pipe.encode_prompt(reg[0], reg[1])
        tokens = cm._encode_prompt_image_emb(faces[k], dev, torch.float16, True).half()  # the product's own Resampler
        specs.append(dict(ctx=torch.cat([n_, e]).half(), pooled=torch.cat([np2, p_]).half(), mask=masks[k],
                          ip=(ipw, 0.8), tokens=tokens))
    for stage in (1, 2):
        tr, cb = tracing()
        extra = dict(region_masks=masks) if stage == 2 else {}
        out = pipe(stage=stage, callback_on_step_end=cb, **extra, **common).images.float()
        ctrl.reset()
        okw = dict(identitynet=idsd, identity_cond=cond[None].repeat(2, 1, 1, 1), identity_scale=0.8)
        t0 = time.perf_counter()
        o32 = run_oracle(torch.float32, prompts, sd, lat0, ctx4, pooled4, tid, specs, stage, steps, 3.0, **okw)
        torch.cuda.synchronize()
        t_or = time.perf_counter() - t0
        oh = run_oracle(torch.float16, prompts, sd, lat0, ctx4, pooled4, tid, specs, stage, steps, 3.0, **okw)
        emit(compare("BASELINE config 3 (InstantID: 2 IP-adapter identities + IdentityNet)", 3, stage, tr, out, o32, oh, t_or))


def config4(steps, emit):
    """4 LoRA concepts (each + style, [0.7, 0.5]) + style LoRA on the main pass + spatial ControlNet (B=4)."""
    cfg = UNetConfig.sdxl()
    sd = synthetic.make_state_dict(cfg, seed=0, device=dev, dtype=torch.float16)
    cnsd = synthetic.make_state_dict(cfg, seed=2, controlnet=True, device=dev, dtype=torch.float16)
    unet = PackedUNet(cfg, sd, device=dev)
    pipe = LoraMultiConceptPipeline(unet, controlnet=PackedUNet(cfg, cnsd, device=dev, controlnet=True))
    prompt = "four friends in a park"
    prompts = [prompt] * 2
    ctrl = AttentionReplace(prompts, 50, {"default_": 1.0}, 0.4, SIZE // 32, SIZE // 32)
    revise_regionally_controlnet_forward(pipe, ctrl)
    style = synthetic.make_lora(cfg, seed=999, rank=32, device=dev, dtype=torch.float16)
    pipe.load_lora_weights(style, adapter_name="style")
    cm = ConceptModels(unet)
    cm.load_lora_weights(style, adapter_name="style")
    names, loras = [], []
    for k in range(4):
        loras.append(synthetic.make_lora(cfg, seed=1000 + k, rank=32, device=dev, dtype=torch.float16))
        cm.load_lora_weights(loras[-1], adapter_name=f"c{k}")
        names.append(f"c{k}")
    masks = []
    for qy in range(2):
        for qx in range(2):
            m = torch.zeros(SIZE, SIZE)
            m[qy * 512 + 32:(qy + 1) * 512 - 32, qx * 512 + 32:(qx + 1) * 512 - 32] = 1
            masks.append(m)
    cond = torch.rand(3, SIZE, SIZE, generator=torch.Generator().manual_seed(5)).half().float()
    regions = [(f"person {k}", "bad") for k in range(4)]
    lat0 = torch.randn(1, 4, SIZE // 8, SIZE // 8, generator=torch.Generator().manual_seed(14)).half()
    common = dict(prompt=[prompts, regions], negative_prompt=["noisy"] * 2, guidance_scale=7.5,
                  num_inference_steps=steps, cross_attention_kwargs={"scale": 0.8}, concept_models=cm, lora_list=names,
                  styleL=True, image=[cond, cond], height=SIZE, width=SIZE, output_type="latent", latents=lat0)
    pe, ne, pp, np_ = pipe.encode_prompt(prompts, ["noisy"] * 2, 0.8)
    ctx4, pooled4 = torch.cat([ne, pe]).half(), torch.cat([np_, pp]).half()
    tid = torch.tensor([[SIZE, SIZE, 0, 0, SIZE, SIZE]], dtype=torch.float32)

    def olora(adapters, gs):
        out = {}
        for lo, w in adapters:
            for name, (a, b, s) in lo.items():
                out.setdefault(name, []).append((a, b, s * w * gs))
        return out

    specs = []
    for k, (rp, rn) in enumerate(regions):
        e, n_, p_, np2 = cm.encode_prompt(rp, negative_prompt=rn)
        specs.append(dict(ctx=torch.cat([n_, e]).half(), pooled=torch.cat([np2, p_]).half(), mask=masks[k],
                          lora=olora([(loras[k], 0.7), (style, 0.5)], 0.8)))
    for stage in (1, 2):
        tr, cb = tracing()
        extra = dict(region_masks=masks) if stage == 2 else {}
        out = pipe(stage=stage, callback_on_step_end=cb, **extra, **common).images.float()
        ctrl.reset()
        okw = dict(controlnet=cnsd, controlnet_cond=cond[None].repeat(4, 1, 1, 1), controlnet_scale=1.0,
                   main_lora=olora([(style, 1.0)], 0.8))
        t0 = time.perf_counter()
        o32 = run_oracle(torch.float32, prompts, sd, lat0, ctx4, pooled4, tid, specs, stage, steps, 7.5, **okw)
        torch.cuda.synchronize()
        t_or = time.perf_counter() - t0
        oh = run_oracle(torch.float16, prompts, sd, lat0, ctx4, pooled4, tid, specs, stage, steps, 7.5, **okw)
        emit(compare("BASELINE config 4 (4 LoRA concepts + style + spatial ControlNet)", 4, stage, tr, out, o32, oh, t_or))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["2"])
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--out", default=None)
    ap.add_argument("--trunk-f32", action="store_true", help="run the CUDA path with the fp32 master copy of the residual trunk")
    a = ap.parse_args()
    if a.trunk_f32:
        os.environ["OMG_TRUNK_F32"] = "1"
    f = open(a.out, "a") if a.out else None

    def emit(rec):
        rec["trunk_f32"] = os.environ.get("OMG_TRUNK_F32", "0") == "1"
        line = json.dumps(rec)
        print(line, flush=True)
        if f:
            f.write(line + "\n")
            f.flush()

    for c in a.configs:
        {"2": config2, "3": config3, "4": config4}[c](a.steps, emit)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
