"""Generate golden vectors from the UNMODIFIED reference modules that are importable in the build container
(/root/reference is not present on the GPU box, hence committed fixtures).  Run:  python tests/golden/make_golden.py

Fixtures (all fp32, fixed seeds):
  p2p_same.pt      AttentionReplace(prompts identical), shipped config of inference_lora.py:156 -> alpha table, mapper,
                   probabilities in/out for self & cross layers at steps inside/outside the self-replace window.
  p2p_edit.pt      AttentionReplace with a one-word edit and a partial cross_replace window (toy tokenizer).
  ip_attn.pt       IPAttnProcessor / IPAttnProcessor2_0 / AttnProcessor under a shim Attention module.
  resampler.pt     Resampler(dim=128, depth=2, heads=4, 16 queries, 512 -> 256).
  cli.pt           prepare_text of both CLIs (inference_lora.py:128-149, inference_instantid.py:233-254) and
                   LoraMultiConceptPipeline.get_region_mask (src/pipelines/lora_pipeline.py:673-681), each extracted from
                   its file by ast (the modules import diffusers) and run on fixed inputs.
  region_attn.pt   RegionControlNet_AttnProcessor (src/pipelines/lora_pipeline.py:61-133; the class is cut out of the
                   module by ast, the module imports diffusers) driven by the reference's own AttentionReplace on a shim
                   Attention: self / cross layers, inside / outside the replace windows, identical and edited prompts.
  fusion.pt        the region noise-fusion + classifier-free-guidance statements of LoraMultiConceptPipeline.__call__
                   (src/pipelines/lora_pipeline.py:568-612): the two `if` nodes are cut out of the method's AST and
                   executed unmodified against stub objects (concept UNet returning prepared noise, adapter-switch log).
  fusion_iid.pt    the same statements of InstantidMultiConceptPipeline.__call__ (src/pipelines/instantid_pipeline.py:
                   618-686): every call the block makes to the IdentityNet and to the concept UNet is recorded
                   (inputs, face / text tokens, condition image, scale, residual hand-over), plus the fused noise.
  sam_encoder.pt   EfficientViTSamImageEncoder (src/efficientvit/models/efficientvit/sam.py:176-192: EfficientViTLargeBackbone
                   with the xl1 block layout res/fmb/fmb/fmb/att@3/att@3 at narrow widths -> SamNeck -> LayerNorm2d), the
                   unmodified modules imported with stubs for the packages their unrelated imports need (timm, onnx,
                   segment_anything, torchvision: training apps and the mask decoder), random BatchNorm statistics,
                   norm eps 1e-6 like sam_model_zoo.py:44; plus one LiteMLA and one EfficientViTBlock on their own.
  kps.npz          draw_kps_multi (inference_instantid.py:127-156, extracted from the file by ast: the module itself
                   imports diffusers) on three faces at 256 x 256.
"""
import os
import sys

import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)


class ToyTokenizer:
    """Whitespace tokenizer with BOS/EOS, enough for seq_aligner/p2p_utils (encode/decode)."""

    def __init__(self):
        self.vocab = {}
        self.inv = {}

    def encode(self, text):
        ids = [0]
        for w in text.split(" "):
            if w not in self.vocab:
                self.vocab[w] = len(self.vocab) + 2
                self.inv[self.vocab[w]] = w
            ids.append(self.vocab[w])
        return ids + [1]

    def decode(self, ids):
        return " ".join(self.inv.get(i, "") for i in ids)


def run_controller(ctrl, layers, steps, batch_heads, seed):
    """Drive the controller like the 140-layer UNet would: list of (is_cross, N, L)."""
    g = torch.Generator().manual_seed(seed)
    rec = []
    ctrl.num_att_layers = len(layers)
    for s in range(steps):
        for (is_cross, n, l) in layers:
            p = torch.softmax(torch.randn(batch_heads, n, l, generator=g), dim=-1)
            inp = p.clone()
            out = ctrl(p, is_cross, "mid")
            rec.append({"step": s, "is_cross": is_cross, "in": inp, "out": out.clone(), "cur_step": ctrl.cur_step,
                        "cur_att_layer": ctrl.cur_att_layer})
    return rec


def make_p2p():
    from src.prompt_attention.p2p_attention import AttentionReplace
    tok = ToyTokenizer()
    prompt = "a photo of a man and a woman on the beach"
    layers = [(False, 16, 16), (True, 16, 77), (False, 32, 32), (True, 32, 77)]
    # shipped configuration: inference_lora.py:156 (num_steps=50, cross 1.0, self 0.4), threshold width*height
    ctrl = AttentionReplace([prompt] * 2, 50, cross_replace_steps={"default_": 1.}, self_replace_steps=0.4,
                            tokenizer=tok, device="cpu", dtype=torch.float32, width=4, height=4)
    rec = run_controller(ctrl, layers, 22, 4, seed=0)  # batch 4 (CFG x 2 images) x 1 head
    keep = [r for r in rec if r["step"] in (0, 19, 20, 21)]
    torch.save({"prompts": [prompt] * 2, "alpha": ctrl.cross_replace_alpha.clone(), "mapper": ctrl.mapper.clone(),
                "num_self_replace": ctrl.num_self_replace, "layers": layers, "records": keep, "width": 4, "height": 4,
                "heads": 1}, os.path.join(OUT, "p2p_same.pt"))
    p2 = "a photo of a dog and a woman on the beach"
    ctrl = AttentionReplace([prompt, p2], 10, cross_replace_steps={"default_": 0.6, "dog": (0.2, 0.9)},
                            self_replace_steps=0.3, tokenizer=tok, device="cpu", dtype=torch.float32, width=4, height=4)
    rec = run_controller(ctrl, layers, 10, 4, seed=1)
    torch.save({"prompts": [prompt, p2], "alpha": ctrl.cross_replace_alpha.clone(), "mapper": ctrl.mapper.clone(),
                "num_self_replace": ctrl.num_self_replace, "layers": layers, "records": rec, "width": 4, "height": 4,
                "heads": 1, "num_steps": 10, "cross": {"default_": 0.6, "dog": (0.2, 0.9)}, "self": 0.3},
               os.path.join(OUT, "p2p_edit.pt"))


class ShimAttention(torch.nn.Module):
    """Minimal stand-in for diffusers Attention with the attributes the reference processors read
    (src/pipelines/lora_pipeline.py:81-131)."""

    def __init__(self, dim, ctx_dim, heads):
        super().__init__()
        self.heads = heads
        self.scale = (dim // heads) ** -0.5
        self.to_q = torch.nn.Linear(dim, dim, bias=False)
        self.to_k = torch.nn.Linear(ctx_dim, dim, bias=False)
        self.to_v = torch.nn.Linear(ctx_dim, dim, bias=False)
        self.to_out = torch.nn.ModuleList([torch.nn.Linear(dim, dim), torch.nn.Dropout(0.0)])
        self.spatial_norm = None
        self.group_norm = None
        self.norm_cross = False
        self.residual_connection = False
        self.rescale_output_factor = 1.0

    def prepare_attention_mask(self, m, *a, **k):
        return m

    def head_to_batch_dim(self, t):
        b, n, c = t.shape
        return t.reshape(b, n, self.heads, c // self.heads).permute(0, 2, 1, 3).reshape(b * self.heads, n, -1)

    def batch_to_head_dim(self, t):
        bh, n, d = t.shape
        return t.reshape(bh // self.heads, self.heads, n, d).permute(0, 2, 1, 3).reshape(bh // self.heads, n, -1)

    def get_attention_scores(self, q, k, mask=None):
        s = torch.baddbmm(torch.empty(q.shape[0], q.shape[1], k.shape[1]), q, k.transpose(-1, -2), beta=0,
                          alpha=self.scale)
        return s.softmax(dim=-1)


def make_ip():
    from src.ip_adapter.attention_processor import AttnProcessor, IPAttnProcessor, IPAttnProcessor2_0
    torch.manual_seed(0)
    dim, ctx_dim, heads = 128, 96, 2
    attn = ShimAttention(dim, ctx_dim, heads)
    ip = IPAttnProcessor(dim, ctx_dim, scale=0.8, num_tokens=16)
    ip2 = IPAttnProcessor2_0(dim, ctx_dim, scale=0.8, num_tokens=16)
    ip2.load_state_dict(ip.state_dict())
    x = torch.randn(2, 64, dim)
    ctx = torch.randn(2, 77 + 16, ctx_dim)
    with torch.no_grad():
        y1 = ip(attn, x, ctx)
        y2 = ip2(attn, x, ctx)
        y_self = AttnProcessor()(attn2 := ShimAttention(dim, dim, heads), x)
    torch.save({"attn": attn.state_dict(), "attn_self": attn2.state_dict(), "ip": ip.state_dict(), "x": x, "ctx": ctx,
                "y_ip": y1, "y_ip2": y2, "y_self": y_self, "dim": dim, "ctx_dim": ctx_dim, "heads": heads,
                "scale": 0.8, "num_tokens": 16}, os.path.join(OUT, "ip_attn.pt"))


def make_resampler():
    from src.ip_adapter.resampler import Resampler
    torch.manual_seed(0)
    r = Resampler(dim=128, depth=2, dim_head=32, heads=4, num_queries=16, embedding_dim=512, output_dim=256, ff_mult=4)
    x = torch.randn(2, 1, 512)
    with torch.no_grad():
        y = r(x)
    torch.save({"sd": r.state_dict(), "x": x, "y": y, "heads": 4, "dim_head": 32}, os.path.join(OUT, "resampler.pt"))


def make_kps():
    import ast
    import math

    import cv2
    import numpy as np
    import PIL.Image
    src = open(os.path.join(REF, "inference_instantid.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "draw_kps_multi")
    ns = {"np": np, "cv2": cv2, "math": math, "PIL": PIL}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "inference_instantid.py", "exec"), ns)
    kps = [[[60 + d, 90], [110 + d, 88], [86 + d, 120], [66 + d, 150], [108 + d, 148]] for d in (0, 70, 120)]
    img = np.asarray(ns["draw_kps_multi"](PIL.Image.new("RGB", (256, 256)), kps))
    np.savez_compressed(os.path.join(OUT, "kps.npz"), kps=np.array(kps), image=img)


def _extract(path, name, cls=None, extra_globals=None):
    import ast
    tree = ast.parse(open(os.path.join(REF, path)).read())
    body = tree.body
    if cls is not None:
        body = next(n for n in body if isinstance(n, ast.ClassDef) and n.name == cls).body
    fn = next(n for n in body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"torch": torch, "F": torch.nn.functional}
    ns.update(extra_globals or {})
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns[name]


REGION_CASES = [(False, False, 0), (False, False, 5), (False, True, 1), (False, True, 7),
                (True, False, 0), (True, False, 5), (True, True, 1), (True, True, 7)]  # (edit, is_cross, step)


def region_case_setup(edit):
    prompts = ["a photo of a man on the beach", "a photo of a dog on the beach"] if edit else ["a b c"] * 2
    cross = {"default_": 0.6, "dog": (0.2, 0.9)} if edit else {"default_": 1.0}
    return prompts, cross


def region_case_inputs(idx, is_cross, dim=64, ctx_dim=32, n=16):
    """Inputs of case idx, regenerated from the seed by the tests (keeps the fixture small)."""
    gen = torch.Generator().manual_seed(100 + idx)
    x = torch.randn(4, n, dim, generator=gen)
    ctx = torch.randn(4, 77, ctx_dim, generator=gen) if is_cross else None
    return x, ctx


def make_region_attn():
    import ast
    from typing import Optional
    from src.prompt_attention.p2p_attention import AttentionReplace
    path = "src/pipelines/lora_pipeline.py"
    tree = ast.parse(open(os.path.join(REF, path)).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "RegionControlNet_AttnProcessor")
    ns = {"torch": torch, "Optional": Optional, "USE_PEFT_BACKEND": True}
    exec(compile(ast.Module(body=[cls], type_ignores=[]), path, "exec"), ns)
    Proc = ns["RegionControlNet_AttnProcessor"]
    dim, ctx_dim, heads = 64, 32, 2
    torch.manual_seed(7)
    attn_self, attn_cross = ShimAttention(dim, dim, heads), ShimAttention(dim, ctx_dim, heads)
    out = {"dim": dim, "ctx_dim": ctx_dim, "heads": heads, "attn_self": attn_self.state_dict(),
           "attn_cross": attn_cross.state_dict(), "cases": []}
    for idx, (edit, is_cross, step) in enumerate(REGION_CASES):
        prompts, cross = region_case_setup(edit)
        x, ctx = region_case_inputs(idx, is_cross, dim, ctx_dim)
        c = AttentionReplace(prompts, 10, dict(cross), 0.3, tokenizer=ToyTokenizer(), width=4, height=4)
        c.num_att_layers = 2
        c.cur_step = step
        with torch.no_grad():
            y = Proc(controller=c, place_in_unet="mid")(attn_cross if is_cross else attn_self, x, ctx)
        out["cases"].append({"edit": edit, "is_cross": is_cross, "step": step, "y": y, "cur_att_layer": c.cur_att_layer})
    torch.save(out, os.path.join(OUT, "region_attn.pt"))


def make_fusion():
    """Execute the reference's own fusion statements: three concepts, one of them without a mask (skipped), two with
    overlapping masks (sum in the overlap), one mask with non-binary values."""
    import ast
    import types
    path = "src/pipelines/lora_pipeline.py"
    tree = ast.parse(open(os.path.join(REF, path)).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "LoraMultiConceptPipeline")
    call = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "__call__")
    nodes = []
    for n in ast.walk(call):
        if isinstance(n, ast.If):
            src = ast.unparse(n.test)
            if src == "i > 15 and stage == 2" or src == "self.do_classifier_free_guidance":
                nodes.append(n)
    fuse_if = next(n for n in nodes if ast.unparse(n.test).startswith("i > 15"))
    cfg_if = [n for n in nodes if ast.unparse(n.test) == "self.do_classifier_free_guidance"
              and "noise_pred.chunk" in ast.unparse(n)][0]
    code = compile(ast.Module(body=[fuse_if, cfg_if], type_ignores=[]), path, "exec")
    region_mask = _extract(path, "get_region_mask", cls="LoraMultiConceptPipeline")
    g = torch.Generator().manual_seed(3)
    h, w, H, W = 16, 24, 64, 96
    noise_pred = torch.randn(4, 4, h, w, generator=g).half().float()   # fp16-representable: the CUDA path stores fp16
    latent_model_input = torch.randn(4, 4, h, w, generator=g).half().float()
    m0 = torch.zeros(H, W)
    m0[8:48, 4:52] = 1
    m1 = torch.zeros(H, W)
    m1[20:60, 40:88] = 1
    m1[20:24] *= 0.5          # non-binary stripe: not part of the region (== 1 tests)
    mask_list = [m0, None, m1]
    region_noise = [torch.randn(2, 4, h, w, generator=g).half().float() for _ in range(3)]
    log = {"adapters": [], "unet_inputs": []}
    calls = iter([region_noise[0], region_noise[2]])  # the concept without a mask never runs

    def unet(sample, t, encoder_hidden_states=None, cross_attention_kwargs=None, added_cond_kwargs=None, return_dict=False):
        log["unet_inputs"].append((sample.clone(), float(t), dict(cross_attention_kwargs)))
        return (next(calls),)

    concept_models = types.SimpleNamespace(_execution_device="cpu", unet=unet,
                                           set_adapters=lambda *a, **k: log["adapters"].append((a, k)))
    self_stub = types.SimpleNamespace(do_classifier_free_guidance=True,
                                      get_region_mask=lambda ml, fh, fw: region_mask(None, ml, fh, fw))
    ns = {"torch": torch, "F": torch.nn.functional, "self": self_stub, "i": 16, "stage": 2, "t": 499.0,
          "noise_pred": noise_pred.clone(), "mask_list": mask_list, "latent_model_input": latent_model_input,
          "region_prompt_embeds_list": [None] * 3, "region_add_text_embeds_list": [None] * 3,
          "add_time_ids_list": [None] * 3, "region_prompts": ["a", "b", "c"], "lora_list": ["A", "B", "C"],
          "styleL": True, "concept_models": concept_models, "guidance_scale": 7.5}
    exec(code, ns)
    torch.save({"noise_pred_in": noise_pred, "latent_model_input": latent_model_input, "masks": mask_list,
                "region_noise": region_noise, "guidance_scale": 7.5, "noise_after_cfg": ns["noise_pred"],
                "new_noise_pred": ns["new_noise_pred"], "adapters": log["adapters"],
                "unet_inputs": log["unet_inputs"]}, os.path.join(OUT, "fusion.pt"))


def make_fusion_instantid():
    import ast
    import types
    path = "src/pipelines/instantid_pipeline.py"
    tree = ast.parse(open(os.path.join(REF, path)).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "InstantidMultiConceptPipeline")
    call = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "__call__")
    ifs = [n for n in ast.walk(call) if isinstance(n, ast.If)]
    fuse_if = next(n for n in ifs if ast.unparse(n.test) == "i > 15 and stage == 2")
    cfg_if = next(n for n in ifs if ast.unparse(n.test) == "self.do_classifier_free_guidance"
                  and "noise_pred.chunk" in ast.unparse(n))
    code = compile(ast.Module(body=[fuse_if, cfg_if], type_ignores=[]), path, "exec")
    region_mask = _extract(path, "get_region_mask", cls="InstantidMultiConceptPipeline")
    g = torch.Generator().manual_seed(11)

    def rn(*s):
        return torch.randn(*s, generator=g)

    h, w, H, W, D, P = 4, 6, 16, 24, 8, 8
    noise_pred, lmi = rn(4, 4, h, w), rn(4, 4, h, w)
    m0 = torch.zeros(H, W)
    m0[2:14, 1:11] = 1
    m1 = torch.zeros(H, W)
    m1[4:15, 9:23] = 1
    mask_list = [m0, None, m1]
    text = [rn(2, 77, D) for _ in range(3)]
    face = [rn(2, 16, D) for _ in range(3)]
    pooled = [rn(2, P) for _ in range(3)]
    tids = [rn(2, 6) for _ in range(3)]
    cond = rn(2, 3, 8 * h, 8 * w)
    cn_out = [([rn(2, 3, h, w), rn(2, 5, h // 2, w // 2)], rn(2, 7, h // 4, w // 4)) for _ in range(2)]
    un_out = [rn(2, 4, h, w) for _ in range(2)]
    log = {"controlnet": [], "unet": []}
    cn_it, un_it = iter(cn_out), iter(un_out)

    def controlnet(sample, t, encoder_hidden_states=None, controlnet_cond=None, conditioning_scale=None,
                   guess_mode=None, added_cond_kwargs=None, return_dict=False):
        log["controlnet"].append({"sample": sample.clone(), "t": float(t), "ctx": encoder_hidden_states.clone(),
                                  "cond": controlnet_cond.clone(), "scale": conditioning_scale, "guess_mode": guess_mode,
                                  "text_embeds": added_cond_kwargs["text_embeds"].clone(),
                                  "time_ids": added_cond_kwargs["time_ids"].clone()})
        return next(cn_it)

    def unet(sample, t, encoder_hidden_states=None, cross_attention_kwargs=None, down_block_additional_residuals=None,
             mid_block_additional_residual=None, added_cond_kwargs=None, return_dict=False):
        log["unet"].append({"sample": sample.clone(), "t": float(t), "ctx": encoder_hidden_states.clone(),
                            "cross_attention_kwargs": cross_attention_kwargs,
                            "down": [d.clone() for d in down_block_additional_residuals],
                            "mid": mid_block_additional_residual.clone(),
                            "text_embeds": added_cond_kwargs["text_embeds"].clone(),
                            "time_ids": added_cond_kwargs["time_ids"].clone()})
        return (next(un_it),)

    self_stub = types.SimpleNamespace(do_classifier_free_guidance=True, controlnet=controlnet,
                                      get_region_mask=lambda ml, fh, fw: region_mask(None, ml, fh, fw))
    ns = {"torch": torch, "F": torch.nn.functional, "self": self_stub, "i": 16, "stage": 2, "t": 433.0,
          "noise_pred": noise_pred.clone(), "mask_list": mask_list, "latent_model_input": lmi,
          "region_prompt_embeds_list": text, "region_add_text_embeds_list": pooled, "add_time_ids_list": tids,
          "region_prompts": ["a", "b", "c"], "image_prompt_image_emb_list": face, "image": cond, "cond_scale": 0.8,
          "guess_mode": False, "concept_models": types.SimpleNamespace(_execution_device="cpu", unet=unet),
          "guidance_scale": 3.0}
    exec(code, ns)
    torch.save({"noise_pred_in": noise_pred, "latent_model_input": lmi, "masks": mask_list, "text": text, "face": face,
                "pooled": pooled, "time_ids": tids, "cond": cond, "cond_scale": 0.8, "guidance_scale": 3.0, "t": 433.0,
                "controlnet_out": cn_out, "unet_out": un_out, "controlnet_calls": log["controlnet"],
                "unet_calls": log["unet"], "noise_after_cfg": ns["noise_pred"]}, os.path.join(OUT, "fusion_iid.pt"))


def make_cli():
    lora_pt = _extract("inference_lora.py", "prepare_text")
    iid_pt = _extract("inference_instantid.py", "prepare_text")
    region_mask = _extract("src/pipelines/lora_pipeline.py", "get_region_mask", cls="LoraMultiConceptPipeline")
    lora_in = ["[a man, smiling]-*-[blurry, ugly]|[a woman [with] hat]-*-[noisy]", "[only one]-*-[neg]|", ""]
    iid_in = ["[a man]-*-[blurry]-*-./a.jpg|[a woman]-*-[noisy]-*-./b.png", ""]
    g = torch.Generator().manual_seed(0)
    masks = []
    for shape in ((64, 64), (64, 64), (48, 80)):
        m = (torch.rand(shape, generator=g) > 0.6).float()
        m[: shape[0] // 4] = 0.5  # non-binary values are NOT part of a region (mask == 1, lora_pipeline.py:680)
        masks.append(m)
    cases = [([masks[0], None, masks[1]], 8, 8), ([masks[2]], 6, 10), ([None, None], 8, 8)]
    out = {"lora_prepare_text": [(s, lora_pt("P", s)) for s in lora_in],
           "instantid_prepare_text": [(s, iid_pt("P", s)) for s in iid_in],
           "region_mask": [{"masks": ml, "h": h, "w": w, "out": region_mask(None, ml, h, w)} for ml, h, w in cases]}
    torch.save(out, os.path.join(OUT, "cli.pt"))


def make_cli_flags():
    """Flag names, defaults and types of both CLIs' parse_args (inference_lora.py:203-222,
    inference_instantid.py:259-286): the function is cut out of the file by ast and run against argparse."""
    import argparse
    import json
    out = {}
    for fname in ("inference_lora.py", "inference_instantid.py"):
        fn = _extract(fname, "parse_args", extra_globals={"argparse": argparse})
        argv, sys.argv = sys.argv, ["x"]
        try:
            ns = fn()
        finally:
            sys.argv = argv
        out[fname] = {k: [v, type(v).__name__] for k, v in sorted(vars(ns).items())}
    with open(os.path.join(OUT, "cli_flags.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


def _import_efficientvit():
    """The reference's EfficientViT modules; packages that only its training apps / mask decoder import are stubbed."""
    import importlib
    import types

    class Stub(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return type(k, (), {})

    importlib.import_module("src")
    importlib.import_module("src.efficientvit")
    for name in ["src.efficientvit.apps.trainer.run_config", "segment_anything.modeling.mask_decoder",
                 "segment_anything.modeling.prompt_encoder", "segment_anything.utils.amg", "segment_anything.utils.transforms",
                 "torchvision.transforms.functional", "timm", "onnx"]:
        parts = name.split(".")
        for i in range(1, len(parts) + 1):
            n = ".".join(parts[:i])
            if n not in sys.modules:
                m = Stub(n)
                m.__path__ = []
                sys.modules[n] = m
    from src.efficientvit.models.efficientvit import sam as S
    from src.efficientvit.models.efficientvit.backbone import EfficientViTLargeBackbone
    from src.efficientvit.models.nn import ops as O
    from src.efficientvit.models.nn.norm import set_norm_eps
    return S, EfficientViTLargeBackbone, O, set_norm_eps


def _randomise_bn(module, g):
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(1.0 + 0.2 * torch.randn(m.weight.shape, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.bias.shape, generator=g))


def make_sam_encoder():
    S, Backbone, O, set_norm_eps = _import_efficientvit()
    g = torch.Generator().manual_seed(0)
    torch.manual_seed(0)
    # the xl1 layout (sam.py:630-652) at narrow widths; every width a multiple of 32 (LiteMLA heads of 32 channels)
    bb = Backbone(width_list=[32, 32, 64, 64, 128, 128], depth_list=[1, 1, 1, 1, 2, 1],
                  block_list=["res", "fmb", "fmb", "fmb", "att@3", "att@3"], expand_list=[1, 4, 4, 4, 4, 6],
                  fewer_norm_list=[False, False, False, False, True, True])
    neck = S.SamNeck(fid_list=["stage5", "stage4", "stage3"], in_channel_list=[128, 128, 64], head_width=64, head_depth=2,
                     expand_ratio=4, middle_op="fmb")
    enc = S.EfficientViTSamImageEncoder(bb, neck).eval()
    set_norm_eps(enc, 1e-6)                                    # sam_model_zoo.py:44
    _randomise_bn(enc, g)
    for p in enc.parameters():                                 # fp16-representable weights (the kernels store fp16)
        p.data = p.data.half().float()
    enc.norm.weight.data.copy_((1.0 + 0.1 * torch.randn(256, generator=g)).half().float())
    enc.norm.bias.data.copy_((0.1 * torch.randn(256, generator=g)).half().float())
    x = torch.randn(1, 3, 384, 384, generator=g).half().float()   # stage3 48^2 / stage4 24^2 / stage5 12^2 -> 64^2 (bicubic up)
    with torch.no_grad():
        feats = bb(x)
        y = enc(x)
    sd = {k: (v.half() if v.is_floating_point() else v) for k, v in enc.state_dict().items() if "num_batches" not in k}
    # one LiteMLA with the 5x5 aggregation of the smaller zoo models, and one EfficientViTBlock, on their own
    mla = O.LiteMLA(64, 64, dim=32, scales=(5,), norm=(None, "bn2d")).eval()
    blk = O.EfficientViTBlock(96, dim=32, expand_ratio=4, scales=(3,), norm="bn2d", act_func="gelu").eval()
    for m in (mla, blk):
        set_norm_eps(m, 1e-6)
        _randomise_bn(m, g)
        for p in m.parameters():
            p.data = p.data.half().float()
    xm = torch.randn(2, 64, 24, 20, generator=g).half().float()
    xb = torch.randn(1, 96, 16, 16, generator=g).half().float()
    with torch.no_grad():
        ym, yb = mla(xm), blk(xb)
    torch.save({"sd": sd, "x": x.half(), "y": y.half(), "stage_shapes": {k: tuple(v.shape) for k, v in feats.items()},
                "stage5": feats["stage5"].half(), "stage3": feats["stage3"].half(),
                "mla_sd": {k: v.half() if v.is_floating_point() else v for k, v in mla.state_dict().items() if "num_batches" not in k},
                "mla_x": xm.half(), "mla_y": ym.half(),
                "blk_sd": {k: v.half() if v.is_floating_point() else v for k, v in blk.state_dict().items() if "num_batches" not in k},
                "blk_x": xb.half(), "blk_y": yb.half()}, os.path.join(OUT, "sam_encoder.pt"))


if __name__ == "__main__":
    make_sam_encoder()
    make_cli_flags()
    make_fusion_instantid()
    make_region_attn()
    make_fusion()
    make_cli()
    make_kps()
    make_p2p()
    make_ip()
    make_resampler()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
