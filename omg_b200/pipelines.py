"""The two-stage OMG denoising pipelines on the B200 kernels, behind the reference call surface.

  LoraMultiConceptPipeline        <- src/pipelines/lora_pipeline.py:154-681
  InstantidMultiConceptPipeline   <- src/pipelines/instantid_pipeline.py:157-767
  ConceptModels                   <- the `concept_models` object the reference passes in (a diffusers
                                     StableDiffusionXLPipeline with LoRA adapters, inference_lora.py:159-170, or
                                     InstantidSingleConceptPipeline with IP-adapter, instantid_single_pieline.py:159-243)

`__call__` keeps the reference keyword arguments (prompt=[[global, global], [(region, region_neg[, ref]), ...]],
negative_prompt, generator, guidance_scale, num_inference_steps, cross_attention_kwargs, controller,
concept_models, stage, region_masks, lora_list, styleL, image, height, width, output_type, ...).  Text encoders,
VAE, segmentation and face analysis sit outside the hot path: prompts are turned into embeddings
by a pluggable `prompt_encoder`, `output_type="latent"` is native and image output needs a `vae_decoder`.

Per step (one iteration of lora_pipeline.py:485-632) the device executes: main UNet (CUDA graph) -> [concept
UNets (CUDA graphs)] -> omg_fuse_step.  No host sync happens inside the loop.
"""
import hashlib
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from .config import UNetConfig
from .prompt_attention import AttentionReplace
from .scheduler import EulerDiscreteSchedule
from .unet import PackedUNet, UNetRunner

FUSION_AFTER_STEP = 15  # `if i > 15 and stage == 2` (lora_pipeline.py:568)


@dataclass
class PipelineOutput:
    images: object


class SyntheticPromptEncoder:
    """Deterministic stand-in for the two CLIP text encoders (out of the hot path; no checkpoints offline):
    a prompt string maps to seeded N(0,1) (77, D) hidden states and a (P,) pooled vector."""

    def __init__(self, cfg: UNetConfig, ctx_len: int = 77):
        self.cfg, self.ctx_len = cfg, ctx_len

    def __call__(self, prompt: str, lora_scale=None):
        seed = int.from_bytes(hashlib.sha256(prompt.encode("utf-8")).digest()[:4], "little")
        g = torch.Generator().manual_seed(seed)
        return (torch.randn(self.ctx_len, self.cfg.cross_attention_dim, generator=g),
                torch.randn(self.cfg.pooled_dim, generator=g))


class ConceptModels:
    """Concept UNet + its adapters (what the reference reaches through `concept_models.unet`, `.set_adapters`,
    `.encode_prompt`, `._execution_device`, `.set_ip_adapter_scale`, `._encode_prompt_image_emb`)."""

    def __init__(self, unet: PackedUNet, prompt_encoder: Optional[Callable] = None):
        self.unet = unet
        self.prompt_encoder = prompt_encoder or SyntheticPromptEncoder(unet.cfg)
        self._loras: Dict[str, dict] = {}
        self._active: Tuple[Tuple[str, ...], Tuple[float, ...]] = ((), ())
        self.image_proj = None  # (state dict, heads, dim_head) of the Resampler

    @property
    def _execution_device(self):
        return self.unet.device

    def load_lora_weights(self, lora: dict, adapter_name: str, **_):
        """lora: Linear path -> (A [r,in], B [out,r], alpha/r) (diffusers-format LoRA, already key-converted)."""
        self._loras[adapter_name] = lora

    def set_adapters(self, adapter_names, adapter_weights=None):
        names = (adapter_names,) if isinstance(adapter_names, str) else tuple(adapter_names)
        weights = tuple(1.0 for _ in names) if adapter_weights is None else tuple(float(w) for w in adapter_weights)
        for n in names:
            if n not in self._loras:
                raise ValueError(f"Adapter name {n} not found")
        self._active = (names, weights)

    def active_lora_key(self, global_scale: float) -> str:
        names, weights = self._active
        key = "|".join(f"{n}:{w:g}" for n, w in zip(names, weights)) + f"@{global_scale:g}"
        if key not in self.unet.lora_sets:
            self.unet.add_lora_set(key, [(self._loras[n], w) for n, w in zip(names, weights)], global_scale)
        return key

    def encode_prompt(self, prompt, negative_prompt=None, lora_scale=None, **_):
        pe, pp = self.prompt_encoder(prompt, lora_scale)
        ne, np_ = self.prompt_encoder(negative_prompt or "", lora_scale)
        return pe[None], ne[None], pp[None], np_[None]

    # --- InstantID pieces (instantid_single_pieline.py:159-243) --------------------------------------------
    def load_ip_adapter_instantid(self, image_proj_sd: dict, ip_weights: dict, heads: int = 20, dim_head: int = 64,
                                  num_tokens: int = 16, scale: float = 0.5):
        self.image_proj = (image_proj_sd, heads, dim_head)
        self.unet.set_ip_adapter(ip_weights, scale, num_tokens)

    def set_ip_adapter_scale(self, scale: float):
        self.unet.set_ip_adapter_scale(scale)

    def _encode_prompt_image_emb(self, prompt_image_emb, device=None, dtype=None, do_classifier_free_guidance=True):
        from .resampler import resampler_forward
        emb = torch.as_tensor(prompt_image_emb, dtype=torch.float32).reshape(1, -1, 512)
        if do_classifier_free_guidance:
            emb = torch.cat([torch.zeros_like(emb), emb], dim=0)
        sd, heads, dim_head = self.image_proj
        return resampler_forward(sd, emb.to(next(iter(sd.values())).device), heads, dim_head)


def _binary_latent_mask(mask: Optional[torch.Tensor], h: int, w: int, device) -> Optional[torch.Tensor]:
    """mask.float() -> nearest resize to (h, w) -> (== 1) (lora_pipeline.py:350,578-580,602,674-681)."""
    if mask is None:
        return None
    m = torch.nn.functional.interpolate(mask[None, None].float(), size=(h, w), mode="nearest")[0, 0]
    return (m == 1).float().reshape(-1).contiguous().to(device)


class _BasePipeline:
    vae_scale_factor = 8

    def __init__(self, unet: PackedUNet, controlnet: Optional[PackedUNet] = None,
                 prompt_encoder: Optional[Callable] = None, vae_decoder: Optional[Callable] = None,
                 use_graphs: bool = True):
        self.unet, self.controlnet = unet, controlnet
        self.controlnet2: Optional[PackedUNet] = None
        self.prompt_encoder = prompt_encoder or SyntheticPromptEncoder(unet.cfg)
        self.vae_decoder = vae_decoder
        self.scheduler = EulerDiscreteSchedule()
        self.use_graphs = use_graphs
        self.controller: Optional[AttentionReplace] = None
        self._runners: Dict[tuple, UNetRunner] = {}
        self.timings: Dict[str, float] = {}

    @property
    def _execution_device(self):
        return self.unet.device

    # ---------------------------------------------------------------------------------------------- helpers
    def _runner(self, tag, model: PackedUNet, batch, h, w, lora_key=None) -> UNetRunner:
        key = (tag, id(model), batch, h, w, lora_key)
        r = self._runners.get(key)
        if r is None:
            r = UNetRunner(model, batch, h, w, lora_key=lora_key, use_graphs=self.use_graphs)
            self._runners[key] = r
        return r

    def encode_prompt(self, prompt, negative_prompt, lora_scale=None):
        """-> prompt_embeds (n,77,D), negative (n,77,D), pooled (n,P), negative pooled (n,P) for a list of prompts."""
        prompts = [prompt] if isinstance(prompt, str) else list(prompt)
        negs = [negative_prompt] * len(prompts) if isinstance(negative_prompt, (str, type(None))) else list(negative_prompt)
        pe, pp, ne, np_ = [], [], [], []
        for p, n in zip(prompts, negs):
            e, pooled = self.prompt_encoder(p, lora_scale)
            e2, pooled2 = self.prompt_encoder(n or "", lora_scale)
            pe.append(e), pp.append(pooled), ne.append(e2), np_.append(pooled2)
        return torch.stack(pe), torch.stack(ne), torch.stack(pp), torch.stack(np_)

    def prepare_latents(self, h, w, generator, latents, dtype=torch.float16):
        """randn((1,4,h,w), generator) * init_noise_sigma, then cat([l, l.clone()]) (lora_pipeline.py:397-409)."""
        dev = self._execution_device
        if latents is None:
            gdev = generator.device if generator is not None else dev
            latents = torch.randn((1, 4, h, w), generator=generator, device=gdev, dtype=dtype).to(dev)
        lat = latents.to(dev).float() * self.scheduler.init_noise_sigma
        lat = torch.cat([lat, lat.clone()])
        return lat.permute(0, 2, 3, 1).contiguous()  # (2, h, w, 4) fp32 channels-last state

    def _p2p_variant(self, runner: UNetRunner, controller: Optional[AttentionReplace], residuals: bool):
        """Edit specification of the main UNet (rows u0,u1,c0,c1) for the controller's current step."""
        v = runner.default_variant()
        v["residuals"] = residuals
        key = ("main", residuals)
        if controller is None:
            return v, key + (False,)
        v["self_replace"] = controller.self_window_active()
        v["self_threshold"] = controller.width * controller.height
        v["self_items"] = [(0, 0, 0, 0), (1, 1, 1, 1), (2, 2, 2, 2), (3, 2, 2, 3)]
        # cross: c1 = P(c0) . V(row 4 = M diag(alpha) ctx_1)  [+ P(c1) . V(row 5 = diag(1-alpha) ctx_1)]
        items0 = [(0, 0, 0, 0), (1, 1, 1, 1), (2, 2, 2, 2), (3, 2, 2, 4)]
        v["cross_items"], v["cross_weights"] = [items0], [1.0]
        two = self._cross_two_terms
        if two:
            v["cross_items"].append([(3, 3, 3, 5)])
            v["cross_weights"].append(1.0)
        return v, key + (True, v["self_replace"], two)

    def _update_p2p_context(self, runner: UNetRunner, controller: AttentionReplace, ctx4: torch.Tensor, first: bool):
        """(Re)build the mixed context rows 4,5 when the step's alpha row differs from the previous step's."""
        coef_base, coef_keep = controller.cross_edit()
        sig = (coef_base.numpy().tobytes(), None if coef_keep is None else coef_keep.numpy().tobytes())
        if not first and sig == self._p2p_sig:
            return
        self._p2p_sig = sig
        dev = runner.dev
        c1 = ctx4[3:4].to(dev, torch.float16).contiguous()
        mix_a = ops.ctx_mix(c1, coef_base.to(dev).contiguous())
        if coef_keep is not None:
            mix_b = ops.ctx_mix(c1, coef_keep.to(dev).contiguous())
        else:
            mix_b = torch.zeros_like(c1)
        self._cross_two_terms = coef_keep is not None
        rows = torch.cat([mix_a, mix_b], dim=0)
        if first:
            self._p2p_rows = rows
        else:
            runner.update_context_rows(4, rows)

    def _finish(self, latents_nhwc: torch.Tensor, output_type: str, return_dict: bool):
        lat = latents_nhwc.permute(0, 3, 1, 2).contiguous().half()  # (2,4,h,w) like the reference's fp16 latents
        if output_type == "latent":
            image = lat
        else:
            if self.vae_decoder is None:
                raise RuntimeError("image output needs a vae_decoder (the VAE is outside the accelerated hot path); "
                                   "use output_type='latent'")
            image = self.vae_decoder(lat, output_type)
        return PipelineOutput(images=image) if return_dict else (image,)


class LoraMultiConceptPipeline(_BasePipeline):
    """src/pipelines/lora_pipeline.py:154-681."""

    def __call__(self, prompt=None, prompt_2=None, image=None, height: Optional[int] = None,
                 width: Optional[int] = None, num_inference_steps: int = 50, guidance_scale: float = 5.0,
                 negative_prompt=None, negative_prompt_2=None, num_images_per_prompt: int = 1, eta: float = 0.0,
                 generator=None, latents=None, prompt_embeds=None, negative_prompt_embeds=None,
                 pooled_prompt_embeds=None, negative_pooled_prompt_embeds=None, output_type: str = "pil",
                 return_dict: bool = True, cross_attention_kwargs=None, controlnet_conditioning_scale=1.0,
                 guess_mode: bool = False, control_guidance_start=0.0, control_guidance_end=1.0, original_size=None,
                 crops_coords_top_left=(0, 0), target_size=None, controller=None, concept_models: ConceptModels = None,
                 stage=None, region_masks=None, lora_list=None, styleL=None, region_prompt_embeds=None, **kwargs):
        dev = self._execution_device
        scale = (cross_attention_kwargs or {}).get("scale", 1.0)
        # 3.1 prompts: prompt = [[global, global], [(region, region_neg), ...]]  (lora_pipeline.py:310-347)
        global_prompt = prompt[0]
        region_prompts = [pt[0] for pt in prompt[1]]
        region_negs = [pt[1] for pt in prompt[1]]
        lora_list = list(lora_list or [])
        if prompt_embeds is None:
            prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds, negative_pooled_prompt_embeds = \
                self.encode_prompt(global_prompt, negative_prompt, scale)
        height = height or 128 * self.vae_scale_factor
        width = width or 128 * self.vae_scale_factor
        h, w = height // self.vae_scale_factor, width // self.vae_scale_factor
        concepts = []
        for k, (lp, rp, rn) in enumerate(zip(lora_list, region_prompts, region_negs)):
            if styleL:
                concept_models.set_adapters([lp, "style"], adapter_weights=[0.7, 0.5])
            else:
                concept_models.set_adapters(lp)
            if region_prompt_embeds is not None:
                pe, ne, pp, np_ = region_prompt_embeds[k]
            else:
                pe, ne, pp, np_ = concept_models.encode_prompt(rp, negative_prompt=rn, lora_scale=scale)
            concepts.append({"ctx": torch.cat([ne, pe], dim=0), "pooled": torch.cat([np_, pp], dim=0),
                             "lora_key": concept_models.active_lora_key(0.8)})  # {'scale': 0.8} (:596)
        masks = [None] * len(concepts)
        if stage == 2:
            masks = [_binary_latent_mask(m, h, w, dev) for m in region_masks]
        # 5/6 timesteps + latents
        ts = self.scheduler.set_timesteps(num_inference_steps)
        sig = self.scheduler.sigmas
        lat = self.prepare_latents(h, w, generator, latents)
        # 7.2 added conditioning
        original_size = original_size or (height, width)
        target_size = target_size or (height, width)
        tid = torch.tensor([list(original_size) + list(crops_coords_top_left) + list(target_size)], dtype=torch.float32)
        ctx4 = torch.cat([negative_prompt_embeds, prompt_embeds], dim=0)          # rows [neg0, neg1, pos0, pos1]
        pooled4 = torch.cat([negative_pooled_prompt_embeds, pooled_prompt_embeds], dim=0)
        # runners
        controller = self.controller  # installed by revise_regionally_controlnet_forward; the kwarg is ignored (:248)
        main = self._runner("main", self.unet, 4, h, w)
        extra = None
        if controller is not None:
            self._update_p2p_context(main, controller, ctx4, first=True)
            extra = self._p2p_rows
        main.set_conditioning(ts, ctx4, pooled4, tid.repeat(4, 1), extra_ctx=extra)
        cn = None
        use_cn = image is not None and self.controlnet is not None
        if use_cn:
            cn = self._runner("cn", self.controlnet, 4, h, w)
            cn.set_conditioning(ts, ctx4, pooled4, tid.repeat(4, 1))
            cn.set_controlnet_cond(self._prepare_image(image, width, height, 4))
        crun = []
        for k, c in enumerate(concepts):
            if stage == 2 and masks[k] is not None:
                r = self._runner(f"concept{k}", concept_models.unet, 2, h, w, lora_key=c["lora_key"])
                r.set_conditioning(ts, c["ctx"], c["pooled"], tid.repeat(2, 1))
                crun.append(r)
            else:
                crun.append(None)
        cn_scale = controlnet_conditioning_scale[0] if isinstance(controlnet_conditioning_scale, list) else controlnet_conditioning_scale
        # initial model inputs: scale_model_input(cat([latents]*2), t0)  (:491-492)
        x0 = (lat * self.scheduler.input_scale(0)).half()
        main.sample_in[..., :4] = torch.cat([x0, x0], dim=0)
        for r in crun:
            if r is not None:
                r.sample_in[..., :4] = torch.cat([x0[1:2], x0[1:2]], dim=0)
        if cn is not None:
            cn.sample_in.copy_(main.sample_in)
        lat = lat.contiguous()
        n_att = self.unet.num_attention_layers()
        for i in range(len(ts)):
            if controller is not None:
                self._update_p2p_context(main, controller, ctx4, first=False)
            variant, key = self._p2p_variant(main, controller, use_cn)
            if cn is not None:
                keep = 1.0 - float(i / len(ts) < control_guidance_start or (i + 1) / len(ts) > control_guidance_end)
                down, mid = cn.forward(i, key=("cn",))
                main.residuals_in = (down, mid, cn_scale * keep)
                key = key + (cn_scale * keep,)  # the scale is baked into the captured launch
            noise = main.forward(i, variant, key=key)
            if controller is not None:
                controller.advance(n_att)
            fuse = i > FUSION_AFTER_STEP and stage == 2
            cn_noise = []
            if fuse:
                for r in crun:
                    cn_noise.append(None if r is None else r.forward(i, key=("concept",)))
            ops.fuse_step(noise, cn_noise if fuse else [], masks if fuse else [], guidance_scale, float(sig[i]),
                          float(sig[i + 1]), lat, main.sample_in,
                          next((r.sample_in for r in crun if r is not None), None))
            # all concept runners read the same input (latent_model_input[3:4] twice, :583-585)
            first = next((r for r in crun if r is not None), None)
            for r in crun:
                if r is not None and r is not first:
                    r.sample_in.copy_(first.sample_in)
            if cn is not None:
                cn.sample_in.copy_(main.sample_in)
        return self._finish(lat, output_type, return_dict)

    def _prepare_image(self, image, width, height, batch):
        """ControlNet condition: list/tensor/PIL -> (batch, 3, H, W) in [0,1] (diffusers prepare_image [3P])."""
        img = image[0] if isinstance(image, (list, tuple)) else image
        if not torch.is_tensor(img):
            import numpy as np
            img = torch.from_numpy(np.asarray(img.convert("RGB").resize((width, height)), dtype="float32") / 255.0)
            img = img.permute(2, 0, 1)
        if img.dim() == 3:
            img = img[None]
        return img.float().repeat(batch // img.shape[0], 1, 1, 1)


def revise_regionally_controlnet_forward(pipe_or_unet, controller: AttentionReplace):
    """src/pipelines/lora_pipeline.py:136-152: hook the controller into every attention layer of the main UNet and
    set controller.num_att_layers = 2 * (#cross-attention layers).  Here the "processors" are the fused kernels, so
    installing means handing the pipeline the controller."""
    pipe = pipe_or_unet
    unet = pipe.unet if hasattr(pipe, "unet") else pipe
    count = unet.num_attention_layers() // 2
    print(f"Number of attention layer registered {count}")
    controller.num_att_layers = count * 2
    if hasattr(pipe, "controller"):
        pipe.controller = controller
    return controller


class InstantidMultiConceptPipeline(_BasePipeline):
    """src/pipelines/instantid_pipeline.py:157-767.  `self.controlnet` is the IdentityNet (used only in the concept
    pass, :638-648), `self.controlnet2` an optional spatial ControlNet for the main pass (:574-616)."""

    def __call__(self, prompt=None, image=None, height=None, width=None, num_inference_steps: int = 50,
                 guidance_scale: float = 5.0, negative_prompt=None, generator=None, latents=None,
                 output_type: str = "pil", return_dict: bool = True, cross_attention_kwargs=None,
                 controlnet_conditioning_scale=1.0, original_size=None, crops_coords_top_left=(0, 0), target_size=None,
                 controller=None, concept_models: ConceptModels = None, stage=None, region_masks=None, face_app=None,
                 t2i_image=None, t2i_controlnet_conditioning_scale=1.0, face_embeds=None, prompt_embeds=None,
                 negative_prompt_embeds=None, pooled_prompt_embeds=None, negative_pooled_prompt_embeds=None,
                 region_prompt_embeds=None, **kwargs):
        dev = self._execution_device
        scale = (cross_attention_kwargs or {}).get("scale", 1.0)
        global_prompt = prompt[0]
        regions = prompt[1]
        if prompt_embeds is None:
            prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds, negative_pooled_prompt_embeds = \
                self.encode_prompt(global_prompt, negative_prompt, scale)
        height = height or 128 * self.vae_scale_factor
        width = width or 128 * self.vae_scale_factor
        h, w = height // self.vae_scale_factor, width // self.vae_scale_factor
        concepts = []
        for k, reg in enumerate(regions):
            if region_prompt_embeds is not None:
                pe, ne, pp, np_ = region_prompt_embeds[k]
            else:
                pe, ne, pp, np_ = self.encode_prompt(reg[0], reg[1], scale)
            c = {"ctx": torch.cat([ne, pe], dim=0), "pooled": torch.cat([np_, pp], dim=0), "tokens": None}
            if stage == 2:
                if face_embeds is not None:
                    emb = face_embeds[k]
                else:
                    emb = self.get_face_embedding(face_app, reg[2])
                c["tokens"] = concept_models._encode_prompt_image_emb(emb, dev, torch.float16, True)  # (2,16,D)
            concepts.append(c)
        masks = [None] * len(concepts)
        if stage == 2:
            masks = [_binary_latent_mask(m, h, w, dev) for m in region_masks]
        ts = self.scheduler.set_timesteps(num_inference_steps)
        sig = self.scheduler.sigmas
        lat = self.prepare_latents(h, w, generator, latents)
        original_size = original_size or (height, width)
        target_size = target_size or (height, width)
        tid = torch.tensor([list(original_size) + list(crops_coords_top_left) + list(target_size)], dtype=torch.float32)
        ctx4 = torch.cat([negative_prompt_embeds, prompt_embeds], dim=0)
        pooled4 = torch.cat([negative_pooled_prompt_embeds, pooled_prompt_embeds], dim=0)
        controller = self.controller
        main = self._runner("main", self.unet, 4, h, w)
        extra = None
        if controller is not None:
            self._update_p2p_context(main, controller, ctx4, first=True)
            extra = self._p2p_rows
        main.set_conditioning(ts, ctx4, pooled4, tid.repeat(4, 1), extra_ctx=extra)
        cn2 = None
        if t2i_image is not None and self.controlnet2 is not None:
            cn2 = self._runner("cn2", self.controlnet2, 4, h, w)
            cn2.set_conditioning(ts, ctx4, pooled4, tid.repeat(4, 1))
            cn2.set_controlnet_cond(LoraMultiConceptPipeline._prepare_image(self, t2i_image, width, height, 4))
        crun, idrun = [], []
        use_id = image is not None and self.controlnet is not None
        for k, c in enumerate(concepts):
            if stage == 2 and masks[k] is not None:
                r = self._runner(f"concept{k}", concept_models.unet, 2, h, w)
                r.set_conditioning(ts, torch.cat([c["ctx"].to(dev), c["tokens"].to(dev)], dim=1), c["pooled"],
                                   tid.repeat(2, 1))
                crun.append(r)
                if use_id:
                    ir = self._runner(f"identity{k}", self.controlnet, 2, h, w)
                    ir.set_conditioning(ts, c["tokens"], c["pooled"], tid.repeat(2, 1))
                    ir.set_controlnet_cond(LoraMultiConceptPipeline._prepare_image(self, image, width, height, 2))
                    idrun.append(ir)
                else:
                    idrun.append(None)
            else:
                crun.append(None)
                idrun.append(None)
        cn_scale = controlnet_conditioning_scale[0] if isinstance(controlnet_conditioning_scale, list) else controlnet_conditioning_scale
        t2i_scale = t2i_controlnet_conditioning_scale[0] if isinstance(t2i_controlnet_conditioning_scale, list) else t2i_controlnet_conditioning_scale
        x0 = (lat * self.scheduler.input_scale(0)).half()
        main.sample_in[..., :4] = torch.cat([x0, x0], dim=0)
        for r in crun + idrun:
            if r is not None:
                r.sample_in[..., :4] = torch.cat([x0[1:2], x0[1:2]], dim=0)
        if cn2 is not None:
            cn2.sample_in.copy_(main.sample_in)
        lat = lat.contiguous()
        n_att = self.unet.num_attention_layers()
        for i in range(len(ts)):
            if controller is not None:
                self._update_p2p_context(main, controller, ctx4, first=False)
            variant, key = self._p2p_variant(main, controller, cn2 is not None)
            if cn2 is not None:
                down, mid = cn2.forward(i, key=("cn2",))
                main.residuals_in = (down, mid, t2i_scale)
                key = key + (t2i_scale,)
            noise = main.forward(i, variant, key=key)
            if controller is not None:
                controller.advance(n_att)
            fuse = i > FUSION_AFTER_STEP and stage == 2
            cn_noise = []
            if fuse:
                for r, ir in zip(crun, idrun):
                    if r is None:
                        cn_noise.append(None)
                        continue
                    v = r.default_variant()
                    if ir is not None:
                        down, mid = ir.forward(i, key=("identity",))
                        r.residuals_in = (down, mid, cn_scale)
                        v["residuals"] = True
                    cn_noise.append(r.forward(i, v, key=("concept", ir is not None, cn_scale)))
            first = next((r for r in crun if r is not None), None)
            ops.fuse_step(noise, cn_noise if fuse else [], masks if fuse else [], guidance_scale, float(sig[i]),
                          float(sig[i + 1]), lat, main.sample_in, None if first is None else first.sample_in)
            for r in crun + idrun:
                if r is not None and r is not first:
                    r.sample_in.copy_(first.sample_in)
            if cn2 is not None:
                cn2.sample_in.copy_(main.sample_in)
        return self._finish(lat, output_type, return_dict)

    def get_face_embedding(self, face_app, ref_image):
        """instantid_pipeline.py:757-767: the reference sorts detections by (x2-x0)*y2 - y1 ascending and takes the
        first (documented quirk); kept as is."""
        import cv2
        import numpy as np
        from PIL import Image
        info = face_app.get(cv2.cvtColor(np.array(Image.open(ref_image).convert("RGB")), cv2.COLOR_RGB2BGR))
        info = sorted(info, key=lambda x: (x["bbox"][2] - x["bbox"][0]) * x["bbox"][3] - x["bbox"][1])[0]
        return info["embedding"]
