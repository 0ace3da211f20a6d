"""Builders for the synthetic workloads of BASELINE.json (random-init SDXL-shaped weights; there is no network for
checkpoints).  Used by bench.py, __graft_entry__.smoke() and the CLIs' --synthetic mode."""
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import synthetic
from .config import UNetConfig
from .pipelines import ConceptModels, LoraMultiConceptPipeline, revise_regionally_controlnet_forward
from .prompt_attention import AttentionReplace
from .unet import PackedUNet

GLOBAL_PROMPT = ("Close-up photo of the cool man and beautiful woman as they accidentally discover a mysterious island "
                 "while on vacation by the sea, facing the camera smiling, 35mm photograph, film, professional, 4k, "
                 "highly detailed.")  # inference_lora.py:211 default


@dataclass
class LoraWorkload:
    pipe: LoraMultiConceptPipeline
    controller: AttentionReplace
    concept_models: ConceptModels
    call_kwargs: dict
    masks: List[torch.Tensor]
    cfg: UNetConfig
    state_dict: Optional[Dict[str, torch.Tensor]] = None


def build_lora_workload(cfg: Optional[UNetConfig] = None, image_size: int = 1024, n_concepts: int = 2,
                        lora_rank: int = 32, num_inference_steps: int = 30, guidance_scale: float = 7.5,
                        seed: int = 0, device="cuda", state_dict=None, use_graphs: bool = True,
                        keep_state_dict: bool = False) -> LoraWorkload:
    """BASELINE config 2: SDXL, 1024^2, 30 steps, 2-concept LoRA fusion (src/prompt_attention path)."""
    cfg = cfg or UNetConfig.sdxl()
    if state_dict is None:
        state_dict = synthetic.make_state_dict(cfg, seed=seed, device=device, dtype=torch.float16)
    unet = PackedUNet(cfg, state_dict, device=device)
    pipe = LoraMultiConceptPipeline(unet, use_graphs=use_graphs)
    lat = image_size // 8
    # inference_lora.py:156,247: AttentionReplace(prompts, 50, {"default_": 1.}, 0.4, width=W//32, height=H//32)
    controller = AttentionReplace([GLOBAL_PROMPT] * 2, 50, cross_replace_steps={"default_": 1.0},
                                  self_replace_steps=0.4, width=image_size // 32, height=image_size // 32)
    revise_regionally_controlnet_forward(pipe, controller)
    # the concept pipeline is a second copy of the same base weights in the reference (inference_lora.py:159); the
    # packed weights are immutable, so it is shared here
    cm = ConceptModels(unet)
    names = []
    for k in range(n_concepts):
        name = f"concept{k}"
        cm.load_lora_weights(synthetic.make_lora(cfg, seed=1000 + k, rank=lora_rank, device=device), adapter_name=name)
        names.append(name)
    regions = [(f"Close-up photo of concept {k}, 35mm photograph, film, professional, 4k, highly detailed.",
                "noisy, blurry, soft, deformed, ugly") for k in range(n_concepts)]
    masks = synthetic.rect_masks(n_concepts, (image_size, image_size))
    kwargs = dict(prompt=[[GLOBAL_PROMPT] * 2, regions], negative_prompt=["noisy, blurry, soft, deformed, ugly"] * 2,
                  guidance_scale=guidance_scale, num_inference_steps=num_inference_steps,
                  cross_attention_kwargs={"scale": 0.8}, concept_models=cm, lora_list=names, styleL=False,
                  height=image_size, width=image_size, output_type="latent")
    return LoraWorkload(pipe, controller, cm, kwargs, masks, cfg, state_dict if keep_state_dict else None)


def run_two_stage(wl: LoraWorkload, seed: int = 14, device="cuda", latents=None, masks=None):
    """The CLI flow of inference_lora.py:262-297 without the segmentation models: stage 1 -> (masks) -> stage 2 from the
    same seed.  Returns (stage-1 latents, stage-2 latents), each (2,4,h,w) fp16."""
    gen = torch.Generator(device).manual_seed(seed)
    out1 = wl.pipe(stage=1, generator=gen, latents=latents, **wl.call_kwargs).images
    wl.controller.reset()
    gen = torch.Generator(device).manual_seed(seed)
    out2 = wl.pipe(stage=2, generator=gen, latents=latents, region_masks=masks if masks is not None else wl.masks,
                   **wl.call_kwargs).images
    wl.controller.reset()
    return out1, out2
