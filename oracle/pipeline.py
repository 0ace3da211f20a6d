"""Restatement of the OMG denoising loops.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows src/pipelines/lora_pipeline.py:397-409 (latent init / duplication), :485-615 (step loop, region noise fusion,
CFG, scheduler step), :674-681 (get_region_mask) and src/pipelines/instantid_pipeline.py:540-690 (same loop; the
concept pass runs IdentityNet + the IP-adapter UNet, the main pass may use a second ControlNet).  Text encoders,
VAE, segmentation and face analysis are outside the hot path: prompt embeddings, masks and face tokens are inputs.
"""
from dataclasses import dataclass
from typing import Callable, List, Optional

import torch
import torch.nn.functional as F

from .scheduler import EulerDiscrete
from .unet import Ctx, controlnet_forward, unet_forward


@dataclass
class Concept:
    """One region: everything the reference prepares before the loop (lora_pipeline.py:336-350,453-454)."""
    prompt_embeds: torch.Tensor          # (2, 77, D)   rows [negative_region, region]
    add_text_embeds: torch.Tensor        # (2, pooled)
    add_time_ids: torch.Tensor           # (2, 6)
    mask: Optional[torch.Tensor]         # (H_img, W_img) {0,1} or None (concept skipped, lora_pipeline.py:577)
    unet: Ctx = None                     # concept UNet context (base weights + this concept's LoRA / IP weights)
    image_tokens: Optional[torch.Tensor] = None  # InstantID: (2, 16, D) face tokens [zero-id, id]


def get_region_mask(mask_list, fh, fw):
    """lora_pipeline.py:674-681."""
    dev = next((m.device for m in mask_list if m is not None), "cpu")  # the reference builds it on the CPU (:675)
    exclusive = torch.zeros((fh, fw), device=dev)
    for mask in mask_list:
        if mask is not None:
            m = F.interpolate(mask[None, None].float(), size=(fh, fw), mode="nearest").squeeze().to(exclusive.dtype)
            exclusive = ((m == 1) | (exclusive == 1)).to(dtype=m.dtype)
    return exclusive


def fuse_noise(noise_pred, region_noises, mask_list):
    """lora_pipeline.py:569-607 with replace_ratio = 1."""
    fh, fw = noise_pred.shape[2], noise_pred.shape[3]
    mask_list = [None if m is None else m.to(noise_pred.device) for m in mask_list]
    region_mask = get_region_mask(mask_list, fh, fw)
    edit = torch.cat([noise_pred[1:2], noise_pred[3:4]], dim=0)
    new = torch.zeros_like(edit)
    new[:, :, region_mask == 0] = edit[:, :, region_mask == 0]
    new[:, :, region_mask != 0] = 0.0 * edit[:, :, region_mask != 0]
    for mask, rn in zip(mask_list, region_noises):
        if mask is None:
            continue
        cm = F.interpolate(mask[None, None].float(), size=(fh, fw), mode="nearest").squeeze()
        new[:, :, cm == 1] += rn[:, :, cm == 1] / cm.reshape(1, 1, *cm.shape)[:, :, cm == 1]
    out = noise_pred.clone()
    out[1] = new[0]
    out[3] = new[1]
    return out


def concept_noises(lmi, t, concepts, identitynet=None, identity_cond=None, identity_scale: float = 1.0):
    """One noise prediction per concept with a mask (None for the others): the concept loop of lora_pipeline.py:576-599
    / instantid_pipeline.py:626-672.  Every concept sees image 1's scaled latent twice; InstantID concepts run the
    IdentityNet on the face tokens alone and the UNet on [text tokens | face tokens]."""
    region_noises = []
    for c in concepts:
        if c.mask is None:
            region_noises.append(None)
            continue
        rl = torch.cat([lmi[3:4].clone()] * 2)                         # :583-585
        ctx = c.prompt_embeds
        rdown = rmid = None
        if identitynet is not None:                                    # instantid_pipeline.py:638-663
            rdown, rmid = controlnet_forward(identitynet, rl, t, c.image_tokens, identity_cond, identity_scale,
                                             c.add_text_embeds, c.add_time_ids)
        if c.image_tokens is not None:
            ctx = torch.cat([c.prompt_embeds, c.image_tokens], dim=1)
        region_noises.append(unet_forward(c.unet, rl, t, ctx, c.add_text_embeds, c.add_time_ids, rdown, rmid))
    return region_noises


def denoise(main: Ctx, latents_1: torch.Tensor, prompt_embeds, add_text_embeds, add_time_ids,
            concepts: List[Concept], stage: int, num_inference_steps: int, guidance_scale: float,
            controlnet: Optional[Ctx] = None, controlnet_cond=None, controlnet_scale: float = 1.0,
            identitynet: Optional[Ctx] = None, identity_cond=None, identity_scale: float = 1.0,
            fusion_after_step: int = 15, trace: Optional[Callable] = None, max_steps: Optional[int] = None):
    """latents_1: (1,4,h,w) unit-variance noise (randn of the seeded generator).  prompt_embeds (4,77,D) rows
    [neg0, neg1, pos0, pos1]; add_text_embeds (4,P); add_time_ids (4,6).  Returns final latents (2,4,h,w)."""
    sched = EulerDiscrete()
    timesteps = sched.set_timesteps(num_inference_steps)
    latents = latents_1 * sched.init_noise_sigma                      # prepare_latents
    latents = torch.cat([latents, latents.clone()])                    # :409
    n = len(timesteps) if max_steps is None else min(max_steps, len(timesteps))
    for i in range(n):
        t = timesteps[i]
        lmi = sched.scale_model_input(torch.cat([latents] * 2), i)     # :491-492
        down = mid = None
        if controlnet is not None and controlnet_cond is not None:     # :519-540 / instantid :574-590
            down, mid = controlnet_forward(controlnet, lmi, t, prompt_embeds, controlnet_cond, controlnet_scale,
                                           add_text_embeds, add_time_ids)
        noise_pred = unet_forward(main, lmi, t, prompt_embeds, add_text_embeds, add_time_ids, down, mid)
        if i > fusion_after_step and stage == 2:                       # :568
            region_noises = concept_noises(lmi, t, concepts, identitynet, identity_cond, identity_scale)
            noise_pred = fuse_noise(noise_pred, region_noises, [c.mask for c in concepts])
        nu, nt = noise_pred.chunk(2)                                    # :610-612
        guided = nu + guidance_scale * (nt - nu)
        latents = sched.step(guided, i, latents)                       # :615
        if trace is not None:
            trace(i, noise_pred, latents)
    return latents
