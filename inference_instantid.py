#!/usr/bin/env python
"""OMG + InstantID multi-identity generation on the B200 path.  Same flags and two-stage flow as the reference CLI
(inference_instantid.py:257-393); additions: --synthetic, --num_inference_steps, --image_size, --tiny.

Face analysis (insightface), segmentation and the VAE / text encoders are outside the accelerated hot path (SURVEY
section 8): in --synthetic mode identities are unit-norm random 512-d embeddings (seeds 1, 2), the IdentityNet
condition is the reference's `draw_kps_multi` rendering of fixed key-points, masks are the config rectangles.
"""
import argparse
import math
import os

import numpy as np
import torch


def draw_kps_multi(image_size, kps_list, color_list=((255, 0, 0), (0, 255, 0), (0, 0, 255), (255, 255, 0),
                                                     (255, 0, 255))):
    """Key-point condition image of the IdentityNet: limbs as filled rotated ellipses towards the nose point, then
    the five points as discs (inference_instantid.py:127-156).  image_size = (w, h); returns an HxWx3 uint8 array."""
    import cv2
    stick_width, limb_seq = 4, np.array([[0, 2], [1, 2], [3, 2], [4, 2]])
    w, h = image_size
    canvas = np.zeros([h, w, 3])
    for kps in kps_list:
        kps = np.array(kps)
        for a, b in limb_seq:
            color = color_list[a]
            xs, ys = kps[[a, b]][:, 0], kps[[a, b]][:, 1]
            length = ((xs[0] - xs[1]) ** 2 + (ys[0] - ys[1]) ** 2) ** 0.5
            angle = math.degrees(math.atan2(ys[0] - ys[1], xs[0] - xs[1]))
            poly = cv2.ellipse2Poly((int(np.mean(xs)), int(np.mean(ys))), (int(length / 2), stick_width), int(angle),
                                    0, 360, 1)
            canvas = cv2.fillConvexPoly(canvas.copy(), poly, color)
        canvas = (canvas * 0.6).astype(np.uint8)  # truncation per face, as the reference does (:148)
        for idx, (x, y) in enumerate(kps):
            canvas = cv2.circle(canvas.copy(), (int(x), int(y)), 10, color_list[idx], -1)
    return canvas.astype(np.uint8)


def prepare_text(prompt, region_prompts):
    """'[prompt]-*-[negative]-*-ref_image|...' -> [(region, negative, ref_image), ...]
    (inference_instantid.py:233-254)."""
    out = []
    for region in region_prompts.split("|"):
        if region == "":
            break
        pos, neg, ref = region.split("-*-")
        out.append((pos.replace("[", "").replace("]", ""), neg.replace("[", "").replace("]", ""), ref))
    return prompt, out


def parse_args():
    p = argparse.ArgumentParser("", add_help=True)
    p.add_argument("--pretrained_model", default="./checkpoint/YamerMIX_v8", type=str)
    p.add_argument("--controlnet_path", default="./checkpoint/InstantID/ControlNetModel", type=str)
    p.add_argument("--face_adapter_path", default="./checkpoint/InstantID/ip-adapter.bin", type=str)
    p.add_argument("--openpose_checkpoint", default="./checkpoint/controlnet-openpose-sdxl-1.0", type=str)
    p.add_argument("--canny_checkpoint", default="./checkpoint/controlnet-canny-sdxl-1.0", type=str)
    p.add_argument("--depth_checkpoint", default="./checkpoint/controlnet-depth-sdxl-1.0", type=str)
    p.add_argument("--efficientViT_checkpoint", default="./checkpoint/sam/xl1.pt", type=str)
    p.add_argument("--dino_checkpoint", default="./checkpoint/GroundingDINO", type=str)
    p.add_argument("--sam_checkpoint", default="./checkpoint/sam/sam_vit_h_4b8939.pth", type=str)
    p.add_argument("--dpt_checkpoint", default="./checkpoint/dpt-hybrid-midas", type=str)
    p.add_argument("--pose_detector_checkpoint", default="./checkpoint/ControlNet/annotator/ckpts/body_pose_model.pth",
                   type=str)
    p.add_argument("--prompt", default="Close-up photo of the happy smiles on the faces of the cool man and beautiful "
                   "woman as they leave the island with the treasure, sail back to the vacation beach, and begin their "
                   "love story, 35mm photograph, film, professional, 4k, highly detailed.", type=str)
    p.add_argument("--negative_prompt", default="noisy, blurry, soft, deformed, ugly", type=str)
    p.add_argument("--prompt_rewrite", type=str,
                   default="[Close-up photo of a man, 35mm photograph, professional, 4k, highly detailed.]-*-[noisy, "
                           "blurry, soft, deformed, ugly]-*-./example/chris-evans.jpg|[Close-up photo of a woman, 35mm "
                           "photograph, professional, 4k, highly detailed.]-*-[noisy, blurry, soft, deformed, ugly]-*-"
                           "./example/TaylorSwift.png")
    p.add_argument("--seed", default=53, type=int)
    p.add_argument("--suffix", default="", type=str)
    p.add_argument("--segment_type", default="yoloworld", type=str)
    p.add_argument("--spatial_condition", type=str, default=None)
    p.add_argument("--t2i_controlnet_conditioning_scale", default=1.0, type=float)
    p.add_argument("--style_lora", default="", type=str)
    p.add_argument("--save_dir", default="results/instantID", type=str)
    p.add_argument("--guidance_scale", default=3.0, type=float)
    p.add_argument("--controlnet_conditioning_scale", default=0.8, type=float)
    p.add_argument("--ip_adapter_scale", default=0.8, type=float)
    # additions
    p.add_argument("--dedup", action="store_true", help="skip work that repeats identical work (same outputs): twin "
                   "rows before the first fusion step, stage-2 steps 0..15")
    p.add_argument("--synthetic", action="store_true")
    p.add_argument("--tiny", action="store_true")
    p.add_argument("--num_inference_steps", default=50, type=int)
    p.add_argument("--image_size", default=1024, type=int)
    return p.parse_args()


def build_synthetic(args, device):
    from omg_b200 import synthetic
    from omg_b200.config import UNetConfig
    from omg_b200.pipelines import ConceptModels, InstantidMultiConceptPipeline, revise_regionally_controlnet_forward
    from omg_b200.prompt_attention import AttentionReplace
    from omg_b200.unet import PackedUNet
    cfg = UNetConfig.tiny() if args.tiny else UNetConfig.sdxl()
    sd = synthetic.make_state_dict(cfg, seed=0, device=device, dtype=torch.float16)
    unet = PackedUNet(cfg, sd, device=device)
    idnet = PackedUNet(cfg, synthetic.make_state_dict(cfg, seed=1, controlnet=True, device=device,
                                                      dtype=torch.float16), device=device, controlnet=True)
    pipe = InstantidMultiConceptPipeline(unet, controlnet=idnet)
    controller = AttentionReplace([args.prompt] * 2, 50, cross_replace_steps={"default_": 1.0}, self_replace_steps=0.4,
                                  width=args.image_size // 32, height=args.image_size // 32)
    revise_regionally_controlnet_forward(pipe, controller)
    cm = ConceptModels(unet)
    # perceiver Resampler of InstantID (dim 1280, depth 4, heads 20, 16 queries, 512 -> cross_attention_dim)
    dim, depth, heads = (1280, 4, 20) if not args.tiny else (128, 2, 4)
    g = torch.Generator().manual_seed(7)
    D = cfg.cross_attention_dim

    def rn(*s, fan=None):
        return torch.randn(*s, generator=g) * ((fan or s[-1]) ** -0.5)

    rs = {"latents": rn(1, 16, dim), "proj_in.weight": rn(dim, 512), "proj_in.bias": torch.zeros(dim),
          "proj_out.weight": rn(D, dim), "proj_out.bias": torch.zeros(D), "norm_out.weight": torch.ones(D),
          "norm_out.bias": torch.zeros(D)}
    for i in range(depth):
        a, f = f"layers.{i}.0", f"layers.{i}.1"
        for n in ("norm1", "norm2"):
            rs[f"{a}.{n}.weight"], rs[f"{a}.{n}.bias"] = torch.ones(dim), torch.zeros(dim)
        rs[f"{a}.to_q.weight"], rs[f"{a}.to_kv.weight"], rs[f"{a}.to_out.weight"] = rn(dim, dim), rn(2 * dim, dim), rn(dim, dim)
        rs[f"{f}.0.weight"], rs[f"{f}.0.bias"] = torch.ones(dim), torch.zeros(dim)
        rs[f"{f}.1.weight"], rs[f"{f}.3.weight"] = rn(4 * dim, dim), rn(dim, 4 * dim)
    cm.load_ip_adapter_instantid(rs, synthetic.make_ip_adapter(cfg, 31, device=device), heads=heads,
                                 dim_head=dim // heads, num_tokens=16)
    return pipe, controller, cm


if __name__ == "__main__":
    args = parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("the B200 path needs a CUDA device (there is no CPU fallback)")
    if not args.synthetic:
        raise SystemExit("loading real InstantID checkpoints needs the encoder / VAE / face-analysis front-end that is "
                         "scheduled after the hot path (SURVEY section 8f); run with --synthetic")
    device = torch.device("cuda")
    from omg_b200 import synthetic
    pipe, controller, cm = build_synthetic(args, device)
    pipe.dedup = args.dedup
    cm.set_ip_adapter_scale(args.ip_adapter_scale)
    size = args.image_size
    prompts = [args.prompt] * 2
    regions = prepare_text(args.prompt, args.prompt_rewrite)[1]
    g = torch.Generator().manual_seed(1)
    faces = [torch.nn.functional.normalize(torch.randn(512, generator=g), dim=0) for _ in regions]
    s = size / 1024.0
    kps = [[[300 * s + dx * s, 380 * s], [400 * s + dx * s, 380 * s], [350 * s + dx * s, 440 * s],
            [310 * s + dx * s, 500 * s], [390 * s + dx * s, 500 * s]] for dx in (0, 380)][: len(regions)]
    cond = torch.from_numpy(draw_kps_multi((size, size), kps)).permute(2, 0, 1).float() / 255.0
    masks = synthetic.rect_masks(len(regions), (size, size))
    common = dict(prompt=[prompts, regions], negative_prompt=[args.negative_prompt] * 2,
                  guidance_scale=args.guidance_scale, num_inference_steps=args.num_inference_steps, concept_models=cm,
                  controller=controller, height=size, width=size, output_type="latent", face_embeds=faces,
                  controlnet_conditioning_scale=args.controlnet_conditioning_scale, image=cond)
    img = pipe(stage=1, generator=torch.Generator(device).manual_seed(args.seed), **common).images
    controller.reset()
    img = pipe(stage=2, generator=torch.Generator(device).manual_seed(args.seed), region_masks=masks, **common).images
    save_dir = os.path.join(args.save_dir, f"seed_{args.seed}")
    os.makedirs(save_dir, exist_ok=True)
    print(f"save to: {save_dir}")
    for idx, name in ((0, "stage-1"), (1, "stage-2")):
        torch.save(img[idx].cpu(), os.path.join(save_dir, name + ".pt"))
