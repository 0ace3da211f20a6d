"""On-disk formats: kohya / SGM / diffusers LoRA key layouts, InstantID ip-adapter.bin numbering,
diffusers safetensors discovery.  Synthetic files written with safetensors / torch.save; no GPU."""
import os

import pytest
import torch

from omg_b200 import checkpoints as ck
from omg_b200 import synthetic
from omg_b200.config import UNetConfig, lora_target_names, transformer_names


def _ref_lora(cfg, rank=4, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for i, (name, fin, fout) in enumerate(lora_target_names(cfg)):
        out[name] = (torch.randn(rank, fin, generator=g), torch.randn(fout, rank, generator=g), 8.0 if i % 2 else None)
    return out


def _same(conv, ref, rank):
    assert set(conv) == set(ref)
    for name, (A, B, alpha) in ref.items():
        a, b, s = conv[name]
        assert torch.equal(a, A) and torch.equal(b, B)
        assert s == pytest.approx((alpha if alpha is not None else rank) / rank)


def test_sgm_block_mapping_sdxl():
    cfg = UNetConfig.sdxl()
    f = ck._sgm_to_diffusers_block
    assert f(cfg, "input_blocks", 0, 0) == "conv_in"
    assert f(cfg, "input_blocks", 1, 0) == "down_blocks.0.resnets.0"
    assert f(cfg, "input_blocks", 3, 0) == "down_blocks.0.downsamplers.0"
    assert f(cfg, "input_blocks", 4, 1) == "down_blocks.1.attentions.0"
    assert f(cfg, "input_blocks", 5, 0) == "down_blocks.1.resnets.1"
    assert f(cfg, "input_blocks", 8, 1) == "down_blocks.2.attentions.1"
    assert f(cfg, "middle_block", 0, 1) == "mid_block.attentions.0"
    assert f(cfg, "output_blocks", 0, 1) == "up_blocks.0.attentions.0"
    assert f(cfg, "output_blocks", 2, 2) == "up_blocks.0.upsamplers.0"
    assert f(cfg, "output_blocks", 5, 1) == "up_blocks.1.attentions.2"
    assert f(cfg, "output_blocks", 8, 0) == "up_blocks.2.resnets.2"


@pytest.mark.parametrize("layout", ["kohya_diffusers", "kohya_sgm", "peft", "diffusers_old", "processor"])
def test_lora_layouts_convert_to_the_same_dict(layout, tmp_path):
    cfg = UNetConfig.sdxl()
    rank = 4
    ref = _ref_lora(cfg, rank)
    sgm_of = {}
    for tname, _c, _l in transformer_names(cfg):
        for part, n in (("input_blocks", 9), ("middle_block", 1), ("output_blocks", 9)):
            for idx in range(n):
                for sub in range(3):
                    if ck._sgm_to_diffusers_block(cfg, part, idx, sub) == tname:
                        sgm_of[tname] = f"{part}_{idx}_{sub}" if part != "middle_block" else f"middle_block_{sub}"
    sd = {}
    for name, (A, B, alpha) in ref.items():
        if layout == "kohya_diffusers":
            stem = "lora_unet_" + name.replace(".", "_")
            sd[stem + ".lora_down.weight"], sd[stem + ".lora_up.weight"] = A, B
            if alpha is not None:
                sd[stem + ".alpha"] = torch.tensor(alpha)
        elif layout == "kohya_sgm":
            t = next(t for t in sgm_of if name.startswith(t + "."))
            stem = "lora_unet_" + sgm_of[t] + "_" + name[len(t) + 1:].replace(".", "_")
            sd[stem + ".lora_down.weight"], sd[stem + ".lora_up.weight"] = A, B
            if alpha is not None:
                sd[stem + ".alpha"] = torch.tensor(alpha)
        elif layout == "peft":
            sd[f"unet.{name}.lora_A.weight"], sd[f"unet.{name}.lora_B.weight"] = A, B
            if alpha is not None:
                sd[f"unet.{name}.alpha"] = torch.tensor(alpha)
        elif layout == "diffusers_old":
            sd[f"unet.{name}.lora.down.weight"], sd[f"unet.{name}.lora.up.weight"] = A, B
            if alpha is not None:
                sd[f"unet.{name}.alpha"] = torch.tensor(alpha)
        else:  # attention-processor spelling exists for the four attention projections only
            m = name.rsplit(".", 2 if name.endswith("to_out.0") else 1)
            if ".attn" in name and (name.endswith(("to_q", "to_k", "to_v")) or name.endswith("to_out.0")):
                proj = "to_out" if name.endswith("to_out.0") else m[-1]
                base = name[: name.index(".to_")]
                sd[f"unet.{base}.processor.{proj}_lora.down.weight"] = A
                sd[f"unet.{base}.processor.{proj}_lora.up.weight"] = B
                if alpha is not None:
                    sd[f"unet.{base}.processor.{proj}_lora.alpha"] = torch.tensor(alpha)
    # text-encoder and unsupported entries ride along
    sd["lora_te1_text_model_encoder_layers_0_self_attn_q_proj.lora_down.weight"] = torch.zeros(rank, 768)
    sd["lora_te1_text_model_encoder_layers_0_self_attn_q_proj.lora_up.weight"] = torch.zeros(768, rank)
    sd["lora_unet_input_blocks_1_0_in_layers_2.lora_down.weight"] = torch.zeros(rank, 320, 3, 3)
    sd["lora_unet_input_blocks_1_0_in_layers_2.lora_up.weight"] = torch.zeros(320, rank, 1, 1)
    from safetensors.torch import save_file
    path = str(tmp_path / "lora.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items()}, path)
    unet_lora, te_lora, skipped = ck.load_lora(path, cfg)
    if layout == "processor":
        ref = {k: v for k, v in ref.items() if ".attn" in k}
    _same(unet_lora, ref, rank)
    assert list(te_lora) == ["te1.text_model_encoder_layers_0_self_attn_q_proj"]
    assert len(skipped) == 2
    with pytest.raises(ValueError):
        ck.load_lora(path, cfg, strict=True)


def test_lora_shape_mismatch_is_rejected():
    cfg = UNetConfig.sdxl()
    name, fin, fout = lora_target_names(cfg)[1]
    sd = {f"unet.{name}.lora_A.weight": torch.zeros(4, fin + 1), f"unet.{name}.lora_B.weight": torch.zeros(fout, 4)}
    with pytest.raises(ValueError):
        ck.convert_lora_state_dict(sd, cfg)


def test_pipeline_accepts_lora_path(tmp_path):
    """ConceptModels.load_lora_weights(path, weight_name=..., adapter_name=...) as in inference_lora.py:163-169."""
    from omg_b200.pipelines import _resolve_lora
    cfg = UNetConfig.tiny()
    lora = synthetic.make_lora(cfg, seed=3, rank=4)
    sd = {}
    for name, (A, B, s) in lora.items():
        stem = "lora_unet_" + name.replace(".", "_")
        sd[stem + ".lora_down.weight"], sd[stem + ".lora_up.weight"] = A.contiguous(), B.contiguous()
        sd[stem + ".alpha"] = torch.tensor(s * A.shape[0])
    from safetensors.torch import save_file
    os.makedirs(tmp_path / "adapter")
    save_file(sd, str(tmp_path / "adapter" / "pytorch_lora_weights.safetensors"))

    class Owner:
        class unet:
            pass
    Owner.unet.cfg = cfg
    o = Owner()
    got = _resolve_lora(o, str(tmp_path / "adapter"), "c0", None)
    assert set(got) == set(lora)
    for k in lora:
        assert torch.equal(got[k][0], lora[k][0]) and got[k][2] == pytest.approx(lora[k][2])
    assert o.text_encoder_loras == {"c0": {}}


def test_ip_adapter_bin_numbering(tmp_path):
    cfg = UNetConfig.sdxl()
    order = ck.attn_processor_order(cfg)
    assert len(order) == 140 and order[0] == "down_blocks.1.attentions.0.transformer_blocks.0.attn1"
    # diffusers registers down_blocks and up_blocks before mid_block
    assert order[48].startswith("up_blocks.0.") and order[120].startswith("mid_block.") and order[139].endswith("attn2")
    ip = {}
    g = torch.Generator().manual_seed(0)
    for i, p in enumerate(order):
        if p.endswith("attn2"):
            c = 1280 if ("blocks.2" in p and p.startswith("down")) or p.startswith(("mid", "up_blocks.0")) else 640
            ip[f"{i}.to_k_ip.weight"] = torch.randn(c, 8, generator=g)
            ip[f"{i}.to_v_ip.weight"] = torch.randn(c, 8, generator=g)
    path = str(tmp_path / "ip-adapter.bin")
    torch.save({"image_proj": {"latents": torch.zeros(1, 16, 1280)}, "ip_adapter": ip}, path)
    image_proj, weights = ck.load_ip_adapter(path, cfg)
    assert set(image_proj) == {"latents"} and len(weights) == 70
    assert torch.equal(weights[order[1]][0], ip["1.to_k_ip.weight"])
    assert torch.equal(weights[order[139]][1], ip["139.to_v_ip.weight"])
    bad = dict(ip)
    bad["0.to_k_ip.weight"] = torch.zeros(1)
    with pytest.raises(ValueError):
        ck.convert_ip_adapter({"ip_adapter": bad}, cfg)


def test_find_diffusers_weights(tmp_path):
    os.makedirs(tmp_path / "unet")
    with pytest.raises(FileNotFoundError):
        ck.find_diffusers_weights(str(tmp_path))
    from safetensors.torch import save_file
    save_file({"conv_in.weight": torch.zeros(1)}, str(tmp_path / "unet" / "diffusion_pytorch_model.safetensors"))
    assert ck.find_diffusers_weights(str(tmp_path)).endswith("diffusion_pytorch_model.safetensors")
    save_file({"conv_in.weight": torch.ones(1)}, str(tmp_path / "unet" / "diffusion_pytorch_model.fp16.safetensors"))
    assert ck.find_diffusers_weights(str(tmp_path)).endswith(".fp16.safetensors")
    assert float(ck.load_unet_weights(str(tmp_path))["conv_in.weight"]) == 1.0
