"""Prompt conditioning front-end: SDXL's two CLIP text encoders behind the `prompt_encoder` callable the
pipelines take.  Once per call and outside the denoising loop, so this is library plumbing (transformers' CLIP
modules on the GPU), not a kernel of this repo.

Restates diffusers 0.25 `StableDiffusionXLPipeline.encode_prompt` as the reference calls it
(src/pipelines/lora_pipeline.py:315-347) [3P]: each tokenizer pads / truncates to 77 tokens, each encoder runs with
`output_hidden_states=True`, the prompt embedding is the concatenation of both encoders' penultimate hidden states
(77 x (768 + 1280)), the pooled embedding is the projected output of the second encoder (1280).  Text-encoder LoRA
(`lora_te1_*` / `lora_te2_*`, `text_encoder.*` entries returned by omg_b200.checkpoints) is merged into the
affected Linear weights for the active adapter set, scaled like peft: weight * adapter_weight * lora_scale * alpha/r.
"""
import os
from typing import Dict, Optional, Sequence, Tuple

import torch


class ClipPromptEncoder:
    supports_adapters = True

    def __init__(self, tokenizers: Sequence, encoders: Sequence[torch.nn.Module], device="cuda",
                 dtype=torch.float16, max_length: int = 77):
        assert len(tokenizers) == len(encoders) >= 1
        self.tokenizers, self.device, self.dtype, self.max_length = list(tokenizers), torch.device(device), dtype, max_length
        self.encoders = [e.to(self.device, dtype).eval() for e in encoders]
        self._flat = []  # per encoder: kohya-flattened module name -> Linear
        for e in self.encoders:
            self._flat.append({n.replace(".", "_"): m for n, m in e.named_modules() if isinstance(m, torch.nn.Linear)})
        self._orig: Dict[int, torch.Tensor] = {}
        self._merged_key = None

    @classmethod
    def from_pretrained(cls, model_dir: str, device="cuda", dtype=torch.float16):
        """`<model_dir>/{tokenizer,tokenizer_2,text_encoder,text_encoder_2}` of an SDXL diffusers checkout."""
        from transformers import CLIPTextModel, CLIPTextModelWithProjection, CLIPTokenizer
        toks = [CLIPTokenizer.from_pretrained(os.path.join(model_dir, "tokenizer")),
                CLIPTokenizer.from_pretrained(os.path.join(model_dir, "tokenizer_2"))]
        encs = [CLIPTextModel.from_pretrained(os.path.join(model_dir, "text_encoder"), torch_dtype=dtype),
                CLIPTextModelWithProjection.from_pretrained(os.path.join(model_dir, "text_encoder_2"), torch_dtype=dtype)]
        return cls(toks, encs, device, dtype)

    @property
    def tokenizer(self):
        """The tokenizer AttentionReplace aligns prompts with (inference_lora.py:156 passes pipe.tokenizer)."""
        return self.tokenizers[0]

    # -------------------------------------------------------------------------------------------- text-encoder LoRA
    def _merge(self, adapters):
        """adapters: (names, weights, {adapter: {'te1.<flat module>': (A, B, alpha/r)}}, lora_scale) or None."""
        key = None
        if adapters is not None:
            names, weights, store, scale = adapters
            key = (tuple(names), tuple(weights), float(1.0 if scale is None else scale),
                   tuple(id(store.get(n)) for n in names))
        if key == self._merged_key:
            return
        for mod_id, (mod, w0) in list(self._orig.items()):  # back to the base weights
            mod.weight.data.copy_(w0)
        if key is not None:
            names, weights, store, scale = adapters
            scale = 1.0 if scale is None else float(scale)
            for n, w in zip(names, weights):
                for full, (A, B, s) in (store.get(n) or {}).items():
                    te, _, flat = full.partition(".")
                    # kohya files flatten the module path with '_', diffusers / peft files keep the dots
                    # (`text_encoder.text_model.encoder.layers.0.self_attn.q_proj.lora_A.weight`): one spelling here
                    flat = flat.replace(".", "_")
                    idx = {"te1": 0, "te2": 1}.get(te)
                    if idx is None or idx >= len(self._flat):
                        continue
                    mod = self._flat[idx].get(flat) or self._flat[idx].get("text_model_" + flat)
                    if mod is None:
                        raise KeyError(f"text-encoder LoRA entry {full} matches no Linear of encoder {idx + 1}")
                    if id(mod) not in self._orig:
                        self._orig[id(mod)] = (mod, mod.weight.data.clone())
                    delta = (B.float() @ A.float()) * (s * w * scale)
                    mod.weight.data.add_(delta.to(mod.weight.device, mod.weight.dtype))
        self._merged_key = key

    # -------------------------------------------------------------------------------------------- encode
    @torch.no_grad()
    def __call__(self, prompt: str, lora_scale: Optional[float] = None, adapters=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (prompt_embeds [77, sum hidden], pooled [proj]) in fp32 on the encoder's device."""
        if adapters is not None:
            names, weights, store = adapters
            self._merge((names, weights, store, lora_scale))
        else:
            self._merge(None)
        hs, pooled = [], None
        for tok, enc in zip(self.tokenizers, self.encoders):
            ids = tok(prompt, padding="max_length", max_length=self.max_length, truncation=True, return_tensors="pt").input_ids
            out = enc(ids.to(self.device), output_hidden_states=True)
            pooled = out[0]  # the last encoder's first output: text_embeds of CLIPTextModelWithProjection
            hs.append(out.hidden_states[-2])
        embeds = torch.cat(hs, dim=-1)[0].float()
        pooled = pooled[0].float() if pooled.ndim == 2 else pooled[0, -1].float()
        return embeds, pooled
