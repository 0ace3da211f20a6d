"""Prompt conditioning front-end: SDXL's two CLIP text encoders behind the `prompt_encoder` callable the
pipelines take.  Once per call and outside the denoising loop.  On a CUDA device the towers run on this repo's kernels
(`PackedClipText`: omg_gemm with bias / quick-gelu / erf-gelu epilogues, omg_layernorm, omg_attention with the causal
mask); transformers supplies the checkpoint loading, the tokenizers and - on the CPU, for tests - the reference modules.

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


class PackedClipText:
    """One CLIP text tower (transformers CLIPTextModel / CLIPTextModelWithProjection weights) executed on the C-ABI
    kernels: pre-LN transformer, causal self-attention over the 77 tokens (head_dim 64 in both SDXL towers),
    quick-gelu (CLIP-L) or erf-gelu (OpenCLIP bigG) MLP.  Returns what encode_prompt reads from the module
    (lora_pipeline.py:315-347 -> diffusers encode_prompt [3P]): hidden_states[-2] and, for the projected tower, text_embeds."""

    def __init__(self, module: torch.nn.Module, device="cuda"):
        from . import _lib as L
        cfg = module.config
        tm = module.text_model
        self.dev = torch.device(device)
        self.heads, self.C, self.layers = cfg.num_attention_heads, cfg.hidden_size, cfg.num_hidden_layers
        if self.C // self.heads != 64:
            raise ValueError("PackedClipText needs head_dim 64 (both SDXL text towers have it)")
        self.eps = float(cfg.layer_norm_eps)
        self.eos = getattr(cfg, "eos_token_id", 2)
        act = cfg.hidden_act
        if act not in ("quick_gelu", "gelu"):
            raise ValueError(f"unsupported CLIP activation {act}")
        self.epi = L.EPI_QUICK_GELU if act == "quick_gelu" else L.EPI_GELU
        h = lambda t: t.detach().to(self.dev, torch.float16).contiguous()  # noqa: E731
        self.tok_emb, self.pos_emb = h(tm.embeddings.token_embedding.weight), h(tm.embeddings.position_embedding.weight)
        self.blocks = []
        for lyr in tm.encoder.layers:
            a = lyr.self_attn
            self.blocks.append({
                "ln1": (h(lyr.layer_norm1.weight), h(lyr.layer_norm1.bias)), "ln2": (h(lyr.layer_norm2.weight), h(lyr.layer_norm2.bias)),
                "qkv_w": h(torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0)),
                "qkv_b": h(torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0)),
                "out_w": h(a.out_proj.weight), "out_b": h(a.out_proj.bias),
                "fc1_w": h(lyr.mlp.fc1.weight), "fc1_b": h(lyr.mlp.fc1.bias),
                "fc2_w": h(lyr.mlp.fc2.weight), "fc2_b": h(lyr.mlp.fc2.bias)})
        self.final_ln = (h(tm.final_layer_norm.weight), h(tm.final_layer_norm.bias))
        proj = getattr(module, "text_projection", None)
        self.proj = None if proj is None else h(proj.weight)

    @torch.no_grad()
    def __call__(self, input_ids: torch.Tensor):
        """input_ids (B, T<=128) -> (hidden_states[-2] (B, T, C) fp16, text_embeds (B, P) fp16 or None)."""
        from . import ops
        ids = input_ids.to(self.dev)
        B, T = ids.shape
        C, M = self.C, B * T
        x = (self.tok_emb[ids] + self.pos_emb[:T][None]).reshape(M, C).contiguous()
        items = [(b, b, b, b) for b in range(B)]
        penult = None
        for li, blk in enumerate(self.blocks):
            if li == self.layers - 1:
                penult = x.clone()
            n1 = ops.layernorm(x, *blk["ln1"], eps=self.eps)
            qkv = ops.linear(n1, blk["qkv_w"], bias=blk["qkv_b"]).view(B, T, 3 * C)
            o = torch.empty(B, T, C, dtype=torch.float16, device=self.dev)
            ops.attention(qkv, qkv, qkv, o, self.heads, T, T, items, 0, C, 2 * C, scale=0.125, causal=True)
            x = ops.linear(o.view(M, C), blk["out_w"], bias=blk["out_b"], residual=x)
            n2 = ops.layernorm(x, *blk["ln2"], eps=self.eps)
            f = ops.linear(n2, blk["fc1_w"], bias=blk["fc1_b"], epilogue=self.epi)
            x = ops.linear(f, blk["fc2_w"], bias=blk["fc2_b"], residual=x)
        pooled = None
        if self.proj is not None:
            last = ops.layernorm(x, *self.final_ln, eps=self.eps).view(B, T, C)
            if self.eos == 2:   # transformers: argmax of the ids (the EOS token has the largest id in CLIP's vocabulary)
                pos = ids.argmax(dim=-1)
            else:
                pos = (ids == self.eos).int().argmax(dim=-1)
            eos_h = last[torch.arange(B, device=self.dev), pos]
            pad = torch.zeros(8, C, dtype=torch.float16, device=self.dev)   # the GEMM's pixel grid wants >= 1 full row box
            pad[:B] = eos_h
            pooled = ops.linear(pad, self.proj)[:B]
        return penult.view(B, T, C), pooled


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
        # CUDA: the towers run on the C-ABI kernels (packed per merged-adapter key); CPU: transformers' modules (tests)
        self.use_kernels = self.device.type == "cuda" and dtype == torch.float16
        self._packed = None
        self._packed_key = "unset"

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
        if self.use_kernels:
            if self._packed is None or self._packed_key != self._merged_key:  # weights changed (text-encoder LoRA merge)
                self._packed = [PackedClipText(e, self.device) for e in self.encoders]
                self._packed_key = self._merged_key
            for tok, enc in zip(self.tokenizers, self._packed):
                ids = tok(prompt, padding="max_length", max_length=self.max_length, truncation=True, return_tensors="pt").input_ids
                h, pl = enc(ids)
                hs.append(h)
                pooled = pl if pl is not None else h[:, -1]
            return torch.cat(hs, dim=-1)[0].float(), pooled[0].float()
        for tok, enc in zip(self.tokenizers, self.encoders):
            ids = tok(prompt, padding="max_length", max_length=self.max_length, truncation=True, return_tensors="pt").input_ids
            out = enc(ids.to(self.device), output_hidden_states=True)
            pooled = out[0]  # the last encoder's first output: text_embeds of CLIPTextModelWithProjection
            hs.append(out.hidden_states[-2])
        embeds = torch.cat(hs, dim=-1)[0].float()
        pooled = pooled[0].float() if pooled.ndim == 2 else pooled[0, -1].float()
        return embeds, pooled
