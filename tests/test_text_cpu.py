"""ClipPromptEncoder on CPU with random-init toy CLIP towers and a stand-in tokenizer: output layout
(penultimate hidden states of both encoders concatenated, pooled = projected output of the second) and the
text-encoder LoRA merge / un-merge."""
import types

import pytest
import torch

transformers = pytest.importorskip("transformers")


class ToyTokenizer:
    def __call__(self, text, padding=None, max_length=77, truncation=True, return_tensors="pt"):
        ids = [1] + [3 + (sum(map(ord, w)) % 90) for w in text.split()][: max_length - 2] + [2]
        ids = ids + [0] * (max_length - len(ids))
        return types.SimpleNamespace(input_ids=torch.tensor([ids]))


def _towers():
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection
    torch.manual_seed(0)
    c1 = CLIPTextConfig(vocab_size=100, hidden_size=32, intermediate_size=64, num_hidden_layers=3, num_attention_heads=4,
                        max_position_embeddings=77, projection_dim=32)
    c2 = CLIPTextConfig(vocab_size=100, hidden_size=48, intermediate_size=96, num_hidden_layers=3, num_attention_heads=4,
                        max_position_embeddings=77, projection_dim=40)
    return CLIPTextModel(c1), CLIPTextModelWithProjection(c2)


def test_layout_and_lora_merge():
    from omg_b200.text import ClipPromptEncoder
    e1, e2 = _towers()
    enc = ClipPromptEncoder([ToyTokenizer(), ToyTokenizer()], [e1, e2], device="cpu", dtype=torch.float32)
    emb, pooled = enc("a man and a woman")
    assert emb.shape == (77, 32 + 48) and pooled.shape == (40,)
    ids = ToyTokenizer()("a man and a woman").input_ids
    o1 = e1(ids, output_hidden_states=True)
    o2 = e2(ids, output_hidden_states=True)
    assert torch.allclose(emb, torch.cat([o1.hidden_states[-2], o2.hidden_states[-2]], -1)[0])
    assert torch.allclose(pooled, o2.text_embeds[0])

    # LoRA on one projection of each tower, kohya-flattened names as omg_b200.checkpoints returns them
    g = torch.Generator().manual_seed(1)
    A1, B1 = torch.randn(2, 32, generator=g) * 0.3, torch.randn(32, 2, generator=g) * 0.3
    A2, B2 = torch.randn(2, 48, generator=g) * 0.3, torch.randn(48, 2, generator=g) * 0.3
    store = {"c0": {"te1.text_model_encoder_layers_0_self_attn_q_proj": (A1, B1, 0.5),
                    "te2.text_model_encoder_layers_1_mlp_fc2": (A2[:, :48], torch.randn(48, 2, generator=g) * 0.3, 1.0)}}
    store["c0"]["te2.text_model_encoder_layers_1_mlp_fc2"] = (torch.randn(2, 96, generator=g) * 0.3, B2, 1.0)
    w0 = e1.text_model.encoder.layers[0].self_attn.q_proj.weight.detach().clone()
    emb_l, _ = enc("a man and a woman", 0.8, adapters=(("c0",), (0.7,), store))
    w1 = e1.text_model.encoder.layers[0].self_attn.q_proj.weight.detach()
    assert torch.allclose(w1, w0 + (B1 @ A1) * (0.5 * 0.7 * 0.8), atol=1e-6)
    assert not torch.allclose(emb_l, emb)
    emb_back, _ = enc("a man and a woman")  # no adapters: base weights restored exactly
    assert torch.equal(e1.text_model.encoder.layers[0].self_attn.q_proj.weight.detach(), w0)
    assert torch.allclose(emb_back, emb)
    with pytest.raises(KeyError):
        enc("x", 1.0, adapters=(("c1",), (1.0,), {"c1": {"te1.no_such_module": (A1, B1, 1.0)}}))


def test_concept_models_pass_active_adapters():
    from omg_b200.pipelines import ConceptModels
    seen = {}

    class Enc:
        supports_adapters = True

        def __call__(self, prompt, lora_scale=None, adapters=None):
            seen["adapters"] = adapters
            return torch.zeros(77, 8), torch.zeros(4)

    cm = ConceptModels.__new__(ConceptModels)
    cm.prompt_encoder, cm._loras, cm._active = Enc(), {"a": {}, "style": {}}, ((), ())
    cm.text_encoder_loras = {"a": {"te1.x": None}}
    cm.set_adapters(["a", "style"], adapter_weights=[0.7, 0.5])
    cm.encode_prompt("p", "n", lora_scale=0.8)
    assert seen["adapters"][0] == ("a", "style") and seen["adapters"][1] == (0.7, 0.5)


def test_peft_layout_text_encoder_lora_merges():
    """diffusers / peft files spell the module path with dots (`text_encoder.text_model.encoder.layers.0.self_attn.
    q_proj.lora_A.weight`); omg_b200.checkpoints returns them as `te1.text_model.encoder...` and the merge must find
    the same Linear as for the kohya spelling."""
    from omg_b200.checkpoints import convert_lora_state_dict
    from omg_b200.text import ClipPromptEncoder
    e1, e2 = _towers()
    enc = ClipPromptEncoder([ToyTokenizer(), ToyTokenizer()], [e1, e2], device="cpu", dtype=torch.float32)
    g = torch.Generator().manual_seed(2)
    A1, B1 = torch.randn(2, 32, generator=g) * 0.3, torch.randn(32, 2, generator=g) * 0.3
    A2, B2 = torch.randn(2, 48, generator=g) * 0.3, torch.randn(48, 2, generator=g) * 0.3
    sd = {"text_encoder.text_model.encoder.layers.0.self_attn.q_proj.lora_A.weight": A1,
          "text_encoder.text_model.encoder.layers.0.self_attn.q_proj.lora_B.weight": B1,
          "text_encoder_2.text_model.encoder.layers.1.self_attn.v_proj.lora_A.weight": A2,
          "text_encoder_2.text_model.encoder.layers.1.self_attn.v_proj.lora_B.weight": B2}
    _unet, te, _skipped = convert_lora_state_dict(sd)
    assert set(te) == {"te1.text_model.encoder.layers.0.self_attn.q_proj", "te2.text_model.encoder.layers.1.self_attn.v_proj"}
    w1 = e1.text_model.encoder.layers[0].self_attn.q_proj.weight.detach().clone()
    w2 = e2.text_model.encoder.layers[1].self_attn.v_proj.weight.detach().clone()
    enc("a b", 0.8, adapters=(("style",), (1.0,), {"style": te}))
    s1, s2 = te["te1.text_model.encoder.layers.0.self_attn.q_proj"][2], te["te2.text_model.encoder.layers.1.self_attn.v_proj"][2]
    assert torch.allclose(e1.text_model.encoder.layers[0].self_attn.q_proj.weight.detach(), w1 + (B1 @ A1) * (s1 * 0.8), atol=1e-6)
    assert torch.allclose(e2.text_model.encoder.layers[1].self_attn.v_proj.weight.detach(), w2 + (B2 @ A2) * (s2 * 0.8), atol=1e-6)
