"""GPU: the diffusers-processor plug-ins (omg_b200.processors) against outputs of the UNMODIFIED reference
processors (tests/golden/ip_attn.pt) and against the reference-pinned controller semantics."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, G)
from make_golden import ShimAttention, ToyTokenizer  # noqa: E402  (pure-torch helpers, no reference import)


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


def test_ip_and_plain_processors_match_reference_outputs():
    from omg_b200.processors import FusedAttnProcessor, FusedIPAttnProcessor
    d = torch.load(os.path.join(G, "ip_attn.pt"))
    attn = ShimAttention(d["dim"], d["ctx_dim"], d["heads"])
    attn.load_state_dict(d["attn"])
    attn = attn.cuda().half()
    proc = FusedIPAttnProcessor(d["dim"], d["ctx_dim"], scale=d["scale"], num_tokens=d["num_tokens"])
    proc.load_state_dict(d["ip"])
    proc = proc.cuda().half()
    with torch.no_grad():
        y = proc(attn, d["x"].cuda().half(), d["ctx"].cuda().half())
    assert y.shape == d["y_ip"].shape
    assert rel(y, d["y_ip"]) < 3e-3 and rel(y, d["y_ip2"]) < 3e-3      # IPAttnProcessor and IPAttnProcessor2_0
    attn2 = ShimAttention(d["dim"], d["dim"], d["heads"])
    attn2.load_state_dict(d["attn_self"])
    attn2 = attn2.cuda().half()
    with torch.no_grad():
        y2 = FusedAttnProcessor()(attn2, d["x"].cuda().half())
    assert rel(y2, d["y_self"]) < 3e-3


@pytest.mark.parametrize("edit", [False, True])
def test_region_processor_applies_controller_edit(edit):
    """Same result as materialising the probabilities and running them through the (reference-pinned) controller."""
    from omg_b200.processors import FusedRegionAttnProcessor
    from omg_b200.prompt_attention import AttentionReplace
    from oracle import p2p as op2p
    torch.manual_seed(0)
    dim, heads, n = 128, 2, 64
    prompts = ["a photo of a man on the beach", "a photo of a dog on the beach"] if edit else ["a b c"] * 2
    cross = {"default_": 0.6, "dog": (0.2, 0.9)} if edit else {"default_": 1.0}
    tok = ToyTokenizer()
    for is_cross, step in ((False, 0), (False, 5), (True, 1), (True, 7)):
        ctx_dim = 96 if is_cross else dim
        attn = ShimAttention(dim, ctx_dim, heads).cuda().half()
        x = torch.randn(4, n, dim).cuda().half()
        ctx = torch.randn(4, 77, ctx_dim).cuda().half() if is_cross else None
        c = AttentionReplace(prompts, 10, dict(cross), 0.3, 8, 8, tokenizer=tok)
        c.num_att_layers = 2
        c.cur_step = step
        with torch.no_grad():
            y = FusedRegionAttnProcessor(controller=c)(attn, x, ctx)
        assert c.cur_att_layer == 1
        # expected: RegionControlNet_AttnProcessor semantics in fp32 with the oracle controller
        oc = op2p.AttentionReplaceOracle(prompts, 10, dict(cross), 0.3, 8, 8, tokenizer=tok)
        oc.num_att_layers = 2
        oc.cur_step = step
        a32 = ShimAttention(dim, ctx_dim, heads)
        a32.load_state_dict({k: v.float().cpu() for k, v in attn.state_dict().items()})
        xs = x.float().cpu()
        src = ctx.float().cpu() if is_cross else xs
        with torch.no_grad():
            q, k, v = a32.to_q(xs), a32.to_k(src), a32.to_v(src)
            p = a32.get_attention_scores(a32.head_to_batch_dim(q), a32.head_to_batch_dim(k))
            p = oc(p, is_cross, "mid")
            ref = a32.to_out[0](a32.batch_to_head_dim(torch.bmm(p, a32.head_to_batch_dim(v))))
        e = rel(y, ref)
        assert e < 3e-3, (is_cross, step, e)
