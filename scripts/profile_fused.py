"""Eager forwards for ncu: main UNet (B=4) then the grouped fusion-step forward (B=8: main + 2 LoRA concepts)."""
import sys
import torch
sys.path.insert(0, ".")
from omg_b200 import factory
from omg_b200.config import UNetConfig
from omg_b200.unet import RowGroup

which = sys.argv[1] if len(sys.argv) > 1 else "main"
wl = factory.build_lora_workload(UNetConfig.sdxl(), 1024, 2, 32, 30, 7.5, use_graphs=False)
pipe = wl.pipe
kw = dict(wl.call_kwargs)
pe, ne, pp, np_ = pipe.encode_prompt(kw["prompt"][0], kw["negative_prompt"], 0.8)
ctx4 = torch.cat([ne, pe]); pooled4 = torch.cat([np_, pp])
tid = torch.tensor([[1024, 1024, 0, 0, 1024, 1024]], dtype=torch.float32)
ts = pipe.scheduler.set_timesteps(30)
cm = wl.concept_models
if which == "main":
    r = pipe._runner("main", pipe.unet, 4, 128, 128, groups=[RowGroup(0, 4, None, False)])
    pipe._update_p2p_context([], wl.controller, ctx4, first=True)
    r.set_conditioning(ts, ctx4, pooled4, tid.repeat(4, 1), extra_ctx=pipe._p2p_rows)
    variant, _ = pipe._p2p_variant(r, wl.controller, False)
else:
    keys = []
    ctx_list = [(ctx4, None, False)]
    pooled = [pooled4]
    groups = [RowGroup(0, 4, None, False)]
    for k, (rp, rn) in enumerate(kw["prompt"][1]):
        cm.set_adapters(kw["lora_list"][k])
        e, n_, p_, np2 = cm.encode_prompt(rp, negative_prompt=rn)
        key = cm.active_lora_key(0.8)
        ctx_list.append((torch.cat([n_, e]), key, False))
        pooled.append(torch.cat([np2, p_]))
        groups.append(RowGroup(4 + 2 * k, 6 + 2 * k, key, False))
    r = pipe._runner("fused", pipe.unet, 8, 128, 128, groups=groups, tag_extra="prof")
    pipe._update_p2p_context([], wl.controller, ctx4, first=True)
    r.set_conditioning(ts, ctx_list, torch.cat(pooled), tid.repeat(8, 1), extra_ctx=pipe._p2p_rows)
    variant, _ = pipe._p2p_variant(r, wl.controller, False)
    variant["self_items"] = variant["self_items"][:4] + [(4 + i, 4 + i, 4 + i, 4 + i) for i in range(4)]
    variant["cross_items"] = [variant["cross_items"][0][:4] + [(4 + i, 4 + i, 6 + i, 6 + i) for i in range(4)]]
    variant["ip_items"] = []
r.sample_in[..., :4] = torch.randn(r.B, 128, 128, 4, device="cuda").half()
torch.cuda.synchronize()
r.forward(0, variant)     # warm-up (allocations, kernel attributes)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("unet_forward")
r.forward(0, variant)
torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print("done")
