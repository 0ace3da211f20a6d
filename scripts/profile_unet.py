"""One eager main-UNet forward (B=4, 128x128 latents, P2P variant) for ncu: `--warm` iterations first."""
import sys
import torch
sys.path.insert(0, ".")
from omg_b200 import factory
from omg_b200.config import UNetConfig

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
wl = factory.build_lora_workload(UNetConfig.sdxl(), 1024, 2, 32, 30, 7.5, use_graphs=False)
pipe = wl.pipe
kw = dict(wl.call_kwargs)
pe, ne, pp, np_ = pipe.encode_prompt(kw["prompt"][0], kw["negative_prompt"], 0.8)
ctx4 = torch.cat([ne, pe]); pooled4 = torch.cat([np_, pp])
tid = torch.tensor([[1024, 1024, 0, 0, 1024, 1024]], dtype=torch.float32)
main = pipe._runner("main", pipe.unet, 4, 128, 128)
pipe._update_p2p_context(main, wl.controller, ctx4, first=True)
main.set_conditioning(pipe.scheduler.set_timesteps(30), ctx4, pooled4, tid.repeat(4, 1), extra_ctx=pipe._p2p_rows)
main.sample_in[..., :4] = torch.randn(4, 128, 128, 4, device="cuda").half()
variant, _ = pipe._p2p_variant(main, wl.controller, False)
torch.cuda.synchronize()
for _ in range(n):
    torch.cuda.nvtx.range_push("unet_forward")
    main.forward(0, variant)
    torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print("done")
