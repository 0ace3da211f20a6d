"""VAE decode timing at the bench workload's size: 2 images, 128x128 latents -> 1024x1024, SDXL widths, synthetic
weights.  Prints one JSON line (ms per 2-image decode, algorithmic TFLOP/s, peak memory)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from omg_b200 import synthetic  # noqa: E402
from omg_b200.vae import PackedVaeDecoder, VaeConfig, vae_decoder_flops  # noqa: E402

cfg = VaeConfig.sdxl()
dec = PackedVaeDecoder(synthetic.make_vae_state_dict(cfg, 0), cfg)
B, h, w = 2, 128, 128
lat = (torch.randn(B, 4, h, w, generator=torch.Generator().manual_seed(0)) * 0.4).half().cuda()
for _ in range(2):
    img = dec.decode(lat)
torch.cuda.synchronize()
torch.cuda.reset_peak_memory_stats()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    img = dec.decode(lat)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
ms = ts[len(ts) // 2]
fl = B * vae_decoder_flops(cfg, h, w)
print(json.dumps({"name": "vae_decode_2x1024", "ms": round(ms, 3), "tflops": round(fl / ms / 1e9, 1),
                  "algorithmic_tflop": round(fl / 1e12, 2), "finite": bool(torch.isfinite(img).all()),
                  "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}))
