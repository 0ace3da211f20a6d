"""EfficientViT-SAM xl1 image encoder (the model inference_lora.py:176 loads) at 1024x1024 on the kernels: ms per image,
with random weights of the reference's shapes (tests/golden/sam_xl1_shapes.json, written by make_golden.py from the
unmodified module; there is no checkpoint offline).  FLOPs are counted from the dense convolutions' shapes."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omg_b200.sam_encoder import PackedSamImageEncoder  # noqa: E402

shapes = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "sam_xl1_shapes.json")))
g = torch.Generator().manual_seed(0)
sd = {}
for k, shp in shapes.items():
    if k.endswith("running_var"):
        t = torch.rand(shp, generator=g) + 0.5
    elif k.endswith("running_mean") or k.endswith(".bias"):
        t = torch.randn(shp, generator=g) * 0.05
    elif k.endswith("norm.weight"):
        t = 1.0 + 0.1 * torch.randn(shp, generator=g)
    else:
        fan_in = 1
        for d in shp[1:]:
            fan_in *= d
        t = torch.randn(shp, generator=g) * fan_in ** -0.5
    sd[k] = t
enc = PackedSamImageEncoder(sd, device="cuda")
x = torch.randn(1, 3, 1024, 1024, generator=g)
for _ in range(3):
    y = enc(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    y = enc(x)
e1.record()
torch.cuda.synchronize()
ms_graph = e0.elapsed_time(e1) / 5
enc.use_graph = False
for _ in range(2):
    enc(x)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    enc(x)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"model": "EfficientViT-SAM xl1 image encoder, 1024x1024, batch 1 (host image in, H2D inside the timing)",
                  "ms": round(ms_graph, 2), "ms_eager_launches": round(e0.elapsed_time(e1) / 5, 2),
                  "out": list(y.shape), "finite": bool(torch.isfinite(y).all()),
                  "params_M": round(sum(v.numel() for k, v in sd.items() if k.endswith("weight")) / 1e6, 1)}))
