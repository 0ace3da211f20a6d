"""Where does the CUDA path's distance to the fp32 oracle come from?  One main-UNet forward (B=2, SDXL widths,
128x128 latents, plain attention) against the fp32 oracle on the GPU, module by module (every ResBlock / Transformer2D
output is compared), under the executor's switches (LayerNorm fold on / off).  Prints one JSON line per variant.

  python scripts/error_budget.py [--latent 128]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omg_b200 import synthetic  # noqa: E402
from omg_b200.config import UNetConfig  # noqa: E402
from omg_b200.unet import PackedUNet, UNetRunner  # noqa: E402
from oracle import unet as ou  # noqa: E402
from oracle.scheduler import EulerDiscrete  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = "cuda"


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--latent", type=int, default=128)
    ap.add_argument("--batch", type=int, default=2)
    a = ap.parse_args()
    cfg = UNetConfig.sdxl()
    H = W = a.latent
    B = a.batch
    sd = synthetic.make_state_dict(cfg, seed=0, device=dev, dtype=torch.float16)
    g = torch.Generator().manual_seed(0)
    sched = EulerDiscrete()
    ts = sched.set_timesteps(30)
    i = 10
    x = (torch.randn(B, 4, H, W, generator=g) * float(sched.sigmas[i]) * sched.scale_model_input(torch.ones(()), i)).half()
    ctx = torch.randn(B, 77, 2048, generator=g).half()
    pooled = torch.randn(B, 1280, generator=g).half()
    tid = torch.tensor([[H * 8, W * 8, 0, 0, H * 8, W * 8]], dtype=torch.float32).repeat(B, 1)
    # ---- fp32 oracle with every block output recorded
    rec = {}
    orig_res, orig_tr = ou.resnet, ou.transformer2d

    def res_hook(c, name, xx, emb):
        o = orig_res(c, name, xx, emb)
        rec[name] = o
        return o

    def tr_hook(c, name, xx, cc, n):
        o = orig_tr(c, name, xx, cc, n)
        rec[name] = o
        return o

    ou.resnet, ou.transformer2d = res_hook, tr_hook
    sd32 = {k: v.float() for k, v in sd.items()}
    with torch.no_grad():
        ref = ou.unet_forward(ou.Ctx(sd32, ou.UNetConfig()), x.float().to(dev), float(ts[i]), ctx.float().to(dev),
                              pooled.float().to(dev), tid.to(dev))
    ref_rec = dict(rec)
    rec.clear()
    sdh = {k: v.half() for k, v in sd.items()}
    with torch.no_grad():
        eager = ou.unet_forward(ou.Ctx(sdh, ou.UNetConfig()), x.to(dev), float(ts[i]), ctx.to(dev), pooled.to(dev), tid.to(dev))
    eager_rec = {k: v.float() for k, v in rec.items()}
    ou.resnet, ou.transformer2d = orig_res, orig_tr
    del sd32, sdh
    print(json.dumps({"variant": "fp16_eager", "noise_rel_l2": rel(eager.float(), ref),
                      "blocks": {k: round(rel(eager_rec[k], ref_rec[k]), 6) for k in ref_rec}}), flush=True)
    unet = PackedUNet(cfg, sd, device=dev)
    for variant, env in (("default", {}), ("trunk_fp32_twins", {"OMG_TRUNK_F32": "1"}), ("ln_fold_off", {"OMG_LN_FOLD": "0"})):
        for k, v in env.items():
            os.environ[k] = v
        r = UNetRunner(unet, B, H, W, use_graphs=False)
        r.set_conditioning([float(ts[i])], ctx, pooled, tid)
        r.sample_in.zero_()
        r.sample_in[..., :4] = x.permute(0, 2, 3, 1).to(dev)
        out = r.forward(0)
        torch.cuda.synchronize()
        blocks = {}
        for name in ref_rec:
            t = r.ws.get(name + ".out")
            if t is not None:
                blocks[name] = round(rel(t.float().permute(0, 3, 1, 2), ref_rec[name]), 6)
        print(json.dumps({"variant": variant, "noise_rel_l2": rel(out[..., :4].permute(0, 3, 1, 2).float(), ref),
                          "blocks": blocks}), flush=True)
        for k in env:
            os.environ.pop(k)
        del r


if __name__ == "__main__":
    main()
