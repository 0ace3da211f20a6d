"""Where one image's wall time goes outside the UNet graphs: per-call hoisted work (set_conditioning: time-embedding
tables, cross-attention K/V, merged LoRA planes lookup), P2P context updates, graph replays, fused step kernel.
Synchronising timers, so the total is slower than bench.py's number; the split is what matters."""
import collections
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from omg_b200 import factory, ops, pipelines, unet as U  # noqa: E402
from omg_b200.config import UNetConfig  # noqa: E402

acc = collections.defaultdict(float)
cnt = collections.Counter()


def wrap(obj, name, label):
    fn = getattr(obj, name)

    def timed(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize()
        acc[label] += time.perf_counter() - t0
        cnt[label] += 1
        return r
    setattr(obj, name, timed)


wl = factory.build_lora_workload(UNetConfig.sdxl(), 1024, 2, 32, 30, 7.5)
for _ in range(2):
    factory.run_two_stage(wl)
wrap(U.UNetRunner, "set_conditioning", "set_conditioning")
_fwd = U.UNetRunner.forward


def fwd_timed(self, *a, **k):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = _fwd(self, *a, **k)
    torch.cuda.synchronize()
    lab = f"unet_forward(graph replay) B={self.B}"
    acc[lab] += time.perf_counter() - t0
    cnt[lab] += 1
    return r


U.UNetRunner.forward = fwd_timed
wrap(pipelines._BasePipeline, "_update_p2p_context", "p2p_context")
wrap(ops, "fuse_step", "fuse_step")
wrap(pipelines._BasePipeline, "encode_prompt", "encode_prompt(synthetic)")
torch.cuda.synchronize()
t0 = time.perf_counter()
factory.run_two_stage(wl)
torch.cuda.synchronize()
total = time.perf_counter() - t0
out = {k: {"ms": round(v * 1e3, 2), "calls": cnt[k]} for k, v in acc.items()}
out["total_ms_with_sync_timers"] = round(total * 1e3, 1)
out["unaccounted_ms"] = round((total - sum(acc.values())) * 1e3, 1)
print(json.dumps(out))
