"""GEMM / conv shapes of the forward, 30 launches back to back per timing (weights differ per launch like in the forward:
a pool of 6 weight matrices is cycled so that the weights come from DRAM, activations stay L2-resident)."""
import json, os, sys, torch
sys.path.insert(0, ".")
from omg_b200 import _lib as L, ops
def rnd(*s, scale=1.0): return (torch.randn(*s, device="cuda") * scale).half()
res = {"env": {k: v for k, v in os.environ.items() if k.startswith("OMG_GEMM")}}
def timeit(fns):
    for f in fns: f()
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            for f in fns: f()
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / (5 * len(fns)))
    ts.sort(); return round(ts[2] * 1e3, 2)
for (M, N, K, epi, tag) in [(4096, 1280, 1280, 0, "out1280"), (4096, 1280, 5120, 0, "ff2_1280"), (4096, 3840, 1280, 0, "qkv1280"),
                            (4096, 10240, 1280, 1, "ff1_geglu_1280"), (16384, 640, 640, 0, "out640"), (16384, 5120, 640, 1, "ff1_geglu_640")]:
    x, r = rnd(M, K), rnd(M, N if not epi else N // 2)
    ws = [rnd(N, K, scale=K ** -0.5) for _ in range(6)]
    out = torch.empty(M, N // 2 if epi else N, device="cuda", dtype=torch.float16)
    fns = [(lambda w=w: ops.linear(x, w, residual=None if epi else r, out=out, epilogue=L.EPI_GEGLU if epi else L.EPI_NONE)) for w in ws]
    us = timeit(fns); res[tag] = {"us": us, "tflops": round(2.0 * M * N * K / us / 1e6, 1)}
for (B, H, Cin, N, tag) in [(4, 128, 320, 320, "conv320@128"), (4, 32, 1280, 1280, "conv1280@32")]:
    x = rnd(B, H, H, Cin); ws = [ops.pack_conv3x3_weight(rnd(N, Cin, 3, 3, scale=(9 * Cin) ** -0.5)) for _ in range(6)]
    out = torch.empty(B, H, H, N, device="cuda", dtype=torch.float16)
    fns = [(lambda w=w: ops.conv3x3(x, w, out=out)) for w in ws]
    us = timeit(fns); res[tag] = {"us": us, "tflops": round(2.0 * B * H * H * N * 9 * Cin / us / 1e6, 1)}
print(json.dumps(res))
