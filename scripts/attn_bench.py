import json, sys, torch
sys.path.insert(0, ".")
from omg_b200 import ops
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
def rnd(*s): return torch.randn(*s, device=dev).half()
for (B, N, heads, tag) in [(4, 4096, 10, "self4096"), (4, 1024, 20, "self1024"), (8, 1024, 20, "self1024_b8")]:
    Cc = heads * 64; qkv = rnd(B, N, 3 * Cc); out = torch.empty(B, N, Cc, device=dev, dtype=torch.float16)
    items = [(b, b, b, b) for b in range(B)]
    ms = timeit(lambda: ops.attention(qkv, qkv, qkv, out, heads, N, N, items, 0, Cc, 2 * Cc))
    print(json.dumps({"name": "attn_" + tag, "ms": round(ms, 4), "tflops": round(4.0 * B * heads * N * N * 64 / ms / 1e9, 1)}))
for (B, N, heads, L, tag) in [(4, 4096, 10, 77, "cross4096"), (4, 1024, 20, 77, "cross1024"), (4, 1024, 20, 16, "ip1024")]:
    Cc = heads * 64; qx, kv = rnd(B, N, Cc), rnd(B, L, 2 * Cc); out = torch.empty(B, N, Cc, device=dev, dtype=torch.float16)
    items = [(b, b, b, b) for b in range(B)]
    ms = timeit(lambda: ops.attention(qx, kv, kv, out, heads, N, L, items, 0, 0, Cc))
    print(json.dumps({"name": "attn_" + tag, "ms": round(ms, 4)}))
