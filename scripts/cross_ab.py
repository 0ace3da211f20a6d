import json, os, sys, torch
sys.path.insert(0, ".")
from omg_b200 import ops
def rnd(*s): return torch.randn(*s, device="cuda").half()
for (B, N, heads, L) in [(4, 1024, 20, 77), (4, 4096, 10, 77), (8, 1024, 20, 77)]:
    Cc = heads * 64; qx, kv = rnd(B, N, Cc), rnd(B, L, 2 * Cc); out = torch.empty(B, N, Cc, device="cuda", dtype=torch.float16)
    items = [(b, b, b, b) for b in range(B)]
    f = lambda: ops.attention(qx, kv, kv, out, heads, N, L, items, 0, 0, Cc)
    for _ in range(5): f()
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): f()
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 50)
    ts.sort()
    print(json.dumps({"hpc": os.environ.get("OMG_ATTN_HPC", "auto"), "shape": [B, N, heads, L], "us": round(ts[2] * 1e3, 2)}))
