import sys, torch
sys.path.insert(0, ".")
from omg_b200 import ops
B, N, heads = 4, 4096, 10
Cc = heads * 64
qkv = (torch.randn(B, N, 3 * Cc, device="cuda")).half()
out = torch.empty(B, N, Cc, device="cuda", dtype=torch.float16)
items = [(b, b, b, b) for b in range(B)]
for _ in range(3):
    ops.attention(qkv, qkv, qkv, out, heads, N, N, items, 0, Cc, 2 * Cc)
torch.cuda.synchronize()
