"""One attention launch of a given shape after two warm-ups (target of `ncu --launch-skip 2 --launch-count 1`)."""
import sys
import torch
sys.path.insert(0, ".")
from omg_b200 import ops
B, N, heads = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
C = heads * 64
qkv = torch.randn(B, N, 3 * C, device="cuda").half()
out = torch.empty(B, N, C, device="cuda", dtype=torch.float16)
items = [(b, b, b, b) for b in range(B)]
for _ in range(3):
    ops.attention(qkv, qkv, qkv, out, heads, N, N, items, 0, C, 2 * C)
torch.cuda.synchronize()
