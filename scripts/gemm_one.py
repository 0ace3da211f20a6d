"""Three launches of one Linear shape (target of `ncu --launch-skip 2 --launch-count 1`): M N K [cta_pair] [block_n]."""
import sys
import torch
sys.path.insert(0, ".")
from omg_b200 import ops
M, N, K = (int(v) for v in sys.argv[1:4])
cp = int(sys.argv[4]) if len(sys.argv) > 4 else 0
bn = int(sys.argv[5]) if len(sys.argv) > 5 else 0
x = torch.randn(M, K, device="cuda").half()
w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
out = torch.empty(M, N, device="cuda", dtype=torch.float16)
for _ in range(3):
    ops.linear(x, w, out=out, block_n=bn, cta_pair=cp)
torch.cuda.synchronize()
