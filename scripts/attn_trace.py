"""Timeline of the persistent self-attention kernel from a diagnostic build (-DOMG_ATT_TRACE, see attn_tc.cu): clock64 stamps of
one softmax warp and of the MMA issuer of CTA 0 (and CTA 148) per KV block.  Prints mean clocks per phase over the steady state.

  nvcc ... -DOMG_ATT_TRACE -c omg_b200/csrc/attn_tc.cu ; link as build/ab/libomg_trace.so
  OMG_B200_LIB=$PWD/build/ab/libomg_trace.so python scripts/attn_trace.py [B N heads]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omg_b200 import ops  # noqa: E402

B, N, heads = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (4, 4096, 10)
C = heads * 64
qkv = torch.randn(B, N, 3 * C, device="cuda").half()
out = torch.empty(B, N, C, device="cuda", dtype=torch.float16)
items = [(b, b, b, b) for b in range(B)]
trace = torch.zeros(2, 2, 512, 8, dtype=torch.int64, device="cuda")
os.environ["OMG_ATT_TRACE_PTR"] = str(trace.data_ptr())
for _ in range(3):
    ops.attention(qkv, qkv, qkv, out, heads, N, N, items, 0, C, 2 * C)
torch.cuda.synchronize()
trace.zero_()
ops.attention(qkv, qkv, qkv, out, heads, N, N, items, 0, C, 2 * C)
torch.cuda.synchronize()
t = trace.cpu()
nkv = N // 64
res = {"shape": [B, N, heads], "blocks_per_tile": nkv}
for cta in range(2):
    s, m = t[cta, 0].double(), t[cta, 1].double()
    n = int((s[:, 0] > 0).sum())
    if n < 40:
        continue
    lo, hi = 8, min(n - 2, 3 * nkv if nkv <= 64 else 200)
    sel = [g for g in range(lo, hi) if g % nkv not in (0, nkv - 1)]   # steady state: not the first / last block of a tile
    idx = torch.tensor(sel)
    sm = {"wait_s_full": s[idx, 1] - s[idx, 0], "tmem_load": s[idx, 2] - s[idx, 1], "row_max_and_vote": s[idx, 3] - s[idx, 2],
          "wait_p_empty": s[idx, 4] - s[idx, 3], "exponentials_pack": s[idx, 5] - s[idx, 4], "tmem_store_wait": s[idx, 6] - s[idx, 5],
          "fence_arrive": s[idx, 7] - s[idx, 6], "period": s[idx + 1, 0] - s[idx, 0]}
    mm = {"wait_p_full": m[idx, 1] - m[idx, 0], "issue_qk(g+2)": m[idx, 2] - m[idx, 1], "issue_pv(g)+commits": m[idx, 3] - m[idx, 2],
          "period": m[idx + 1, 0] - m[idx, 0],
          "softmax_arrive_to_mma_wake": m[idx, 1] - s[idx, 7],
          "qk_issued_to_softmax_sees_S(g+2)": s[idx + 2, 1] - m[idx, 2]}
    res[f"cta{0 if cta == 0 else 148}"] = {"softmax_warp": {k: round(float(v.mean()), 1) for k, v in sm.items()},
                                          "mma_issuer": {k: round(float(v.mean()), 1) for k, v in mm.items()}, "blocks": len(sel)}
print(json.dumps(res))
