"""Self-attention kernel timing for same-box A/B of library builds (OMG_B200_LIB=<path> selects the build): 20 launches
back to back per measurement (CUDA events have ~2 us granularity), median of 7, SDXL shapes; --sdpa adds torch's
scaled_dot_product_attention on the same tensors as the library reference of the box.  One JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omg_b200 import ops  # noqa: E402

dev = "cuda"


def timeit(fn, reps=20, iters=7, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    out = {"lib": os.environ.get("OMG_B200_LIB", "default")}
    torch.manual_seed(0)
    for (B, N, heads, tag) in [(4, 4096, 10, "self4096"), (4, 1024, 20, "self1024"), (8, 1024, 20, "self1024_b8"), (8, 4096, 10, "self4096_b8")]:
        C = heads * 64
        qkv = torch.randn(B, N, 3 * C, device=dev).half()
        o = torch.empty(B, N, C, device=dev, dtype=torch.float16)
        items = [(b, b, b, b) for b in range(B)]
        ms = timeit(lambda: ops.attention(qkv, qkv, qkv, o, heads, N, N, items, 0, C, 2 * C))
        fl = 4.0 * B * heads * N * N * 64
        out[tag] = {"us": round(ms * 1e3, 1), "tflops": round(fl / ms / 1e9, 1)}
        if "--sdpa" in sys.argv:
            q, k, v = [t.reshape(B, N, heads, 64).transpose(1, 2) for t in qkv.split(C, dim=-1)]
            ms2 = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
            out[tag]["sdpa_us"] = round(ms2 * 1e3, 1)
            out[tag]["sdpa_tflops"] = round(fl / ms2 / 1e9, 1)
            ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float()).transpose(1, 2).reshape(B, N, C)
            out[tag]["rel_l2"] = float((o.float() - ref).norm() / ref.norm())
    for (B, N, heads, Lk, tag) in [(4, 1024, 20, 77, "cross1024"), (4, 4096, 10, 77, "cross4096"), (4, 1024, 20, 16, "ip1024")]:
        C = heads * 64
        qx, kv = torch.randn(B, N, C, device=dev).half(), torch.randn(B, Lk, 2 * C, device=dev).half()
        o = torch.empty(B, N, C, device=dev, dtype=torch.float16)
        items = [(b, b, b, b) for b in range(B)]
        ms = timeit(lambda: ops.attention(qx, kv, kv, o, heads, N, Lk, items, 0, 0, C))
        out[tag] = {"us": round(ms * 1e3, 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
