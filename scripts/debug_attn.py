import sys, torch
sys.path.insert(0, ".")
from omg_b200 import ops
torch.manual_seed(0)
def run(N, Lk, heads=1, B=1, q_zero=True):
    C = heads*64
    q = torch.zeros(B, N, C, device="cuda", dtype=torch.float16) if q_zero else torch.randn(B,N,C,device="cuda").half()
    k = torch.randn(B, Lk, C, device="cuda").half()
    v = torch.zeros(B, Lk, C, device="cuda", dtype=torch.float16)
    nb = (Lk+63)//64
    for j in range(nb):
        v[:, j*64:(j+1)*64, j % 64] = 1.0      # channel j counts keys of block j
    v[..., 63] = torch.arange(Lk, device="cuda").half()[None, :] / Lk
    out = torch.empty(B, N, C, device="cuda", dtype=torch.float16)
    ops.attention(q, k, v, out, heads, N, Lk, [(b,b,b,b) for b in range(B)])
    qh = q.float(); p = torch.softmax(qh @ k.float().transpose(1,2) * 0.125, -1); ref = p @ v.float()
    torch.cuda.synchronize()
    err = (out.float()-ref).abs()
    print(f"N={N} Lk={Lk} q_zero={q_zero} max err {err.max().item():.4f}")
    print(" out row0 block-mass:", [round(x,3) for x in out[0,0,:nb].float().tolist()], " ref:", [round(x,3) for x in ref[0,0,:nb].tolist()])
    print(" out row200 block-mass:", [round(x,3) for x in out[0,min(200,N-1),:nb].float().tolist()])
    rows = err.amax(dim=2)[0]
    print(" err by 32-row block:", [round(rows[i:i+32].max().item(),3) for i in range(0, N, 32)])
for (N, Lk) in [(256,64),(256,128),(256,192),(256,256),(256,512)]:
    run(N, Lk)
run(256, 256, q_zero=False)
