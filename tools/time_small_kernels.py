"""Run RoPE and the attention backward once at Llama-3-8B layer shapes (for `ncu --metrics gpu__time_duration.sum -k regex:...`)."""
import sys, torch
sys.path.insert(0, ".")
from automodel_b200 import ops
T, Hq, Hkv, D = 4096, 32, 8, 128
qkv = torch.randn(T, (Hq + 2 * Hkv) * D, device="cuda").bfloat16()
pos = torch.arange(T, dtype=torch.int32, device="cuda")
inv = 1.0 / (500000.0 ** (torch.arange(0, D, 2).float() / D))
emb = torch.outer(torch.arange(T).float(), inv); emb = torch.cat((emb, emb), -1)
cos, sin = emb.cos().bfloat16().cuda(), emb.sin().bfloat16().cuda()
cu = torch.tensor([0, T], dtype=torch.int32, device="cuda")
q, k, v = qkv[:, :Hq * D], qkv[:, Hq * D:(Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:]
for _ in range(3):
    ops.rope_(qkv, cos, sin, pos, Hq + Hkv, D)
    o, lse = ops.attn_fwd(q, k, v, cu, T, Hq, Hkv, D)
    do = torch.randn_like(o); dqkv = torch.empty_like(qkv)
    ops.attn_bwd(q, k, v, o, do, lse, cu, T, Hq, Hkv, D, dqkv[:, :Hq * D], dqkv[:, Hq * D:(Hq + Hkv) * D], dqkv[:, (Hq + Hkv) * D:])
    ops.rope_(dqkv, cos, sin, pos, Hq + Hkv, D, backward=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    ops.rope_(qkv, cos, sin, pos, Hq + Hkv, D)
e1.record(); torch.cuda.synchronize()
print(f"rope (events, back to back, data 50 MB in L2 partly): {e0.elapsed_time(e1) / 50 * 1e3:.1f} us")
