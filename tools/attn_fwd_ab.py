import sys, torch
sys.path.insert(0, ".")
from automodel_b200 import ops
T, Hq, Hkv, D = 4096, 32, 8, 128
qkv = torch.randn(T, (Hq + 2 * Hkv) * D, device="cuda").bfloat16()
cu = torch.tensor([0, T], dtype=torch.int32, device="cuda")
q, k, v = qkv[:, :Hq * D], qkv[:, Hq * D:(Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:]
outs = {}
for var in (0, 1):
    ops.set_option("attn_fwd_variant", var)
    for _ in range(3): o, lse = ops.attn_fwd(q, k, v, cu, T, Hq, Hkv, D)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): o, lse = ops.attn_fwd(q, k, v, cu, T, Hq, Hkv, D)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    fl = 4 * T * T * D * Hq / 2
    print(f"attn_fwd variant {var}: {ms*1e3:.1f} us  {fl/ms/1e9:.0f} TFLOP/s")
    outs[var] = (o.clone(), lse.clone())
print("max |o0-o1|", (outs[0][0].float() - outs[1][0].float()).abs().max().item(), "max |lse0-lse1|", (outs[0][1] - outs[1][1]).abs().max().item())
