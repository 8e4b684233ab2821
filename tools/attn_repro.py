#!/usr/bin/env python
"""How often does the attention forward differ between launches on identical inputs (8B shapes)?  python tools/attn_repro.py [launches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from automodel_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
T, Hq, Hkv, D = 4096, 32, 8, 128
g = torch.Generator(device="cuda").manual_seed(7)
qkv = torch.randn(T, (Hq + 2 * Hkv) * D, device="cuda", generator=g).bfloat16()
cu = torch.tensor([0, T], dtype=torch.int32, device="cuda")
q, k, v = qkv[:, :Hq * D], qkv[:, Hq * D:(Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:]
for variant in [int(x) for x in os.environ.get('VARIANTS', '1,2,3').split(',')]:
    ops.set_option("attn_fwd_variant", variant)
    o0, l0 = ops.attn_fwd(q, k, v, cu, T, Hq, Hkv, D); o0, l0 = o0.clone(), l0.clone()
    bad, worst, rows = 0, 0.0, set()
    for _ in range(n):
        o, l = ops.attn_fwd(q, k, v, cu, T, Hq, Hkv, D)
        if not (torch.equal(o, o0) and torch.equal(l, l0)):
            bad += 1
            d = (o.float() - o0.float()).abs()
            worst = max(worst, float(d.max()))
            idx = torch.nonzero(d.view(T, Hq, D).amax(-1))     # (token, head) pairs that differ
            rows.update((int(a) // 128, int(b)) for a, b in idx[:64].tolist())
    print(f"variant {variant}: {bad} of {n} launches differ; max |do| {worst:.3e}; (q tile, head) of differing rows: {sorted(rows)[:12]} lse differs: {not torch.equal(l, l0)}")
