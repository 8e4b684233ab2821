import sys, torch
sys.path.insert(0, ".")
from automodel_b200 import ops
for kind, (M, N, K) in [(0, (4096, 6144, 4096)), (0, (4096, 28672, 4096)), (0, (4096, 4096, 14336)), (1, (4096, 4096, 28672)), (2, (28672, 4096, 4096)), (0, (4096, 128256, 4096))]:
    a = torch.randn((M, K) if kind != 2 else (K, M), device="cuda").bfloat16(); b = torch.randn((N, K) if kind == 0 else (K, N), device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    res = []
    for gm in (1, 2, 4, 8, 16):
        fn = lambda: ops.gemm(kind, a, b, out=out, group_m=gm)
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        res.append((gm, round(2 * M * N * K / (e0.elapsed_time(e1) / 10) / 1e9)))
    print(f"kind={kind} {M}x{N}x{K}:", res, flush=True)
