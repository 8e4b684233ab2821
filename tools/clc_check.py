"""Bring-up check of the cluster-launch-control tile scheduler of the CTA-pair GEMM: exact-integer results, then time vs static."""
import sys, torch
sys.path.insert(0, ".")
from automodel_b200 import ops
g = torch.Generator(device="cuda").manual_seed(1)
mk = lambda s: torch.randint(-3, 4, s, device="cuda", generator=g).float().bfloat16()
ok = True
for kind, M, N, K in [(0, 512, 512, 128), (2, 1000, 776, 328), (1, 4096, 6144, 4096), (0, 4096, 28672, 4096)]:
    a = mk((M, K) if kind != 2 else (K, M)); b = mk((N, K) if kind == 0 else (K, N))
    af, bfl = a.float(), b.float()
    ref = (af @ bfl.t() if kind == 0 else af @ bfl if kind == 1 else af.t() @ bfl).bfloat16()
    for mode in (0, 1):
        ops.set_option("gemm_sched", mode)
        out = ops.gemm(kind, a, b); torch.cuda.synchronize()
        bad = int((out != ref).sum())
        print(f"sched={mode} kind={kind} {M}x{N}x{K} mismatches={bad}", flush=True)
        ok &= bad == 0
for kind, M, N, K in [(0, 4096, 4096, 14336), (0, 4096, 28672, 4096), (2, 6144, 4096, 4096)]:
    a = torch.randn((M, K) if kind != 2 else (K, M), device="cuda").bfloat16(); b = torch.randn((N, K) if kind == 0 else (K, N), device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for mode in (0, 1):
        ops.set_option("gemm_sched", mode)
        for _ in range(3): ops.gemm(kind, a, b, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): ops.gemm(kind, a, b, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"BENCH sched={mode} kind={kind} {M}x{N}x{K}: {ms:.3f} ms {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
ops.set_option("gemm_sched", 0)
print("CLC_OK" if ok else "CLC_MISMATCH")
