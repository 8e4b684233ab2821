"""Standalone tcgen05 GEMM bring-up: each case in its own subprocess with a timeout so a deadlocked pipeline cannot take
the whole GPU session down.  Usage: python tools/gemm_selftest.py [--bench]"""
import os, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASE = r'''
import sys, torch
sys.path.insert(0, %r)
from automodel_b200 import ops
kind, M, N, K, pair = %d, %d, %d, %d, %d
ops.set_option("gemm_2cta", pair)
g = torch.Generator(device="cuda").manual_seed(1)
mk = lambda s: torch.randint(-3, 4, s, device="cuda", generator=g).float().bfloat16()
a = mk((M, K) if kind != 2 else (K, M)); b = mk((N, K) if kind == 0 else (K, N))
out = ops.gemm(kind, a, b); torch.cuda.synchronize()
af, bfl = a.float(), b.float()
ref = (af @ bfl.t() if kind == 0 else af @ bfl if kind == 1 else af.t() @ bfl).bfloat16()
bad = (out != ref)
print("RESULT pair=%%d kind=%%d %%dx%%dx%%d mismatches=%%d/%%d" %% (pair, kind, M, N, K, int(bad.sum()), bad.numel()))
if bad.any():
    idx = bad.nonzero()[:8].tolist()
    print("  first bad:", [(i, j, out[i, j].item(), ref[i, j].item()) for i, j in idx])
    rows = bad.any(1).nonzero().flatten(); cols = bad.any(0).nonzero().flatten()
    print("  bad rows: n=%%d min=%%d max=%%d ; bad cols: n=%%d min=%%d max=%%d" %% (len(rows), rows.min(), rows.max(), len(cols), cols.min(), cols.max()))
'''
cases = [(k, *s, 1) for k in (0, 1, 2) for s in [(256, 256, 64), (256, 256, 256), (512, 512, 128), (384, 320, 192), (1000, 776, 328), (4096, 6144, 4096)]]
if "--all" in sys.argv:
    cases += [(k, *s, 0) for k in (0, 1, 2) for s in [(128, 256, 64), (256, 512, 128), (128, 128, 64), (384, 320, 192), (4096, 6144, 4096)]]
fails = 0
for c in cases:
    t0 = time.time()
    try:
        r = subprocess.run([sys.executable, "-c", CASE % ((ROOT,) + c)], capture_output=True, text=True, timeout=120)
        out = (r.stdout + r.stderr).strip().splitlines()
        keep = [l for l in out if l.startswith("RESULT") or l.startswith("  ") or "rror" in l]
        print("\n".join(keep[-6:]) if keep else "\n".join(out[-5:]), f"[{time.time()-t0:.1f}s rc={r.returncode}]", flush=True)
        if r.returncode != 0 or "mismatches=0/" not in r.stdout:
            fails += 1
    except subprocess.TimeoutExpired:
        print("TIMEOUT", c, flush=True)
        fails += 1
print("FAILS", fails)
if "--bench" in sys.argv and fails == 0:
    import torch
    sys.path.insert(0, ROOT)
    from automodel_b200 import ops
    for kind, (M, N, K) in [(0, (4096, 6144, 4096)), (0, (4096, 28672, 4096)), (0, (4096, 4096, 14336)), (1, (4096, 4096, 6144)), (1, (4096, 14336, 4096)),
                            (2, (6144, 4096, 4096)), (2, (28672, 4096, 4096)), (2, (4096, 14336, 4096)), (0, (8192, 8192, 8192))]:
        a = torch.randn((M, K) if kind != 2 else (K, M), device="cuda").bfloat16(); b = torch.randn((N, K) if kind == 0 else (K, N), device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        def mk(bn):
            def f():
                ops.set_option("gemm_bn", bn); ops.gemm(kind, a, b, out=out); ops.set_option("gemm_bn", 0)
            return f
        def pair():
            ops.set_option("gemm_2cta", 1); ops.gemm(kind, a, b, out=out); ops.set_option("gemm_2cta", 0)
        for name, fn in [("tcgen05_1cta", lambda: ops.gemm(kind, a, b, out=out)), ("tcgen05_2cta", pair),
                         ("cublasLt", lambda: ops.gemm_cublaslt(kind, a, b, out=out))]:
            for _ in range(3): fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(10): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"BENCH kind={kind} {M}x{N}x{K} {name}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
sys.exit(1 if fails else 0)
