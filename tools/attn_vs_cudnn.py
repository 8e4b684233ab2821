#!/usr/bin/env python
"""Attention forward / backward of this repository (b200_attn_fwd / b200_attn_bwd through the C ABI) against torch SDPA restricted to
the cuDNN fused attention backend (and, for orientation, the flash backend) on the SAME B200, Llama-3-8B shapes: S=4096 causal,
32 query heads / 8 kv heads (GQA 4:1), head_dim 128, bf16.  What the reference uses on this path: `attn_implementation: sdpa`
(components/models/llama/model.py:135-148 -> transformers' sdpa_attention_forward -> F.scaled_dot_product_attention).

  python tools/attn_vs_cudnn.py [--seq 4096] [--iters 30] [--md profiles/r2_attn_vs_cudnn.md]

Timing: CUDA events around `iters` back-to-back launches after warm-up (inputs 100 MB >> nothing cached matters: the kernels are
compute-bound); FLOPs = 4*S^2*D*Hq/2 forward, 2.5x that backward (causal)."""
import argparse
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

from automodel_b200 import ops


def timeit(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seq", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--md", default=None)
    ap.add_argument("--variant", type=int, default=None, help="b200_set_option attn_fwd_variant (A/B of forward kernels)")
    ap.add_argument("--ours-only", action="store_true", help="skip the SDPA arms (ncu captures, quick A/Bs)")
    a = ap.parse_args()
    if a.variant is not None:
        ops.set_option("attn_fwd_variant", a.variant)
    T, Hq, Hkv, D = a.seq, 32, 8, 128
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = torch.randn(T, (Hq + 2 * Hkv) * D, device=dev, generator=g).bfloat16()
    cu = torch.tensor([0, T], dtype=torch.int32, device=dev)
    q, k, v = qkv[:, :Hq * D], qkv[:, Hq * D:(Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:]
    fl_f = 4.0 * T * T * D * Hq / 2
    fl_b = 2.5 * fl_f
    rows = []

    # ---- ours
    o, lse = ops.attn_fwd(q, k, v, cu, T, Hq, Hkv, D)
    do = torch.randn(T, Hq * D, device=dev, generator=g).bfloat16()
    dqkv = torch.empty_like(qkv)
    ms_f = timeit(lambda: ops.attn_fwd(q, k, v, cu, T, Hq, Hkv, D, out=o, lse=lse), a.iters)
    ms_b = timeit(lambda: ops.attn_bwd(q, k, v, o, do, lse, cu, T, Hq, Hkv, D, dqkv[:, :Hq * D], dqkv[:, Hq * D:(Hq + Hkv) * D], dqkv[:, (Hq + Hkv) * D:]), a.iters)
    rows.append((f"b200_attn_fwd / b200_attn_bwd (this repo{'' if a.variant is None else f', fwd variant {a.variant}'})", ms_f, ms_b))
    if a.ours_only:
        print(f"ours: fwd {ms_f * 1e3:.1f} us = {fl_f / ms_f / 1e9:.0f} TFLOP/s, bwd {ms_b * 1e3:.1f} us = {fl_b / ms_b / 1e9:.0f} TFLOP/s")
        return

    # ---- torch SDPA backends, [b, h, S, d] layout as the reference's model passes it; GQA via enable_gqa (cuDNN supports it natively)
    q4 = q.view(1, T, Hq, D).transpose(1, 2).detach().clone().requires_grad_(True)
    k4 = k.view(1, T, Hkv, D).transpose(1, 2).detach().clone().requires_grad_(True)
    v4 = v.view(1, T, Hkv, D).transpose(1, 2).detach().clone().requires_grad_(True)
    do4 = do.view(1, T, Hq, D).transpose(1, 2)
    ref_o = None
    for name, be in (("torch SDPA, cuDNN fused attention", SDPBackend.CUDNN_ATTENTION), ("torch SDPA, flash-attention 2", SDPBackend.FLASH_ATTENTION)):
        try:
            with sdpa_kernel(be):
                def fwd():
                    return F.scaled_dot_product_attention(q4, k4, v4, is_causal=True, enable_gqa=True)
                out = fwd()
                if ref_o is None:
                    ref_o = out.detach()
                ms_f = timeit(lambda: fwd(), a.iters)

                def fb():
                    oo = fwd()
                    oo.backward(do4)
                ms_fb = timeit(fb, a.iters)
            rows.append((name, ms_f, ms_fb - ms_f))
        except Exception as e:  # noqa: BLE001
            rows.append((name + f" - unavailable ({type(e).__name__}: {str(e)[:80]})", float("nan"), float("nan")))
    if ref_o is not None:
        ours = o.view(T, Hq, D).float()
        theirs = ref_o[0].transpose(0, 1).float()
        print(f"max |o_ours - o_sdpa| = {(ours - theirs).abs().max().item():.3e}")
    lines = [f"### attention, S={T}, Hq={Hq}, Hkv={Hkv}, D={D}, causal, bf16 ({torch.cuda.get_device_name(0)}, torch {torch.__version__})", "",
             "| implementation | fwd us | fwd TFLOP/s | bwd us | bwd TFLOP/s |", "|---|---:|---:|---:|---:|"]
    for name, f_, b_ in rows:
        lines.append(f"| {name} | {f_ * 1e3:.1f} | {fl_f / f_ / 1e9:.0f} | {b_ * 1e3:.1f} | {fl_b / b_ / 1e9:.0f} |")
    text = "\n".join(lines)
    print(text)
    if a.md:
        with open(a.md, "a") as f:
            f.write(text + "\n\n")


if __name__ == "__main__":
    main()
