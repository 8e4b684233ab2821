#!/usr/bin/env python
"""Step time and memory of the Llama-3-8B config with reshard_after_forward (+ activation checkpointing) vs the resident layout, one GPU.
At world 1 the all-gathers are shard -> pool-slot copies, so this isolates the cost of the schedule itself (2 x 16 GB of copies per step)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from automodel_b200.engine import ShardedLlamaEngine
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import LLAMA3_8B, SEQ
g = torch.Generator().manual_seed(1)
ids = torch.randint(0, 128256, (1, SEQ), generator=g)
lab = torch.full_like(ids, -100); lab[:, :-1] = ids[:, 1:]
for rs, ac in ((False, False), (True, False), (True, True)):
    torch.cuda.reset_peak_memory_stats()
    e = ShardedLlamaEngine(dict(LLAMA3_8B), "cuda", max_tokens=SEQ, lr=1e-5, adam_mode=1, max_positions=SEQ, reshard_after_forward=rs, activation_checkpointing=ac)
    e.init_random_(seed=3)
    st = [e.stage(ids, lab)]
    for _ in range(3):
        e.train_step(None, 1.0, num_label_tokens=SEQ - 1, staged=st)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(6):
        l, gn = e.train_step(None, 1.0, num_label_tokens=SEQ - 1, staged=st)
    e.sync_params(); t1.record(); torch.cuda.synchronize()
    print(f"reshard={rs} ac={ac}: {t0.elapsed_time(t1) / 6:.1f} ms/step  loss {float(l):.4f} gnorm {float(gn):.3f}  peak {torch.cuda.max_memory_allocated() / 1e9:.1f} GB", flush=True)
    e.close(); del e, st
    torch.cuda.empty_cache()
