"""Self-checks of the multi-GPU data path that run ON the ranks of a live job (bench.py runs them before its timed region and prints
the result in its JSON line; tests/test_dist_gpu.py asserts on the same functions).  They compare this repository's per-unit collectives
and the N-rank sharded step against independent computations on the same GPUs - an fp32 NCCL all-reduce, and a single-rank engine that
accumulates the N ranks' micro-batches - so a driver that only sees bench.py's output still sees whether N > 1 is CORRECT, not just fast.
No reference or oracle code is imported here."""
import torch
import torch.distributed as dist

from .engine import ShardedLlamaEngine


def _bf16_ulp_distance(a: torch.Tensor, b: torch.Tensor) -> int:
    """max distance in units of bf16 representable values (monotone integer mapping of the bit patterns)."""
    def key(t):
        i = t.contiguous().view(torch.int16).to(torch.int32)
        return torch.where(i < 0, -(i & 0x7FFF), i)
    return int((key(a) - key(b)).abs().max().item()) if a.numel() else 0


@torch.no_grad()
def check_collectives(eng: ShardedLlamaEngine, unit_index: int = 1, seed: int = 99):
    """Reduce-scatter and all-gather of ONE unit of `eng` (default: decoder layer 0 - at the benchmark config a 218 M-element unit) on
    random bf16 data, against fp32 references computed with plain NCCL collectives.  Clobbers that unit's gradient buffer and parameter
    buffer: call before the weights are initialised.  Returns a dict (same on every rank)."""
    assert eng.world > 1 and eng.streams.cuda
    ui, dev, pg = unit_index, eng.device, eng.pg
    n = eng.units[ui].padded
    a, b = eng.units[ui].shard_range(eng.rank, eng.world)
    g = torch.Generator(device=dev).manual_seed(seed + eng.rank)
    # gradients with a wide dynamic range: products of two normals, so that partial sums cancel and bf16 per-hop rounding would show
    x = (torch.randn(n, generator=g, device=dev) * torch.randn(n, generator=g, device=dev)).to(torch.bfloat16)
    eng.g_full[ui].copy_(x)
    ref32 = x.float()
    dist.all_reduce(ref32, op=dist.ReduceOp.SUM, group=pg)           # independent path: NCCL fp32 sum of the N bf16 tensors
    mag32 = x.float().abs()
    dist.all_reduce(mag32, op=dist.ReduceOp.SUM, group=pg)           # sum of |addends|: the scale of fp32 accumulation error
    want = ref32[a:b].to(torch.bfloat16)
    torch.cuda.synchronize(dev)
    dist.barrier(group=pg)
    eng._rs_started = False
    eng._reduce_scatter_unit(ui)
    eng.streams.wait(eng.ev_rs[ui])
    eng.ev_rs[ui] = None
    torch.cuda.synchronize(dev)
    got = eng.g_full[ui][a:b]
    rs_ulp = _bf16_ulp_distance(got, want)
    rs_exact = float((got == want).float().mean().item())
    # error model of "fp32 accumulation in ANY order, then ONE rounding to bf16": |got - sum| <= 2^-8 |sum| (one bf16 ulp of the result)
    # + 2^-21 sum|addends| (N fp32 roundings of partial sums; shows when the addends cancel).  A bf16 ring (one rounding per hop) is
    # outside it by orders of magnitude.
    err = (got.float() - ref32[a:b]).abs()
    bound = ref32[a:b].abs() * 2.0 ** -8 + mag32[a:b] * 2.0 ** -21
    rs_excess = float((err / bound.clamp_min(1e-30)).max().item())
    # a few mismatching elements of rank 0's slice with every rank's addend (how does the reducing hardware round?)
    per = (b - a)
    idx = torch.nonzero(got != want).flatten()[:4] if eng.rank == 0 else torch.zeros(0, dtype=torch.int64, device=dev)
    pick = torch.full((4,), -1, dtype=torch.int64, device=dev)
    pick[:idx.numel()] = idx
    dist.broadcast(pick, src=dist.get_global_rank(pg, 0), group=pg)
    sel = pick.clamp_min(0)
    addends = [torch.empty(4, dtype=torch.float32, device=dev) for _ in range(eng.world)]
    dist.all_gather(addends, x[sel].float(), group=pg)            # rank 0's slice starts at element 0 of the unit
    examples = []
    if eng.rank == 0:
        for k in range(4):
            if int(pick[k]) >= 0:
                examples.append({"addends": [float(t[k]) for t in addends], "fp32_sum": float(ref32[sel[k]]), "got": float(got[sel[k]]), "rne": float(want[sel[k]])})
    # the grad-norm partial that the reduce-scatter path accumulated for this shard
    norm_rel = abs(float(eng.norm_sq[0]) - float(got.float().pow(2).sum())) / max(float(got.float().pow(2).sum()), 1e-30)
    eng._rs_started = False
    # all-gather: every rank writes a rank-specific pattern into its own slice, garbage elsewhere
    y = torch.randn(n, generator=g, device=dev).to(torch.bfloat16)
    eng.p_full[ui].copy_(y)
    shards = [torch.empty(b - a, dtype=torch.bfloat16, device=dev) for _ in range(eng.world)]
    dist.all_gather(shards, y[a:b].contiguous(), group=pg)
    want_full = torch.cat(shards)
    torch.cuda.synchronize(dev)
    dist.barrier(group=pg)
    eng._all_gather_unit(ui)
    eng.streams.wait(eng.ev_ag[ui])
    eng.ev_ag[ui] = None
    torch.cuda.synchronize(dev)
    ag_equal = bool(torch.equal(eng.p_full[ui], want_full))
    stats = torch.tensor([float(rs_ulp), 1.0 - rs_exact, norm_rel, 0.0 if ag_equal else 1.0, rs_excess], dtype=torch.float64, device=dev)
    dist.all_reduce(stats, op=dist.ReduceOp.MAX, group=pg)
    rs_ulp, inexact, norm_rel, ag_bad, rs_excess = stats.tolist()
    return {"unit_elems": int(n), "comm": eng.comm_kind, "reduce_dtype": eng.reduce_dtype,
            "rs_max_bf16_ulp_vs_fp32_allreduce": int(rs_ulp), "rs_frac_not_bit_equal": inexact, "rs_norm_sq_rel_err": norm_rel,
            "rs_err_over_fp32_accumulate_bound": rs_excess,
            "ag_bit_exact": ag_bad == 0.0, "rs_mismatch_examples_rank0": examples}


PARITY_CFG = {"vocab_size": 2048, "hidden_size": 512, "intermediate_size": 1024, "num_hidden_layers": 2, "num_attention_heads": 4,
              "num_key_value_heads": 1, "max_position_embeddings": 512, "rms_norm_eps": 1e-5, "rope_theta": 500000.0,
              "rope_scaling": {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                               "original_max_position_embeddings": 128}}


def check_sharded_step_parity(pg, device, steps: int = 10, seq: int = 512, replica_group=None, **engine_kw):
    """`steps` optimizer steps of a small Llama (head_dim 128, GQA 4:1, llama3 RoPE - the benchmark config's layer structure) run twice on
    the same data: sharded over the ranks of `pg` (one sequence per rank per step, the production collectives), and on ONE rank
    accumulating all ranks' sequences as micro-batches (no collectives; this single-rank path is what the -m gpu tests pin against the
    reference's fixtures).  Returns max |dloss|, max relative grad-norm difference, whether all ranks hold identical parameters after
    the last all-gather, and the largest parameter difference between the two runs."""
    dev = torch.device(device)
    world = dist.get_world_size(pg) * (dist.get_world_size(replica_group) if replica_group is not None else 1)
    grank = dist.get_rank()     # data index: global rank
    kw = dict(max_tokens=seq, lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1, adam_mode=1, max_positions=seq)
    eng_n = ShardedLlamaEngine(PARITY_CFG, dev, process_group=pg, replica_group=replica_group, **kw, **engine_kw)
    eng_1 = ShardedLlamaEngine(PARITY_CFG, dev, process_group=None, **kw)
    eng_n.init_random_(seed=7)
    eng_1.init_random_(seed=7)
    V = PARITY_CFG["vocab_size"]
    max_dl, max_dg = 0.0, 0.0
    for s in range(steps):
        rows = []
        for r in range(world):
            g = torch.Generator().manual_seed(1000 * s + r)
            ids = torch.randint(0, V, (1, seq), generator=g, dtype=torch.int64)
            lab = torch.full((1, seq), -100, dtype=torch.int64)
            lab[:, :-1] = ids[:, 1:]
            rows.append({"input_ids": ids, "labels": lab})
        ln, gn = eng_n.train_step([rows[grank]], 1.0)
        l1, g1 = eng_1.train_step(rows, 1.0)
        ln, gn, l1, g1 = float(ln), float(gn), float(l1), float(g1)
        max_dl = max(max_dl, abs(ln - l1))
        max_dg = max(max_dg, abs(gn - g1) / max(abs(g1), 1e-12))
    flat_n = torch.cat([p.float().reshape(-1) for p in eng_n.state_dict().values()])
    flat_1 = torch.cat([p.float().reshape(-1) for p in eng_1.state_dict().values()])
    lo, hi = flat_n.clone(), flat_n.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    agree = bool(torch.equal(lo, hi))
    dparam = float((flat_n - flat_1).abs().max())
    stats = torch.tensor([max_dl, max_dg, dparam], dtype=torch.float64, device=dev)
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    eng_n.close(); eng_1.close()
    return {"steps": steps, "world": world, "comm": eng_n.comm_kind, "max_abs_dloss": stats[0].item(), "max_rel_dgnorm": stats[1].item(),
            "ranks_agree": agree, "max_abs_dparam_vs_single_rank": stats[2].item(), "final_loss": ln,
            "against": "one rank accumulating the same sequences (no collectives)"}
