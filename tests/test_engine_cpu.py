"""Host-side orchestration of the sharded step, exercised on CPU with the torch stand-in kernels (tests/cpu_kernels.py):
flat layout and fused-weight views, accumulate flags, grad-norm/clip/AdamW plumbing, and - over gloo, world size 2 - the
in-place reduce-scatter / all-gather schedule.  Parity targets: the reference-generated fixtures and the numpy oracle."""
import os
import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from automodel_b200.engine import ShardedLlamaEngine, cu_seqlens_from_position_ids
from automodel_b200.layout import LlamaDims, build_layout, total_params
from tests import cpu_kernels
from tests.golden_utils import load, model_cfg, init_params, batches, check_summary, check_rel_l2


def _mb(b):
    return {"input_ids": torch.from_numpy(b["input_ids"]), "labels": torch.from_numpy(b["labels"])}


def test_layout_matches_survey_constants():
    """Llama-3-8B = 8,030,261,248 params, 218,112,000 per layer; tiny = 1,705,216."""
    d = LlamaDims(hidden=4096, ffn=14336, layers=32, heads=32, kv_heads=8, head_dim=128, vocab=128256)
    units = build_layout(d, 8)
    assert total_params(units) == 8_030_261_248
    assert units[1].numel == 218_112_000
    assert units[0].numel + units[-1].numel == 1_050_677_248
    for u in units:
        assert u.padded % (8 * 8) == 0 and u.padded - u.numel < 64
        a, b = u.shard_range(3, 8)
        assert (b - a) * 8 == u.padded and a % 8 == 0
    t = build_layout(LlamaDims(hidden=256, ffn=512, layers=2, heads=4, kv_heads=2, head_dim=64, vocab=1024), 1)
    assert total_params(t) == 1_705_216


def test_cu_seqlens_from_position_ids():
    pos = np.array([[0, 1, 2, 0, 1, 0], [0, 1, 2, 3, 4, 5]])
    cu, mx = cu_seqlens_from_position_ids(pos)
    assert cu.tolist() == [0, 3, 5, 6, 12] and mx == 6


@pytest.mark.parametrize("name,prec,tol", [("tiny_bf16", "bf16", 1e-3), ("hd128_fp32", "bf16", 4e-3), ("qwen2_tiny_bf16", "bf16", 1e-3), ("hd128_bf16", "bf16", 1e-3), ("mistral_tiny_bf16", "bf16", 1e-3),
                                           ("qwen2_tiny_fp32", "bf16", 4e-3)])
def test_engine_tracks_reference_fixture(name, prec, tol):
    """bf16 engine (CPU stand-in kernels) vs the reference run: loss / grad_norm per step, step-0 grads, weights."""
    z, meta = load(name)
    cfg = model_cfg(meta)
    oc = meta["optimizer"]
    eng = ShardedLlamaEngine(cfg, "cpu", max_tokens=meta["config"]["lbs"] * meta["config"]["seq"], lr=oc["lr"], betas=tuple(oc["betas"]),
                             eps=oc["eps"], weight_decay=oc["weight_decay"], adam_mode=1, ops=cpu_kernels)
    eng.load_state_dict(init_params(meta))
    for s in range(min(3, len(meta["loss"]))):
        mbs = [_mb(b) for b in batches(z, meta, s)]
        if s == 0:
            n = sum(int((m["labels"] != -100).sum()) for m in mbs)
            eng.loss_dev.zero_()
            for i, m in enumerate(mbs):
                eng.forward_backward(m["input_ids"], m["labels"], None, n, first_micro=i == 0, last_micro=i == len(mbs) - 1)
            for k, g in eng.named_grads().items():
                # k_proj.bias: softmax is invariant to a constant added to every key, so this gradient is what RoPE's position dependence leaves
                # of an exact zero - a sum of cancelling bf16 terms (relative noise ~10x that of the other parameters)
                check_rel_l2(z, "grad0", k, g.float().numpy(), tol=0.15 if k.endswith("k_proj.bias") else (3e-2 if name == "hd128_bf16" else 2e-2))   # hd128_bf16: q/k weight grads of layer 0 sit at 2.6 % (two bf16 runs, two micro-batches)
            loss = float(eng.loss_dev[0]); gn = float(eng.optimizer_step(meta["max_grad_norm"]).sqrt())
        else:
            l, g = eng.train_step(mbs, meta["max_grad_norm"])
            loss, gn = float(l), float(g)
        assert abs(loss - meta["loss"][s]) < tol, (s, loss, meta["loss"][s])
        assert abs(gn - meta["grad_norm"][s]) < 2e-2 * meta["grad_norm"][s], (s, gn, meta["grad_norm"][s])
        if s in (0, 2) and name == "tiny_bf16":
            for k, p in eng.state_dict().items():
                check_summary(z, f"after{s}", k, p.float().numpy(), rtol=2 ** -7, atol=1e-3, outlier_frac=0.01, outlier_atol=2 * oc["lr"] * (s + 1) + 1e-3)


def _worker(rank, world, port, name, out_q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z, meta = load(name)
    cfg = model_cfg(meta)
    oc = meta["optimizer"]
    eng = ShardedLlamaEngine(cfg, "cpu", process_group=dist.group.WORLD, max_tokens=meta["config"]["seq"], lr=oc["lr"], betas=tuple(oc["betas"]),
                             eps=oc["eps"], weight_decay=oc["weight_decay"], adam_mode=1, ops=cpu_kernels)
    eng.load_state_dict(init_params(meta))
    res = []
    for s in range(2):
        b = batches(z, meta, s)[0]
        # each rank takes one sequence of the reference's 2-sequence global batch (data parallel over sequences)
        mb = {"input_ids": torch.from_numpy(b["input_ids"][rank:rank + 1]), "labels": torch.from_numpy(b["labels"][rank:rank + 1])}
        l, g = eng.train_step([mb], meta["max_grad_norm"])
        res.append((float(l), float(g)))
    # after the step every rank must hold the same, fully gathered parameters
    flat = torch.cat([p.float().reshape(-1) for p in eng.state_dict().values()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    if rank == 0:
        out_q.put((res, same, {k: v.float().numpy() for k, v in eng.state_dict().items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_matches_reference_and_single_rank():
    """world_size 2 over gloo: sharded states, in-place reduce-scatter/all-gather.  The 2-rank run of the reference's global
    batch must reproduce the reference's loss/grad_norm (fixture: world 1, same global batch) and the single-rank engine."""
    name = "tiny_bf16"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    res, same, params2 = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert same, "ranks disagree on the gathered parameters"
    z, meta = load(name)
    for s, (l, g) in enumerate(res):
        assert abs(l - meta["loss"][s]) < 1e-3, (s, l, meta["loss"][s])
        assert abs(g - meta["grad_norm"][s]) < 2e-2 * meta["grad_norm"][s]
    # single-rank engine, same data as two micro-batches (gradient accumulation)
    cfg = model_cfg(meta); oc = meta["optimizer"]
    eng = ShardedLlamaEngine(cfg, "cpu", max_tokens=meta["config"]["seq"], lr=oc["lr"], betas=tuple(oc["betas"]), eps=oc["eps"],
                             weight_decay=oc["weight_decay"], adam_mode=1, ops=cpu_kernels)
    eng.load_state_dict(init_params(meta))
    for s in range(2):
        b = batches(z, meta, s)[0]
        mbs = [{"input_ids": torch.from_numpy(b["input_ids"][r:r + 1]), "labels": torch.from_numpy(b["labels"][r:r + 1])} for r in range(2)]
        l, g = eng.train_step(mbs, meta["max_grad_norm"])
        assert abs(float(l) - res[s][0]) < 2e-4 and abs(float(g) - res[s][1]) < 5e-3 * res[s][1]
    worst = 0.0
    for k, p in eng.state_dict().items():
        worst = max(worst, float(np.abs(p.float().numpy() - params2[k]).max()))
    assert worst <= 2 * 2 * oc["lr"] + 1e-3, worst


def test_activation_checkpointing_is_bit_identical():
    """with activation_checkpointing the layer activations live in one shared buffer set and are recomputed from the
    saved layer inputs before each layer's backward (distributed/parallelizer.py:237-268 semantics) - same kernels, same values:
    losses, gradients and updated weights are identical to the run that keeps everything, including with gradient accumulation."""
    z, meta = load("hd128_fp32")       # this fixture runs two micro-batches per step
    cfg = model_cfg(meta)
    oc = meta["optimizer"]
    engs = []
    for ac in (False, True):
        e = ShardedLlamaEngine(cfg, "cpu", max_tokens=meta["config"]["lbs"] * meta["config"]["seq"], lr=oc["lr"], betas=tuple(oc["betas"]), eps=oc["eps"],
                               weight_decay=oc["weight_decay"], adam_mode=1, ops=cpu_kernels, activation_checkpointing=ac)
        e.load_state_dict(init_params(meta))
        engs.append(e)
    assert engs[1].act["gu"][0] is engs[1].act["gu"][-1] and engs[0].act["gu"][0] is not engs[0].act["gu"][-1]
    for s in range(2):
        mbs = [_mb(b) for b in batches(z, meta, s)]
        assert len(mbs) > 1
        res = [e.train_step(mbs, meta["max_grad_norm"]) for e in engs]
        assert float(res[0][0]) == float(res[1][0]) and float(res[0][1]) == float(res[1][1])
        for k in engs[0].named_grads():
            assert torch.equal(engs[0].named_grads()[k], engs[1].named_grads()[k]), k
    for k, p in engs[0].state_dict().items():
        assert torch.equal(p, engs[1].state_dict()[k]), k


def test_rope_config_resolution_follows_the_reference():
    """components/models/llama/rope_utils.py:90-109: `rope_parameters` (transformers >= 5: theta and scaling in one dict) wins over the
    legacy `rope_theta` / `rope_scaling` pair.  A real LlamaConfig and its to_dict() must give the dims the plain dict gives."""
    import transformers
    from automodel_b200.engine import _rope_inv_freq
    base = dict(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1,
                max_position_embeddings=256, rms_norm_eps=1e-5)
    legacy = LlamaDims.from_hf(dict(base, rope_theta=500000.0))
    hf = transformers.LlamaConfig(**base, rope_theta=500000.0)
    for cfg in (hf, hf.to_dict()):
        d = LlamaDims.from_hf(cfg)
        assert d.rope_theta == 500000.0
        assert torch.equal(_rope_inv_freq(d), _rope_inv_freq(legacy))
    l3 = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0, "original_max_position_embeddings": 64}
    a = _rope_inv_freq(LlamaDims.from_hf(dict(base, rope_theta=500000.0, rope_scaling=l3)))
    b = _rope_inv_freq(LlamaDims.from_hf(dict(base, rope_parameters=dict(l3, rope_theta=500000.0))))
    assert torch.equal(a, b) and not torch.equal(a, _rope_inv_freq(legacy))


def test_rope_tables_equal_the_reference_golden():
    """engine._rope_inv_freq / rope_tables vs the reference's own `_compute_default_inv_freq` / `_compute_llama3_inv_freq` and
    `LlamaRotaryEmbedding._build_cache` (components/models/llama/rope_utils.py:112-150, 191-205): golden vectors written by
    tests/golden/gen_rope_golden.py from the unmodified reference, bit-exact (fp32 inv_freq, bf16 cos/sin).  The llama3_8b case is the
    benchmark config's RoPE."""
    import numpy as np
    from automodel_b200.engine import _rope_inv_freq, rope_tables
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rope_golden.npz"))
    l3 = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0}
    cases = {
        "llama3_8b": dict(hidden_size=4096, num_attention_heads=32, max_position_embeddings=8192, rope_theta=500000.0,
                          rope_scaling=dict(l3, original_max_position_embeddings=8192)),
        "default_hd64": dict(hidden_size=256, num_attention_heads=4, max_position_embeddings=512, rope_theta=10000.0),
        "llama3_small_ctx": dict(hidden_size=256, num_attention_heads=2, max_position_embeddings=256, rope_theta=500000.0,
                                 rope_scaling=dict(l3, original_max_position_embeddings=64)),
    }
    for name, kw in cases.items():
        d = LlamaDims.from_hf(dict(vocab_size=512, intermediate_size=512, num_hidden_layers=1, num_key_value_heads=kw["num_attention_heads"], **kw))
        assert np.array_equal(_rope_inv_freq(d).numpy(), z[name + "/inv_freq"]), name
        cos, sin = rope_tables(d, 512, torch.device("cpu"))
        assert cos.dtype == torch.bfloat16
        assert np.array_equal(cos.float().numpy(), z[name + "/cos"]) and np.array_equal(sin.float().numpy(), z[name + "/sin"]), name


def _hsdp_worker(rank, port, out_q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=4)
    # mesh (dp_replicate=2, dp_shard=2): rows are shard groups, columns are replica groups
    shard_groups = [dist.new_group([0, 1]), dist.new_group([2, 3])]
    replica_groups = [dist.new_group([0, 2]), dist.new_group([1, 3])]
    pg, rpg = shard_groups[rank // 2], replica_groups[rank % 2]
    z, meta = load("hd128_fp32")
    oc = meta["optimizer"]
    eng = ShardedLlamaEngine(model_cfg(meta), "cpu", process_group=pg, replica_group=rpg, max_tokens=meta["config"]["seq"], lr=oc["lr"],
                             betas=tuple(oc["betas"]), eps=oc["eps"], weight_decay=oc["weight_decay"], adam_mode=1, ops=cpu_kernels)
    assert (eng.world, eng.replicas, eng.rank, eng.replica_rank) == (2, 2, rank % 2, rank // 2)
    eng.load_state_dict(init_params(meta))
    res = []
    for s in range(2):
        bs = batches(z, meta, s)                       # two micro-batches of one sequence each
        b = bs[rank % 2]
        ids = torch.from_numpy(b["input_ids"]).roll(7 * (rank // 2), dims=1)    # replicas see different tokens
        lab = torch.full_like(ids, -100); lab[:, :-1] = ids[:, 1:]
        l, g = eng.train_step([{"input_ids": ids, "labels": lab}], meta["max_grad_norm"])
        res.append((float(l), float(g)))
    flat = torch.cat([p.float().reshape(-1) for p in eng.state_dict().values()])
    gathered = [torch.empty_like(flat) for _ in range(4)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        out_q.put((res, all(torch.equal(gathered[0], t) for t in gathered)))
    dist.barrier()
    dist.destroy_process_group()


def test_hsdp_2x2_gloo_equals_single_rank_accumulation():
    """SURVEY §8f N4 (HSDP, reference dp_replicate_size): 2 replicas x 2 shards over gloo.  Gradient shards are reduce-scattered inside
    the shard group and all-reduced across the replicas; the step equals one rank accumulating the four micro-batches, and all four
    ranks end with the same parameters."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_hsdp_worker, args=(r, port, q)) for r in range(4)]
    for p in procs:
        p.start()
    res, same = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert same, "ranks disagree on the parameters after the step"
    z, meta = load("hd128_fp32")
    oc = meta["optimizer"]
    eng = ShardedLlamaEngine(model_cfg(meta), "cpu", max_tokens=meta["config"]["seq"], lr=oc["lr"], betas=tuple(oc["betas"]), eps=oc["eps"],
                             weight_decay=oc["weight_decay"], adam_mode=1, ops=cpu_kernels)
    eng.load_state_dict(init_params(meta))
    for s in range(2):
        bs = batches(z, meta, s)
        mbs = []
        for rank in range(4):
            ids = torch.from_numpy(bs[rank % 2]["input_ids"]).roll(7 * (rank // 2), dims=1)
            lab = torch.full_like(ids, -100); lab[:, :-1] = ids[:, 1:]
            mbs.append({"input_ids": ids, "labels": lab})
        l, g = eng.train_step(mbs, meta["max_grad_norm"])
        assert abs(float(l) - res[s][0]) < 1e-3, (s, float(l), res[s])
        assert abs(float(g) - res[s][1]) < 5e-3 * float(g), (s, float(g), res[s])


@pytest.mark.parametrize("packed", [False, True])
def test_device_side_input_staging_writes_what_host_staging_writes(packed):
    """ShardedLlamaEngine._fill_input_buffer (batches that already live on the engine's device) against the pinned-host staging path:
    same [ids | labels | pos | cu_seqlens] words, same (nseq, max_len)."""
    z, meta = load("hd128_fp32")
    eng = ShardedLlamaEngine(model_cfg(meta), "cpu", max_tokens=2 * meta["config"]["seq"], ops=cpu_kernels)
    g = torch.Generator().manual_seed(0)
    S = meta["config"]["seq"]
    ids = torch.randint(0, 512, (2, S), generator=g)
    lab = torch.full_like(ids, -100); lab[:, :-1] = ids[:, 1:]
    pos = None
    if packed:
        pos = torch.stack([torch.cat([torch.arange(100), torch.arange(S - 100)]), torch.arange(S)])
    k, T, nseq, max_len = eng.stage(ids, lab, pos)
    want = eng._in_dev[k][:3 * T + nseq + 1].clone()
    buf = torch.full_like(eng._in_dev[k], 12345)
    got = ShardedLlamaEngine._fill_input_buffer(buf, ids, lab, pos)
    assert got == (nseq, max_len)
    assert torch.equal(buf[:3 * T + nseq + 1], want)
    buf2 = torch.full_like(buf, 12345)
    ShardedLlamaEngine._fill_input_buffer(buf2, ids, None, pos)
    assert (buf2[T:2 * T] == -100).all() and torch.equal(buf2[:T], want[:T])


def test_unsupported_configs_are_rejected_loudly():
    base = dict(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1,
                max_position_embeddings=256, rms_norm_eps=1e-5)
    from automodel_b200.engine import _rope_inv_freq
    for bad in (dict(model_type="gemma2"), dict(sliding_window=4096), dict(attention_bias=True), dict(mlp_bias=True), dict(hidden_act="gelu"),
                dict(attention_dropout=0.1), dict(model_type="qwen2", use_sliding_window=True, sliding_window=128),
                dict(model_type="qwen2", layer_types=["full_attention", "sliding_attention"])):
        with pytest.raises(ValueError):
            LlamaDims.from_hf(dict(base, **bad))
    # the Llama-family variants the kernels do cover: Qwen2's q/k/v bias, tied embeddings (Qwen2 <= 1.5B, Llama-3.2-1B)
    q = LlamaDims.from_hf(dict(base, model_type="qwen2", sliding_window=4096, use_sliding_window=False, tie_word_embeddings=True))
    assert q.qkv_bias and q.tied
    names = [sl.name for u in build_layout(q, 2) for sl in u.slots]
    assert "lm_head.weight" not in names and "model.layers.1.self_attn.v_proj.bias" in names
    t = LlamaDims.from_hf(dict(base, tie_word_embeddings=True))
    assert t.tied and not t.qkv_bias
    with pytest.raises(ValueError):
        _rope_inv_freq(LlamaDims.from_hf(dict(base, rope_scaling={"rope_type": "yarn", "factor": 4.0})))
    LlamaDims.from_hf(dict(base, model_type="mistral", sliding_window=None))


def _reshard_worker(rank, world, port, out_q, ga, ac):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z, meta = load("tiny_bf16")
    cfg = dict(model_cfg(meta), num_hidden_layers=5)     # 5 layers through a pool of 2: every slot is reused, in forward and in backward
    oc = meta["optimizer"]
    from oracle.portable_init import llama_param_shapes, portable_state_dict
    params = portable_state_dict(llama_param_shapes(cfg), seed=4)
    engs = [ShardedLlamaEngine(cfg, "cpu", process_group=dist.group.WORLD, max_tokens=meta["config"]["seq"], lr=oc["lr"], betas=tuple(oc["betas"]), eps=oc["eps"],
                               weight_decay=oc["weight_decay"], adam_mode=1, ops=cpu_kernels, reshard_after_forward=rs, activation_checkpointing=ac) for rs in (False, True)]
    for e in engs:
        e.load_state_dict(params)
    assert engs[1].p_full[1].data_ptr() == engs[1].p_full[3].data_ptr() != engs[1].p_full[2].data_ptr()     # layers share pool slots
    res = []
    for s in range(3):
        mbs = []
        for j in range(ga):
            b = batches(z, meta, s * ga + j)[0]
            mbs.append({"input_ids": torch.from_numpy(b["input_ids"][rank:rank + 1]), "labels": torch.from_numpy(b["labels"][rank:rank + 1])})
        res.append([tuple(float(x) for x in e.train_step(mbs, meta["max_grad_norm"])) for e in engs])
    sd = [e.state_dict() for e in engs]
    dmax = max(float((sd[0][k].float() - sd[1][k].float()).abs().max()) for k in sd[0])
    same = all(torch.equal(sd[0][k], sd[1][k]) for k in sd[0])
    ost = [e.gather_optimizer_state() for e in engs]
    same_opt = all(torch.equal(ost[0][k][0], ost[1][k][0]) and torch.equal(ost[0][k][1], ost[1][k][1]) for k in ost[0])
    if rank == 0:
        out_q.put((res, same, same_opt, dmax))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("ga,ac", [(1, False), (1, True), (2, False)])
def test_reshard_after_forward_equals_resident_parameters(ga, ac):
    """The reference's FSDP2 schedule (parallelizer.py:858-872: layers unsharded only while they compute; all-gather before forward and
    again before backward, reduce-scatter after each layer's backward) over gloo, world 2, 5 layers through a 2-slot pool: with one
    micro-batch per step every loss, grad norm, weight and Adam moment equals the resident-parameter engine BIT FOR BIT (also with
    activation checkpointing); with gradient accumulation the reduced shards are added instead of the unsharded gradients (the
    reference does the same), which moves roundings only."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90 + ga * 3 + int(ac)
    procs = [ctx.Process(target=_reshard_worker, args=(r, 2, port, q, ga, ac)) for r in range(2)]
    for p in procs:
        p.start()
    res, same, same_opt, dmax = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for (l0, g0), (l1, g1) in res:
        if ga == 1:
            assert l0 == l1 and g0 == g1, res
        else:
            assert abs(l0 - l1) < 2e-4 and abs(g0 - g1) < 5e-3 * g0, res
    if ga == 1:
        assert same and same_opt, dmax
    else:
        assert dmax <= 2 * 3 * 1e-3 + 1e-3, dmax


def test_memory_plan_70b_needs_resharding_and_checkpointing():
    """BASELINE config 5 (Llama-3-70B, seq 8192, 8 x B200 with 180 GB): the resident layout cannot hold it, reshard_after_forward +
    activation checkpointing can; the 8B headline config fits resident on one GPU (what bench.py runs)."""
    from automodel_b200.layout import memory_plan
    d70 = LlamaDims.from_hf(dict(vocab_size=128256, hidden_size=8192, intermediate_size=28672, num_hidden_layers=80, num_attention_heads=64,
                                 num_key_value_heads=8, max_position_embeddings=8192, rms_norm_eps=1e-5, rope_theta=500000.0))
    GB = 1e9
    resident = memory_plan(d70, 8, 8192)
    assert resident["params"] / GB > 140 and resident["total"] / GB > 400
    fit = memory_plan(d70, 8, 8192, reshard_after_forward=True, activation_checkpointing=True)
    assert fit["total"] / GB < 150, fit
    assert fit["params"] / GB < 30 and fit["optimizer"] / GB < 40
    no_ac = memory_plan(d70, 8, 8192, reshard_after_forward=True)
    assert no_ac["total"] / GB > 180, no_ac          # 80 layers of saved activations at 8192 tokens do not fit without checkpointing
    d8 = LlamaDims.from_hf(dict(vocab_size=128256, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32, num_attention_heads=32,
                                num_key_value_heads=8, max_position_embeddings=8192, rms_norm_eps=1e-5, rope_theta=500000.0))
    one = memory_plan(d8, 1, 4096)
    assert 75 < one["total"] / GB < 100, one          # ~85 GB
