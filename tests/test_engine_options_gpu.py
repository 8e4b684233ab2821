"""GPU checks of the engine's optional schedules: activation checkpointing (reference: distributed.activation_checkpointing,
components/distributed/parallelizer.py:222-286), the cluster-launch-control GEMM tile scheduler inside the step, the SwiGLU GEMM epilogue.
First run green on a B200 in round 2 (gpurun_out/r2_experimental.log, 7 passed); part of the default `-m gpu` gate since."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from automodel_b200 import ops  # noqa: E402
from automodel_b200.engine import ShardedLlamaEngine  # noqa: E402
from tests.golden_utils import load, model_cfg, init_params, batches  # noqa: E402


def _mb(b):
    return {"input_ids": torch.from_numpy(b["input_ids"]), "labels": torch.from_numpy(b["labels"])}


def _engine(meta, **kw):
    oc = meta["optimizer"]
    e = ShardedLlamaEngine(model_cfg(meta), "cuda", max_tokens=meta["config"]["lbs"] * meta["config"]["seq"], lr=oc["lr"], betas=tuple(oc["betas"]),
                           eps=oc["eps"], weight_decay=oc["weight_decay"], adam_mode=1, **kw)
    e.load_state_dict(init_params(meta))
    return e


def test_activation_checkpointing_reproduces_the_reference_curve():
    z, meta = load("tiny_bf16")
    e = _engine(meta, activation_checkpointing=True)
    for s in range(10):
        l, g = e.train_step([_mb(b) for b in batches(z, meta, s)], meta["max_grad_norm"])
        assert abs(float(l) - meta["loss"][s]) < 1e-3, (s, float(l), meta["loss"][s])
        assert abs(float(g) - meta["grad_norm"][s]) < 2e-2 * meta["grad_norm"][s]


def test_activation_checkpointing_equals_plain_run_at_8b_layer_dims():
    cfg = {"vocab_size": 32768, "hidden_size": 4096, "intermediate_size": 14336, "num_hidden_layers": 4, "num_attention_heads": 32,
           "num_key_value_heads": 8, "max_position_embeddings": 8192, "rms_norm_eps": 1e-5, "rope_theta": 500000.0}
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, 32768, (1, 4096), generator=g)
    lab = torch.full_like(ids, -100); lab[:, :-1] = ids[:, 1:]
    out = []
    for ac in (False, True):
        e = ShardedLlamaEngine(cfg, "cuda", max_tokens=4096, lr=1e-5, adam_mode=1, max_positions=4096, activation_checkpointing=ac)
        e.init_random_(seed=5)
        e.loss_dev.zero_()
        e.forward_backward(ids, lab, None, 4095)
        torch.cuda.synchronize()
        out.append((float(e.loss_dev[0]), {k: v.clone() for k, v in e.named_grads().items()}))
        del e
        torch.cuda.empty_cache()
    assert out[0][0] == out[1][0]
    for k, a in out[0][1].items():
        assert (a.float() - out[1][1][k].float()).norm() <= 5e-2 * a.float().norm() + 1e-12, k


def test_clc_gemm_scheduler_inside_the_step():
    """The whole step with the cluster-launch-control GEMM scheduler, wgrad stream on: same curve as the reference."""
    z, meta = load("tiny_bf16")
    ops.set_option("gemm_sched", 1)
    try:
        e = _engine(meta)
        for s in range(5):
            l, g = e.train_step([_mb(b) for b in batches(z, meta, s)], meta["max_grad_norm"])
            assert abs(float(l) - meta["loss"][s]) < 1e-3, (s, float(l), meta["loss"][s])
        torch.cuda.synchronize()
    finally:
        ops.set_option("gemm_sched", 0)


@pytest.mark.parametrize("M,F,K", [(512, 512, 256), (1000, 640, 328), (4096, 14336, 4096)])
def test_gemm_swiglu_epilogue_is_bit_identical_to_the_two_kernel_path(M, F, K):
    g = torch.Generator(device="cuda").manual_seed(M + F)
    x = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(2 * F, K, device="cuda", generator=g) * 0.05).bfloat16()
    gu_ref = ops.gemm(ops.NT, x, w)
    a_ref = ops.swiglu_fwd(gu_ref)
    gu, a = ops.gemm_swiglu(x, w)
    torch.cuda.synchronize()
    assert torch.equal(gu, gu_ref), int((gu != gu_ref).sum())
    assert torch.equal(a, a_ref), int((a != a_ref).sum())


def test_fused_swiglu_inside_the_step(monkeypatch):
    monkeypatch.setenv("B200_FUSE_SWIGLU", "1")
    z, meta = load("tiny_bf16")
    e = _engine(meta)
    assert e._fuse_swiglu
    before = ops.LAUNCHES
    for s in range(5):
        l, g = e.train_step([_mb(b) for b in batches(z, meta, s)], meta["max_grad_norm"])
        assert abs(float(l) - meta["loss"][s]) < 1e-3, (s, float(l), meta["loss"][s])
    torch.cuda.synchronize()
    assert ops.LAUNCHES > before


@pytest.mark.parametrize("ac", [False, True])
def test_reshard_after_forward_equals_resident_on_one_gpu(ac):
    """reshard_after_forward on the CUDA engine at world 1 (no collectives, but the real pool / prefetch / event logic next to the
    weight-gradient and optimizer side streams): 5 layers through the 2-slot pool give the resident engine's losses, grad norms and
    weights bit for bit over 4 steps, with two micro-batches in the last two (accumulation on the shard: rounding only)."""
    z, meta = load("tiny_bf16")
    cfg = dict(model_cfg(meta), num_hidden_layers=5)
    oc = meta["optimizer"]
    from oracle.portable_init import llama_param_shapes, portable_state_dict
    params = portable_state_dict(llama_param_shapes(cfg), seed=4)
    engs = [ShardedLlamaEngine(cfg, "cuda", max_tokens=meta["config"]["lbs"] * meta["config"]["seq"], lr=oc["lr"], betas=tuple(oc["betas"]), eps=oc["eps"],
                               weight_decay=oc["weight_decay"], adam_mode=1, reshard_after_forward=rs, activation_checkpointing=ac) for rs in (False, True)]
    for e in engs:
        e.load_state_dict(params)
    for s in range(4):
        mbs = [_mb(b) for b in batches(z, meta, s)]
        if s >= 2:
            mbs = mbs + [_mb(b) for b in batches(z, meta, s + 10)]
        res = [e.train_step(mbs, meta["max_grad_norm"]) for e in engs]
        l0, g0, l1, g1 = float(res[0][0]), float(res[0][1]), float(res[1][0]), float(res[1][1])
        if s < 2:
            assert l0 == l1 and g0 == g1, (s, l0, l1, g0, g1)
        else:
            assert abs(l0 - l1) < 2e-4 and abs(g0 - g1) < 5e-3 * g0, (s, l0, l1, g0, g1)
        if s == 1:
            sd0, sd1 = engs[0].state_dict(), engs[1].state_dict()
            for k in sd0:
                assert torch.equal(sd0[k], sd1[k]), k
    torch.cuda.synchronize()
