"""Pin the CPU oracle (oracle/llama_step.py) against the reference's own run (tests/golden/*.npz)."""
import numpy as np
import pytest

from oracle import llama_step as O
from tests.golden_utils import load, model_cfg, init_params, batches, check_summary


@pytest.mark.parametrize("name", ["tiny_fp32", "hd128_fp32", "qwen2_tiny_fp32"])
def test_oracle_fp32_matches_reference(name):
    z, meta = load(name)
    cfg = model_cfg(meta)
    params = init_params(meta)
    oc = meta["optimizer"]
    opt = O.AdamW(lr=oc["lr"], betas=tuple(oc["betas"]), eps=oc["eps"], weight_decay=oc["weight_decay"], prec="fp32")
    nsteps = len(meta["loss"])
    for s in range(min(nsteps, 3)):
        loss, gn, pre = O.train_step(params, opt, cfg, batches(z, meta, s), prec="fp32", max_grad_norm=meta["max_grad_norm"])
        assert abs(loss - meta["loss"][s]) < 2e-5, (s, loss, meta["loss"][s])
        assert abs(gn - meta["grad_norm"][s]) < 1e-4 * meta["grad_norm"][s], (s, gn, meta["grad_norm"][s])
        if s == 0:
            for k in pre:
                check_summary(z, "grad0", k, pre[k], rtol=2e-3, atol=2e-6)
        if s in meta["snap_steps"]:
            for k in params:
                check_summary(z, f"after{s}", k, params[k], rtol=1e-4, atol=1e-4)  # Adam sign-noise where g~0: |dp|<=lr


def test_oracle_bf16_tracks_reference_curve():
    """bf16 emulation (round at every materialised tensor) follows the reference's bf16 CPU run."""
    z, meta = load("tiny_bf16")
    cfg = model_cfg(meta)
    params = init_params(meta)
    oc = meta["optimizer"]
    opt = O.AdamW(lr=oc["lr"], betas=tuple(oc["betas"]), eps=oc["eps"], weight_decay=oc["weight_decay"], prec="bf16")
    for s in range(3):
        loss, gn, pre = O.train_step(params, opt, cfg, batches(z, meta, s), prec="bf16", max_grad_norm=meta["max_grad_norm"])
        assert abs(loss - meta["loss"][s]) < 1e-3, (s, loss, meta["loss"][s])
        assert abs(gn - meta["grad_norm"][s]) < 1e-2 * meta["grad_norm"][s], (s, gn, meta["grad_norm"][s])
        if s in (0, 2):
            worst = 0.0
            for k in params:
                # bf16 weights: agree to within a couple of bf16 ulps of the weight scale
                worst = max(worst, check_summary(z, f"after{s}", k, params[k], rtol=2 ** -7, atol=1e-3,
                                                 outlier_frac=0.01, outlier_atol=2 * oc["lr"] * (s + 1) + 1e-3))


def test_masked_ce_known_answers():
    """Mirrors reference tests/unit_tests/loss/test_masked_ce.py:22-128 (sum/N normalisation, ignore_index, zero-label -> 0)."""
    rng = np.random.default_rng(0)
    logits = rng.standard_normal((2, 5, 11)).astype(np.float32)
    labels = rng.integers(0, 11, (2, 5))
    labels[0, 1] = -100; labels[1, 4] = -100
    loss, d = O.masked_ce_fwd_bwd(logits, labels, 8, O.Prec("fp32"))
    z = logits.reshape(-1, 11).astype(np.float64); y = labels.reshape(-1)
    lse = np.log(np.exp(z).sum(-1))
    ref = sum(lse[i] - z[i, y[i]] for i in range(10) if y[i] != -100) / 8
    assert abs(loss - ref) < 1e-6
    assert np.all(d.reshape(-1, 11)[y == -100] == 0)
    loss0, d0 = O.masked_ce_fwd_bwd(logits, np.full((2, 5), -100), 0, O.Prec("fp32"))
    assert loss0 == 0.0 and not d0.any()


def test_backward_matches_finite_difference():
    """The hand-derived backward agrees with central differences of the fp64 forward."""
    cfg = {"vocab_size": 37, "hidden_size": 32, "intermediate_size": 48, "num_hidden_layers": 2, "num_attention_heads": 4,
           "num_key_value_heads": 2, "max_position_embeddings": 16, "rms_norm_eps": 1e-5, "rope_theta": 10000.0}
    from oracle.portable_init import llama_param_shapes, portable_state_dict
    params = {k: v.astype(np.float64) * 3 for k, v in portable_state_dict(llama_param_shapes(cfg), seed=3).items()}
    rng = np.random.default_rng(1)
    ids = rng.integers(0, 37, (2, 12)); lab = O.mock_labels(ids)
    pos = np.array([[0, 1, 2, 3, 4, 0, 1, 2, 3, 0, 1, 2], list(range(12))])  # first row packed (3 documents)
    n = int((lab != -100).sum())
    _, grads = O.forward_backward(params, cfg, ids, lab, n, "fp64", position_ids=pos)
    for name in ["model.layers.0.self_attn.k_proj.weight", "model.layers.1.mlp.gate_proj.weight",
                 "model.layers.0.input_layernorm.weight", "model.embed_tokens.weight", "lm_head.weight", "model.norm.weight"]:
        w = params[name]
        for _ in range(3):
            idx = tuple(rng.integers(0, s) for s in w.shape)
            if name == "model.embed_tokens.weight":
                idx = (int(ids[0, 0]), idx[1])
            old = w[idx]; h = 1e-5
            w[idx] = old + h; lp, _ = O.forward_backward(params, cfg, ids, lab, n, "fp64", position_ids=pos, compute_grads=False)
            w[idx] = old - h; lm, _ = O.forward_backward(params, cfg, ids, lab, n, "fp64", position_ids=pos, compute_grads=False)
            w[idx] = old
            fd = (lp - lm) / (2 * h)
            assert abs(fd - grads[name][idx]) < 1e-6 + 1e-4 * abs(fd), (name, idx, fd, grads[name][idx])
