"""the flat sharded layout round-trips through HF-named safetensors and a resumed run continues bit-identically."""
import numpy as np
import torch

from automodel_b200.checkpoint import save_checkpoint, load_checkpoint, load_model
from automodel_b200.engine import ShardedLlamaEngine
from tests import cpu_kernels
from tests.golden_utils import load, model_cfg, init_params, batches


def _eng(meta, cfg):
    oc = meta["optimizer"]
    return ShardedLlamaEngine(cfg, "cpu", max_tokens=meta["config"]["lbs"] * meta["config"]["seq"], lr=oc["lr"], betas=tuple(oc["betas"]), eps=oc["eps"],
                              weight_decay=oc["weight_decay"], adam_mode=1, ops=cpu_kernels)


def _mb(b):
    return {"input_ids": torch.from_numpy(b["input_ids"]), "labels": torch.from_numpy(b["labels"])}


def test_checkpoint_resume_is_bit_identical(tmp_path):
    z, meta = load("tiny_bf16")
    cfg = model_cfg(meta)
    a = _eng(meta, cfg); a.load_state_dict(init_params(meta))
    for s in range(2):
        a.train_step([_mb(b) for b in batches(z, meta, s)], meta["max_grad_norm"])
    save_checkpoint(a, str(tmp_path / "ckpt"))
    la, _ = a.train_step([_mb(b) for b in batches(z, meta, 2)], meta["max_grad_norm"])
    b_ = _eng(meta, cfg)
    load_checkpoint(b_, str(tmp_path / "ckpt"))
    assert b_.step_count == 2
    lb, _ = b_.train_step([_mb(b) for b in batches(z, meta, 2)], meta["max_grad_norm"])
    assert float(la) == float(lb)
    for k, p in a.state_dict().items():
        assert torch.equal(p, b_.state_dict()[k]), k


def test_hf_named_export_loads_like_from_pretrained(tmp_path):
    """The exported file holds exactly the HF Llama names/shapes the reference's loaders expect
    (components/models/llama/model.py q/k/v/o, gate/up/down, norms, untied lm_head)."""
    from safetensors.torch import load_file
    z, meta = load("hd128_fp32")
    cfg = model_cfg(meta)
    a = _eng(meta, cfg); a.load_state_dict(init_params(meta))
    save_checkpoint(a, str(tmp_path / "c"))
    sd = load_file(str(tmp_path / "c" / "model.safetensors"))
    assert sd["model.layers.1.self_attn.k_proj.weight"].shape == (cfg["num_key_value_heads"] * 128, cfg["hidden_size"])
    assert sd["model.layers.0.mlp.gate_proj.weight"].shape == (cfg["intermediate_size"], cfg["hidden_size"])
    assert set(sd) == set(init_params(meta))
    for k, v in init_params(meta).items():
        assert np.array_equal(sd[k].float().numpy(), v), k
    b_ = _eng(meta, cfg)
    load_model(b_, str(tmp_path / "c"))
    for k in sd:
        assert torch.equal(b_.state_dict()[k], a.state_dict()[k])
