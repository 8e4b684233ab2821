"""the flat sharded layout round-trips through HF-named safetensors and a resumed run continues bit-identically."""
import numpy as np
import pytest
import torch

from automodel_b200.checkpoint import save_checkpoint, load_checkpoint, load_model
from automodel_b200.engine import ShardedLlamaEngine
from tests import cpu_kernels
from tests.golden_utils import load, model_cfg, init_params, batches


def _eng(meta, cfg, device="cpu"):
    oc = meta["optimizer"]
    return ShardedLlamaEngine(cfg, device, max_tokens=meta["config"]["lbs"] * meta["config"]["seq"], lr=oc["lr"], betas=tuple(oc["betas"]), eps=oc["eps"],
                              weight_decay=oc["weight_decay"], adam_mode=1, ops=cpu_kernels if device == "cpu" else None)


DEVICES = ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)]     # cuda: the real kernels and side streams (save while an optimizer sweep is in flight)


def _mb(b):
    return {"input_ids": torch.from_numpy(b["input_ids"]), "labels": torch.from_numpy(b["labels"])}


@pytest.mark.parametrize("device", DEVICES)
def test_checkpoint_resume_is_bit_identical(tmp_path, device):
    z, meta = load("tiny_bf16")
    cfg = model_cfg(meta)
    a = _eng(meta, cfg, device); a.load_state_dict(init_params(meta))
    for s in range(2):
        a.train_step([_mb(b) for b in batches(z, meta, s)], meta["max_grad_norm"])
    save_checkpoint(a, str(tmp_path / "ckpt"))
    la, _ = a.train_step([_mb(b) for b in batches(z, meta, 2)], meta["max_grad_norm"])
    b_ = _eng(meta, cfg, device)
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


def _reshard_worker(rank, world, port, ckpt, out_q, mode):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z, meta = load("tiny_bf16")
    cfg = model_cfg(meta); oc = meta["optimizer"]
    e = ShardedLlamaEngine(cfg, "cpu", process_group=dist.group.WORLD, max_tokens=meta["config"]["seq"], lr=oc["lr"], betas=tuple(oc["betas"]), eps=oc["eps"],
                           weight_decay=oc["weight_decay"], adam_mode=1, ops=cpu_kernels)
    if mode == "save":
        e.load_state_dict(init_params(meta))
        for s in range(2):
            b = batches(z, meta, s)[0]
            e.train_step([{"input_ids": torch.from_numpy(b["input_ids"][rank:rank + 1]), "labels": torch.from_numpy(b["labels"][rank:rank + 1])}], meta["max_grad_norm"])
        save_checkpoint(e, ckpt)
    else:
        load_checkpoint(e, ckpt)
    st = e.gather_optimizer_state()
    if rank == 0:
        out_q.put({k: (m.float().numpy().copy(), v.float().numpy().copy()) for k, (m, v) in st.items()})   # numpy: plain pickling, no fd passing
    dist.barrier()
    dist.destroy_process_group()


def test_optimizer_state_is_resharded_across_world_sizes(tmp_path):
    """A checkpoint written by 2 ranks resumes on 1 rank (and a 1-rank checkpoint on 2 ranks) with every Adam moment in place."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    z, meta = load("tiny_bf16")
    cfg = model_cfg(meta)

    def run(world, mode, ckpt, port):
        q = ctx.Queue()
        ps = [ctx.Process(target=_reshard_worker, args=(r, world, port, ckpt, q, mode)) for r in range(world)]
        for p in ps:
            p.start()
        out = q.get(timeout=300)
        for p in ps:
            p.join(timeout=60)
            assert p.exitcode == 0
        return out

    import os
    port = 29800 + os.getpid() % 500
    want = run(2, "save", str(tmp_path / "w2"), port)
    one = _eng(meta, cfg)
    load_checkpoint(one, str(tmp_path / "w2"))
    assert one.step_count == 2
    got = one.gather_optimizer_state()
    for k, (m, v) in want.items():
        assert np.array_equal(got[k][0].float().numpy(), m) and np.array_equal(got[k][1].float().numpy(), v), k
        assert float(np.abs(m).sum()) > 0
    save_checkpoint(one, str(tmp_path / "w1"))
    back = run(2, "load", str(tmp_path / "w1"), port + 1)
    for k, (m, v) in want.items():
        assert np.array_equal(back[k][0], m) and np.array_equal(back[k][1], v), k
