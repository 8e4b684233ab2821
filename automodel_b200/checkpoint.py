"""State-dict / checkpoint round trip for the flat sharded layout.

On disk the model is presented in the reference's format - HF-named tensors in `model.safetensors`
(components/checkpoint/checkpointing.py:256-345 consolidated HF export; components/models/llama/state_dict_adapter.py)
so a checkpoint written here loads into the reference / HF and vice versa; optimizer state is saved per rank as flat shards
(`optim_rank{r}.safetensors`: m, v [, master] per unit + step), the analogue of the reference's per-rank DCP optimizer files.
Host glue only (no device code).
"""
import json
import os

import torch
from safetensors.torch import load_file, save_file


def save_checkpoint(engine, path):
    """Rank 0 writes the full HF-named model (parameters are gathered on every rank in this design); every rank writes its
    optimizer shards."""
    os.makedirs(path, exist_ok=True)
    engine.sync_params()
    if engine.device.type == "cuda":
        torch.cuda.synchronize(engine.device)
    if engine.replica_rank != 0:
        return       # HSDP: the replicas hold identical shards; replica 0 writes
    if engine.rank == 0:
        sd = {k: v.detach().to("cpu").contiguous() for k, v in engine.state_dict().items()}
        save_file(sd, os.path.join(path, "model.safetensors"), metadata={"format": "pt"})
        with open(os.path.join(path, "b200_meta.json"), "w") as f:
            json.dump({"world": engine.world, "step_count": engine.step_count, "adam_mode": engine.adam_mode,
                       "units": [[u.name, u.numel, u.padded] for u in engine.units]}, f)
    opt = {}
    for ui, u in enumerate(engine.units):
        opt[f"m.{u.name}"] = engine.m[ui].detach().to("cpu").contiguous()
        opt[f"v.{u.name}"] = engine.v[ui].detach().to("cpu").contiguous()
        if engine.master is not None:
            opt[f"master.{u.name}"] = engine.master[ui].detach().to("cpu").contiguous()
    opt["step_count"] = torch.tensor([engine.step_count], dtype=torch.int64)
    save_file(opt, os.path.join(path, f"optim_rank{engine.rank}.safetensors"))


def load_model(engine, path_or_file):
    """from_pretrained into the flat buffers: any HF Llama `model.safetensors` (or a directory holding one / several shards)."""
    files = [path_or_file]
    if os.path.isdir(path_or_file):
        files = sorted(os.path.join(path_or_file, f) for f in os.listdir(path_or_file) if f.endswith(".safetensors") and not f.startswith("optim_rank"))
    sd = {}
    for f in files:
        sd.update(load_file(f))
    missing = [k for k in engine.P if k not in sd]
    if missing:
        raise KeyError(f"checkpoint lacks {len(missing)} tensors, e.g. {missing[:3]}")
    engine.load_state_dict(sd)


def load_checkpoint(engine, path):
    """Resume: model + this rank's optimizer shards + step counter.  The shard layout must match (same world size)."""
    meta = json.load(open(os.path.join(path, "b200_meta.json")))
    if meta["world"] != engine.world:
        raise ValueError(f"checkpoint was written with world size {meta['world']}, engine has {engine.world} (re-sharding optimizer state is not implemented)")
    load_model(engine, os.path.join(path, "model.safetensors"))
    opt = load_file(os.path.join(path, f"optim_rank{engine.rank}.safetensors"))
    with torch.no_grad():
        for ui, u in enumerate(engine.units):
            engine.m[ui].copy_(opt[f"m.{u.name}"])
            engine.v[ui].copy_(opt[f"v.{u.name}"])
            if engine.master is not None:
                engine.master[ui].copy_(opt[f"master.{u.name}"])
    engine.step_count = int(opt["step_count"][0])
