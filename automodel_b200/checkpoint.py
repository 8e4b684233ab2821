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
    """Resume: model + this rank's optimizer shards + step counter.  A checkpoint written at another world size is re-sharded: the flat
    optimizer state of a unit is the concatenation of the old ranks' shards (minus the old padding), of which this rank takes its new
    contiguous slice (the reference gets the same from DCP's resharding load, components/checkpoint/checkpointing.py:382-572)."""
    from safetensors import safe_open
    meta = json.load(open(os.path.join(path, "b200_meta.json")))
    load_model(engine, os.path.join(path, "model.safetensors"))
    old_world = int(meta["world"])
    kinds = ["m", "v"] + (["master"] if engine.master is not None else [])
    dst = {"m": engine.m, "v": engine.v, "master": engine.master}
    if old_world == engine.world:
        opt = load_file(os.path.join(path, f"optim_rank{engine.rank}.safetensors"))
        with torch.no_grad():
            for ui, u in enumerate(engine.units):
                for k in kinds:
                    dst[k][ui].copy_(opt[f"{k}.{u.name}"])
        engine.step_count = int(opt["step_count"][0])
        return
    old_padded = {name: padded for name, _, padded in meta["units"]}
    files = [safe_open(os.path.join(path, f"optim_rank{r}.safetensors"), framework="pt") for r in range(old_world)]
    with torch.no_grad():
        for ui, u in enumerate(engine.units):
            a, b = u.shard_range(engine.rank, engine.world)          # this rank's new slice of the unit's flat index space
            per_old = old_padded[u.name] // old_world
            for k in kinds:
                key = f"{k}.{u.name}"
                if key not in files[0].keys():
                    raise KeyError(f"checkpoint has no {key} (written without fp32 master weights?)")
                out = dst[k][ui]
                out.zero_()                                          # elements beyond u.numel are padding
                lo, hi = a, min(b, u.numel)
                pos = lo
                while pos < hi:
                    r_old, off = divmod(pos, per_old)
                    n = min(hi - pos, per_old - off)
                    out[pos - a:pos - a + n].copy_(files[r_old].get_slice(key)[off:off + n])
                    pos += n
    engine.step_count = int(files[0].get_tensor("step_count")[0])
