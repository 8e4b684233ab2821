"""The drop-in claim, end to end: the reference's own recipe (unmodified, imported from /root/reference) trains through this
repository's `distributed.strategy: b200_sharded` + optimizer `_target_` after `automodel_b200.integration.register()`, and
reproduces the loss / grad-norm curve the same recipe produced with FSDP2 (tests/golden/*.npz).  CPU only (stand-in kernels): this
tests the boundary, the kernels are tested on the GPU.  Skipped where the reference is not present (e.g. the GPU box)."""
import json
import os
import subprocess
import sys
import tempfile

import pytest

from tests.golden_utils import load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("B200_REFERENCE_PATH", "/root/reference")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "nemo_automodel")), reason="reference checkout not present")


def _run(name, loss_kind, steps, **extra_env):
    env = dict(os.environ, PYTHONPATH=ROOT, TORCHDYNAMO_DISABLE="1", MASTER_ADDR="127.0.0.1", **extra_env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "run_reference_recipe_b200.py"), name, loss_kind, str(steps)],
                       capture_output=True, text=True, timeout=900, env=env, cwd=tempfile.mkdtemp(prefix="b200_dropin_"))
    lines = [l for l in r.stdout.splitlines() if l.startswith("B200_DROPIN_RESULT ")]
    assert r.returncode == 0 and lines, (r.stdout[-2000:], r.stderr[-4000:])
    return json.loads(lines[-1][len("B200_DROPIN_RESULT "):])


def _run_world2(name, loss_kind, steps, **extra_env):
    return _run_world(2, name, loss_kind, steps, **extra_env)


def _run_world(world, name, loss_kind, steps, **extra_env):
    """`world` ranks of the runner over gloo; returns their result records."""
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = []
    scratch = tempfile.mkdtemp(prefix="b200_dropin_")     # the recipe writes checkpoints/*.jsonl metric logs into its working directory
    for rank in range(world):
        env = dict(os.environ, PYTHONPATH=ROOT, TORCHDYNAMO_DISABLE="1", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                   LOCAL_RANK=str(rank), WORLD_SIZE=str(world), **extra_env)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "golden", "run_reference_recipe_b200.py"), name, loss_kind, str(steps)],
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=scratch))
    recs = []
    for p in procs:
        out, err = p.communicate(timeout=900)
        lines = [l for l in out.splitlines() if l.startswith("B200_DROPIN_RESULT ")]
        assert p.returncode == 0 and lines, (out[-2000:], err[-4000:])
        recs.append(json.loads(lines[-1][len("B200_DROPIN_RESULT "):]))
    return recs


@pytest.mark.parametrize("name,loss_kind,steps,loss_tol", [("hd128_fp32", "reference_loss", 2, 4e-3), ("tiny_bf16", "fused_loss", 8, 1e-3),
                                                          ("qwen2_tiny_fp32", "reference_loss", 3, 4e-3), ("qwen2_tiny_bf16", "fused_loss", 8, 1e-3)])
def test_reference_recipe_trains_through_b200_strategy(name, loss_kind, steps, loss_tol):
    """Training curve vs the FSDP2 fixture; the fused-loss run also goes through the recipe's validation loop after every step
    (model.eval(), is_train=False, loss_fn called with num_label_tokens=None = plain sum)."""
    val = loss_kind == "fused_loss"
    rec = _run(name, loss_kind, steps, **({"B200_DROPIN_VAL": "1"} if val else {}))
    _, meta = load(name)
    assert rec["model_class"] == "B200CausalLM" and rec["optimizer_class"] == "B200FusedAdamW"
    assert rec["loss_class"] == ("B200MaskedCrossEntropy" if loss_kind == "fused_loss" else "MaskedCrossEntropy")
    assert rec["recipe_init"][0] > 0 and rec["recipe_init"][1] > 0, "the recipe's own (meta-device -> initialize_weights) init left zero tensors"
    assert all(rec["ids_match"]), "the reference data loader fed different batches than in the fixture run"
    assert rec["max_grad_norm"] == meta["max_grad_norm"]
    n = len(rec["loss"])
    assert n == min(steps, len(meta["loss"]))
    for s in range(n):
        assert rec["num_label_tokens"][s] == meta["num_label_tokens"][s]
        assert abs(rec["loss"][s] - meta["loss"][s]) < loss_tol, (s, rec["loss"][s], meta["loss"][s])
        assert abs(rec["grad_norm"][s] - meta["grad_norm"][s]) < 2e-2 * meta["grad_norm"][s], (s, rec["grad_norm"][s], meta["grad_norm"][s])
    if val:
        c = meta["config"]
        assert rec["val_tokens"] == [3 * c["lbs"] * (c["seq"] - 1)] * n and not any(rec["training_flag_during_val"])
        # random-init model on random tokens: the held-out loss sits at ln(vocab) like the training loss, resolved to better than 1e-3
        assert all(abs(v - t) < 0.1 for v, t in zip(rec["val_loss"], rec["loss"])), (rec["val_loss"], rec["loss"])
        assert len({round(v, 3) for v in rec["val_loss"]}) > 1, "validation loss is quantised (accumulator precision)"


@pytest.mark.parametrize("gbs,sync_hook", [(4, True), (4, False)], ids=["ga2_sync_hook", "ga2_lazy_reduce_scatter"])
def test_reference_recipe_world2_equals_single_rank_accumulation(tmp_path, gbs, sync_hook):
    """Two reference-recipe processes (gloo) over the b200_sharded strategy: the recipe's data-parallel arithmetic - global label-token
    count, `(local_loss * dp_group_size).backward()`, its loss all-reduce, the clip utility on sharded gradients - must give the step
    a single rank gives when it accumulates the same micro-batches (what the two ranks were fed is recorded and replayed).
    With two micro-batches per rank the reduce-scatter must happen once, after the last one: either announced by the recipe's
    get_sync_ctx (patched to recognise the facade) or, without that hook, performed lazily when the clip utility asks for the norm."""
    recs = _run_world2("hd128_fp32", "reference_loss", 2, B200_DROPIN_DUMP=str(tmp_path / "fed"), B200_DROPIN_GBS=str(gbs),
                       B200_DROPIN_NO_SYNC_HOOK="0" if sync_hook else "1")
    assert recs[0]["loss"] == recs[1]["loss"] and recs[0]["grad_norm"] == recs[1]["grad_norm"]
    assert recs[0]["num_micro"] == [gbs // 2] * len(recs[0]["loss"])
    # the recipe's own initialisation (meta device -> parallelize -> Checkpointer.initialize_model_weights -> facade.initialize_weights):
    # non-zero and identical on both ranks
    assert recs[0]["recipe_init"] == recs[1]["recipe_init"] and recs[0]["recipe_init"][0] > 0 and recs[0]["recipe_init"][1] > 0
    _replay_and_compare(tmp_path, recs, 2)


def _replay_and_compare(tmp_path, recs, world):
    """One rank (CPU stand-in kernels) accumulates every micro-batch the `world` recipe ranks were fed; loss and grad norm must agree."""
    import numpy as np
    import torch
    from automodel_b200.engine import ShardedLlamaEngine
    from tests import cpu_kernels
    from tests.golden_utils import model_cfg, init_params
    fed = [np.load(str(tmp_path / f"fed.rank{r}.npz")) for r in range(world)]
    assert not np.array_equal(fed[0]["0/0/input_ids"], fed[1]["0/0/input_ids"]), "both ranks were fed the same samples"
    _, meta = load("hd128_fp32")
    oc = meta["optimizer"]
    eng = ShardedLlamaEngine(model_cfg(meta), "cpu", max_tokens=meta["config"]["lbs"] * meta["config"]["seq"], lr=oc["lr"], betas=tuple(oc["betas"]),
                             eps=oc["eps"], weight_decay=oc["weight_decay"], adam_mode=1, ops=cpu_kernels)
    eng.load_state_dict(init_params(meta))
    for s in range(len(recs[0]["loss"])):
        mbs = []
        for f in fed:
            j = 0
            while f"{s}/{j}/input_ids" in f:
                mbs.append({"input_ids": torch.from_numpy(f[f"{s}/{j}/input_ids"]), "labels": torch.from_numpy(f[f"{s}/{j}/labels"])})
                j += 1
        loss, gn = eng.train_step(mbs, meta["max_grad_norm"])
        assert abs(float(loss) - recs[0]["loss"][s]) < 1e-3, (s, float(loss), recs[0]["loss"][s])
        assert abs(float(gn) - recs[0]["grad_norm"][s]) < 5e-3 * float(gn), (s, float(gn), recs[0]["grad_norm"][s])


def test_reference_recipe_hsdp_2x2(tmp_path):
    """`distributed.dp_replicate_size: 2` on four ranks (the reference's HSDP mesh (dp_replicate, dp_shard) = (2, 2)): the facade shards
    inside dp_shard, all-reduces gradient shards across dp_replicate, and scales by the recipe's full dp_group_size."""
    recs = _run_world(4, "hd128_fp32", "reference_loss", 2, B200_DROPIN_DUMP=str(tmp_path / "fed"), B200_DROPIN_GBS="4", B200_DROPIN_REPLICATE="2")
    assert all((r["world"], r["replicas"]) == (2, 2) for r in recs)
    assert all(r["loss"] == recs[0]["loss"] and r["grad_norm"] == recs[0]["grad_norm"] for r in recs)
    _replay_and_compare(tmp_path, recs, 4)


def test_reference_checkpointer_saves_and_resumes_the_b200_strategy(tmp_path):
    """The reference's own Checkpointer (recipes/base_recipe.py save_checkpoint / load_checkpoint: DCP get/set_model_state_dict and
    get/set_optimizer_state_dict, safetensors + consolidated HF export) on the facade: the HF-shaped module tree gives it the
    reference's FQNs, B200FusedAdamW exposes step / exp_avg / exp_avg_sq per parameter.  After `restore_from` the weights, both Adam
    moments and the step counter are exactly those of the uninterrupted run at that step."""
    import shutil
    from safetensors.torch import load_file
    ck = tmp_path / "ck"
    full = _run("tiny_bf16", "reference_loss", 4, B200_DROPIN_CKPT=str(ck))
    assert [p[3] for p in full["pre_state"]] == [0, 1, 2, 3]
    saved = sorted(d for d in os.listdir(ck) if d.startswith("epoch_"))
    assert saved == ["epoch_0_step_1", "epoch_0_step_3"], saved
    sd = load_file(str(ck / "epoch_0_step_1" / "model" / "consolidated" / "model-00001-of-00001.safetensors"))
    _, meta = load("tiny_bf16")
    from tests.golden_utils import init_params
    assert set(sd) == set(init_params(meta)), "consolidated export does not carry the HF Llama names"
    shutil.rmtree(ck / "epoch_0_step_3")
    for f in ("LATEST", "latest"):
        if os.path.lexists(ck / f):
            os.remove(ck / f)
    res = _run("tiny_bf16", "reference_loss", 4, B200_DROPIN_CKPT=f"{ck}:epoch_0_step_1")
    assert len(res["loss"]) == 2, "the resumed run should execute steps 2 and 3 only"
    assert res["pre_state"][0] == full["pre_state"][2], (res["pre_state"][0], full["pre_state"][2])


def test_reference_checkpointer_world2_restores_sharded_optimizer_state(tmp_path):
    """Same at world size 2: the Adam moments live as 1/N flat shards per rank; for the reference's DCP optimizer state dict they are
    all-gathered into per-parameter tensors at save time and sliced back into the shards on load.  Every rank's shard fingerprints
    after `restore_from` equal the uninterrupted run's."""
    import shutil
    ck = tmp_path / "ck"
    full = _run_world2("tiny_bf16", "reference_loss", 4, B200_DROPIN_CKPT=str(ck), B200_DROPIN_GBS="4")
    assert full[0]["pre_state"][2][1] != full[1]["pre_state"][2][1], "both ranks report the same optimizer shard"
    shutil.rmtree(ck / "epoch_0_step_3")
    for f in ("LATEST", "latest"):
        if os.path.lexists(ck / f):
            os.remove(ck / f)
    res = _run_world2("tiny_bf16", "reference_loss", 4, B200_DROPIN_CKPT=f"{ck}:epoch_0_step_1", B200_DROPIN_GBS="4")
    for r in range(2):
        assert len(res[r]["loss"]) == 2
        assert res[r]["pre_state"][0] == full[r]["pre_state"][2], (r, res[r]["pre_state"][0], full[r]["pre_state"][2])


def test_reference_benchmark_recipe_runs_over_the_b200_strategy():
    """BenchmarkingRecipeForNextTokenPrediction (recipes/llm/benchmark.py: the reference's throughput harness - its own timers, no clip
    utility, loss_fn called with num_label_tokens=None) over the facade with the fused loss: four iterations, four optimizer steps, and
    the per-iteration loss it prints follows the FSDP2 fixture curve (Adam is invariant to the missing normalisation and clip)."""
    rec = _run("tiny_bf16", "fused_loss", 4, B200_DROPIN_RECIPE="benchmark")
    _, meta = load("tiny_bf16")
    assert rec["model_class"] == "B200CausalLM" and rec["optimizer_class"] == "B200FusedAdamW" and rec["loss_class"] == "B200MaskedCrossEntropy"
    assert rec["engine_steps"] == 4 and len(rec["loss"]) == 4
    for s in range(4):
        assert abs(rec["loss"][s] - meta["loss"][s]) < 2e-3, (s, rec["loss"][s], meta["loss"][s])     # printed with 4 decimals


def test_from_pretrained_through_the_reference_loader(tmp_path):
    """SURVEY §8f N1, load direction: `NeMoAutoModelForCausalLM.from_pretrained(<HF checkpoint>)` + `strategy: b200_sharded` -
    every tensor of the checkpoint arrives bit-exactly in the flat unit buffers (HF names are views of them)."""
    import torch
    import transformers
    _, meta = load("hd128_fp32")
    c = meta["config"]
    cfg = transformers.LlamaConfig(vocab_size=c["vocab"], hidden_size=c["hidden"], intermediate_size=c["ffn"], num_hidden_layers=c["layers"],
                                   num_attention_heads=c["heads"], num_key_value_heads=c["kv"], max_position_embeddings=c["seq"],
                                   rms_norm_eps=1e-5, rope_theta=c["theta"], tie_word_embeddings=False)
    torch.manual_seed(3)
    transformers.LlamaForCausalLM(cfg).to(torch.bfloat16).save_pretrained(str(tmp_path / "hf"))
    rec = _run("hd128_fp32", "reference_loss", 1, B200_DROPIN_PRETRAINED=str(tmp_path / "hf"))
    assert rec["model_class"] == "B200CausalLM" and rec["pretrained_tensors"] == 2 + 9 * c["layers"] + 1
    assert rec["pretrained_mismatch"] == []


@pytest.mark.parametrize("loss_kind", ["reference_loss", "fused_loss"])
def test_unsupported_config_runs_on_the_reference_path(loss_kind):
    """north_star: "any HF config the reference accepts runs unchanged".  A Llama variant the engine refuses (bias on every attention
    projection) under `strategy: b200_sharded` + the B200 optimizer / loss `_target_`s trains on the reference's OWN model and FSDP2 path:
    the strategy hands the model back, B200FusedAdamW degrades to torch.optim.AdamW, B200MaskedCrossEntropy to the reference formula."""
    res = _run("tiny_bf16", loss_kind, 4, B200_DROPIN_FALLBACK="attention_bias: true")
    assert res["model_class"] == "LlamaForCausalLM" and not res["has_engine"]
    assert res["optimizer_class"] == "AdamW"
    assert len(res["loss"]) == 4 and all(6.5 < l < 7.5 for l in res["loss"])
