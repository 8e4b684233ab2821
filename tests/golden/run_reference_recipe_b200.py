#!/usr/bin/env python
"""Drive this repository's `b200_sharded` strategy with the UNMODIFIED reference recipe
(`TrainFinetuneRecipeForNextTokenPrediction`, /root/reference/nemo_automodel/recipes/llm/train_ft.py) on CPU.

Test infrastructure (needs /root/reference; run as a subprocess by tests/test_reference_dropin.py because tests/golden/_ref_env.py
patches torch globally).  Same YAML as the fixture generator (tests/golden/gen_fixtures.py) except for exactly the lines
INTEGRATION.md names: `distributed.strategy: b200_sharded`, the optimizer `_target_`, and optionally the loss `_target_`.
The engine runs on the CPU stand-in kernels (tests/cpu_kernels.py): what is under test is the boundary - that the reference's own
setup(), data loader, gradient-accumulation loop, loss, clip utility, optimizer/scheduler calls and metric logging work against the
facade unchanged and reproduce the curve the reference produced with FSDP2.

usage: run_reference_recipe_b200.py <fixture name> <reference_loss|fused_loss> <max steps>   -> one JSON line on stdout
"""
import json, os, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_env  # noqa: F401,E402
import torch  # noqa: E402

from tests import cpu_kernels  # noqa: E402
from tests.golden_utils import load, init_params  # noqa: E402
import automodel_b200.integration as b200  # noqa: E402

b200.register(ops=cpu_kernels, device=torch.device("cpu"), patch_sync_ctx=os.environ.get("B200_DROPIN_NO_SYNC_HOOK") != "1")

from nemo_automodel.components.config._arg_parser import parse_args_and_load_config  # noqa: E402
from nemo_automodel.recipes.llm.train_ft import TrainFinetuneRecipeForNextTokenPrediction  # noqa: E402
import gen_fixtures as gf  # noqa: E402


def main(name, loss_kind, steps):
    c = dict(gf.CONFIGS[name])
    c["steps"] = min(int(steps), c["steps"])
    if os.environ.get("B200_DROPIN_GBS"):
        c["gbs"] = int(os.environ["B200_DROPIN_GBS"])     # more micro-batches per step than the fixture run (gradient accumulation per rank)
    y = gf.YAML.format(**c)
    y = y.replace("strategy: fsdp2", "strategy: b200_sharded, max_tokens: %d, reference_rounding: true" % (c["lbs"] * c["seq"]))
    y = y.replace("_target_: torch.optim.AdamW", "_target_: automodel_b200.recipe.B200FusedAdamW")
    if os.environ.get("B200_DROPIN_REPLICATE"):    # HSDP: distributed.dp_replicate_size
        y = y.replace("dp_size: none", "dp_size: none, dp_replicate_size: " + os.environ["B200_DROPIN_REPLICATE"])
        assert "dp_replicate_size" in y
    if os.environ.get("B200_DROPIN_VAL"):          # a validation pass every step (the recipe's eval loop: model.eval(), is_train=False)
        y = y.replace("max_steps: %d}" % c["steps"], "max_steps: %d, val_every_steps: 1}" % c["steps"])
        y += (
            "validation_dataset:\n  _target_: nemo_automodel.components.datasets.llm.mock_iterable_dataset.MockIterableDataset\n"
            "  vocab_size: %d\n  seq_len: %d\n  num_samples: 3\n  batch_size: %d\n"
            "validation_dataloader: {_target_: torch.utils.data.DataLoader, batch_size: null}\n" % (c["vocab"], c["seq"], c["lbs"]))
        assert "val_every_steps" in y
    if loss_kind == "fused_loss":
        y = y.replace("_target_: nemo_automodel.components.loss.masked_ce.MaskedCrossEntropy", "_target_: automodel_b200.recipe.B200MaskedCrossEntropy")
    assert "b200_sharded" in y and "B200FusedAdamW" in y
    if os.environ.get("B200_DROPIN_FALLBACK"):       # e.g. "attention_bias: true": a Llama variant the engine refuses
        y = y.replace("    tie_word_embeddings: false\n", "    tie_word_embeddings: false\n    " + os.environ["B200_DROPIN_FALLBACK"] + "\n")
        assert os.environ["B200_DROPIN_FALLBACK"] in y
    ck = os.environ.get("B200_DROPIN_CKPT")     # "<dir>" or "<dir>:<restore_from>": the reference's own Checkpointer, every 2 steps
    if ck:
        ck_dir, _, restore = ck.partition(":")
        y = y.replace("checkpoint: {enabled: false}", "checkpoint: {enabled: true, checkpoint_dir: %s, model_save_format: safetensors, save_consolidated: true%s}"
                      % (ck_dir, (", restore_from: " + restore) if restore else ""))
        y = y.replace("ckpt_every_steps: 100000", "ckpt_every_steps: 2")
        assert "enabled: true" in y and "ckpt_every_steps: 2" in y
    with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as f:
        f.write(y)
        path = f.name
    pre = os.environ.get("B200_DROPIN_PRETRAINED")      # model section = NeMoAutoModelForCausalLM.from_pretrained(<local HF checkpoint dir>)
    if pre:
        i, j = y.index("model:\n"), y.index("checkpoint:")
        y = (y[:i] + "model:\n  _target_: nemo_automodel.NeMoAutoModelForCausalLM.from_pretrained\n  pretrained_model_name_or_path: %s\n"
             "  torch_dtype: bfloat16\n  attn_implementation: sdpa\n  use_liger_kernel: false\n" % pre + y[j:])
        with open(path, "w") as f:
            f.write(y)
    bench = os.environ.get("B200_DROPIN_RECIPE") == "benchmark"     # the reference's BenchmarkingRecipeForNextTokenPrediction instead
    if bench:
        with open(path, "a") as f:
            f.write("benchmark: {warmup_steps: 1, peak_tflops: 1471.7, nsys_start: -1, nsys_end: -1, nsys_ranks: []}\n")
    cfg = parse_args_and_load_config(path, argv=[])
    if bench:
        from nemo_automodel.recipes.llm.benchmark import BenchmarkingRecipeForNextTokenPrediction
        r = BenchmarkingRecipeForNextTokenPrediction(cfg)
    else:
        r = TrainFinetuneRecipeForNextTokenPrediction(cfg)
    r.setup()
    model = r.model_parts[0]
    if os.environ.get("B200_DROPIN_FALLBACK"):
        # a config the engine does not implement: the strategy must have routed the model to the reference's own path
        losses = []
        orig_fb = r._run_train_optim_step

        def spy_fb(batches, max_grad_norm=None):
            m = orig_fb(batches, max_grad_norm)
            losses.append(float(m.metrics["loss"]))
            return m

        r._run_train_optim_step = spy_fb
        r.run_train_validation_loop()
        sys.stdout.write("\nB200_DROPIN_RESULT " + json.dumps({"loss": losses, "model_class": type(model).__name__, "optimizer_class": type(r.optimizer[0]).__name__,
                                                               "loss_class": type(r.loss_fn).__name__, "has_engine": hasattr(model, "engine")}) + "\n")
        return
    z, meta = load(name)
    # what the recipe's own initialisation left in the flat buffers (from_config: Checkpointer.initialize_model_weights, before or after
    # sharding depending on the world size)
    init_fp = [float(sum(p.double().abs().sum() for p in model.engine.p_full)), float(min(float(p.float().abs().max()) for n_, p in model.engine.P.items() if not n_.endswith(".bias")))]     # (HF zero-initialises biases)
    if pre:
        from safetensors.torch import load_file
        want = {}
        for fn in sorted(os.listdir(pre)):
            if fn.endswith(".safetensors"):
                want.update(load_file(os.path.join(pre, fn)))
        got = model.state_dict()
        mism = [k for k in want if k not in got or not torch.equal(got[k].cpu(), want[k])]
        sys.stdout.write("\nB200_DROPIN_RESULT " + json.dumps({"pretrained_tensors": len(want), "pretrained_mismatch": mism,
                                                               "model_class": type(model).__name__}) + "\n")
        return
    if not (ck and restore):
        model.engine.load_state_dict(init_params(meta))     # the snapshot the fixture run started from
    rec = {"loss": [], "grad_norm": [], "num_label_tokens": [], "ids_match": [], "pre_state": [], "recipe_init": init_fp}
    dump = {}
    step_i = [0]
    orig = r._run_train_optim_step

    def spy(batches, max_grad_norm=None):
        s = step_i[0]
        rec["ids_match"].append(all(f"batch/{s}/{j}/input_ids" in z and b["input_ids"].shape == z[f"batch/{s}/{j}/input_ids"].shape and bool((b["input_ids"].numpy() == z[f"batch/{s}/{j}/input_ids"]).all())
                                    for j, b in enumerate(batches)))
        rec.setdefault("num_micro", []).append(len(batches))
        e = model.engine     # exact fingerprints of weights / Adam moments / step counter before this step
        rec["pre_state"].append([float(sum(p.double().abs().sum() for p in e.p_full)), float(sum(x.double().abs().sum() for x in e.m)),
                                 float(sum(x.double().abs().sum() for x in e.v)), e.step_count])
        for j, b in enumerate(batches):
            dump[f"{s}/{j}/input_ids"] = b["input_ids"].numpy().copy()
            dump[f"{s}/{j}/labels"] = b["labels"].numpy().copy()
        m = orig(batches, max_grad_norm)
        rec["loss"].append(float(m.metrics["loss"])); rec["grad_norm"].append(float(m.metrics["grad_norm"]))
        rec["num_label_tokens"].append(int(m.metrics["num_label_tokens"]))
        rec["max_grad_norm"] = max_grad_norm
        step_i[0] += 1
        return m

    if bench:
        # no _run_train_optim_step here: the benchmark loop calls _forward_backward_step / optimizer.step() itself (no clip utility,
        # loss_fn with num_label_tokens=None) and prints "num_label_tokens=... | loss=..." per iteration
        import contextlib, io, re
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            r.run_benchmark()
        rec["loss"] = [float(x) for x in re.findall(r"loss=([0-9.]+)", buf.getvalue())]
        rec["engine_steps"] = model.engine.step_count
        rec["model_class"] = type(model).__name__
        rec["optimizer_class"] = type(r.optimizer[0]).__name__
        rec["loss_class"] = type(r.loss_fn).__name__
        sys.stdout.write("\nB200_DROPIN_RESULT " + json.dumps(rec) + "\n")
        return
    r._run_train_optim_step = spy
    if os.environ.get("B200_DROPIN_VAL"):
        orig_val = r._run_validation_epoch

        def val_spy(dl):
            m = orig_val(dl)
            rec.setdefault("val_loss", []).append(float(m.metrics["val_loss"]))
            rec.setdefault("val_tokens", []).append(int(m.metrics["num_label_tokens"]))
            rec.setdefault("training_flag_during_val", []).append(bool(model.training))
            return m

        r._run_validation_epoch = val_spy
    r.run_train_validation_loop()
    if os.environ.get("B200_DROPIN_DUMP"):   # per-rank record of what the reference's data loader fed (world-size > 1 check)
        import numpy as np
        np.savez(os.environ["B200_DROPIN_DUMP"] + f".rank{int(os.environ.get('RANK', '0'))}.npz", **dump)
    rec["world"], rec["replicas"] = model.engine.world, model.engine.replicas
    rec["model_class"] = type(model).__name__
    rec["optimizer_class"] = type(r.optimizer[0]).__name__
    rec["loss_class"] = type(r.loss_fn).__name__
    sys.stdout.write("\nB200_DROPIN_RESULT " + json.dumps(rec) + "\n")


if __name__ == "__main__":
    main(*sys.argv[1:4])
