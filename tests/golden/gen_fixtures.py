#!/usr/bin/env python
"""Generate golden parity fixtures by running the UNMODIFIED reference recipe on CPU/gloo.

Test infrastructure only.  Runs in the build container (needs /root/reference, which
does NOT exist on the GPU box); the .npz files it writes under tests/golden/ are
committed and are what the -m gpu tests and the oracle tests read.

What is recorded (per config): the initial full state_dict, every micro-batch fed to
`TrainFinetuneRecipeForNextTokenPrediction._run_train_optim_step`
(/root/reference/nemo_automodel/recipes/llm/train_ft.py:1482), the returned loss and
grad_norm of each step, the per-parameter gradients of step 0 as seen by
`scale_grads_and_clip_grad_norm` (components/training/utils.py:290, i.e. before
clipping), and the weights after the last step.

Usage:  TORCHDYNAMO_DISABLE=1 python tests/golden/gen_fixtures.py [tiny_fp32|tiny_bf16|tiny_bf16_100|all]
"""
import os, sys, tempfile, json

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _ref_env  # noqa: F401,E402  (stubs + CPU shims + sys.path for /root/reference; must precede the reference imports)
import numpy as np
import torch
import transformers
from oracle.portable_init import portable_state_dict

from nemo_automodel.components.config._arg_parser import parse_args_and_load_config
import nemo_automodel.recipes.llm.train_ft as train_ft
from nemo_automodel.recipes.llm.train_ft import TrainFinetuneRecipeForNextTokenPrediction

YAML = """
recipe: TrainFinetuneRecipeForNextTokenPrediction
seed: 1234
step_scheduler: {{global_batch_size: {gbs}, local_batch_size: {lbs}, ckpt_every_steps: 100000, num_epochs: 1, max_steps: {steps}}}
dist_env: {{backend: gloo, timeout_minutes: 5}}
model:
  _target_: nemo_automodel.NeMoAutoModelForCausalLM.from_config
  config:
    _target_: transformers.{cfg_class}
    vocab_size: {vocab}
    hidden_size: {hidden}
    intermediate_size: {ffn}
    num_hidden_layers: {layers}
    num_attention_heads: {heads}
    num_key_value_heads: {kv}
    max_position_embeddings: {seq}
    rms_norm_eps: 1.0e-5
    rope_theta: {theta}
{rope_scaling}    tie_word_embeddings: {tied}
    architectures: [{arch}]
  torch_dtype: {dtype}
  attn_implementation: sdpa
  use_liger_kernel: false
checkpoint: {{enabled: false}}
distributed: {{strategy: fsdp2, backend: gloo, dp_size: none, tp_size: 1, cp_size: 1}}
loss_fn: {{_target_: nemo_automodel.components.loss.masked_ce.MaskedCrossEntropy}}
dataset:
  _target_: nemo_automodel.components.datasets.llm.mock_iterable_dataset.MockIterableDataset
  vocab_size: {vocab}
  seq_len: {seq}
  num_samples: 100000
  batch_size: {lbs}
dataloader: {{_target_: torch.utils.data.DataLoader, batch_size: null}}
optimizer: {{_target_: torch.optim.AdamW, lr: {lr}, betas: [0.9, 0.95], eps: 1.0e-8, weight_decay: 0.1{opt_extra}}}
"""

_LLAMA = dict(cfg_class="LlamaConfig", arch="LlamaForCausalLM", tied="false", rope_scaling="")
_LLAMA3_ROPE = ("    rope_scaling: {rope_type: llama3, factor: 8.0, low_freq_factor: 1.0, high_freq_factor: 4.0, "
                "original_max_position_embeddings: 64}\n")
CONFIGS = {
    # BASELINE.json configs[0]: d_model 256, L 2, seq 512, world 1 (fp32: the exact-math pin for the oracle)
    "tiny_fp32": dict(gbs=2, lbs=2, steps=3, vocab=1024, hidden=256, ffn=512, layers=2, heads=4, kv=2, seq=512,
                      theta=10000.0, dtype="float32", lr="1.0e-3", opt_extra="", **_LLAMA),
    # same shapes in bf16 (params, compute and AdamW states bf16 as torch.optim.AdamW does on bf16 params)
    # 100 steps: the north_star's "step-loss within 1e-3 over 100 steps" curve
    "tiny_bf16": dict(gbs=2, lbs=2, steps=100, vocab=1024, hidden=256, ffn=512, layers=2, heads=4, kv=2, seq=512,
                      theta=10000.0, dtype="bfloat16", lr="1.0e-3", opt_extra="", **_LLAMA),
    # head_dim 128 / GQA 4:1 like Llama-3-8B, 2 micro-batches per step (grad accumulation), fp32
    "hd128_fp32": dict(gbs=2, lbs=1, steps=2, vocab=512, hidden=256, ffn=512, layers=2, heads=2, kv=1, seq=256,
                       theta=500000.0, dtype="float32", lr="1.0e-3", opt_extra="", **_LLAMA),
    # the 8B layer structure in bf16: head_dim 128, GQA 4:1, llama3 RoPE scaling (original context 64 < seq 256 so all three frequency
    # bands of the scaling are exercised), 2 micro-batches per step, 20 steps - the like-for-like 1e-3 check the fp32 fixture above cannot give
    "hd128_bf16": dict(gbs=2, lbs=1, steps=20, vocab=512, hidden=512, ffn=1024, layers=2, heads=4, kv=1, seq=256,
                       theta=500000.0, dtype="bfloat16", lr="1.0e-3", opt_extra="", **dict(_LLAMA, rope_scaling=_LLAMA3_ROPE)),
    # Qwen2 (components/models/qwen2/model.py: q/k/v bias, tied embeddings) through the same kernels: head_dim 64, GQA 2:1, bf16, 2 micro-batches
    "qwen2_tiny_bf16": dict(gbs=2, lbs=1, steps=20, vocab=1024, hidden=256, ffn=512, layers=2, heads=4, kv=2, seq=256,
                            theta=1000000.0, dtype="bfloat16", lr="1.0e-3", opt_extra="", cfg_class="Qwen2Config", arch="Qwen2ForCausalLM", tied="true", rope_scaling=""),
    # Mistral (HF MistralForCausalLM: the reference has no custom class for it) without sliding window: Llama parameter set, HF's RMSNorm
    # (weight multiply AFTER the down-cast) - a third family through the same kernels
    "mistral_tiny_bf16": dict(gbs=2, lbs=2, steps=20, vocab=1024, hidden=256, ffn=512, layers=2, heads=4, kv=2, seq=256,
                              theta=10000.0, dtype="bfloat16", lr="1.0e-3", opt_extra="", cfg_class="MistralConfig", arch="MistralForCausalLM", tied="false",
                              rope_scaling="    sliding_window: null\n"),
    # the same model untied, fp32: the exact-math pin of the bias path for the oracle
    "qwen2_tiny_fp32": dict(gbs=2, lbs=2, steps=3, vocab=512, hidden=256, ffn=512, layers=2, heads=2, kv=1, seq=128,
                            theta=1000000.0, dtype="float32", lr="1.0e-3", opt_extra="", cfg_class="Qwen2Config", arch="Qwen2ForCausalLM", tied="false", rope_scaling=""),
}


def _np(t):
    t = t.detach()
    if hasattr(t, "full_tensor"):
        t = t.full_tensor()
    return t.to(torch.float32).cpu().numpy()


SEED = 7


def _summ(out, prefix, name, a):
    """Compact per-parameter record: strided sample (<=4096 values) + sum + sum of squares."""
    flat = a.reshape(-1).astype(np.float64)
    stride = max(1, flat.size // 4096)
    out[f"{prefix}/{name}/sample"] = flat[::stride][:4096].astype(np.float32)
    out[f"{prefix}/{name}/stats"] = np.array([flat.sum(), (flat * flat).sum(), stride], dtype=np.float64)


def run(name):
    c = CONFIGS[name]
    with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as f:
        f.write(YAML.format(**c))
        path = f.name
    cfg = parse_args_and_load_config(path, argv=[])
    r = TrainFinetuneRecipeForNextTokenPrediction(cfg)
    r.setup()
    model = r.model_parts[0]
    # ---- both sides start from the same portable snapshot (oracle/portable_init.py)
    retied = None
    if c.get("tied") == "true" and model.lm_head.weight.data_ptr() != model.model.embed_tokens.weight.data_ptr():
        # The reference ties lm_head to the embedding when the config says so (components/models/qwen2/model.py:397-400, and again after
        # weight loading, components/checkpoint/checkpointing.py:720-722).  Under this image's transformers (5.5.0; the reference pins
        # 5.8.1) the recipe's from_config path leaves them untied, so the declared state is established here: first through the model's own
        # tie_weights(), else by sharing the parameter - exactly what tie_weights() does under the pinned version.
        model.tie_weights()
        retied = "tie_weights()"
        if model.lm_head.weight.data_ptr() != model.model.embed_tokens.weight.data_ptr():
            model.lm_head.weight = model.model.embed_tokens.weight
            retied = "parameter shared by assignment"
        # the optimizer was built from the untied parameter list: rebuild its groups from the tied model
        opt0 = r.optimizer[0]
        keep = {id(p) for p in model.parameters()}
        for g_ in opt0.param_groups:
            g_["params"] = [p for p in g_["params"] if id(p) in keep]
    sd, seen = {}, set()
    for k, v in model.state_dict().items():     # tied embeddings list the shared matrix under both names: initialise it once (first name)
        if v.data_ptr() not in seen:
            seen.add(v.data_ptr())
            sd[k] = v
    init = portable_state_dict({k: tuple(v.shape) for k, v in sd.items()}, seed=SEED)
    with torch.no_grad():
        for k, v in sd.items():
            v.copy_(torch.from_numpy(init[k]).to(v.dtype))
            assert torch.equal(v.float(), torch.from_numpy(init[k])), k   # bf16-representable => exact in both dtypes
    out = {}
    rec = {"loss": [], "grad_norm": [], "num_label_tokens": [], "lr": []}
    step_i = [0]
    snap_steps = {0, 2, c["steps"] - 1}   # weight summaries after these steps

    orig_clip = train_ft.scale_grads_and_clip_grad_norm

    def clip_spy(max_grad_norm, model_parts, **kw):
        if step_i[0] == 0:
            for n, p in model_parts[0].named_parameters():
                if p.grad is not None:
                    _summ(out, "grad0", n, _np(p.grad))
        return orig_clip(max_grad_norm, model_parts, **kw)

    train_ft.scale_grads_and_clip_grad_norm = clip_spy
    orig_step = r._run_train_optim_step

    def step_spy(batches, max_grad_norm=None):
        s = step_i[0]
        for j, b in enumerate(batches):
            ids = b["input_ids"].numpy()
            assert ids.max() < 65536
            out[f"batch/{s}/{j}/input_ids"] = ids.astype(np.uint16)
            lab = b["labels"].numpy(); pos = b["position_ids"].numpy()
            # MockIterableDataset (components/datasets/llm/mock_iterable_dataset.py:41-59): labels = shift-left ++ -100
            assert (lab[:, :-1] == ids[:, 1:]).all() and (lab[:, -1] == -100).all()
            assert (pos == np.arange(ids.shape[1])[None]).all()
        m = orig_step(batches, max_grad_norm)
        rec["loss"].append(float(m.metrics["loss"]))
        rec["grad_norm"].append(float(m.metrics["grad_norm"]))
        rec["num_label_tokens"].append(int(m.metrics["num_label_tokens"]))
        rec["lr"].append(float(m.metrics["lr"]))
        rec["max_grad_norm"] = max_grad_norm
        rec["num_micro"] = len(batches)
        if s in snap_steps:
            for k, v in model.state_dict().items():
                _summ(out, f"after{s}", k, _np(v))
        step_i[0] += 1
        return m

    r._run_train_optim_step = step_spy
    r.run_train_validation_loop()
    opt = r.optimizer[0]
    out["meta"] = np.frombuffer(json.dumps({
        "config": c, "init_seed": SEED, "loss": rec["loss"], "grad_norm": rec["grad_norm"],
        "num_label_tokens": rec["num_label_tokens"], "lr": rec["lr"], "max_grad_norm": rec.get("max_grad_norm"),
        "num_micro": rec.get("num_micro"), "snap_steps": sorted(snap_steps),
        "torch": torch.__version__, "transformers": transformers.__version__,
        "model_class": type(model).__name__, "retied": retied, "tied": bool(model.lm_head.weight.data_ptr() == model.model.embed_tokens.weight.data_ptr()), "norm_class": type(model.model.norm).__name__,
        "optimizer_class": type(opt).__name__,
        "optimizer": {k: (list(v) if isinstance(v, tuple) else v) for k, v in opt.param_groups[0].items() if k != "params"},
    }, default=str).encode(), dtype=np.uint8)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz")
    np.savez_compressed(dst, **out)
    print(name, "loss", rec["loss"][:3], "grad_norm", rec["grad_norm"][:3], "->", dst, os.path.getsize(dst) // 1024, "KiB")
    train_ft.scale_grads_and_clip_grad_norm = orig_clip


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    names = list(CONFIGS) if which == "all" else [which]
    # one recipe per process (the reference initialises a process group in setup())
    if len(names) > 1:
        import subprocess
        for n in names:
            subprocess.check_call([sys.executable, __file__, n])
    else:
        run(names[0])
