#!/usr/bin/env python
"""Time the UNMODIFIED reference recipe (`TrainFinetuneRecipeForNextTokenPrediction`, baseline/_ref install of /root/reference) on the
host cores (CPU/gloo, world size 1) on a BOUNDED sample of the benchmark workload, for bench.py's `cpu_baseline` / `--impl reference`:

  Llama-3-8B layer dimensions (hidden 4096, ffn 14336, 32/8 heads of 128, llama3 RoPE), ONE decoder layer, vocab 2048, seq 512, b=1,
  fp32 (bf16 GEMMs on a CPU without AMX run an order of magnitude below the fp32 path, which would flatter the GPU), MaskedCrossEntropy,
  AdamW(lr 1e-5, betas 0.9/0.95, wd 0.1), clip 1.0, MockIterableDataset.  Full optimizer steps, every host thread torch can use.

Runs in its own process because tests/golden/_ref_env.py patches torch globally (CPU shims for the recipe's CUDA-hard-coded call
sites).  Prints one JSON line: per-step seconds after warm-up, threads, model FLOPs per step."""
import json, os, sys, tempfile, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("B200_REFERENCE_PATH", os.path.join(ROOT, "baseline", "_ref"))
os.environ["TORCHDYNAMO_DISABLE"] = "1"      # Float32RMSNorm's @torch.compile needs Inductor's C++/OpenMP toolchain on CPU
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import _ref_env  # noqa: F401,E402
import torch  # noqa: E402
from nemo_automodel.components.config._arg_parser import parse_args_and_load_config  # noqa: E402
from nemo_automodel.recipes.llm.train_ft import TrainFinetuneRecipeForNextTokenPrediction  # noqa: E402

S, V, L = 512, 2048, 1
YAML = f"""
recipe: TrainFinetuneRecipeForNextTokenPrediction
seed: 1234
step_scheduler: {{global_batch_size: 1, local_batch_size: 1, ckpt_every_steps: 100000, num_epochs: 1, max_steps: STEPS}}
dist_env: {{backend: gloo, timeout_minutes: 5}}
model:
  _target_: nemo_automodel.NeMoAutoModelForCausalLM.from_config
  config:
    _target_: transformers.LlamaConfig
    vocab_size: {V}
    hidden_size: 4096
    intermediate_size: 14336
    num_hidden_layers: {L}
    num_attention_heads: 32
    num_key_value_heads: 8
    max_position_embeddings: 8192
    rms_norm_eps: 1.0e-5
    rope_theta: 500000.0
    rope_scaling: {{rope_type: llama3, factor: 8.0, low_freq_factor: 1.0, high_freq_factor: 4.0, original_max_position_embeddings: 8192}}
    tie_word_embeddings: false
    architectures: [LlamaForCausalLM]
  torch_dtype: float32
  attn_implementation: sdpa
  use_liger_kernel: false
checkpoint: {{enabled: false}}
distributed: {{strategy: fsdp2, backend: gloo, dp_size: none, tp_size: 1, cp_size: 1}}
loss_fn: {{_target_: nemo_automodel.components.loss.masked_ce.MaskedCrossEntropy}}
dataset:
  _target_: nemo_automodel.components.datasets.llm.mock_iterable_dataset.MockIterableDataset
  vocab_size: {V}
  seq_len: {S}
  num_samples: 100000
  batch_size: 1
dataloader: {{_target_: torch.utils.data.DataLoader, batch_size: null}}
optimizer: {{_target_: torch.optim.AdamW, lr: 1.0e-5, betas: [0.9, 0.95], eps: 1.0e-8, weight_decay: 0.1}}
"""


def main(steps, warmup, budget_s):
    with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as f:
        f.write(YAML.replace("STEPS", str(steps + warmup)))
        path = f.name
    os.chdir(tempfile.mkdtemp(prefix="ref_cpu_"))      # the recipe writes metric logs into its working directory
    cfg = parse_args_and_load_config(path, argv=[])
    r = TrainFinetuneRecipeForNextTokenPrediction(cfg)
    r.setup()
    times, losses = [], []
    orig = r._run_train_optim_step
    t_start = time.perf_counter()

    class _Stop(Exception):
        pass

    def spy(batches, max_grad_norm=None):
        t0 = time.perf_counter()
        m = orig(batches, max_grad_norm)
        times.append(time.perf_counter() - t0)
        losses.append(float(m.metrics["loss"]))
        if len(times) > warmup and time.perf_counter() - t_start > budget_s:
            raise _Stop()
        return m

    r._run_train_optim_step = spy
    try:
        r.run_train_validation_loop()
    except _Stop:
        pass
    timed = times[warmup:] if len(times) > warmup else times[-1:]
    h, ffn, heads, kv = 4096, 14336, 32, 8
    f_tok = L * h * h * (12 + 12 * kv / heads + 18 * ffn / h + 6 * S / h + 6 * V / (L * h))     # components/utils/flops_utils.py:51-78
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count()
    sys.stdout.write("\nREF_CPU_RESULT " + json.dumps({
        "step_s": timed, "mean_step_s": sum(timed) / len(timed), "steps_timed": len(timed), "warmup": min(warmup, len(times) - len(timed)),
        "threads": torch.get_num_threads(), "cores": cores, "flops_per_step": f_tok * S, "seq": S, "vocab": V, "layers": L,
        "model_class": type(r.model_parts[0]).__name__, "optimizer_class": type(r.optimizer[0]).__name__, "loss": losses[:3],
        "torch": torch.__version__}) + "\n")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 5, int(sys.argv[2]) if len(sys.argv) > 2 else 1, float(sys.argv[3]) if len(sys.argv) > 3 else 60.0)
