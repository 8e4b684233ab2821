#!/usr/bin/env python
"""Run the UNMODIFIED reference recipe (`TrainFinetuneRecipeForNextTokenPrediction`, recipes/llm/train_ft.py:1482-1635) on the GPU box,
either over its own DTensor/FSDP2 path (`--strategy fsdp2`: cuBLAS + SDPA + NCCL, the number to beat and the parity oracle of
north_star) or over this repository's strategy (`--strategy b200_sharded`: the same YAML with exactly the lines INTEGRATION.md names
changed - `distributed.strategy`, the optimizer `_target_`, optionally the loss `_target_`).  Measurement infrastructure: imports the
reference from baseline/_ref (git-ignored install, travels with gpurun), never from /root/reference.

Both arms start from the same deterministic initial weights (`det_init`: per-parameter Philox stream keyed by the HF name, generated on
the GPU - bit-identical across processes on the same torch build) and, because both run the reference's own MockIterableDataset /
StepScheduler with the same seed, see byte-identical batches (a checksum per step is recorded and compared).

  python tools/ref_gpu_run.py --strategy fsdp2 --config 8b --steps 100 --out gpurun_out/ref_8b_n1.json
  torchrun --nproc-per-node 8 ... tools/ref_gpu_run.py --strategy b200_sharded --config 8b --steps 100 --out gpurun_out/b200_8b_n8.json
  python tools/ref_gpu_run.py --compare gpurun_out/ref_8b_n1.json gpurun_out/b200_8b_n1.json      -> markdown table + pass/fail
"""
import argparse
import json
import os
import sys
import tempfile
import time
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

YAML = """
recipe: TrainFinetuneRecipeForNextTokenPrediction
seed: 1234
step_scheduler: {{global_batch_size: {gbs}, local_batch_size: {lbs}, ckpt_every_steps: 100000, num_epochs: 1, max_steps: {steps}}}
dist_env: {{backend: nccl, timeout_minutes: 10}}
model:
  _target_: nemo_automodel.NeMoAutoModelForCausalLM.from_config
  config:
    _target_: transformers.LlamaConfig
    vocab_size: {vocab}
    hidden_size: {hidden}
    intermediate_size: {ffn}
    num_hidden_layers: {layers}
    num_attention_heads: {heads}
    num_key_value_heads: {kv}
    max_position_embeddings: {max_pos}
    rms_norm_eps: 1.0e-5
    rope_theta: {theta}
{rope_scaling}    tie_word_embeddings: false
    architectures: [LlamaForCausalLM]
  torch_dtype: {dtype}
  attn_implementation: {attn}
  use_liger_kernel: false
checkpoint: {{enabled: false}}
distributed: {{strategy: fsdp2, dp_size: none, tp_size: 1, cp_size: 1}}
loss_fn: {{_target_: nemo_automodel.components.loss.masked_ce.MaskedCrossEntropy}}
dataset:
  _target_: nemo_automodel.components.datasets.llm.mock_iterable_dataset.MockIterableDataset
  vocab_size: {vocab}
  seq_len: {seq}
  num_samples: 1000000
  batch_size: {lbs}
dataloader: {{_target_: torch.utils.data.DataLoader, batch_size: null}}
optimizer: {{_target_: torch.optim.AdamW, lr: {lr}, betas: [0.9, 0.95], eps: 1.0e-8, weight_decay: 0.1}}
"""

LLAMA3_ROPE = ("    rope_scaling: {rope_type: llama3, factor: 8.0, low_freq_factor: 1.0, high_freq_factor: 4.0, "
               "original_max_position_embeddings: 8192}\n")

CONFIGS = {
    # shake-out size (BASELINE.json configs[0] shapes)
    "tiny": dict(lbs=2, vocab=1024, hidden=256, ffn=512, layers=2, heads=4, kv=2, seq=512, max_pos=512, theta=10000.0, lr="1.0e-3",
                 rope_scaling=""),
    # head_dim 128, GQA 4:1, llama3 RoPE scaling: the 8B layer structure at a size that runs in seconds
    "hd128": dict(lbs=1, vocab=4096, hidden=1024, ffn=3584, layers=4, heads=8, kv=2, seq=2048, max_pos=8192, theta=500000.0, lr="1.0e-4",
                  rope_scaling=LLAMA3_ROPE),
    # BASELINE.json configs[1]: the headline config
    "8b": dict(lbs=1, vocab=128256, hidden=4096, ffn=14336, layers=32, heads=32, kv=8, seq=4096, max_pos=8192, theta=500000.0, lr="1.0e-5",
               rope_scaling=LLAMA3_ROPE),
}


def det_init(name, shape, device, seed=1234, std=0.02):
    """fp32 initial value of one HF-named parameter: N(0, std) for matrices, ones for the norms (HF `_init_weights` semantics), from a
    Philox stream keyed by the parameter name."""
    import torch
    if name.endswith("norm.weight") or "layernorm" in name:
        return torch.ones(shape, dtype=torch.float32, device=device)
    g = torch.Generator(device=device).manual_seed((zlib.crc32(name.encode()) ^ seed) & 0x7FFFFFFF)
    return torch.randn(tuple(shape), generator=g, device=device, dtype=torch.float32) * std


def _sample(t, n=4096):
    flat = t.reshape(-1)
    stride = max(1, flat.numel() // n)
    return flat[::stride][:n].float().cpu().numpy()


def run(args):
    import ref_env_gpu  # noqa: F401  (sys.path for baseline/_ref + stubs)
    import numpy as np
    import torch
    from nemo_automodel.components.config._arg_parser import parse_args_and_load_config
    from nemo_automodel.recipes.llm.train_ft import TrainFinetuneRecipeForNextTokenPrediction

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if args.sdpa_backend:      # restrict torch SDPA to one backend (reference self-noise runs: same model, another attention kernel)
        torch.backends.cuda.enable_flash_sdp(args.sdpa_backend == "flash")
        torch.backends.cuda.enable_cudnn_sdp(args.sdpa_backend == "cudnn")
        torch.backends.cuda.enable_mem_efficient_sdp(args.sdpa_backend == "efficient")
        torch.backends.cuda.enable_math_sdp(args.sdpa_backend == "math")
    c = dict(CONFIGS[args.config])
    c.update(steps=args.steps, gbs=c["lbs"] * world * args.grad_accum, attn=args.attn, dtype=args.dtype)
    y = YAML.format(**c)
    if args.strategy == "b200_sharded":
        import automodel_b200.integration as b200
        b200.register()
        extra = ", max_tokens: %d, reference_rounding: true" % (c["lbs"] * c["seq"])
        if args.reduce_dtype:
            extra += ", reduce_dtype: " + args.reduce_dtype
        y = y.replace("strategy: fsdp2", "strategy: b200_sharded" + extra)
        y = y.replace("_target_: torch.optim.AdamW", "_target_: automodel_b200.recipe.B200FusedAdamW")
        if args.loss == "fused":
            y = y.replace("_target_: nemo_automodel.components.loss.masked_ce.MaskedCrossEntropy", "_target_: automodel_b200.recipe.B200MaskedCrossEntropy")
        assert "b200_sharded" in y and "B200FusedAdamW" in y
    with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as f:
        f.write(y)
        path = f.name
    cfg = parse_args_and_load_config(path, argv=[])
    t_setup = time.perf_counter()
    r = TrainFinetuneRecipeForNextTokenPrediction(cfg)
    r.setup()
    model = r.model_parts[0]
    dev = torch.device("cuda", torch.cuda.current_device())

    # ---- identical deterministic initial weights on both arms
    with torch.no_grad():
        if args.strategy == "b200_sharded":
            eng = model.engine
            eng.sync_params()
            for name, dst in eng.P.items():
                dst.copy_(det_init(name, dst.shape, dev).to(torch.bfloat16).to(dst.dtype))
            eng.refresh_master_()
            for t in eng.m + eng.v:
                t.zero_()
        else:
            from torch.distributed.tensor import DTensor, distribute_tensor
            # guard: the custom LlamaRotaryEmbedding keeps inv_freq in a non-persistent buffer; after a meta-device build nothing
            # recomputes it (HF's _init_weights only handles modules with `original_inv_freq`).  Record and repair if it is not the formula's.
            from nemo_automodel.components.models.llama import rope_utils as _ru
            rot = model.model.rotary_emb
            _, sc = _ru._get_rope_config(model.config)
            fn = _ru._compute_llama3_inv_freq if sc.get("rope_type", sc.get("type", "default")) == "llama3" else _ru._compute_default_inv_freq
            want = fn(model.config)[0].to(dev)
            inv_ok = bool(rot.inv_freq.device.type == "cuda" and torch.equal(rot.inv_freq.float(), want))
            have = rot.inv_freq.float().to(dev)
            globals()["_INV_DEV"] = float(((have - want).abs() / want.abs()).max()) if have.shape == want.shape else float("nan")
            globals()["_INV_HAVE"] = {"dtype": str(rot.inv_freq.dtype), "device": str(rot.inv_freq.device), "stock": [float(x) for x in have.cpu()],
                                      "formula": [float(x) for x in want.cpu()], "stock_equals_bf16_rounded_formula": bool(torch.equal(have, want.to(torch.bfloat16).float()))}
            if not inv_ok and not args.keep_stock_inv_freq:
                rot.inv_freq = want
                rot._cos_cache = rot._sin_cache = None
                rot.max_seq_len_cached = 0
            globals()["_INV_OK"] = inv_ok
            for name, p in model.named_parameters():
                full = det_init(name, p.shape, dev).to(torch.bfloat16).to(p.dtype)   # bf16-representable in every dtype
                if isinstance(p, DTensor):
                    p.copy_(distribute_tensor(full, p.device_mesh, p.placements))
                else:
                    p.copy_(full)
                del full
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup

    if args.fwd_only:
        return fwd_only(args, r, model, c, dev)

    rec = {"ref_inv_freq_was_initialised": globals().get("_INV_OK"), "ref_inv_freq_max_rel_dev_from_formula": globals().get("_INV_DEV"),
           "ref_inv_freq": globals().get("_INV_HAVE"), "kept_stock_inv_freq": bool(args.keep_stock_inv_freq), "loss": [], "grad_norm": [], "step_ms": [], "tps": [], "ids_crc": [], "num_label_tokens": [], "mem_gb": []}
    orig = r._run_train_optim_step

    def spy(batches, max_grad_norm=None):
        crc = 0
        for b in batches:
            crc = zlib.crc32(b["input_ids"].cpu().numpy().tobytes(), crc)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m = orig(batches, max_grad_norm)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        rec["loss"].append(float(m.metrics["loss"]))
        rec["grad_norm"].append(float(m.metrics["grad_norm"]))
        rec["step_ms"].append(dt * 1e3)
        rec["tps"].append(float(m.metrics["num_tokens_per_step"]) / dt)
        rec["ids_crc"].append(crc)
        rec["num_label_tokens"].append(int(m.metrics["num_label_tokens"]))
        rec["mem_gb"].append(float(m.metrics["mem"]))
        rec["max_grad_norm"] = max_grad_norm
        if rank == 0 and (len(rec["loss"]) <= 3 or len(rec["loss"]) % 20 == 0):
            print(f"[{args.strategy}] step {len(rec['loss']) - 1}: loss {rec['loss'][-1]:.5f} gnorm {rec['grad_norm'][-1]:.4f} {dt * 1e3:.1f} ms", file=sys.stderr, flush=True)
        return m

    r._run_train_optim_step = spy
    r.run_train_validation_loop()
    torch.cuda.synchronize()

    # ---- weights after the last step: strided samples + norms of a spread of parameters
    L = c["layers"]
    want_layers = sorted({0, L // 2, L - 1})
    samples = {}
    if args.strategy == "b200_sharded":
        named = model.engine.state_dict()
    else:
        named = dict(model.named_parameters())
    for name, p in named.items():
        keep = (".layers." not in name) or any(f".layers.{l}." in name for l in want_layers)
        if not keep:
            continue
        t = p.detach()
        if hasattr(t, "full_tensor"):
            t = t.full_tensor()
        if rank == 0:
            samples[name] = _sample(t)
            samples[name + "/l2"] = np.array([float(t.float().norm())])
        del t
    if rank == 0:
        skip = min(5, max(0, len(rec["step_ms"]) - 1))
        ms = rec["step_ms"][skip:]
        rec.update({
            "strategy": args.strategy, "config": args.config, "world": world, "grad_accum": args.grad_accum, "attn": args.attn, "sdpa_backend": args.sdpa_backend, "loss_kind": args.loss,
            "setup_s": setup_s, "mean_step_ms": sum(ms) / len(ms), "median_step_ms": sorted(ms)[len(ms) // 2], "skipped_steps": skip,
            "tokens_per_step": c["lbs"] * c["seq"] * world * args.grad_accum,
            "model_class": type(model).__name__, "optimizer_class": type(r.optimizer[0]).__name__, "loss_class": type(r.loss_fn).__name__,
            "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0),
            "sdpa_backends": {"flash": torch.backends.cuda.flash_sdp_enabled(), "cudnn": torch.backends.cuda.cudnn_sdp_enabled(),
                              "mem_efficient": torch.backends.cuda.mem_efficient_sdp_enabled()},
        })
        rec["tokens_per_s"] = rec["tokens_per_step"] / (rec["mean_step_ms"] / 1e3)
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(rec, open(args.out, "w"))
        np.savez_compressed(os.path.splitext(args.out)[0] + "_weights.npz", **samples)
        print(f"[{args.strategy}] {args.config} world {world}: {rec['mean_step_ms']:.1f} ms/step = {rec['tokens_per_s']:.0f} tok/s; "
              f"loss[0,-1] {rec['loss'][0]:.5f} {rec['loss'][-1]:.5f}; mem {max(rec['mem_gb']):.1f} GB", file=sys.stderr, flush=True)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


def fwd_only(args, r, model, c, dev):
    """Noise-floor experiment: the loss of the INITIAL weights on the first N batches of the recipe's data loader, forward only.  Run for the
    reference in bf16, the reference in fp32 (same bf16-representable weights: the exact-arithmetic answer) and ours; the per-token NLLs
    are saved so that |bf16 reference - fp32| and |ours - fp32| can be compared token by token."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    out = {"loss": [], "strategy": args.strategy, "dtype": args.dtype, "config": args.config}
    nll_all = []
    it = iter(r.dataloader)
    model.eval()
    with torch.no_grad():
        for i in range(args.fwd_only):
            batch = next(it)
            ids = batch["input_ids"].to(dev)
            labels = batch["labels"].to(dev)
            o = model(input_ids=ids, position_ids=batch["position_ids"].to(dev))
            logits = o.logits if hasattr(o, "logits") else o
            nll = F.cross_entropy(logits.float().view(-1, logits.shape[-1]), labels.view(-1), ignore_index=-100, reduction="none")
            n = int((labels != -100).sum())
            out["loss"].append(float(nll.sum() / n))
            nll_all.append(nll.cpu().numpy())
            out.setdefault("ids_crc", []).append(zlib.crc32(batch["input_ids"].numpy().tobytes()))
            del o, logits
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    json.dump(out, open(args.out, "w"))
    np.savez_compressed(os.path.splitext(args.out)[0] + "_nll.npz", nll=np.stack(nll_all))
    print(f"[{args.strategy} {args.dtype}] fwd-only losses: {[round(x, 5) for x in out['loss']]}", file=sys.stderr, flush=True)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


def compare(a_path, b_path, md=None):
    """a = reference (fsdp2), b = b200_sharded.  Prints a markdown summary; exit status 1 when north_star's tolerance is exceeded."""
    import numpy as np
    a, b = json.load(open(a_path)), json.load(open(b_path))
    n = min(len(a["loss"]), len(b["loss"]))
    la, lb = np.array(a["loss"][:n]), np.array(b["loss"][:n])
    ga, gb = np.array(a["grad_norm"][:n]), np.array(b["grad_norm"][:n])
    dl = np.abs(la - lb)
    dg = np.abs(ga - gb) / np.maximum(np.abs(ga), 1e-12)
    same_data = a["ids_crc"][:n] == b["ids_crc"][:n]
    lines = []
    P = lines.append
    P(f"### {a['config']} config, world {a['world']} (reference) vs world {b['world']} (b200_sharded), {n} steps, grad-accum {a['grad_accum']}")
    P("")
    P(f"* reference: `{a['model_class']}` + `{a['optimizer_class']}` + `{a['loss_class']}`, attn `{a['attn']}` (SDPA backends enabled: {a['sdpa_backends']}), torch {a['torch']}, {a['gpu']}")
    P(f"* ours: `{b['model_class']}` + `{b['optimizer_class']}` + `{b['loss_class']}` under the same unmodified recipe")
    P(f"* identical batches on both arms (crc32 of every step's input_ids): **{same_data}**")
    P(f"* max |Δloss| over {n} steps: **{dl.max():.2e}** (mean {dl.mean():.2e}; tolerance 1e-3); at step {int(dl.argmax())}")
    P(f"* max rel Δgrad_norm: **{dg.max():.2e}** (mean {dg.mean():.2e}; tolerance 1e-2 … 2e-2 over the curve)")
    P(f"* throughput: reference **{a['tokens_per_s']:.0f} tok/s** ({a['mean_step_ms']:.1f} ms/step, peak mem {max(a['mem_gb']):.1f} GB) vs ours "
      f"**{b['tokens_per_s']:.0f} tok/s** ({b['mean_step_ms']:.1f} ms/step, peak torch mem {max(b['mem_gb']):.1f} GB): **{b['tokens_per_s'] / a['tokens_per_s']:.2f}x**")
    P("")
    P("| step | ref loss | our loss | Δ | ref gnorm | our gnorm | rel Δ |")
    P("|---:|---:|---:|---:|---:|---:|---:|")
    for s in sorted(set(list(range(0, n, max(1, n // 10))) + [n - 1])):
        P(f"| {s} | {la[s]:.5f} | {lb[s]:.5f} | {dl[s]:.1e} | {ga[s]:.4f} | {gb[s]:.4f} | {dg[s]:.1e} |")
    wa_p, wb_p = os.path.splitext(a_path)[0] + "_weights.npz", os.path.splitext(b_path)[0] + "_weights.npz"
    if os.path.exists(wa_p) and os.path.exists(wb_p):
        wa, wb = np.load(wa_p), np.load(wb_p)
        worst, worst_name, rels = 0.0, "", []
        for k in wa.files:
            if k.endswith("/l2") or k not in wb.files:
                continue
            x, y_ = wa[k].astype(np.float64), wb[k].astype(np.float64)
            rel = np.linalg.norm(x - y_) / max(np.linalg.norm(x), 1e-12)
            rels.append(rel)
            if rel > worst:
                worst, worst_name = rel, k
        P("")
        P(f"* weights after step {n - 1} ({len(rels)} parameters sampled: embed, head, final norm, layers first/middle/last; 4096 strided values each): "
          f"max relative L2 difference **{worst:.2e}** ({worst_name}), median {float(np.median(rels)):.2e}")
    ok = bool(same_data and dl.max() <= 1e-3 and dg.max() <= 2e-2)
    P("")
    P(f"**{'PASS' if ok else 'FAIL'}** (north_star: step-loss within 1e-3 over the run)")
    text = "\n".join(lines)
    print(text)
    if md:
        with open(md, "a") as f:
            f.write(text + "\n\n")
    return 0 if ok else 1


def noise(truth_p, ref_p, ours_p, md=None):
    """Noise floor of the bf16 forward: per-token NLLs of the same weights on the same batches from (fp32 reference = exact arithmetic,
    bf16 reference, ours)."""
    import numpy as np
    t, a, b = (np.load(os.path.splitext(p)[0] + "_nll.npz")["nll"].astype(np.float64) for p in (truth_p, ref_p, ours_p))
    mask = t != 0
    nb = t.shape[0]
    lines = []
    P = lines.append
    P(f"### forward noise floor, initial weights, {nb} batches ({int(mask.sum() // nb)} label tokens each)")
    P("")
    P("| batch | fp32 reference loss | bf16 reference - fp32 | ours - fp32 | ours - bf16 reference |")
    P("|---:|---:|---:|---:|---:|")
    rows = []
    for i in range(nb):
        m = mask[i]
        lt, la, lb = t[i][m].mean(), a[i][m].mean(), b[i][m].mean()
        rows.append((la - lt, lb - lt, lb - la))
        P(f"| {i} | {lt:.5f} | {la - lt:+.2e} | {lb - lt:+.2e} | {lb - la:+.2e} |")
    r = np.array(rows)
    P(f"| rms | | {np.sqrt((r[:, 0] ** 2).mean()):.2e} | {np.sqrt((r[:, 1] ** 2).mean()):.2e} | {np.sqrt((r[:, 2] ** 2).mean()):.2e} |")
    P("")
    P(f"* per-token |NLL - fp32|: bf16 reference rms {np.sqrt(((a - t)[mask] ** 2).mean()):.3e}, ours rms {np.sqrt(((b - t)[mask] ** 2).mean()):.3e}; "
      f"ours vs bf16 reference rms {np.sqrt(((b - a)[mask] ** 2).mean()):.3e}")
    text = "\n".join(lines)
    print(text)
    if md:
        with open(md, "a") as f:
            f.write(text + "\n\n")
    return 0


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--strategy", default="fsdp2", choices=["fsdp2", "b200_sharded"])
    ap.add_argument("--config", default="tiny", choices=list(CONFIGS))
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--grad-accum", type=int, default=1)
    ap.add_argument("--attn", default="sdpa")
    ap.add_argument("--loss", default="reference", choices=["reference", "fused"])
    ap.add_argument("--reduce-dtype", default=None)
    ap.add_argument("--keep-stock-inv-freq", action="store_true", help="do not repair the reference's RoPE inv_freq buffer (see the guard in run())")
    ap.add_argument("--sdpa-backend", default=None, choices=["flash", "cudnn", "efficient", "math"])
    ap.add_argument("--dtype", default="bfloat16", help="torch_dtype of the reference model (float32: exact-arithmetic run of the same weights)")
    ap.add_argument("--fwd-only", type=int, default=0, help="noise-floor mode: loss of the initial weights on N batches, no training")
    ap.add_argument("--out", default="gpurun_out/ref_run.json")
    ap.add_argument("--compare", nargs=2, default=None)
    ap.add_argument("--noise", nargs=3, default=None, metavar=("FP32_REF", "BF16_REF", "OURS"))
    ap.add_argument("--md", default=None)
    a = ap.parse_args()
    if a.noise:
        sys.exit(noise(a.noise[0], a.noise[1], a.noise[2], a.md))
    if a.compare:
        sys.exit(compare(a.compare[0], a.compare[1], a.md))
    run(a)
