"""Reference-facing surface of the B200 sharded step: what the `automodel` recipe (unchanged) binds to.

Mirrors the reference's extension points for this path:

  * `B200ShardedConfig` / `B200ShardedManager.parallelize(model)`  <->  FSDP2Config / FSDP2Manager.parallelize
        (components/distributed/config.py:49-136, fsdp2.py:105-142; registered under distributed.strategy: b200_sharded)
  * `B200CausalLM` (nn.Module)  <->  the FSDP-wrapped LlamaForCausalLM the recipe calls as `model(**batch).logits`
        (recipes/llm/train_ft.py:1443-1457), with HF-named parameters / .grad views, `set_requires_gradient_sync`
        (components/distributed/utils.py:222-247) and the grad-norm hook the clip utility needs for sharded flat grads
        (components/training/utils.py:290-359)
  * `B200MaskedCrossEntropy`  <->  loss_fn `_target_` called as `loss_fn(logits=, labels=, num_label_tokens=)`
        (components/loss/utils.py:84-104, masked_ce.py:41-90) - fused with the engine's CE kernel
  * `B200FusedAdamW`  <->  optimizer `_target_` called as `target(params=trainable_params, **yaml)` (train_ft.py:368); `.step()`,
        `.zero_grad()`, mutable `.param_groups[0]["lr" | "weight_decay"]` for OptimizerParamScheduler (optim/scheduler.py:257-261)

Host glue only: every device operation goes through ShardedLlamaEngine -> the C ABI.
"""
import weakref
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn as nn

from .engine import ShardedLlamaEngine, IGNORE_INDEX


@dataclass
class B200ShardedConfig:
    """YAML `distributed:` section for `strategy: b200_sharded` (same role as FSDP2Config)."""
    max_tokens: int = 4096
    adam_mode: int = 1               # 1: torch.optim.AdamW bf16 op sequence (reference default optimizer); 0: fp32 math
    master_weights: bool = False
    reference_rounding: bool = True  # bf16(bf16(acc) + residual), as the reference's two eager ops
    max_positions: Optional[int] = None
    activation_checkpointing: bool = False   # same key as FSDP2Config.activation_checkpointing (distributed/config.py:49-136)
    # Packed rows (several documents per row, position_ids restarting at each): True = always derive the documents from position_ids (one
    # device->host read per micro-batch when the batch already lives on the GPU), False = never, None = auto (CPU position_ids are
    # inspected for free; CUDA position_ids are inspected for the first micro-batches and afterwards only checked asynchronously, so a
    # plain SFT run keeps the host running ahead of the GPU and a packed batch that appears later still fails loudly one step later).
    packed_sequences: Optional[bool] = None
    backend: str = "nccl"                    # accepted for YAML compatibility (`distributed.backend`); torch.distributed is set up by the recipe
    # --- FSDP2Config's remaining keys (components/distributed/config.py:106-120), so that an FSDP2 YAML only needs `strategy:` changed.
    # Scheduling hints are accepted and ignored (this engine has its own overlap schedule); options that would change what is computed
    # are refused in __post_init__.
    defer_fsdp_grad_sync: bool = True        # gradients always accumulate unsharded and are reduce-scattered once per optimizer step
    enable_fsdp2_prefetch: bool = False
    fsdp2_backward_prefetch_depth: int = 2
    fsdp2_forward_prefetch_depth: int = 1
    enable_compile: bool = False
    patch_is_packed_sequence: bool = False
    mp_policy: Optional[object] = None       # MixedPrecisionPolicy: param_dtype must be bf16; reduce_dtype (float32 | bfloat16) is honoured
    reduce_dtype: Optional[str] = None       # gradient reduction precision; None = mp_policy.reduce_dtype, else float32 (FSDP2Config's default policy)
    reshard_after_forward: bool = False      # the reference's per-layer gather/free schedule: ENGINE-LEVEL ONLY for now (see __post_init__)
    comm: Optional[str] = None               # per-unit collectives: nvls (own kernels on symmetric memory) | p2p | nccl; None = B200_COMM / default
    sequence_parallel: bool = False
    tp_plan: Optional[dict] = None
    offload_policy: Optional[object] = None
    autocast_dtype: Optional[object] = None
    enable_async_tensor_parallel: bool = False

    def __post_init__(self):
        bad = [k for k in ("sequence_parallel", "enable_async_tensor_parallel", "enable_compile") if getattr(self, k)]
        bad += [k for k in ("tp_plan", "offload_policy", "autocast_dtype") if getattr(self, k) is not None]
        if bad:
            raise ValueError(f"strategy b200_sharded does not support {bad} (data-parallel bf16 training without offload / compile)")
        if self.reshard_after_forward:
            # ShardedLlamaEngine(reshard_after_forward=True) exists (layers live as shards + a 2-slot pool, verified over gloo), but the
            # facade exposes HF-named nn.Parameters as views of resident unsharded buffers, which that mode does not have
            raise ValueError("strategy b200_sharded: reshard_after_forward is not available through the recipe facade yet (engine-level only)")
        pd = getattr(self.mp_policy, "param_dtype", None)
        if pd is not None and pd != torch.bfloat16:
            raise ValueError(f"strategy b200_sharded computes in bf16; mp_policy.param_dtype={pd} is not supported")
        # components/distributed/config.py:121-132: the default policy reduces gradients in fp32
        names = {torch.float32: "float32", torch.bfloat16: "bfloat16", "float32": "float32", "bfloat16": "bfloat16", "fp32": "float32", "bf16": "bfloat16"}
        rd = self.reduce_dtype if self.reduce_dtype is not None else getattr(self.mp_policy, "reduce_dtype", None)
        if rd is None:
            rd = "float32"
        if rd not in names:
            raise ValueError(f"strategy b200_sharded reduces gradients in float32 or bfloat16; reduce_dtype={rd} is not supported")
        self.reduce_dtype = names[rd]
        od = getattr(self.mp_policy, "output_dtype", None)
        if od is not None and od != torch.bfloat16:
            raise ValueError(f"strategy b200_sharded produces bf16 activations; mp_policy.output_dtype={od} is not supported")


class _Fwd(torch.autograd.Function):
    """The whole decoder stack as ONE autograd node: forward runs the engine up to the logits, backward consumes dlogits and runs the
    engine backward (wgrads straight into the flat gradient buffers, reduce-scatter per unit).  Parameter gradients therefore do
    not flow through autograd: they appear on the nn.Parameters as views of the flat buffers."""

    @staticmethod
    def forward(ctx, anchor, model, handle, b, S):
        eng = model.engine
        logits = eng.forward_logits(handle)
        ctx.model, ctx.handle = model, handle
        return logits.view(b, S, -1)

    @staticmethod
    def backward(ctx, dlogits):
        model, eng = ctx.model, ctx.model.engine
        T = ctx.handle[1]
        buf = eng.logits[:T]
        d2 = dlogits.reshape(T, -1)
        # the recipe back-propagates local_loss * dp_group_size (train_ft.py:1473) because FSDP2 averages; our reduce-scatter sums
        scale = 1.0 / (eng.world * eng.replicas)
        if d2.data_ptr() != buf.data_ptr():
            buf.copy_(d2 if scale == 1.0 else d2 * scale)
        else:
            # fused loss: the buffer already holds dlogits for a unit upstream gradient; fold the upstream scalar (a DEVICE tensor: reading
            # it on the host would stall the launch queue once per micro-batch) and 1/dp in with one in-place multiply
            up = model._upstream_grad
            model._upstream_grad = None
            t = scale if up is None else up.to(torch.float32) * scale
            if up is not None or scale != 1.0:
                buf.mul_(t)
        eng.backward_from_dlogits(ctx.handle, first_micro=model._first_micro, last_micro=bool(model._sync_grads))
        model._first_micro = False
        return None, None, None, None, None


class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, model, num_label_tokens):
        eng = model.engine
        eng.loss_dev.zero_()      # one value per call (the recipe sums the micro-batch losses itself); never difference a growing accumulator
        eng.fused_loss(model._last_handle, num_label_tokens)   # logits buffer now holds dlogits for a unit upstream gradient
        ctx.model = model
        return eng.loss_dev[0].clone()

    @staticmethod
    def backward(ctx, g):
        # dlogits are already in the engine buffer (scaled by 1/num_label_tokens); the upstream scalar is applied by _Fwd.backward
        model = ctx.model
        model._upstream_grad = g.reshape(()) if g.numel() == 1 else None
        T = model._last_handle[1]
        return model.engine.logits[:T].view(model._last_shape[0], model._last_shape[1], -1), None, None


class _Node(nn.Module):
    """Empty container: one node of the HF module tree the facade mirrors (no forward of its own)."""


class B200CausalLM(nn.Module):
    """nn.Module facade over ShardedLlamaEngine with the surface the recipe uses."""

    def __init__(self, config, engine: ShardedLlamaEngine):
        super().__init__()
        self.config = config
        self.engine = engine
        self._sync_grads = None    # None: nobody told us which micro-batch is the last (no get_sync_ctx hook) -> reduce-scatter lazily
        self._first_micro = True
        self._last_handle = None
        self._last_shape = None
        self._upstream_grad = None
        self.packed_sequences = None          # see B200ShardedConfig.packed_sequences (set by B200ShardedManager.parallelize)
        self._pack_probes_left = 3
        self._pack_flag = None                 # device bool: "a CUDA batch skipped by the auto mode had restarting position_ids"
        self._pack_check = None                # (pinned host copy, event) of the flag taken at the previous optimizer step
        engine._facade = weakref.ref(self)   # lets the optimizer built from `model.parameters()` find the module that owns the step state
        self._anchor = nn.Parameter(torch.zeros((), device=engine.device), requires_grad=True)  # keeps the autograd node alive
        # HF-named parameters as views of the flat buffers; .grad = views of the flat gradient buffers.  They hang on a skeleton of empty
        # container modules that mirrors the HF module tree (model.layers.<i>.self_attn.q_proj.weight ...), so named_parameters(),
        # state_dict() and FQN walkers such as torch.distributed.checkpoint.state_dict.get_model_state_dict (what the reference's
        # Checkpointer uses, components/checkpoint/stateful_wrappers.py:278) see the names and shapes of the model they replace.
        self._hf = {}
        grads = engine.named_grads()
        for name, p in engine.state_dict().items():
            param = nn.Parameter(p, requires_grad=True)
            param.grad = grads[name]
            param._b200_engine = engine
            self._hf[name] = param
            node = self
            *path, leaf = name.split(".")
            for part in path:
                if part not in node._modules:
                    node.add_module(part, _Node())
                node = node._modules[part]
            node.register_parameter(leaf, param)

    # ---- reference surface
    def named_parameters(self, prefix="", recurse=True, remove_duplicate=True):
        for name, p in super().named_parameters(prefix=prefix, recurse=recurse, remove_duplicate=remove_duplicate):
            if not name.endswith("_anchor"):
                yield name, p

    def state_dict(self, *a, **k):
        self.engine.sync_params()
        sd = super().state_dict(*a, **k)
        sd.pop(k.get("prefix", "") + "_anchor", None)
        return sd

    def load_state_dict(self, state_dict, strict=True, assign=False):
        """In-place copy into the flat buffers (from_pretrained / checkpoint resume); optimizer state is untouched."""
        missing = [n for n in self._hf if n not in state_dict]
        unexpected = [n for n in state_dict if n not in self._hf and n != "_anchor"]
        if strict and (missing or unexpected):
            raise RuntimeError(f"B200CausalLM.load_state_dict: missing {missing[:3]} unexpected {unexpected[:3]}")
        self.engine.sync_params()
        with torch.no_grad():
            for n, p in self._hf.items():
                if n in state_dict:
                    p.copy_(state_dict[n])
        self.engine.refresh_master_()
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def initialize_weights(self):
        """Random initialisation AFTER sharding: what `Checkpointer.initialize_model_weights` calls on the parallelized model when the
        recipe built it on the meta device (from_config, and from_pretrained before the checkpoint overwrites it;
        components/checkpoint/checkpointing.py:574-676, _transformers/infrastructure.py:528-552).  HF `_init_weights` semantics: N(0,
        config.initializer_range) for projections and embeddings, ones for the norms.  Every data-parallel rank must end with the same
        parameters, so the seed is the global rank 0's torch seed."""
        import torch.distributed as dist
        seed = torch.tensor([torch.initial_seed() % (2 ** 31)], dtype=torch.int64)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            seed = seed.to(self.engine.device if dist.get_backend() == "nccl" else "cpu")
            dist.broadcast(seed, src=0)
        cfg = self.config
        std = (cfg.get("initializer_range", 0.02) if isinstance(cfg, dict) else getattr(cfg, "initializer_range", 0.02)) or 0.02
        self.engine.init_random_(seed=int(seed.item()), std=float(std))

    def set_requires_gradient_sync(self, flag: bool, recurse: bool = True):
        """FSDPModule API used by get_sync_ctx: False on all but the last micro-batch (defer_fsdp_grad_sync)."""
        self._sync_grads = bool(flag)

    def forward(self, input_ids, position_ids=None, labels=None, attention_mask=None, logits_to_keep=None, **_ignored):
        # `logits_to_keep` is part of the HF causal-LM signature; the recipe keeps a non-default loss_fn only for models that accept
        # it (train_ft.py:1056-1058, _supports_logits_to_keep).  Training always needs every position.
        if logits_to_keep not in (None, 0):
            raise NotImplementedError("B200CausalLM computes logits for all positions (logits_to_keep must be None or 0)")
        eng = self.engine
        if getattr(self, "_norm_ready", False) and self.training:
            # the clip utility ran but B200FusedAdamW.step() did not: another optimizer is stepping the parameter views.  That would skip
            # the clip (it is applied inside the fused AdamW) and, with more than one rank, update from reduce-scattered gradient buffers
            raise RuntimeError("B200CausalLM must be optimised by automodel_b200.recipe.B200FusedAdamW "
                               "(optimizer: {_target_: automodel_b200.recipe.B200FusedAdamW, ...}); another optimizer stepped it")
        handle = eng.stage(input_ids, labels, self._document_position_ids(position_ids))   # labels=None: the loss function gets them (set_labels)
        self._last_handle, self._last_shape = handle, tuple(input_ids.shape)
        logits = _Fwd.apply(self._anchor, self, handle, input_ids.shape[0], input_ids.shape[1])
        logits._b200_model = self
        return SimpleNamespace(logits=logits)

    def _document_position_ids(self, position_ids):
        """position_ids only matter as document delimiters of packed rows (RoPE rotates by the index inside the row, like the reference's
        rotary module).  Returns them when they must be inspected, None when the batch is to be treated as one document per row."""
        if position_ids is None or self.packed_sequences is False:
            return None
        if self.packed_sequences is True or position_ids.device.type == "cpu":
            return position_ids
        if self._pack_probes_left > 0:        # auto mode, CUDA tensors: look (with a sync) at the first few micro-batches
            self._pack_probes_left -= 1
            if bool((position_ids[:, 1:] <= position_ids[:, :-1]).any()):
                self.packed_sequences = True
            return position_ids
        flag = (position_ids[:, 1:] <= position_ids[:, :-1]).any().reshape(1)
        self._pack_flag = flag if self._pack_flag is None else (self._pack_flag | flag)
        return None

    def _poll_pack_check(self):
        """Auto mode: the asynchronous verdict on the batches that were NOT inspected (no sync: the copy was issued one step ago)."""
        if self._pack_check is not None:
            host, ev = self._pack_check
            if ev.query():
                self._pack_check = None
                if bool(host[0]):
                    raise RuntimeError("a packed batch (position_ids restarting inside a row) arrived after the first micro-batches; set "
                                       "distributed.packed_sequences: true so documents are always derived from position_ids")
        if self._pack_flag is not None and self._pack_check is None:
            host = torch.empty(1, dtype=torch.bool, pin_memory=True)
            host.copy_(self._pack_flag, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._pack_check, self._pack_flag = (host, ev), None

    # ---- hooks for the clip utility / optimizer
    def b200_clip_grad_norm(self, max_norm: Optional[float]):
        """components/training/utils.py:290-359 equivalent for the flat sharded grads: returns the global grad norm (device scalar);
        the clip itself is applied inside the fused AdamW with this max_norm."""
        self._max_norm = max_norm
        self._norm_ready = True
        return self.engine.compute_grad_norm_sq().sqrt()[0]

    def b200_optimizer_step(self, lr=None):
        if not getattr(self, "_norm_ready", False):
            self.engine.compute_grad_norm_sq()
        self._poll_pack_check()
        self.engine.apply_adamw(getattr(self, "_max_norm", None), lr=lr)
        self._max_norm = None          # a clip threshold applies to the step it was requested for (the benchmark recipe never clips)
        self._norm_ready = False
        self._first_micro = True


class B200MaskedCrossEntropy(nn.Module):
    """Drop-in for components/loss/masked_ce.py:MaskedCrossEntropy when the model is a B200CausalLM: same call signature and
    value (sum of token NLL / num_label_tokens, ignore_index -100, exactly 0 when there are no label tokens), computed by the
    fused CE kernel on the engine's logits buffer."""

    def __init__(self, fp32_upcast: bool = True, ignore_index: int = IGNORE_INDEX, reduction: str = "sum"):
        super().__init__()
        assert ignore_index == IGNORE_INDEX and reduction == "sum"

    def forward(self, logits, labels, mask=None, num_label_tokens: Optional[int] = None):
        model = getattr(logits, "_b200_model", None)
        if model is None:
            # logits of a model that runs on the reference's own path (B200ShardedManager.fallback): the reference's MaskedCrossEntropy
            # computation itself (components/loss/masked_ce.py:73-89) - fp32 upcast, sum over label tokens / num_label_tokens
            import torch.nn.functional as F
            if mask is not None:
                labels = labels.masked_fill(mask == 0, IGNORE_INDEX)
            loss = F.cross_entropy(logits.float().view(-1, logits.shape[-1]), labels.view(-1), ignore_index=IGNORE_INDEX, reduction="sum")
            return loss / num_label_tokens if num_label_tokens is not None else loss
        if mask is not None:
            raise NotImplementedError("mask= is not supported; pre-mask the labels with -100 (what the reference does internally)")
        if num_label_tokens is None:
            num_label_tokens = 1      # reduction="sum" without normalisation, as the reference returns it (masked_ce.py:84-89; validation loop)
        model.engine.set_labels(model._last_handle, labels)   # the recipe passes labels to the loss, not to the model
        return _FusedLoss.apply(logits, model, int(num_label_tokens))


class B200FusedAdamW(torch.optim.Optimizer):
    """optimizer `_target_`: fused AdamW on the flat shards of the engine that owns `params`.  Parameters that belong to NO engine (the
    strategy routed the model to the reference's FSDP2 path, see B200ShardedManager.fallback) get a plain torch.optim.AdamW with the same
    hyper-parameters - the reference's default optimizer - so the YAML keeps working unchanged."""

    def __new__(cls, params=None, *args, **kwargs):
        if params is not None:
            params = list(params)
            flat = [p for g in params for p in g["params"]] if params and isinstance(params[0], dict) else params
            if flat and all(getattr(p, "_b200_engine", None) is None for p in flat):
                kw = {k: v for k, v in kwargs.items() if k in ("lr", "betas", "eps", "weight_decay", "amsgrad", "maximize", "foreach", "fused", "capturable", "differentiable")}
                return torch.optim.AdamW(params, *args, **kw)        # not an instance of cls: __init__ below is skipped
        obj = super().__new__(cls)
        obj._materialised_params = params      # `params` may be a generator: __init__ receives the same (now exhausted) object
        return obj

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, maximize=False, foreach=None,
                 fused=None, capturable=False, differentiable=False):
        # torch.optim.AdamW's keyword set AND defaults, so a YAML written for it only needs the `_target_` changed: implementation hints (foreach,
        # fused, capturable) mean nothing here, options that change the mathematics are refused
        if amsgrad or maximize or differentiable:
            raise NotImplementedError("B200FusedAdamW implements plain AdamW (amsgrad / maximize / differentiable are not supported)")
        params = self.__dict__.pop("_materialised_params", None) or list(params)
        engines = {id(getattr(p, "_b200_engine", None)): getattr(p, "_b200_engine", None) for p in params}
        if len(engines) != 1 or None in engines.values():
            raise TypeError("B200FusedAdamW needs the parameters of ONE B200CausalLM")
        self.engine = next(iter(engines.values()))
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._model = None
        self._refresh_state()

    def attach(self, model: B200CausalLM):
        """Optional: the owning module is found through the engine (the recipe builds the optimizer from parameters only)."""
        self._model = model
        return self

    def _owner(self):
        if self._model is not None:
            return self._model
        ref = getattr(self.engine, "_facade", None)
        return ref() if ref is not None else None

    # ---- optimizer state in torch.optim form (what torch.distributed.checkpoint's get/set_optimizer_state_dict and therefore the
    # reference's Checkpointer read and write, components/checkpoint/stateful_wrappers.py): per-parameter "step" / "exp_avg" /
    # "exp_avg_sq".  `self.state` must never be empty, or DCP "initialises" it by running a throw-away optimizer.step().
    def _refresh_state(self, full=False):
        """World 1: live views of the flat moment shards (zero copy).  World N: `self.state` holds only per-parameter placeholders (the
        shared step counter + zero-size moments) so that nothing replicated stays resident - the full moments (2 x model bytes) are
        all-gathered only for the duration of a `state_dict()` call (checkpoint time) and live exactly as long as the caller keeps the
        returned dict."""
        self._step_t = torch.tensor(float(self.engine.step_count))   # ONE tensor shared by every parameter's state entry
        owner = self._owner()
        by_param = {id(p): n for n, p in owner._hf.items()} if owner is not None else {}
        sharded = self.engine.world > 1
        named = self.engine.gather_optimizer_state() if (full or not sharded) else None
        for group in self.param_groups:
            for p in group["params"]:
                name = by_param.get(id(p))
                if name is None:
                    continue
                if named is not None:
                    m, v = named[name]
                else:
                    m = v = p.new_empty(0)
                self.state[p] = {"step": self._step_t, "exp_avg": m, "exp_avg_sq": v}

    def state_dict(self):
        self._refresh_state(full=True)
        sd = super().state_dict()
        if self.engine.world > 1:
            self._refresh_state(full=False)      # drop this object's references to the gathered copies
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        owner = self._owner()
        named, step = {}, self.engine.step_count
        for n, p in (owner._hf.items() if owner is not None else []):
            st = self.state.get(p)
            if st and "exp_avg" in st and st["exp_avg"].numel() == p.numel():
                named[n] = (st["exp_avg"], st["exp_avg_sq"])
                step = int(float(st["step"]))
        self.engine.load_optimizer_state(named, step)
        self._refresh_state()     # world 1: back to live views of the flat shards; world N: placeholders

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        e = self.engine
        e.betas, e.eps, e.wd = tuple(g["betas"]), g["eps"], g["weight_decay"]
        owner = self._owner()
        if owner is not None:
            owner.b200_optimizer_step(lr=g["lr"])     # uses the max_norm of the preceding clip call and re-arms the accumulation window
        else:
            e.optimizer_step(None, lr=g["lr"])
        self._step_t += 1

    def zero_grad(self, set_to_none: bool = True):
        return None  # wgrad epilogues overwrite the flat buffers on the first micro-batch of the next step


class B200ShardedManager:
    """`.parallelize(model)` of the new distributed strategy."""

    def __init__(self, config: B200ShardedConfig, process_group=None, device=None, ops=None, replica_group=None, fallback=None):
        self.config, self.pg, self.device, self.ops, self.rpg = config, process_group, device, ops, replica_group
        # fallback: callable returning the reference's own manager (FSDP2Manager over the same mesh).  north_star: "any HF config the
        # reference accepts runs unchanged" - a model this engine does not implement (another architecture, PEFT adapters, frozen
        # parameters, an unsupported RoPE variant ...) is handed to the unmodified reference path instead of failing the job.
        self.fallback = fallback
        self.used_fallback = None     # reason string once parallelize() has routed a model to the reference path

    def parallelize(self, model, optimizer_defaults=None):
        if self.fallback is None:
            return self._parallelize(model, optimizer_defaults)
        try:
            return self._parallelize(model, optimizer_defaults)
        except (ValueError, NotImplementedError) as e:
            import logging
            self.used_fallback = f"{type(e).__name__}: {e}"
            logging.getLogger(__name__).warning("strategy b200_sharded: %s -> training this model on the reference's FSDP2 path", self.used_fallback)
            return self.fallback().parallelize(model)

    def _parallelize(self, model, optimizer_defaults=None):
        cfg = model.config if hasattr(model, "config") else model
        dev = self.device if self.device is not None else torch.device("cuda", torch.cuda.current_device())
        od = optimizer_defaults or {}
        cfg_d = cfg.to_dict() if hasattr(cfg, "to_dict") else cfg
        if hasattr(model, "named_parameters"):
            # the engine trains exactly the HF Llama parameter set, all of it: anything else must fail here (before any device memory is
            # taken), not train something different
            from .layout import LlamaDims, build_layout
            expected = {sl.name for u in build_layout(LlamaDims.from_hf(cfg_d), 1) for sl in u.slots}
            names = {n for n, _ in model.named_parameters()}
            extra, missing = sorted(names - expected), sorted(expected - names)
            if cfg_d.get("tie_word_embeddings") and extra == ["lm_head.weight"]:
                # the config ties lm_head to the embedding but the module handed over still lists a separate lm_head parameter (meta-device
                # builds before tie_weights(), or a transformers version whose tie_weights() does not act on the reference's custom class):
                # the engine follows the CONFIG - one shared matrix, as the reference does after loading (checkpointing.py:720-722)
                extra = []
            if extra or missing:
                raise NotImplementedError(f"strategy b200_sharded: the model's parameters are not the Llama set the engine implements "
                                          f"(unexpected {extra[:3]}, missing {missing[:3]}); PEFT adapters / other architectures are not supported")
            frozen = [n for n, p in model.named_parameters() if not p.requires_grad]
            if frozen:
                raise NotImplementedError(f"strategy b200_sharded trains every parameter; frozen parameters are not supported ({frozen[:3]})")
        eng = ShardedLlamaEngine(cfg_d, dev, process_group=self.pg, max_tokens=self.config.max_tokens,
                                 adam_mode=self.config.adam_mode, master_weights=self.config.master_weights,
                                 reference_rounding=self.config.reference_rounding, max_positions=self.config.max_positions,
                                 activation_checkpointing=self.config.activation_checkpointing, replica_group=self.rpg, ops=self.ops,
                                 reduce_dtype=self.config.reduce_dtype, comm=self.config.comm, **od)
        if hasattr(model, "named_parameters"):
            sd = model.state_dict()
            if sd and all(getattr(v, "device", torch.device("cpu")).type != "meta" for v in sd.values()):
                eng.load_state_dict(sd)      # materialised model (single-GPU load-before-shard); a meta model is initialised / loaded afterwards
        model = B200CausalLM(cfg, eng)
        model.packed_sequences = self.config.packed_sequences
        return model
