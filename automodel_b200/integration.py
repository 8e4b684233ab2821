"""Run-time application of INTEGRATION.md to an importable `nemo_automodel` (the reference): registers `distributed.strategy:
b200_sharded` and routes the grad-clip utility to the sharded flat gradients.  This is the patch a maintainer would commit upstream,
expressed as monkey-patches so the UNMODIFIED reference recipe can drive this repository's engine:

    import automodel_b200.integration as b200; b200.register()
    # YAML: distributed: {strategy: b200_sharded, ...}; optimizer: {_target_: automodel_b200.recipe.B200FusedAdamW, ...}

Touch points (reference file:line):
  * components/distributed/mesh.py:48-52        STRATEGY_MAP gets "b200_sharded" -> B200ShardedConfig
  * components/distributed/mesh_utils.py:46-113 create_device_mesh builds the FSDP2 mesh (pp, dp_replicate, dp_shard, cp, tp) for it
  * _transformers/infrastructure.py:152-182     _instantiate_distributed returns a B200ShardedManager for that config
  * components/training/utils.py:290-359        scale_grads_and_clip_grad_norm dispatches to model.b200_clip_grad_norm
  * components/distributed/utils.py:222-247     get_sync_ctx also tells a B200CausalLM which micro-batch is the last one (it only does so
                                                for FSDPModule instances); without it the facade still trains correctly, it just
                                                reduce-scatters at clip time instead of under the last backward
Host glue only; imports the reference lazily (it is not a dependency of this package).
"""
from .recipe import B200ShardedConfig, B200ShardedManager

_registered = False


def register(ops=None, device=None, patch_sync_ctx=True):
    """Idempotent.  `ops` / `device` are test hooks (CPU stand-in kernels); production leaves them None."""
    global _registered
    import nemo_automodel.components.distributed.mesh as _mesh
    import nemo_automodel._transformers.infrastructure as _infra
    import nemo_automodel.components.training.utils as _tu

    _mesh.STRATEGY_MAP["b200_sharded"] = B200ShardedConfig
    if _registered:
        return
    _registered = True

    import nemo_automodel.components.distributed.mesh_utils as _mu
    orig_mesh = _mu.create_device_mesh

    def create_device_mesh(distributed_config, **kw):
        if isinstance(distributed_config, B200ShardedConfig):
            if (kw.get("tp_size") or 1) > 1 or (kw.get("pp_size") or 1) > 1 or (kw.get("cp_size") or 1) > 1 or (kw.get("ep_size") or 1) > 1:
                raise ValueError("strategy b200_sharded is data parallel only (tp/pp/cp/ep sizes must be 1)")
            return _mu._create_fsdp2_device_mesh(dp_size=kw.get("dp_size"), dp_replicate_size=kw.get("dp_replicate_size"), tp_size=1, pp_size=1,
                                                 cp_size=1, ep_size=1, world_size=kw["world_size"], backend=distributed_config.backend)
        return orig_mesh(distributed_config, **kw)

    _mu.create_device_mesh = create_device_mesh

    orig_inst = _infra._instantiate_distributed

    def _instantiate_distributed(config, mesh):
        if isinstance(config, B200ShardedConfig):
            pg = rpg = None
            dm = getattr(mesh, "device_mesh", None)
            if dm is not None and dm.size() > 1:
                names = dm.mesh_dim_names or ()
                if "dp_shard" in names:
                    pg = dm["dp_shard"].get_group() if dm["dp_shard"].size() > 1 else None
                    if "dp_replicate" in names and dm["dp_replicate"].size() > 1:     # HSDP: distributed.dp_replicate_size
                        rpg = dm["dp_replicate"].get_group()
                else:
                    pg = dm.get_group()
            def reference_manager():
                """The reference's own FSDP2 manager over the same mesh, configured from the keys the two configs share."""
                import dataclasses
                from nemo_automodel.components.distributed.config import FSDP2Config
                shared = {f.name: getattr(config, f.name) for f in dataclasses.fields(FSDP2Config) if hasattr(config, f.name)}
                return orig_inst(FSDP2Config(**shared), mesh)

            return B200ShardedManager(config, process_group=pg, replica_group=rpg, device=device, ops=ops, fallback=reference_manager)
        return orig_inst(config, mesh)

    _infra._instantiate_distributed = _instantiate_distributed

    orig_clip = _tu.scale_grads_and_clip_grad_norm

    def scale_grads_and_clip_grad_norm(max_grad_norm, model_parts, *args, **kwargs):
        m = model_parts[0] if isinstance(model_parts, (list, tuple)) else model_parts
        if hasattr(m, "b200_clip_grad_norm"):
            return m.b200_clip_grad_norm(max_grad_norm)
        return orig_clip(max_grad_norm, model_parts, *args, **kwargs)

    _tu.scale_grads_and_clip_grad_norm = scale_grads_and_clip_grad_norm

    import nemo_automodel.components.distributed.utils as _du
    orig_sync = _du.get_sync_ctx

    def get_sync_ctx(model, is_optim_step, defer_fsdp_grad_sync=True):
        if hasattr(model, "b200_clip_grad_norm"):
            model.set_requires_gradient_sync(is_optim_step)   # gradients accumulate unsharded: sync exactly once, on the last micro-batch
            from contextlib import nullcontext
            return nullcontext()
        return orig_sync(model, is_optim_step, defer_fsdp_grad_sync)

    if patch_sync_ctx:
        _du.get_sync_ctx = get_sync_ctx
    # modules that imported the symbols by name keep their own reference: rebind the recipe's
    try:
        import nemo_automodel.recipes.llm.train_ft as _ft
        if getattr(_ft, "scale_grads_and_clip_grad_norm", None) is orig_clip:
            _ft.scale_grads_and_clip_grad_norm = scale_grads_and_clip_grad_norm
        if patch_sync_ctx and getattr(_ft, "get_sync_ctx", None) is orig_sync:
            _ft.get_sync_ctx = get_sync_ctx
    except Exception:  # the recipe module is optional at registration time
        pass
