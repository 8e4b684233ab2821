"""Environment for importing and running the UNMODIFIED reference (NVIDIA-NeMo/Automodel, /root/reference) on a GPU-less host:
stubs for modules it imports unconditionally but that are absent here (mlflow, torchao) and CPU shims for CUDA-hard-coded call sites
(SURVEY.md appendix A).  Test infrastructure only: imported by tests/golden/gen_fixtures.py (fixture generation) and
tests/test_reference_dropin.py (the reference recipe driving this repository's strategy on CPU).  Never imported by the product or by
-m gpu tests (the GPU box has no /root/reference)."""
import os, sys, types, importlib.machinery

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
REF = os.environ.get("B200_REFERENCE_PATH", "/root/reference")
sys.path.insert(0, REF)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import transformers, transformers.utils.import_utils as iu  # noqa: F401

iu.is_torchao_available()  # cache False before the stub exists


def _mod(name):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    sys.modules[name] = m
    return m


# ---- stubs for modules the reference imports unconditionally but that are absent here
ml = _mod("mlflow"); ml.active_run = lambda: None; ml.log_metrics = lambda *a, **k: None
ta = _mod("torchao"); f8 = _mod("torchao.float8"); f8.precompute_float8_dynamic_scale_for_fsdp = lambda m: None
q = _mod("torchao.quantization"); qq = _mod("torchao.quantization.qat"); ql = _mod("torchao.quantization.qat.linear")


class _Dummy:
    def __init__(self, *a, **k):
        pass


for n in ["Int4WeightOnlyQATQuantizer", "Int8DynActInt4WeightQATQuantizer", "FakeQuantizeConfig", "IntxFakeQuantizeConfig",
          "QATConfig", "FromIntXQuantizationAwareTrainingConfig", "IntXQuantizationAwareTrainingConfig"]:
    setattr(qq, n, _Dummy); setattr(ql, n, _Dummy)
for n in ["disable_4w_fake_quant", "disable_8da4w_fake_quant", "enable_4w_fake_quant", "enable_8da4w_fake_quant"]:
    setattr(ql, n, lambda *a, **k: None)
ta.float8 = f8; ta.quantization = q; q.qat = qq; qq.linear = ql
fu = _mod("torchao.float8.fsdp_utils")


class WeightWithDynamicFloat8CastTensor:
    pass


fu.WeightWithDynamicFloat8CastTensor = WeightWithDynamicFloat8CastTensor; f8.fsdp_utils = fu

import torch

# ---- CPU shims for CUDA-hard-coded call sites (GPU-less host only)
torch.cuda.reset_peak_memory_stats = lambda *a, **k: None
torch.cuda.max_memory_allocated = lambda *a, **k: 0
torch.cuda.empty_cache = lambda *a, **k: None
torch.cuda.current_device = lambda: 0
torch.Tensor.cuda = lambda self, *a, **k: self


def _fix(d):
    if isinstance(d, int) and not isinstance(d, bool):
        return torch.device("cpu")
    if isinstance(d, torch.device) and d.type == "cuda":
        return torch.device("cpu")
    if isinstance(d, str) and d.startswith("cuda"):
        return torch.device("cpu")
    return d


_to = torch.nn.Module.to
torch.nn.Module.to = lambda self, *a, **k: _to(self, *tuple(_fix(x) for x in a), **{kk: _fix(v) for kk, v in k.items()})
_te = torch.nn.Module.to_empty
torch.nn.Module.to_empty = lambda self, *, device, recurse=True: _te(self, device=_fix(device), recurse=recurse)
_tt = torch.Tensor.to
torch.Tensor.to = lambda self, *a, **k: _tt(
    self, *tuple(x if isinstance(x, (torch.dtype, torch.Tensor)) else _fix(x) for x in a),
    **{kk: (_fix(v) if kk == "device" else v) for kk, v in k.items()})
_el = torch.empty_like


def _empty_like(t, *a, **k):
    if "device" in k:
        k["device"] = _fix(k["device"])
    return _el(t, *a, **k)


torch.empty_like = _empty_like

# the benchmark recipe's timers (components/training/timers.py:197, 361) synchronise and allocate on "the current CUDA device"
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.nvtx.range_push = lambda *a, **k: None
torch.cuda.nvtx.range_pop = lambda *a, **k: None
_zr = torch.zeros


def _zeros(*a, **k):
    if "device" in k:
        k["device"] = _fix(k["device"])
    return _zr(*a, **k)


torch.zeros = _zeros

