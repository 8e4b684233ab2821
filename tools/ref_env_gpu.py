"""Import environment for running the UNMODIFIED reference (installed under baseline/_ref, see DESIGN.md "Reference arm") on the GPU box.
Only the two stub modules of SURVEY.md appendix A (mlflow, torchao: imported unconditionally by recipes/llm/train_ft.py:36,42 but absent from
this image) - none of the CPU shims of tests/golden/_ref_env.py.  Measurement / test infrastructure; never imported by the product."""
import importlib.machinery
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("B200_REFERENCE_PATH", os.path.join(ROOT, "baseline", "_ref"))
if not os.path.isdir(os.path.join(REF, "nemo_automodel")):
    raise ImportError(f"reference not installed under {REF} (python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>)")
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import transformers.utils.import_utils as iu  # noqa: E402

iu.is_torchao_available()  # cache the real answer before a stub exists


def _mod(name):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    sys.modules[name] = m
    return m


if importlib.util.find_spec("mlflow") is None:
    ml = _mod("mlflow"); ml.active_run = lambda: None; ml.log_metrics = lambda *a, **k: None
if importlib.util.find_spec("torchao") is None:
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
