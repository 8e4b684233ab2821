"""The drop-in claim, end to end: the reference's own recipe (unmodified, imported from /root/reference) trains through this
repository's `distributed.strategy: b200_sharded` + optimizer `_target_` after `automodel_b200.integration.register()`, and
reproduces the loss / grad-norm curve the same recipe produced with FSDP2 (tests/golden/*.npz).  CPU only (stand-in kernels): this
tests the boundary, the kernels are tested on the GPU.  Skipped where the reference is not present (e.g. the GPU box)."""
import json
import os
import subprocess
import sys

import pytest

from tests.golden_utils import load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("B200_REFERENCE_PATH", "/root/reference")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "nemo_automodel")), reason="reference checkout not present")


def _run(name, loss_kind, steps):
    env = dict(os.environ, PYTHONPATH=ROOT, TORCHDYNAMO_DISABLE="1", MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "run_reference_recipe_b200.py"), name, loss_kind, str(steps)],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("B200_DROPIN_RESULT ")]
    assert r.returncode == 0 and lines, (r.stdout[-2000:], r.stderr[-4000:])
    return json.loads(lines[-1][len("B200_DROPIN_RESULT "):])


@pytest.mark.parametrize("name,loss_kind,steps,loss_tol", [("hd128_fp32", "reference_loss", 2, 4e-3), ("tiny_bf16", "fused_loss", 8, 1e-3)])
def test_reference_recipe_trains_through_b200_strategy(name, loss_kind, steps, loss_tol):
    rec = _run(name, loss_kind, steps)
    _, meta = load(name)
    assert rec["model_class"] == "B200CausalLM" and rec["optimizer_class"] == "B200FusedAdamW"
    assert rec["loss_class"] == ("B200MaskedCrossEntropy" if loss_kind == "fused_loss" else "MaskedCrossEntropy")
    assert all(rec["ids_match"]), "the reference data loader fed different batches than in the fixture run"
    assert rec["max_grad_norm"] == meta["max_grad_norm"]
    n = len(rec["loss"])
    assert n == min(steps, len(meta["loss"]))
    for s in range(n):
        assert rec["num_label_tokens"][s] == meta["num_label_tokens"][s]
        assert abs(rec["loss"][s] - meta["loss"][s]) < loss_tol, (s, rec["loss"][s], meta["loss"][s])
        assert abs(rec["grad_norm"][s] - meta["grad_norm"][s]) < 2e-2 * meta["grad_norm"][s], (s, rec["grad_norm"][s], meta["grad_norm"][s])
