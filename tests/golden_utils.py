"""Helpers to read tests/golden/*.npz (written by tests/golden/gen_fixtures.py from the reference run)."""
import json, os
import numpy as np

from oracle.portable_init import llama_param_shapes, portable_state_dict
from oracle.llama_step import mock_labels

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def model_cfg(meta):
    c = meta["config"]
    out = {
        "vocab_size": c["vocab"], "hidden_size": c["hidden"], "intermediate_size": c["ffn"],
        "num_hidden_layers": c["layers"], "num_attention_heads": c["heads"], "num_key_value_heads": c["kv"],
        "max_position_embeddings": c["seq"], "rms_norm_eps": 1e-5, "rope_theta": c["theta"],
    }
    if "llama3" in (c.get("rope_scaling") or ""):
        out["rope_scaling"] = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0, "original_max_position_embeddings": 64}
    if c.get("cfg_class") == "MistralConfig":
        out.update(model_type="mistral", sliding_window=None)
    if c.get("cfg_class") == "Qwen2Config":
        out.update(model_type="qwen2", tie_word_embeddings=c.get("tied") == "true")
    return out


def init_params(meta):
    cfg = model_cfg(meta)
    return portable_state_dict(llama_param_shapes(cfg), seed=meta["init_seed"])


def batches(z, meta, step):
    out = []
    for j in range(meta["num_micro"]):
        ids = z[f"batch/{step}/{j}/input_ids"].astype(np.int64)
        out.append({"input_ids": ids, "labels": mock_labels(ids)})
    return out


def check_summary(z, prefix, name, arr, rtol, atol, outlier_frac=0.0, outlier_atol=None):
    """Compare an array with the fixture's strided sample + sum/sumsq record.  Returns max abs err of the sample.
    ``outlier_frac``: fraction of sample elements allowed outside (rtol, atol) provided they stay within
    ``outlier_atol`` (Adam's normalised update flips sign on near-zero gradients, moving a weight by up to lr per step)."""
    flat = np.asarray(arr, dtype=np.float64).reshape(-1)
    stats = z[f"{prefix}/{name}/stats"]
    stride = int(stats[2])
    ref = z[f"{prefix}/{name}/sample"].astype(np.float64)
    got = flat[::stride][:4096]
    if outlier_frac > 0:
        bad = np.abs(got - ref) > atol + rtol * np.abs(ref)
        assert bad.mean() <= outlier_frac, f"{prefix}/{name}: {bad.mean():.4f} of sample outside tolerance"
        assert np.abs(got - ref).max() <= outlier_atol, f"{prefix}/{name}: max err {np.abs(got - ref).max()}"
    else:
        np.testing.assert_allclose(got, ref, rtol=rtol, atol=atol, err_msg=f"{prefix}/{name} sample")
    ref_l2 = np.sqrt(stats[1]); got_l2 = np.sqrt((flat * flat).sum())
    assert abs(got_l2 - ref_l2) <= rtol * ref_l2 + atol * np.sqrt(flat.size), f"{prefix}/{name} l2 {got_l2} vs {ref_l2}"
    return np.abs(got - ref).max()


def check_rel_l2(z, prefix, name, arr, tol):
    """Relative L2 error of the strided sample and of the full-tensor norm (for bf16 gradients, where per-element
    tolerances are meaningless on sparse / near-zero rows)."""
    flat = np.asarray(arr, dtype=np.float64).reshape(-1)
    stats = z[f"{prefix}/{name}/stats"]
    stride = int(stats[2])
    ref = z[f"{prefix}/{name}/sample"].astype(np.float64)
    got = flat[::stride][:4096]
    den = max(np.linalg.norm(ref), 1e-12)
    rel = np.linalg.norm(got - ref) / den
    assert rel <= tol, f"{prefix}/{name}: sample rel-l2 {rel:.4f} > {tol}"
    ref_l2 = np.sqrt(stats[1]); got_l2 = np.sqrt((flat * flat).sum())
    assert abs(got_l2 - ref_l2) <= tol * ref_l2 + 1e-12, f"{prefix}/{name}: l2 {got_l2} vs {ref_l2}"
    return rel
