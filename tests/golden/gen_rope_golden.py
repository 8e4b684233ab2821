#!/usr/bin/env python
"""Golden RoPE vectors from the UNMODIFIED reference (components/models/llama/rope_utils.py:112-150, 191-205): inv_freq and the bf16
cos/sin tables of `LlamaRotaryEmbedding` for the default and the llama3-scaled (Llama-3-8B benchmark config) RoPE.  Test infrastructure;
needs /root/reference.  Writes tests/golden/rope_golden.npz (read by tests/test_engine_cpu.py on any host)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _ref_env  # noqa: F401
import numpy as np
import torch
import transformers
from nemo_automodel.components.models.llama.rope_utils import LlamaRotaryEmbedding

CASES = {
    "llama3_8b": dict(hidden_size=4096, num_attention_heads=32, max_position_embeddings=8192, rope_theta=500000.0,
                      rope_scaling={"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                                    "original_max_position_embeddings": 8192}),
    "default_hd64": dict(hidden_size=256, num_attention_heads=4, max_position_embeddings=512, rope_theta=10000.0),
    "llama3_small_ctx": dict(hidden_size=256, num_attention_heads=2, max_position_embeddings=256, rope_theta=500000.0,
                             rope_scaling={"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                                           "original_max_position_embeddings": 64}),
}
out = {}
for name, kw in CASES.items():
    cfg = transformers.LlamaConfig(vocab_size=512, intermediate_size=512, num_hidden_layers=1, num_key_value_heads=kw["num_attention_heads"], **kw)
    cfg.torch_dtype = torch.bfloat16
    rot = LlamaRotaryEmbedding(cfg)
    S = 512
    x = torch.zeros(1, S, 8, dtype=torch.bfloat16)
    cos, sin = rot(x, torch.arange(S)[None])
    assert cos.dtype == torch.bfloat16
    out[name + "/inv_freq"] = rot.inv_freq.float().numpy()
    out[name + "/cos"] = cos[0].float().numpy()
    out[name + "/sin"] = sin[0].float().numpy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rope_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
