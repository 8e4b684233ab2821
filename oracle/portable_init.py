"""Portable, bit-reproducible parameter initialisation for parity fixtures (test infrastructure).

The reference initialises weights *after* sharding with a rank-offset torch Philox stream
(/root/reference/nemo_automodel/_transformers/infrastructure.py:546-552,
 components/checkpoint/checkpointing.py:574-676), which cannot be reproduced outside that exact
torch build.  Parity therefore starts from a weight snapshot that BOTH sides load.  Instead of
committing megabytes of snapshot, the snapshot is a pure integer function of (param name, index):
splitmix64 -> four 16-bit uniforms -> Irwin-Hall sum (approximately normal) -> scaled by
``std`` (HF ``initializer_range`` 0.02, matching ``PreTrainedModel.initialize_weights``) and
rounded to a bf16-representable value, so fp32 and bf16 runs share the same start point and no
libm call is involved (bit-identical on any IEEE-754 host).
Norm weights are 1 + 0.1*z (instead of exactly 1) so the norm-weight multiply and its gradient
are exercised by the fixtures.
"""
import zlib
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def round_to_bf16(a):
    """Round-to-nearest-even fp32 -> bf16, returned as fp32 (numpy has no bf16)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32)
    bias = np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))
    r = ((u + bias) & np.uint32(0xFFFF0000)).astype(np.uint32)
    return r.view(np.float32).reshape(a.shape)


def portable_normal(name, shape, seed=0):
    """~N(0,1) samples (Irwin-Hall of four 16-bit uniforms), exact integer arithmetic, float64 out."""
    n = int(np.prod(shape))
    key = np.uint64(zlib.crc32(name.encode()) & 0xFFFFFFFF) << np.uint64(32)
    key = key ^ np.uint64(seed & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64)
        h = _splitmix64(_splitmix64(idx ^ key) + key)
    s = ((h & np.uint64(0xFFFF)).astype(np.int64) + ((h >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.int64)
         + ((h >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.int64) + (h >> np.uint64(48)).astype(np.int64) - 131070)
    # std of the sum of four U{0..65535} is 65536/sqrt(3) = 37837.22...; use the exact integer 37837
    return (s.astype(np.float64) / 37837.0).reshape(shape)


def portable_state_dict(shapes, seed=0, std=0.02):
    """shapes: dict name -> shape (HF Llama names).  Returns dict name -> fp32 array (bf16-representable)."""
    out = {}
    for name, shape in shapes.items():
        z = portable_normal(name, shape, seed)
        if name.endswith("norm.weight") or "layernorm" in name:
            w = 1.0 + 0.1 * z
        else:
            w = std * z
        out[name] = round_to_bf16(w.astype(np.float32))
    return out


def llama_param_shapes(cfg):
    """HF-layout parameter names/shapes of LlamaForCausalLM (untied), as in
    /root/reference/nemo_automodel/components/models/llama/model.py:85-101,162-166,282-288,440-444."""
    h, f, v = cfg["hidden_size"], cfg["intermediate_size"], cfg["vocab_size"]
    nh, nkv = cfg["num_attention_heads"], cfg["num_key_value_heads"]
    d = cfg.get("head_dim") or h // nh
    s = {"model.embed_tokens.weight": (v, h)}
    for l in range(cfg["num_hidden_layers"]):
        p = f"model.layers.{l}."
        s[p + "self_attn.q_proj.weight"] = (nh * d, h)
        s[p + "self_attn.k_proj.weight"] = (nkv * d, h)
        s[p + "self_attn.v_proj.weight"] = (nkv * d, h)
        if cfg.get("model_type") == "qwen2":       # components/models/qwen2/model.py:80-82: q/k/v projections with bias
            s[p + "self_attn.q_proj.bias"] = (nh * d,)
            s[p + "self_attn.k_proj.bias"] = (nkv * d,)
            s[p + "self_attn.v_proj.bias"] = (nkv * d,)
        s[p + "self_attn.o_proj.weight"] = (h, nh * d)
        s[p + "mlp.gate_proj.weight"] = (f, h)
        s[p + "mlp.up_proj.weight"] = (f, h)
        s[p + "mlp.down_proj.weight"] = (h, f)
        s[p + "input_layernorm.weight"] = (h,)
        s[p + "post_attention_layernorm.weight"] = (h,)
    s["model.norm.weight"] = (h,)
    if not cfg.get("tie_word_embeddings", False):
        s["lm_head.weight"] = (v, h)
    return s
