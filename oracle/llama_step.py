"""CPU oracle: numpy restatement of the reference's Llama training step.  TEST INFRASTRUCTURE ONLY.

Explicit forward + hand-derived backward + grad-norm/clip + AdamW for the custom Llama the
reference trains on its hot path.  Each function cites the reference code it restates
(paths relative to /root/reference/nemo_automodel/).  Pinned against the reference itself:
``tests/golden/*.npz`` are produced by ``tests/golden/gen_fixtures.py`` running the unmodified
reference recipe on CPU, and ``tests/test_oracle.py`` checks this file against them
(loss, grad_norm, per-parameter gradients, updated weights).

``prec``:
  "fp64"/"fp32" - all math in that precision (fp32 reference run: torch_dtype float32).
  "bf16"        - fp32 math with a round-to-bf16 at every tensor boundary where the reference's
                  eager bf16 execution materialises a bf16 tensor (autograd included).

Nothing under ``automodel_b200/`` imports this module.
"""
import math
import numpy as np

from .portable_init import round_to_bf16

IGNORE_INDEX = -100  # components/loss/masked_ce.py:30


class Prec:
    def __init__(self, prec):
        self.name = prec
        self.dt = np.float64 if prec == "fp64" else np.float32
        self.bf16 = prec == "bf16"

    def r(self, a):
        """Materialise a tensor in the model dtype."""
        a = np.asarray(a, dtype=self.dt)
        return round_to_bf16(a) if self.bf16 else a


# ----------------------------------------------------------------------------- rope
def rope_inv_freq(cfg):
    """components/models/llama/rope_utils.py:108-150 (_compute_default_inv_freq / _compute_llama3_inv_freq)."""
    d = cfg.get("head_dim") or cfg["hidden_size"] // cfg["num_attention_heads"]
    base = cfg.get("rope_theta", 10000.0)
    inv = (1.0 / (np.float32(base) ** (np.arange(0, d, 2, dtype=np.float32) / np.float32(d)))).astype(np.float32)
    sc = cfg.get("rope_scaling") or {}
    rtype = sc.get("rope_type", sc.get("type", "default"))
    if rtype == "default":
        return inv
    factor = sc.get("factor", 1.0)
    lo, hi = sc.get("low_freq_factor", 1.0), sc.get("high_freq_factor", 4.0)
    old = sc.get("original_max_position_embeddings", cfg["max_position_embeddings"])
    low_wl, high_wl = old / lo, old / hi
    wavelen = (2 * math.pi / inv).astype(np.float32)
    inv_l = np.where(wavelen > low_wl, inv / factor, inv).astype(np.float32)
    smooth = (old / wavelen - lo) / (hi - lo)
    smoothed = (1 - smooth) * inv_l / factor + smooth * inv_l
    med = (~(wavelen < high_wl)) & (~(wavelen > low_wl))
    return np.where(med, smoothed, inv_l).astype(np.float32)


def rope_tables(cfg, seq_len, P):
    """cos/sin cache [S, d] in the model dtype (rope_utils.py:191-205: fp32 math, then .to(dtype))."""
    inv = rope_inv_freq(cfg)
    t = np.arange(seq_len, dtype=np.float32)
    freqs = np.outer(t, inv).astype(np.float32)
    emb = np.concatenate([freqs, freqs], axis=-1)
    return P.r(np.cos(emb)), P.r(np.sin(emb))


def rotate_half(x):  # rope_utils.py:39-43
    h = x.shape[-1] // 2
    return np.concatenate([-x[..., h:], x[..., :h]], axis=-1)


def rotate_half_T(y):  # adjoint of rotate_half
    h = y.shape[-1] // 2
    return np.concatenate([y[..., h:], -y[..., :h]], axis=-1)


def rope_fwd(x, cos, sin, P):
    """x [b,H,S,d]; rope_utils.py:46-67: (q*cos) + (rotate_half(q)*sin), each op materialised."""
    return P.r(P.r(x * cos) + P.r(rotate_half(x) * sin))


def rope_bwd(dy, cos, sin, P):
    return P.r(P.r(dy * cos) + rotate_half_T(P.r(dy * sin)))


# ----------------------------------------------------------------------------- rmsnorm
def rmsnorm_fwd(x, w, eps, P):
    """components/models/common/utils.py:250-256 (Float32RMSNorm): fp32 norm, (w * xhat) in fp32, one down-cast."""
    x = x.astype(P.dt)
    r = 1.0 / np.sqrt((x * x).mean(-1, keepdims=True) + P.dt(eps))
    return P.r(w * (x * r)), r


def rmsnorm_bwd(dy, x, w, r, P):
    x = x.astype(P.dt)
    xhat = x * r
    dw = P.r((dy * xhat).reshape(-1, x.shape[-1]).sum(0))
    dxhat = dy * w
    dx = r * (dxhat - xhat * (dxhat * xhat).mean(-1, keepdims=True))
    return P.r(dx), dw


# ----------------------------------------------------------------------------- attention
def attention_fwd(q, k, v, scale, seg, P):
    """Causal GQA softmax(QK^T*scale)V (models/llama/model.py:135-148 via the HF sdpa interface).
    q [b,Hq,S,d], k/v [b,Hkv,S,d].  seg: [b,S] int document ids (packed sequences attend within a
    document only) or None."""
    b, Hq, S, d = q.shape
    g = Hq // k.shape[1]
    kk = np.repeat(k, g, axis=1)
    vv = np.repeat(v, g, axis=1)
    s = (q @ kk.transpose(0, 1, 3, 2)).astype(P.dt) * P.dt(scale)
    mask = np.tril(np.ones((S, S), dtype=bool))[None, None]
    if seg is not None:
        mask = mask & (seg[:, None, :, None] == seg[:, None, None, :])
    s = np.where(mask, s, -np.inf)
    m = s.max(-1, keepdims=True)
    p = np.exp(s - m)
    p = p / p.sum(-1, keepdims=True)
    o = p @ vv
    return P.r(o), p


def attention_bwd(do, q, k, v, p, scale, P):
    b, Hq, S, d = q.shape
    Hkv = k.shape[1]
    g = Hq // Hkv
    kk = np.repeat(k, g, axis=1)
    vv = np.repeat(v, g, axis=1)
    dv = p.transpose(0, 1, 3, 2) @ do
    dp = do @ vv.transpose(0, 1, 3, 2)
    ds = p * (dp - (dp * p).sum(-1, keepdims=True)) * P.dt(scale)
    dq = ds @ kk
    dk = ds.transpose(0, 1, 3, 2) @ q
    dk = dk.reshape(b, Hkv, g, S, d).sum(2)
    dv = dv.reshape(b, Hkv, g, S, d).sum(2)
    return P.r(dq), P.r(dk), P.r(dv)


# ----------------------------------------------------------------------------- mlp
def silu(x):
    return x / (1.0 + np.exp(-x))


def swiglu_fwd(g, u, P):
    """models/llama/model.py:170: act_fn(gate) * up, two materialised ops."""
    return P.r(P.r(silu(g)) * u)


def swiglu_bwd(da, g, u, P):
    sg = 1.0 / (1.0 + np.exp(-g))
    s = P.r(g * sg)
    du = P.r(da * s)
    ds = P.r(da * u)
    dg = P.r(ds * (sg * (1.0 + g * (1.0 - sg))))
    return dg, du


# ----------------------------------------------------------------------------- loss
def masked_ce_fwd_bwd(logits, labels, num_label_tokens, P):
    """components/loss/masked_ce.py:73-89: logits.float(); F.cross_entropy(sum, ignore_index=-100) / num_label_tokens.
    Returns (loss, dlogits) with dlogits = d loss / d logits (model dtype)."""
    V = logits.shape[-1]
    z = logits.reshape(-1, V).astype(P.dt)
    y = labels.reshape(-1)
    valid = y != IGNORE_INDEX
    m = z.max(-1, keepdims=True)
    e = np.exp(z - m)
    se = e.sum(-1, keepdims=True)
    lse = (np.log(se) + m)[:, 0]
    ysafe = np.where(valid, y, 0)
    nll = lse - z[np.arange(z.shape[0]), ysafe]
    if num_label_tokens == 0:
        return 0.0, np.zeros_like(logits)
    loss = float((nll * valid).sum(dtype=np.float64) / num_label_tokens)
    dz = e / se
    dz[np.arange(z.shape[0]), ysafe] -= 1.0
    dz = dz * (valid[:, None] / P.dt(num_label_tokens))
    return loss, P.r(dz.reshape(logits.shape))


# ----------------------------------------------------------------------------- model
def _mm(a, b):
    return a @ b


def linear_fwd(x, w, P, b=None):
    """nn.Linear: y = x W^T (+ b) (W [out,in], HF layout).  With a bias (Qwen2 q/k/v, components/models/qwen2/model.py:80-82) the add
    happens before the one down-cast (addmm)."""
    y = _mm(x, w.T)
    return P.r(y if b is None else y + b)


def linear_bwd(dy, x, w, P):
    dx = P.r(_mm(dy, w))
    dw = P.r(_mm(dy.reshape(-1, dy.shape[-1]).T, x.reshape(-1, x.shape[-1])))
    return dx, dw


def bias_bwd(dy, P):
    return P.r(dy.reshape(-1, dy.shape[-1]).sum(0))


def segments_from_position_ids(position_ids):
    """Packed batches restart position_ids at each document (components/datasets/llm/packed_sequence.py:37-110);
    a new document starts wherever position_ids == 0."""
    starts = (position_ids == 0)
    return np.cumsum(starts, axis=1)


def forward_backward(params, cfg, input_ids, labels, num_label_tokens, prec="fp32", position_ids=None,
                     grads=None, compute_grads=True):
    """One micro-batch: loss (already divided by the GLOBAL label-token count, recipes/llm/train_ft.py:1449-1473)
    and parameter gradients accumulated into ``grads`` (dict name -> array)."""
    P = Prec(prec)
    H, Hkv = cfg["num_attention_heads"], cfg["num_key_value_heads"]
    hsz = cfg["hidden_size"]
    d = cfg.get("head_dim") or hsz // H
    L = cfg["num_hidden_layers"]
    eps = cfg.get("rms_norm_eps", 1e-5)
    b, S = input_ids.shape
    scale = d ** -0.5
    W = {k: np.asarray(v, dtype=P.dt) for k, v in params.items()}
    if position_ids is None:
        position_ids = np.broadcast_to(np.arange(S), (b, S))
    # The reference's LlamaRotaryEmbedding takes only the LENGTH from position_ids and returns the tables of positions [0, S)
    # (components/models/llama/rope_utils.py:212-235): a packed row is rotated by its row index, not by the restarting position_ids.
    # Within a document this is the same rotation up to the rounding of the table entries (RoPE is relative); position_ids only
    # delimit the documents (HF flash-attention derives cu_seqlens from them).
    row_pos = np.broadcast_to(np.arange(S), (b, S))
    cos_t, sin_t = rope_tables(cfg, S, P)
    cos = cos_t[row_pos][:, None]  # [b,1,S,d]
    sin = sin_t[row_pos][:, None]
    packed = bool((position_ids[:, 1:] <= position_ids[:, :-1]).any())
    seg = segments_from_position_ids(position_ids) if packed else None

    # ---------------- forward (models/llama/model.py:293-388, 203-234, 101-152)
    h = W["model.embed_tokens.weight"][input_ids]  # [b,S,h]
    saved = []
    for l in range(L):
        p = f"model.layers.{l}."
        x1, r1 = rmsnorm_fwd(h, W[p + "input_layernorm.weight"], eps, P)
        q = linear_fwd(x1, W[p + "self_attn.q_proj.weight"], P, W.get(p + "self_attn.q_proj.bias")).reshape(b, S, H, d).transpose(0, 2, 1, 3)
        k = linear_fwd(x1, W[p + "self_attn.k_proj.weight"], P, W.get(p + "self_attn.k_proj.bias")).reshape(b, S, Hkv, d).transpose(0, 2, 1, 3)
        v = linear_fwd(x1, W[p + "self_attn.v_proj.weight"], P, W.get(p + "self_attn.v_proj.bias")).reshape(b, S, Hkv, d).transpose(0, 2, 1, 3)
        qr, kr = rope_fwd(q, cos, sin, P), rope_fwd(k, cos, sin, P)
        o, pr = attention_fwd(qr, kr, v, scale, seg, P)
        o2 = o.transpose(0, 2, 1, 3).reshape(b, S, H * d)
        h1 = P.r(h + linear_fwd(o2, W[p + "self_attn.o_proj.weight"], P))
        x2, r2 = rmsnorm_fwd(h1, W[p + "post_attention_layernorm.weight"], eps, P)
        g = linear_fwd(x2, W[p + "mlp.gate_proj.weight"], P)
        u = linear_fwd(x2, W[p + "mlp.up_proj.weight"], P)
        a = swiglu_fwd(g, u, P)
        h2 = P.r(h1 + linear_fwd(a, W[p + "mlp.down_proj.weight"], P))
        saved.append((h, x1, r1, qr, kr, v, pr, o2, h1, x2, r2, g, u, a))
        h = h2
    xf, rf = rmsnorm_fwd(h, W["model.norm.weight"], eps, P)
    head = "lm_head.weight" if "lm_head.weight" in W else "model.embed_tokens.weight"    # tie_word_embeddings: one shared matrix
    logits = linear_fwd(xf, W[head], P)
    loss, dlogits = masked_ce_fwd_bwd(logits, labels, num_label_tokens, P)
    if not compute_grads:
        return loss, logits

    # ---------------- backward
    if grads is None:
        grads = {}

    def acc(name, g_):
        grads[name] = P.r(grads[name] + g_) if name in grads else g_

    dxf, dw = linear_bwd(dlogits, xf, W[head], P)
    acc(head, dw)
    dh, dw = rmsnorm_bwd(dxf, h, W["model.norm.weight"], rf, P)
    acc("model.norm.weight", dw)
    for l in reversed(range(L)):
        p = f"model.layers.{l}."
        (h0, x1, r1, qr, kr, v, pr, o2, h1, x2, r2, g, u, a) = saved[l]
        da, dw = linear_bwd(dh, a, W[p + "mlp.down_proj.weight"], P)
        acc(p + "mlp.down_proj.weight", dw)
        dg, du = swiglu_bwd(da, g, u, P)
        dx2g, dw = linear_bwd(dg, x2, W[p + "mlp.gate_proj.weight"], P)
        acc(p + "mlp.gate_proj.weight", dw)
        dx2u, dw = linear_bwd(du, x2, W[p + "mlp.up_proj.weight"], P)
        acc(p + "mlp.up_proj.weight", dw)
        dx2 = P.r(dx2g + dx2u)
        dh1n, dw = rmsnorm_bwd(dx2, h1, W[p + "post_attention_layernorm.weight"], r2, P)
        acc(p + "post_attention_layernorm.weight", dw)
        dh1 = P.r(dh + dh1n)
        do2, dw = linear_bwd(dh1, o2, W[p + "self_attn.o_proj.weight"], P)
        acc(p + "self_attn.o_proj.weight", dw)
        do = do2.reshape(b, S, H, d).transpose(0, 2, 1, 3)
        dqr, dkr, dv = attention_bwd(do, qr, kr, v, pr, scale, P)
        dq, dk = rope_bwd(dqr, cos, sin, P), rope_bwd(dkr, cos, sin, P)
        dq2 = dq.transpose(0, 2, 1, 3).reshape(b, S, H * d)
        dk2 = dk.transpose(0, 2, 1, 3).reshape(b, S, Hkv * d)
        dv2 = dv.transpose(0, 2, 1, 3).reshape(b, S, Hkv * d)
        dx1q, dw = linear_bwd(dq2, x1, W[p + "self_attn.q_proj.weight"], P)
        acc(p + "self_attn.q_proj.weight", dw)
        dx1k, dw = linear_bwd(dk2, x1, W[p + "self_attn.k_proj.weight"], P)
        acc(p + "self_attn.k_proj.weight", dw)
        dx1v, dw = linear_bwd(dv2, x1, W[p + "self_attn.v_proj.weight"], P)
        acc(p + "self_attn.v_proj.weight", dw)
        if p + "self_attn.q_proj.bias" in W:
            acc(p + "self_attn.q_proj.bias", bias_bwd(dq2, P))
            acc(p + "self_attn.k_proj.bias", bias_bwd(dk2, P))
            acc(p + "self_attn.v_proj.bias", bias_bwd(dv2, P))
        dx1 = P.r(P.r(dx1q + dx1k) + dx1v)
        dh0n, dw = rmsnorm_bwd(dx1, h0, W[p + "input_layernorm.weight"], r1, P)
        acc(p + "input_layernorm.weight", dw)
        dh = P.r(dh1 + dh0n)
    demb = np.zeros_like(W["model.embed_tokens.weight"])
    np.add.at(demb, input_ids.reshape(-1), dh.reshape(-1, hsz))
    acc("model.embed_tokens.weight", P.r(demb))
    return loss, grads


# ----------------------------------------------------------------------------- clip + optimizer
def grad_norm_and_clip(grads, max_norm, prec="fp32"):
    """components/training/utils.py:122-141,168-169: sqrt(sum_p sum(|g.float()|^2)) accumulated in fp32,
    then torch clip_grads_with_norm_: coef = max_norm/(norm+1e-6) clamped to 1, grads *= coef (in grad dtype)."""
    P = Prec(prec)
    tot = np.float32(0.0)
    for g in grads.values():
        tot = np.float32(tot + np.float32((g.astype(np.float32) ** 2).sum(dtype=np.float32)))
    total_norm = float(np.sqrt(tot))
    if max_norm is not None:
        coef = np.float32(min(np.float32(max_norm) / (np.float32(total_norm) + np.float32(1e-6)), 1.0))
        for k in grads:
            grads[k] = P.r(grads[k] * coef)
    return total_norm


class AdamW:
    """torch.optim.AdamW (decoupled weight decay) on the local shard (recipes/llm/train_ft.py:1556-1558).

    prec "bf16" follows torch's op sequence for bf16 params/states (each foreach op materialises a
    bf16 tensor: mul_, lerp_, mul_, addcmul_, sqrt, div_, add_, addcdiv_  - torch/optim/adamw.py ->
    adam.py _single_tensor_adam/_multi_tensor_adam), which is what the reference's default YAML
    optimizer does on bf16 parameters."""

    def __init__(self, lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1, prec="fp32"):
        self.lr, self.b1, self.b2, self.eps, self.wd = lr, betas[0], betas[1], eps, weight_decay
        self.P = Prec(prec)
        self.t = 0
        self.m, self.v = {}, {}

    def _step_inplace_threaded(self, params, grads, step_size, bc2_sqrt):
        """Same update as the loop in step() for fp32/fp64 (no bf16 rounding points), done in place on ~8M-element chunks
        across a thread pool (numpy releases the GIL), so the timed CPU baseline uses all host cores like torch's CPU ops do."""
        import os
        from concurrent.futures import ThreadPoolExecutor
        f = self.P.dt
        decay, w1, b2, w2, eps = f(1 - self.lr * self.wd), f(1 - self.b1), f(self.b2), f(1 - self.b2), f(self.eps)
        ss, bs = f(step_size), f(bc2_sqrt)
        jobs = []
        for k, p in params.items():
            if k not in self.m:
                self.m[k] = np.zeros_like(p, dtype=f)
                self.v[k] = np.zeros_like(p, dtype=f)
            if p.dtype != f or not p.flags.c_contiguous:
                params[k] = p = np.ascontiguousarray(p, dtype=f)
            g = np.ascontiguousarray(grads[k], dtype=f)
            pf, gf, mf, vf = p.reshape(-1), g.reshape(-1), self.m[k].reshape(-1), self.v[k].reshape(-1)
            for a in range(0, pf.size, 1 << 23):
                jobs.append((pf[a:a + (1 << 23)], gf[a:a + (1 << 23)], mf[a:a + (1 << 23)], vf[a:a + (1 << 23)]))

        def work(j):
            p, g, m, v = j
            t = np.empty_like(p)
            p *= decay
            np.subtract(g, m, out=t); t *= w1; m += t
            v *= b2
            np.multiply(g, g, out=t); t *= w2; v += t
            np.sqrt(v, out=t); t /= bs; t += eps
            np.divide(m, t, out=t); t *= ss
            p -= t

        try:
            nthr = len(os.sched_getaffinity(0))
        except Exception:
            nthr = os.cpu_count() or 1
        with ThreadPoolExecutor(max_workers=max(1, nthr)) as ex:
            list(ex.map(work, jobs))
        return params

    def step(self, params, grads, fast=False):
        P = self.P
        self.t += 1
        f = P.dt
        bc1 = 1 - self.b1 ** self.t
        bc2 = 1 - self.b2 ** self.t
        step_size = self.lr / bc1
        bc2_sqrt = math.sqrt(bc2)
        if fast and not P.bf16:
            return self._step_inplace_threaded(params, grads, step_size, bc2_sqrt)
        for k, p in params.items():
            g = grads[k].astype(f)
            if k not in self.m:
                self.m[k] = np.zeros_like(p, dtype=f)
                self.v[k] = np.zeros_like(p, dtype=f)
            m, v = self.m[k], self.v[k]
            p = P.r(p.astype(f) * f(1 - self.lr * self.wd))
            # lerp_(grad, 1-beta1): weight < 0.5 -> m + w*(g-m)
            w = f(1 - self.b1)
            m = P.r(m + w * (g - m)) if w < 0.5 else P.r(g - (g - m) * (1 - w))
            v = P.r(v * f(self.b2))
            v = P.r(v + f(1 - self.b2) * g * g)
            den = P.r(np.sqrt(v))
            den = P.r(den / f(bc2_sqrt))
            den = P.r(den + f(self.eps))
            p = P.r(p + f(-step_size) * (m / den))
            params[k] = p
            self.m[k], self.v[k] = m, v
        return params


def train_step(params, opt, cfg, micro_batches, prec="fp32", max_grad_norm=1.0, timing=False):
    """One optimizer step over a list of micro-batches (recipes/llm/train_ft.py:1482-1635), world size 1.
    micro_batches: list of dicts with input_ids, labels (and optional position_ids).  Returns (loss, grad_norm, grads_pre_clip)."""
    n_lab = int(sum((mb["labels"] != IGNORE_INDEX).sum() for mb in micro_batches))
    grads = {}
    loss = 0.0
    for mb in micro_batches:
        l, grads = forward_backward(params, cfg, mb["input_ids"], mb["labels"], n_lab, prec,
                                    position_ids=mb.get("position_ids"), grads=grads)
        loss += l
    pre = None if timing else {k: v.copy() for k, v in grads.items()}   # timing=True: CPU-baseline run, skip the debug copy
    gn = grad_norm_and_clip(grads, max_grad_norm, prec)
    opt.step(params, grads, fast=timing)
    return loss, gn, pre


def mock_labels(input_ids):
    """components/datasets/llm/mock_iterable_dataset.py:41-59: labels = tokens shifted left, last position -100."""
    lab = np.full_like(input_ids, IGNORE_INDEX, dtype=np.int64)
    lab[:, :-1] = input_ids[:, 1:]
    return lab
