"""CPU stand-in for automodel_b200.ops (TEST INFRASTRUCTURE ONLY - never imported by the product).

Same call signatures as the CUDA wrappers, implemented with plain torch CPU ops (bf16 storage, fp32 math, the same
rounding points as the kernels).  It lets the engine's orchestration - flat layout, fused-weight views, accumulate flags,
reduce-scatter / all-gather ordering over gloo - be tested here without a GPU, and is an executable spec of each kernel.
"""
import math
import torch

NT, NN, TN = 0, 1, 2
BF = torch.bfloat16


def _r(x):
    return x.to(BF)


def gemm(kind, a, b, out=None, residual=None, round_before_add=True, group_m=0, max_ctas=0):
    af, bf_ = a.float(), b.float()
    acc = af @ bf_.t() if kind == NT else af @ bf_ if kind == NN else af.t() @ bf_
    if residual is not None:
        acc = (_r(acc).float() if round_before_add else acc) + residual.float()
    res = _r(acc)
    if out is None:
        return res
    out.copy_(res)
    return out


def rmsnorm_fwd(x, w, eps, out=None, rstd=None):
    xf = x.float()
    r = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    y = _r(w.float() * (xf * r))
    if out is not None:
        out.copy_(y); y = out
    if rstd is not None:
        rstd.copy_(r[:, 0]); r2 = rstd
    else:
        r2 = r[:, 0].contiguous()
    return y, r2


def rmsnorm_bwd(dy, x, w, rstd, dres=None, dx=None, dw=None, accumulate_dw=False, workspace=None):
    xf, r = x.float(), rstd[:, None]
    xhat = xf * r
    dxhat = dy.float() * w.float()
    d = r * (dxhat - xhat * (dxhat * xhat).mean(-1, keepdim=True))
    if dres is not None:
        d = _r(d).float() + dres.float()
    d = _r(d)
    g = (dy.float() * xhat).sum(0)
    if accumulate_dw:
        g = _r(g).float() + dw.float()
    g = _r(g)
    if dx is not None:
        dx.copy_(d); d = dx
    if dw is not None:
        dw.copy_(g); g = dw
    return d, g


def rope_(qk, cos, sin, pos, heads, head_dim, backward=False):
    T = qk.shape[0]
    D = head_dim
    x = qk[:, :heads * D].reshape(T, heads, D).float()
    c, s = cos[pos.long()][:, None].float(), sin[pos.long()][:, None].float()
    if not backward:
        rot = torch.cat([-x[..., D // 2:], x[..., :D // 2]], -1)
        y = _r(_r(x * c).float() + _r(rot * s).float())
    else:
        t1 = _r(x * s).float()
        rotT = torch.cat([t1[..., D // 2:], -t1[..., :D // 2]], -1)
        y = _r(_r(x * c).float() + rotT)
    qk[:, :heads * D] = y.reshape(T, heads * D)
    return qk


def bias_rope_(qkv, bias, cos, sin, pos, rope_heads, heads, head_dim):
    qkv.copy_(_r(qkv.float() + bias.float()[None, :]))
    return rope_(qkv, cos, sin, pos, rope_heads, head_dim)


def colsum_(x, out, accumulate=False):
    t = x.float().sum(0)
    out.copy_(_r(_r(t).float() + out.float()) if accumulate else _r(t))
    return out


def swiglu_fwd(gu, out=None):
    F = gu.shape[1] // 2
    g, u = gu[:, :F].float(), gu[:, F:].float()
    a = _r(_r(g * torch.sigmoid(g)).float() * u)
    if out is not None:
        out.copy_(a); a = out
    return a


def swiglu_bwd(da, gu, out=None):
    F = gu.shape[1] // 2
    g, u, d = gu[:, :F].float(), gu[:, F:].float(), da.float()
    sg = torch.sigmoid(g)
    du = _r(d * _r(g * sg).float())
    dg = _r(_r(d * u).float() * (sg * (1 + g * (1 - sg))))
    res = torch.cat([dg, du], 1)
    if out is not None:
        out.copy_(res); res = out
    return res


def embed_fwd(ids, W, out=None):
    r = W[ids.long()]
    if out is not None:
        out.copy_(r); r = out
    return r


def embed_bwd(ids, dh, dW, accumulate=False, workspace=None):
    acc = torch.zeros(dW.shape, dtype=torch.float32).index_add_(0, ids.long(), dh.float())
    touched = torch.zeros(dW.shape[0], dtype=torch.bool)
    touched[ids.long()] = True
    new = _r(acc).float() + dW.float() if accumulate else acc
    dW[touched] = _r(new)[touched]
    return dW


def _segments(cu):
    cu = cu.tolist()
    return list(zip(cu[:-1], cu[1:]))


def attn_fwd(q, k, v, cu_seqlens, max_seqlen, Hq, Hkv, D, scale=None, out=None, lse=None):
    T = q.shape[0]
    scale = scale if scale is not None else D ** -0.5
    g = Hq // Hkv
    o = torch.empty(T, Hq * D, dtype=torch.float32)
    l = torch.empty(Hq, T, dtype=torch.float32)
    for a, b in _segments(cu_seqlens):
        L = b - a
        qs = q[a:b].float().reshape(L, Hq, D).transpose(0, 1)
        ks = k[a:b].float().reshape(L, Hkv, D).transpose(0, 1).repeat_interleave(g, 0)
        vs = v[a:b].float().reshape(L, Hkv, D).transpose(0, 1).repeat_interleave(g, 0)
        s = (qs @ ks.transpose(1, 2)) * scale
        s = s.masked_fill(~torch.tril(torch.ones(L, L, dtype=torch.bool)), float("-inf"))
        l[:, a:b] = torch.logsumexp(s, -1)
        p = _r(torch.softmax(s, -1)).float()  # kernel rounds P to bf16 for the PV product
        o[a:b] = (p @ vs).transpose(0, 1).reshape(L, Hq * D)
    o = _r(o)
    if out is not None:
        out.copy_(o); o = out
    if lse is not None:
        lse[:, :T].copy_(l) if lse.shape[1] != T else lse.copy_(l)
        l = lse
    return o, l


def attn_bwd(q, k, v, o, dout, lse, cu_seqlens, max_seqlen, Hq, Hkv, D, dq, dk, dv, scale=None, workspace=None):
    T = q.shape[0]
    scale = scale if scale is not None else D ** -0.5
    g = Hq // Hkv
    for a, b in _segments(cu_seqlens):
        L = b - a
        qs = q[a:b].float().reshape(L, Hq, D).transpose(0, 1)
        ks = k[a:b].float().reshape(L, Hkv, D).transpose(0, 1).repeat_interleave(g, 0)
        vs = v[a:b].float().reshape(L, Hkv, D).transpose(0, 1).repeat_interleave(g, 0)
        do = dout[a:b].float().reshape(L, Hq, D).transpose(0, 1)
        os_ = o[a:b].float().reshape(L, Hq, D).transpose(0, 1)
        s = (qs @ ks.transpose(1, 2)) * scale
        mask = torch.tril(torch.ones(L, L, dtype=torch.bool))
        p = torch.exp(s - lse[:, a:b, None]).masked_fill(~mask, 0.0)
        delta = (do * os_).sum(-1, keepdim=True)
        dvh = _r(p).float().transpose(1, 2) @ do
        dp = do @ vs.transpose(1, 2)
        ds = _r(p * (dp - delta)).float()
        dqh = (ds @ ks) * scale
        dkh = (ds.transpose(1, 2) @ qs) * scale
        dq[a:b] = _r(dqh.transpose(0, 1).reshape(L, Hq * D))
        dk[a:b] = _r(dkh.reshape(Hkv, g, L, D).sum(1).transpose(0, 1).reshape(L, Hkv * D))
        dv[a:b] = _r(dvh.reshape(Hkv, g, L, D).sum(1).transpose(0, 1).reshape(L, Hkv * D))
    return dq, dk, dv


def ce_fwd_bwd_(logits, labels, num_label_tokens, loss_out, accumulate=False, row_loss=None):
    z = logits.float()
    valid = labels != -100
    lse = torch.logsumexp(z, -1)
    y = labels.long().clamp_min(0)
    nll = (lse - z.gather(1, y[:, None])[:, 0]) * valid
    inv = 1.0 / num_label_tokens if num_label_tokens > 0 else 0.0
    d = torch.softmax(z, -1)
    d[torch.arange(z.shape[0]), y] -= 1.0
    d = d * (valid[:, None] * inv)
    logits.copy_(_r(d))
    tot = nll.sum() * inv
    loss_out[0] = (loss_out[0] if accumulate else 0.0) + tot
    return loss_out


def sumsq_(g, out, accumulate=False):
    s = g.float().pow(2).sum()
    out[0] = (out[0] if accumulate else 0.0) + s
    return out


def adamw_step_(p, g, m, v, lr, beta1, beta2, eps, wd, step, max_grad_norm=0.0, grad_norm_sq=None, mode=0, master=None):
    f32 = lambda x: torch.tensor(x, dtype=torch.float64).float()
    coef = 1.0
    if grad_norm_sq is not None and max_grad_norm and max_grad_norm > 0:
        coef = min(max_grad_norm / (float(grad_norm_sq[0].sqrt()) + 1e-6), 1.0)
    coef = torch.tensor(coef, dtype=torch.float32)
    decay, w1, w2 = f32(1.0 - lr * wd), f32(1.0 - beta1), f32(1.0 - beta2)
    step_size = f32(lr / (1.0 - beta1 ** step))
    bc2 = f32(math.sqrt(1.0 - beta2 ** step))
    pf = master.clone() if master is not None else p.float()
    gf, mf, vf = g.float(), m.float(), v.float()
    if mode == 1:
        R = lambda x: _r(x).float()
        gg = R(gf * coef)
        pp = R(pf * decay)
        mm = R(mf + w1 * (gg - mf))
        vv = R(vf * f32(beta2))
        vv = R(vv + w2 * gg * gg)
        den = R(torch.sqrt(vv)); den = R(den / bc2); den = R(den + f32(eps))
        pp = R(pp + (-step_size) * (mm / den))
    else:
        gg = gf * coef
        mm = f32(beta1) * mf + w1 * gg
        vv = f32(beta2) * vf + w2 * gg * gg
        den = torch.sqrt(vv) / bc2 + f32(eps)
        pp = pf * decay - step_size * (mm / den)
    p.copy_(_r(pp)); m.copy_(_r(mm)); v.copy_(_r(vv))
    if master is not None:
        master.copy_(pp)


def add_(dst, src):
    dst.copy_(_r(dst.float() + src.float()))
    return dst
