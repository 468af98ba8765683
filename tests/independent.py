"""Independent fp64 torch-autograd implementation of the six CTR models.

Written from the papers' formulas (FM: Rendle 2010; DeepFM: Guo 2017; xDeepFM/CIN:
Lian 2018; DCN: Wang 2017; IPNN: Qu 2016) with the reference's documented quirks
(second-order term averaged over K; scalar cross / product biases; no bias on the
CIN / DCN output heads) -- NOT from oracle/refport.py.  It exists to pin the oracle:
the reference itself ships no tests and cannot run here.
Gradients come from autograd, so no hand-derived backward formula is shared with
the oracle.  Used by tests/ and tests/golden/make_golden.py only.
"""
import numpy as np
import torch


def _lin(params, off, i, o, bias=True):
    w = params[off:off + i * o].view(o, i)
    off += i * o
    b = None
    if bias:
        b = params[off:off + o]
        off += o
    return w, b, off


def _mlp(x, mats, off, in_dim, dims, head):
    d = in_dim
    for o in dims:
        w, b, off = _lin(mats, off, d, o)
        x = torch.relu(x @ w.t() + b)
        d = o
    if head:
        w, b, off = _lin(mats, off, d, 1)
        x = x @ w.t() + b
    return x, off


def logits(kind, B, F, K, index, w, bias, emb, mats, fc=(), cin=(), depth=0):
    """All args torch fp64 tensors (index: int64).  Returns logit [B]."""
    first = torch.zeros(B, dtype=w.dtype).index_add(0, index, w)
    out = first + bias[0]
    if kind == "lr":
        return out
    V = emb.view(B, F, K)
    if kind in ("fm", "deepfm"):
        s = V.sum(1)
        out = out + 0.5 * ((s * s - (V * V).sum(1)).mean(1))
    X = emb.view(B, F * K)
    if kind == "deepfm":
        h, _ = _mlp(X, mats, 0, F * K, fc, True)
        out = out + h[:, 0]
    elif kind == "xdeepfm":
        dnn, off = _mlp(X, mats, 0, F * K, fc, False)
        x0 = V                                    # [B,F,K]
        xl, H = V, F
        pooled = []
        for C in cin:
            Wl, bl, off = _lin(mats, off, F * H, C)
            W3 = Wl.view(C, F, H)
            # x^l[b,c,k] = relu( sum_{i,j} W[c,i,j] x0[b,i,k] x^{l-1}[b,j,k] + b_c )
            xl = torch.relu(torch.einsum("cij,bik,bjk->bck", W3, x0, xl) + bl[None, :, None])
            pooled.append(xl.sum(2))
            H = C
        j = torch.cat(pooled + [dnn], 1)
        wo = mats[off:off + j.shape[1]]
        out = out + j @ wo
    elif kind == "dcn":
        D = F * K
        ws = [mats[l * D:(l + 1) * D] for l in range(depth)]
        cs = mats[depth * D:depth * D + depth]
        x = X
        for l in range(depth):
            x = X * (x @ ws[l])[:, None] + x + cs[l]
        dnn, off = _mlp(X, mats, depth * D + depth, D, fc, False)
        j = torch.cat([x, dnn], 1)
        out = out + j @ mats[off:off + j.shape[1]]
    elif kind == "pnn":
        D, O, P = F * K, fc[0], F * (F - 1) // 2
        wz = mats[:D * O].view(O, D)
        wp = mats[D * O:D * O + P * O].view(O, P)
        c = mats[D * O + P * O]
        iu = torch.triu_indices(F, F, 1)
        G = V @ V.transpose(1, 2)                  # [B,F,F] Gram
        ip = G[:, iu[0], iu[1]]                    # lexicographic i<j
        h = torch.relu(X @ wz.t() + ip @ wp.t() + c)
        m, _ = _mlp(h, mats, D * O + P * O + 1, O, list(fc[1:]), True)
        out = out + m[:, 0]
    return out


def run(kind, B, F, K, index, weights, bias, embedding, mats, targets, fc=(), cin=(), depth=0):
    """fp64 forward+backward.  Returns dict(pred, loss, gw, gb, ge, gm) as numpy float64."""
    t64 = lambda a: None if a is None else torch.tensor(np.asarray(a, np.float64), requires_grad=True)
    w, b, e, m = t64(weights), t64(bias), t64(embedding), t64(mats)
    if e is None:
        e = torch.zeros(0, dtype=torch.float64, requires_grad=True)
    if m is None or m.numel() == 0:
        m = torch.zeros(1, dtype=torch.float64, requires_grad=True)
    idx = torch.tensor(np.asarray(index, np.int64))
    z = logits(kind, B, F, K, idx, w, b, e, m, list(fc), list(cin), depth)
    p = torch.sigmoid(z)
    t = torch.tensor((np.asarray(targets) > 0).astype(np.float64))
    loss = -(t * torch.log(p) + (1 - t) * torch.log(1 - p)).mean()
    loss.backward()
    g = lambda x: None if x.grad is None else x.grad.numpy().copy()
    return dict(pred=p.detach().numpy().copy(), loss=float(loss.detach()), gw=g(w), gb=g(b), ge=g(e), gm=g(m))
