"""torch-CPU autograd twin of the reference's training step.

TEST INFRASTRUCTURE ONLY.  The dense graph is written op-for-op like the reference
(``matrix_factorization.py:130-176``, ``loss_graphs.py``, ``embedding_graphs.py``) so that
``loss_vector.sum().backward()`` plays the role of ``tape.gradient(loss_fn, vars)``
(gradient of the SUM of a vector target [TF-sem]).  It validates the hand-derived
gradients of ``oracle/mf_oracle.py`` and is the "reference-faithful dense" CPU baseline
timed by ``bench.py --impl reference`` (TensorFlow itself is not installable here).
"""
from __future__ import annotations

import math

import numpy as np
import torch


def _t(x, dtype, grad=False):
    out = torch.as_tensor(np.asarray(x), dtype=dtype).clone()
    return out.requires_grad_(grad)


def _gather_matrix_indices(inp, idx):  # utils.py:94-105
    return torch.gather(inp, 1, idx)


def _embed(kind, X, p):
    if kind == "linear":  # embedding_graphs.py:38
        return X @ p["W"]
    if kind == "biased":  # :58
        return X @ p["W"] + p["b"]
    if kind == "relu":  # :85-87
        return torch.relu(X @ p["Wr"] + p["br"]) @ p["W"]
    raise ValueError(kind)


def _ndtr(z):
    return 0.5 * torch.erfc(-z / math.sqrt(2.0))


def loss_vector(loss, P, rows, cols, vals, random_ind, n_items, n_samples):
    if loss == "mse":  # loss_graphs.py:47-52
        return torch.square(vals - P[rows, cols])
    serial = P[rows, cols]  # matrix_factorization.py:154,160
    if loss == "wmrb":  # loss_graphs.py:74-88
        sample_pred = _gather_matrix_indices(P, random_ind)  # :153
        mask = vals > 0
        pos_rows = rows[mask]
        pos_pred = serial[mask]
        mapped = sample_pred[pos_rows]
        # torch.clamp's sub-gradient at 0 is 1 like TF's maximum(x, 0) (x >= 0 -> x) [TF-sem]
        summation = torch.clamp(1.0 - pos_pred[:, None] + mapped, min=0.0)
        return torch.log(1.0 + (n_items / n_samples) * summation.sum(dim=1))
    if loss == "kl":  # loss_graphs.py:111-122
        pos, neg = vals > 0, vals <= 0
        pp, pn = serial[pos], serial[neg]
        mp, vp = pp.mean(), pp.var(unbiased=False)
        mn, vn = pn.mean(), pn.var(unbiased=False)
        scale = torch.sqrt(vp + vn)
        return (1.0 - _ndtr((0.0 - (mn - mp)) / scale)).reshape(1)
    raise ValueError(loss)


def train_step(loss, Xu, Xi, kind_u, kind_i, params_u, params_i, rows, cols, vals,
               random_ind=None, n_items=None, n_samples=None, lr=1e-2, dtype=torch.float32, update=True):
    """One reference step with autograd.  numpy in, numpy out:
    ``(loss_vector, grads_u, grads_i, new_params_u, new_params_i)``."""
    Xu_t = _t(Xu.toarray() if hasattr(Xu, "toarray") else Xu, dtype)
    Xi_t = _t(Xi.toarray() if hasattr(Xi, "toarray") else Xi, dtype)
    pu = {k: _t(v, dtype, True) for k, v in params_u.items()}
    pi = {k: _t(v, dtype, True) for k, v in params_i.items()}
    rows_t = torch.as_tensor(np.asarray(rows), dtype=torch.int64)
    cols_t = torch.as_tensor(np.asarray(cols), dtype=torch.int64)
    vals_t = _t(vals, dtype)
    ri = None if random_ind is None else torch.as_tensor(np.asarray(random_ind), dtype=torch.int64)
    P = _embed(kind_u, Xu_t, pu) @ _embed(kind_i, Xi_t, pi).T  # :149
    lvec = loss_vector(loss, P, rows_t, cols_t, vals_t, ri, n_items, n_samples)
    lvec.sum().backward()
    gu = {k: v.grad.numpy().copy() for k, v in pu.items()}
    gi = {k: v.grad.numpy().copy() for k, v in pi.items()}
    out_l = lvec.detach().numpy().copy()
    if loss == "kl":
        out_l = out_l[0]
    if not update:
        return out_l, gu, gi, params_u, params_i
    # a brand-new Keras Adam every step (:176) => step t=1 from zero moments.  NOT torch.optim.Adam:
    # torch adds eps to sqrt(v)/sqrt(1-b2), Keras/TF adds it to sqrt(v) with the bias correction folded
    # into the step size [TF-sem], so the effective epsilon differs (1e-7 vs 3.16e-6).
    b1, b2, eps = 0.9, 0.999, 1e-7
    nu, ni = {}, {}
    for src, dst in ((pu, nu), (pi, ni)):
        for k, v in src.items():
            g = v.grad
            one = torch.ones((), dtype=dtype)
            m = g * (one - torch.tensor(b1, dtype=dtype))
            vv = (g * g) * (one - torch.tensor(b2, dtype=dtype))
            alpha = torch.tensor(lr, dtype=dtype) * torch.sqrt(one - torch.tensor(b2, dtype=dtype)) / (
                one - torch.tensor(b1, dtype=dtype))
            dst[k] = (v.detach() - alpha * m / (torch.sqrt(vv) + torch.tensor(eps, dtype=dtype))).numpy().copy()
    return out_l, gu, gi, nu, ni
