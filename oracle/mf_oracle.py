"""NumPy restatement of the reference matrix-factorization hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  All ``file:line``
citations are relative to ``/root/reference/``.  Statements tagged [TF-sem] come
from TensorFlow's documented behaviour (TF is not installed here): parity for
those is *unpinned* except where noted.

Conventions
  * interactions are COO triples ``(rows int64[nnz], cols int64[nnz], vals f32[nnz])``
    in the stored order of ``tf_interactions.indices`` / ``.values``
    (``loss_graphs.py:47-52``),
  * features ``X`` are dense ndarrays or scipy CSR (the reference only accepts
    dense, ``embedding_graphs.py:38``; the sparse form is the same product),
  * every function takes ``dtype`` (float32 = parity twin, float64 = error budget).
"""
from __future__ import annotations

import numpy as np
from scipy import sparse, special

F32 = np.float32

# --------------------------------------------------------------------------
# mf/utils.py
# --------------------------------------------------------------------------


def gather_matrix_indices(input_arr, index_arr):
    """``out[r, c] = input[r, index[r, c]]`` -- utils.py:94-105 (== torch.gather dim=1).

    PINNED by the reference's known-answer test ``test/test_utils.py:47-61``.
    """
    input_arr = np.asarray(input_arr)
    index_arr = np.asarray(index_arr)
    return np.take_along_axis(input_arr, index_arr.astype(np.int64), axis=1)


def random_sampler(n_items, n_users, n_samples, replace=False):
    """Per-user item samples from numpy's *global* RNG -- utils.py:20-22."""
    items_per_user = [np.random.choice(a=n_items, size=n_samples, replace=replace) for _ in range(n_users)]
    return np.array(items_per_user, dtype=np.int64).reshape(n_users, n_samples)


def generate_random_interaction(n_users, n_items, min_val=0.0, max_val=5.0, density=0.50, random_state=None):
    """utils.py:37-59.  Returns ``(rows, cols, vals), A_dense`` (row-major sorted, zeros dropped)."""
    p = sparse.random(n_users, n_items, density=density, random_state=random_state)
    p = (max_val - min_val) * p + min_val * p.ceil()
    random_arr = np.round(p.toarray())
    csr = sparse.csr_matrix(random_arr)
    rows, cols = csr.nonzero()
    vals = np.asarray(csr[rows, cols]).ravel().astype(F32) if rows.size else np.zeros(0, F32)
    return (rows.astype(np.int64), cols.astype(np.int64), vals), random_arr.astype(F32)


# --------------------------------------------------------------------------
# mf/initializer_graphs.py
# --------------------------------------------------------------------------


def l2_normalize_global(Z, eps=1e-12):
    """``tf.math.l2_normalize(Z)`` with axis=None -- initializer_graphs.py:34,51.

    [TF-sem] ``x * rsqrt(maximum(reduce_sum(x*x), eps))`` over the WHOLE tensor.
    """
    Z = np.asarray(Z)
    t = Z.dtype.type
    ss = np.sum(Z.astype(np.float64) ** 2)
    return (Z * t(1.0 / np.sqrt(max(ss, eps)))).astype(Z.dtype)


def normal_initializer(n_features, n_components, rng):
    """initializer_graphs.py:27-35 (RNG stream is numpy's, not TF's)."""
    return l2_normalize_global(rng.standard_normal((n_features, n_components)).astype(F32))


def uniform_initializer(n_features, n_components, rng):
    """initializer_graphs.py:43-52."""
    return l2_normalize_global(rng.random((n_features, n_components)).astype(F32))


# --------------------------------------------------------------------------
# mf/embedding_graphs.py  (forward + hand-derived backward)
# --------------------------------------------------------------------------

LINEAR, BIASED, RELU = "linear", "biased", "relu"


def _mm(X, W):
    out = X @ W
    return np.asarray(out)


def embed_forward(kind, X, params):
    """``get_repr`` of Linear (:38), BiasedLinear (:58), ReLU (:85-87) embeddings.

    ``params``: dict with ``W`` and, per kind, ``b`` ([1,r]) or ``Wr`` ([F,5r]), ``br`` ([1,5r]).
    Returns ``(E, cache)``.
    """
    W = params["W"]
    if kind == LINEAR:
        return _mm(X, W), {}
    if kind == BIASED:
        return _mm(X, W) + params["b"], {}
    if kind == RELU:
        Z = _mm(X, params["Wr"]) + params["br"]
        H = np.maximum(Z, Z.dtype.type(0))
        return _mm(H, W), {"Z": Z, "H": H}
    raise ValueError(kind)


def embed_backward(kind, X, params, cache, dE):
    """Gradients of ``sum(loss)`` w.r.t. the trainables returned by ``get_repr``."""
    XT = X.T
    if kind == LINEAR:
        return {"W": _mm(XT, dE)}
    if kind == BIASED:
        return {"W": _mm(XT, dE), "b": dE.sum(axis=0, keepdims=True)}
    if kind == RELU:
        H, Z = cache["H"], cache["Z"]
        dW = H.T @ dE
        dH = dE @ params["W"].T
        dZ = dH * (Z > 0)  # tf.nn.relu gradient is strict (> 0) [TF-sem]
        return {"W": dW, "Wr": _mm(XT, dZ), "br": dZ.sum(axis=0, keepdims=True)}
    raise ValueError(kind)


# --------------------------------------------------------------------------
# mf/predict_graphs.py and scoring
# --------------------------------------------------------------------------


def dot_product_prediction(user_embedding, item_embedding):
    """predict_graphs.py:35 / matrix_factorization.py:149,195 -- ``U @ V^T`` (BLAS order)."""
    return user_embedding @ item_embedding.T


def canonical_scores(U, V, block=2048):
    """Canonical fp32 score used for bit-exact top-k.

    ``score(u,i) = fp32( sum_{c=0}^{r-1} fma_fp64(U[u,c], V[i,c]) )`` in index order.
    The product of two fp32 values is exact in fp64, so mul-then-add in fp64 equals
    the fused chain bit for bit.  fp32 BLAS sums are blocking-dependent, so this is
    the order-independent definition both the oracle and the CUDA rerank implement.
    """
    U64 = np.asarray(U, dtype=np.float64)
    V64 = np.asarray(V, dtype=np.float64)
    n_u, r = U64.shape
    n_i = V64.shape[0]
    out = np.empty((n_u, n_i), dtype=F32)
    for a in range(0, n_u, block):
        acc = np.zeros((min(block, n_u - a), n_i), dtype=np.float64)
        for c in range(r):
            acc += U64[a:a + block, c:c + 1] * V64[None, :, c]
        out[a:a + block] = acc.astype(F32)
    return out


def canonical_pair_scores(U, V, rows, cols):
    """Canonical score for explicit (row, col) pairs."""
    U64 = np.asarray(U, dtype=np.float64)[rows]
    V64 = np.asarray(V, dtype=np.float64)[cols]
    acc = np.zeros(U64.shape[0], dtype=np.float64)
    for c in range(U64.shape[1]):
        acc += U64[:, c] * V64[:, c]
    return acc.astype(F32)


def topk_stable(P, k):
    """``tf.math.top_k(P, k).indices`` -- descending, ties -> lower index first [TF-sem].

    Stable argsort of the negated scores gives exactly that order.
    """
    P = np.asarray(P)
    if P.ndim == 1:
        return np.argsort(-P, kind="stable")[:k].astype(np.int32)
    return np.argsort(-P, axis=1, kind="stable")[:, :k].astype(np.int32)


# --------------------------------------------------------------------------
# mf/loss_graphs.py  (forward at the reference's get_loss boundary)
# --------------------------------------------------------------------------


def mse_loss(rows, cols, vals, predictions):
    """loss_graphs.py:47-52: ``square(values - gather_nd(predictions, indices))``."""
    p = predictions[rows, cols]
    return (vals.astype(p.dtype) - p) ** 2


def wmrb_loss(rows, vals, sample_predictions, prediction_serial, n_items, n_samples):
    """loss_graphs.py:74-88.  Output length = #positives, stored order."""
    pos = vals > 0
    prow = rows[pos]
    ppred = prediction_serial[pos]
    t = ppred.dtype.type
    mapped = sample_predictions[prow]  # [P, S]
    summation = np.maximum(t(1.0) - ppred[:, None] + mapped, t(0.0))
    margin_rank = t(n_items / n_samples) * summation.sum(axis=1)
    return np.log(t(1.0) + margin_rank)


def kl_loss(prediction_serial, vals):
    """loss_graphs.py:111-122 -- scalar ``1 - Normal(mu_neg - mu_pos, sqrt(var_pos+var_neg)).cdf(0)``.

    ``tf.nn.moments`` = population variance [TF-sem]; TFP ``Normal.cdf`` = ndtr [TF-sem].
    """
    pos, neg = vals > 0, vals <= 0
    pp, pn = prediction_serial[pos], prediction_serial[neg]
    t = prediction_serial.dtype.type
    with np.errstate(all="ignore"):
        mp, vp = pp.mean(dtype=t), pp.var(dtype=t)
        mn, vn = pn.mean(dtype=t), pn.var(dtype=t)
        scale = np.sqrt(vp + vn)
        z = (t(0.0) - (mn - mp)) / scale
        return t(1.0) - t(special.ndtr(np.float64(z)))


# --------------------------------------------------------------------------
# d(sum loss)/d(prediction): SURVEY App. A.2-A.4 (grad of the SUM of the loss
# vector, matrix_factorization.py:170-171 [TF-sem])
# --------------------------------------------------------------------------


def mse_coef(vals, p):
    """``c_k = -2 (a_k - p_k)`` (A.2)."""
    return p.dtype.type(-2.0) * (vals.astype(p.dtype) - p)


def wmrb_coefs(rows, vals, p, sample_scores, n_items, n_samples, chunk=1 << 15):
    """A.3.  Returns ``(loss[P], c[nnz], G[n_u, S])``.

    ``h_kj = (1 - p_k) + s_{u_k, j}``; indicator ``h >= 0`` ([TF-sem]: ``maximum(x, 0)``
    routes the gradient to ``x`` when ``x >= 0``).
    """
    t = p.dtype.type
    scale = t(n_items / n_samples)
    pos_idx = np.nonzero(vals > 0)[0]
    c = np.zeros_like(p)
    G = np.zeros_like(sample_scores)
    loss = np.empty(pos_idx.size, dtype=p.dtype)
    for a in range(0, pos_idx.size, chunk):
        kk = pos_idx[a:a + chunk]
        u = rows[kk]
        h = (t(1.0) - p[kk])[:, None] + sample_scores[u]
        m = scale * np.maximum(h, t(0.0)).sum(axis=1)
        loss[a:a + chunk] = np.log(t(1.0) + m)
        w = scale / (t(1.0) + m)
        ind = h >= 0
        c[kk] = -w * ind.sum(axis=1).astype(p.dtype)
        np.add.at(G, u, w[:, None] * ind)
    return loss, c, G


def kl_coef(vals, p):
    """A.4 closed-form gradient of the scalar KL loss w.r.t. each prediction."""
    t = p.dtype.type
    pos, neg = vals > 0, vals <= 0
    npos, nneg = int(pos.sum()), int(neg.sum())
    with np.errstate(all="ignore"):
        mp, vp = p[pos].mean(dtype=t), p[pos].var(dtype=t)
        mn, vn = p[neg].mean(dtype=t), p[neg].var(dtype=t)
        s = np.sqrt(vp + vn)
        z = (mp - mn) / s
        phi = t(np.exp(-0.5 * np.float64(z) ** 2) / np.sqrt(2.0 * np.pi))
        c = np.zeros_like(p)
        c[pos] = -phi / (s * t(npos)) * (t(1.0) - z * (p[pos] - mp) / s)
        c[neg] = -phi / (s * t(nneg)) * (t(-1.0) - z * (p[neg] - mn) / s)
    return c


# --------------------------------------------------------------------------
# optimizer: a brand-new Adam every step (matrix_factorization.py:176)
# --------------------------------------------------------------------------

ADAM_BETA1, ADAM_BETA2, ADAM_EPS = 0.9, 0.999, 1e-7


def adam_step1(w, g, lr):
    """Adam step t=1 from zero moments (A.6) [TF-sem], evaluated in ``w.dtype``.

    m = (1-b1) g ; v = (1-b2) g^2 ; alpha = lr sqrt(1-b2)/(1-b1) ; w -= alpha m / (sqrt(v)+eps)
    """
    t = w.dtype.type
    g = g.astype(w.dtype)
    m = g * (t(1.0) - t(ADAM_BETA1))
    v = (g * g) * (t(1.0) - t(ADAM_BETA2))
    alpha = t(lr) * np.sqrt(t(1.0) - t(ADAM_BETA2)) / (t(1.0) - t(ADAM_BETA1))
    return w - alpha * m / (np.sqrt(v) + t(ADAM_EPS))


# --------------------------------------------------------------------------
# one training step (matrix_factorization.py:128-176)
# --------------------------------------------------------------------------

MSE, WMRB, KL = "mse", "wmrb", "kl"


def adam_step(w, g, m, v, t, lr):
    """Keras Adam with PERSISTENT moments (beta_1 0.9, beta_2 0.999, epsilon 1e-7) [TF-sem]: what one long-lived
    ``tf.keras.optimizers.Adam`` would do.  NOT reference behaviour (the reference builds a new optimizer every epoch,
    matrix_factorization.py:176 -> ``adam_step1``); oracle of the ``fit(optimizer="adam")`` extension.
    Returns ``(w, m, v)`` after step ``t`` (>= 1)."""
    dt = w.dtype.type
    b1, b2, eps = dt(0.9), dt(0.999), dt(1e-7)
    m = b1 * m + g * (dt(1) - b1)
    v = b2 * v + (g * g) * (dt(1) - b2)
    alpha = dt(lr) * np.sqrt(dt(1) - b2 ** dt(t)) / (dt(1) - b1 ** dt(t))
    return (w - alpha * m / (np.sqrt(v) + eps)).astype(w.dtype), m, v


def _loss_and_coefs(loss, rows, cols, vals, p, sample_scores, n_items, n_samples):
    if loss == MSE:
        return (vals.astype(p.dtype) - p) ** 2, mse_coef(vals, p), None
    if loss == WMRB:
        return wmrb_coefs(rows, vals, p, sample_scores, n_items, n_samples)
    if loss == KL:
        return kl_loss(p, vals), kl_coef(vals, p), None
    raise ValueError(loss)


def train_step_dense(loss, Xu, Xi, kind_u, kind_i, params_u, params_i, rows, cols, vals,
                     random_ind=None, n_items=None, n_samples=None, lr=1e-2, update=True):
    """Reference-faithful DENSE step: full ``U V^T`` (:149), gathers (:153-154,:160), loss
    (:165-167), gradient of the summed loss routed through a dense ``dP`` (what the TF tape
    does for ``gather_nd``: ``scatter_nd`` into zeros), embedding backward, Adam step 1 (:176).

    Returns ``(loss_vector, grads_u, grads_i, new_params_u, new_params_i)``.
    """
    Eu, cu = embed_forward(kind_u, Xu, params_u)
    Ei, ci = embed_forward(kind_i, Xi, params_i)
    P = dot_product_prediction(Eu, Ei)
    p = P[rows, cols]
    ss = gather_matrix_indices(P, random_ind) if loss == WMRB else None
    lvec, c, G = _loss_and_coefs(loss, rows, cols, vals, p, ss, n_items, n_samples)
    dP = np.zeros_like(P)
    np.add.at(dP, (rows, cols), c)
    if G is not None:
        ur = np.repeat(np.arange(P.shape[0]), random_ind.shape[1])
        np.add.at(dP, (ur, random_ind.ravel()), G.ravel())
    dEu = dP @ Ei
    dEi = dP.T @ Eu
    gu = embed_backward(kind_u, Xu, params_u, cu, dEu)
    gi = embed_backward(kind_i, Xi, params_i, ci, dEi)
    if not update:
        return lvec, gu, gi, params_u, params_i
    nu = {k: adam_step1(v, gu[k], lr) for k, v in params_u.items()}
    ni = {k: adam_step1(v, gi[k], lr) for k, v in params_i.items()}
    return lvec, gu, gi, nu, ni


def _segment_rows(idx, coef, src, n_out):
    """``out[idx[e]] += coef[e] * src_row[e]`` (deterministic in e order per numpy add.at)."""
    out = np.zeros((n_out, src.shape[1]), dtype=src.dtype)
    np.add.at(out, idx, coef[:, None] * src)
    return out


def train_step_sparse(loss, Xu, Xi, kind_u, kind_i, params_u, params_i, rows, cols, vals,
                      random_ind=None, n_items=None, n_samples=None, lr=1e-2, update=True):
    """Sparse restatement of the same step: only the needed dot products are formed
    (``nnz + n_u*S`` of them), gradients are row segment-sums.  Same math as
    ``train_step_dense`` up to fp32 summation order.
    """
    Eu, cu = embed_forward(kind_u, Xu, params_u)
    Ei, ci = embed_forward(kind_i, Xi, params_i)
    n_u, n_i = Eu.shape[0], Ei.shape[0]
    p = np.einsum("kc,kc->k", Eu[rows], Ei[cols]).astype(Eu.dtype)
    ss = None
    if loss == WMRB:
        S = random_ind.shape[1]
        ss = np.einsum("uc,usc->us", Eu, Ei[random_ind.ravel()].reshape(n_u, S, -1)).astype(Eu.dtype)
    lvec, c, G = _loss_and_coefs(loss, rows, cols, vals, p, ss, n_items, n_samples)
    dEu = _segment_rows(rows, c, Ei[cols], n_u)
    dEi = _segment_rows(cols, c, Eu[rows], n_i)
    if G is not None:
        S = random_ind.shape[1]
        ur = np.repeat(np.arange(n_u), S)
        flat = random_ind.ravel()
        dEu += _segment_rows(ur, G.ravel(), Ei[flat], n_u)
        dEi += _segment_rows(flat, G.ravel(), Eu[ur], n_i)
    gu = embed_backward(kind_u, Xu, params_u, cu, dEu)
    gi = embed_backward(kind_i, Xi, params_i, ci, dEi)
    if not update:
        return lvec, gu, gi, params_u, params_i
    nu = {k: adam_step1(v, gu[k], lr) for k, v in params_u.items()}
    ni = {k: adam_step1(v, gi[k], lr) for k, v in params_i.items()}
    return lvec, gu, gi, nu, ni


def fit(epochs, loss, Xu, Xi, kind_u, kind_i, params_u, params_i, rows, cols, vals,
        random_ind=None, n_items=None, n_samples=None, lr=1e-2, dense=True):
    """The epoch loop of ``MatrixFactorization.fit`` (:128-187) from given initial params.

    Returns ``(params_u, params_i, Eu, Ei, [mean loss per epoch])``.
    """
    step = train_step_dense if dense else train_step_sparse
    hist = []
    for _ in range(epochs):
        lvec, _, _, params_u, params_i = step(loss, Xu, Xi, kind_u, kind_i, params_u, params_i, rows, cols,
                                              vals, random_ind, n_items, n_samples, lr)
        hist.append(float(np.mean(lvec)))  # reduce_mean(loss_fn), :179
    Eu, _ = embed_forward(kind_u, Xu, params_u)
    Ei, _ = embed_forward(kind_i, Xi, params_i)
    return params_u, params_i, Eu, Ei, hist


# --------------------------------------------------------------------------
# prediction / evaluation methods of MatrixFactorization
# --------------------------------------------------------------------------


def predict(Eu, Ei, A=None, canonical=True):
    """matrix_factorization.py:189-201."""
    P = canonical_scores(Eu, Ei) if canonical else dot_product_prediction(Eu, Ei)
    if A is not None:
        return P, P[np.asarray(A) == 0]  # gather_nd(P, where(A == 0)) -- row-major order
    return P


def predict_ranks(Eu, Ei, A, canonical=True):
    """:203-216 -- descending argsort of the unobserved predictions (ties -> lower index)."""
    _, tf_predictions = predict(Eu, Ei, A, canonical)
    return topk_stable(tf_predictions, tf_predictions.shape[0])


def _topk_clamped(P, k):
    Ppos = np.where(P > 0, P, F32(0.0))  # :237
    return topk_stable(Ppos, k).astype(np.int64)  # :245


def _hits_relevant(P, A, k):
    A = np.asarray(A)
    top = _topk_clamped(P, k)
    res = gather_matrix_indices(A, top)  # :248 gathers from A (any sign)
    relevant = np.count_nonzero(np.where(A > 0, A, 0), axis=1).astype(F32)  # :251
    hits = np.count_nonzero(res, axis=1).astype(F32)  # :254
    return hits, relevant


def recall_at_k(P, A, k=10, preserve_rows=False):
    """:218-269."""
    hits, relevant = _hits_relevant(P, A, k)
    if not preserve_rows:
        m = relevant != 0
        return hits[m] / relevant[m]
    with np.errstate(all="ignore"):
        rec = hits / relevant
    return np.where(np.isnan(rec), F32(0.0), rec)


def precision_at_k(P, A, k=10, preserve_rows=False):
    """:271-304."""
    hits, relevant = _hits_relevant(P, A, k)
    if not preserve_rows:
        return hits[relevant != 0] / F32(k)
    return hits / F32(k)


def f1_at_k(P, A, k=10, beta=1.0):
    """:306-318 (formula kept as written: beta^2 multiplies the whole sum)."""
    prec = np.mean(precision_at_k(P, A, k), dtype=F32)
    rec = np.mean(recall_at_k(P, A, k), dtype=F32)
    with np.errstate(all="ignore"):
        return F32(((1 + beta ** 2) * prec * rec) / (beta ** 2 * (prec + rec)))


def _discount(n):
    return (np.log1p(np.arange(1, n + 1, dtype=F32)) / np.log(F32(2.0))).astype(F32)  # :342-346


def dcg_at_k(P, A, k=10):
    """:320-351 -- full ranking of RAW scores, gains ``2^A - 1``."""
    A = np.asarray(A, dtype=F32)
    n_i = P.shape[1]
    ranks = topk_stable(P, n_i).astype(np.int64)
    num = np.power(F32(2.0), gather_matrix_indices(A, ranks)) - F32(1.0)
    return (num / _discount(n_i)[None, :])[:, :k].sum(axis=1, dtype=F32)


def idcg_at_k(P, A, k=10):
    """:353-384 -- same gains, sorted descending (:373)."""
    A = np.asarray(A, dtype=F32)
    n_i = P.shape[1]
    num = np.power(F32(2.0), A) - F32(1.0)
    ideal = -np.sort(-num, axis=1, kind="stable")
    return (ideal / _discount(n_i)[None, :])[:, :k].sum(axis=1, dtype=F32)


def ndcg_at_k(P, A, k=10, preserve_rows=False):
    """:386-413."""
    with np.errstate(all="ignore"):
        ndcg = dcg_at_k(P, A, k) / idcg_at_k(P, A, k)
    if not preserve_rows:
        return ndcg[np.count_nonzero(np.asarray(A), axis=1) > 0]
    return np.where(~np.isnan(ndcg), ndcg, F32(0.0))


def retrieve_user_recs(P, user=None, k=None):
    """:416-438 -- top-k on RAW scores, int32."""
    n_i = P.shape[1]
    if user is None:
        return topk_stable(P, k if k is not None else n_i)
    return topk_stable(P[user], k if k is not None else n_i)
