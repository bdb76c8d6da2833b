"""CPU tests pinning the oracle: reference known-answer vector, the independent
scalar-loop golden vectors, the torch-autograd twin, and TF-semantics tables."""
import json
import os

import numpy as np
import pytest
from scipy import sparse

from oracle import autograd_twin as tw
from oracle import mf_oracle as o

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


def _inter(case):
    it = case["inter"]
    rows = np.array([x[0] for x in it], dtype=np.int64)
    cols = np.array([x[1] for x in it], dtype=np.int64)
    vals = np.array([x[2] for x in it], dtype=np.float32)
    return rows, cols, vals


def test_gather_matrix_indices_reference_vector():
    g = G["gather_matrix_indices"]  # reference test/test_utils.py:47-61
    out = o.gather_matrix_indices(np.array(g["input"], np.float32), np.array(g["index"], np.int64))
    assert np.array_equal(out, np.array(g["expected"], np.float32))


@pytest.mark.parametrize("name,loss", [("wmrb_3x4", "wmrb"), ("mse_2x2", "mse"), ("kl_2p2n", "kl"),
                                       ("rand_mse", "mse"), ("rand_wmrb", "wmrb"), ("rand_kl", "kl")])
@pytest.mark.parametrize("dense", [True, False])
def test_step_against_scalar_golden(name, loss, dense):
    c = G[name]
    rows, cols, vals = _inter(c)
    U, V = np.array(c["U"], np.float64), np.array(c["V"], np.float64)
    samp = np.array(c["samp"], np.int64) if "samp" in c else None
    step = o.train_step_dense if dense else o.train_step_sparse
    Xu, Xi = np.eye(U.shape[0]), np.eye(V.shape[0])
    lvec, gu, gi, _, _ = step(loss, Xu, Xi, "linear", "linear", {"W": U}, {"W": V}, rows, cols,
                              vals.astype(np.float64), samp, c.get("n_items"), c.get("n_samples"), update=False)
    np.testing.assert_allclose(lvec, np.array(c["loss"]), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(gu["W"], np.array(c["dU"]), rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(gi["W"], np.array(c["dV"]), rtol=1e-11, atol=1e-13)
    # fp32 twin within the 1e-5 relative bar of north_star
    l32, gu32, gi32, _, _ = step(loss, Xu.astype(np.float32), Xi.astype(np.float32), "linear", "linear",
                                 {"W": U.astype(np.float32)}, {"W": V.astype(np.float32)}, rows, cols, vals,
                                 samp, c.get("n_items"), c.get("n_samples"), update=False)
    np.testing.assert_allclose(l32, np.array(c["loss"]), rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(gu32["W"], np.array(c["dU"]), rtol=2e-5, atol=2e-6)


def test_wmrb_exact_zero_hinge_takes_gradient():
    # golden case: user 1's positive (1,0) has h == 0 against sample item 1 -> indicator is 1 (>=)
    c = G["wmrb_3x4"]
    rows, cols, vals = _inter(c)
    U, V = np.array(c["U"], np.float32), np.array(c["V"], np.float32)
    P = U @ V.T
    ss = o.gather_matrix_indices(P, np.array(c["samp"]))
    _, cc, Gm = o.wmrb_coefs(rows, vals, P[rows, cols], ss, 4, 2)
    assert 1.0 - P[1, 0] + P[1, 1] == 0.0
    assert Gm[1, 0] > 0  # the zero hinge contributed
    assert cc[3] == 0.0  # negative-valued interaction ignored


@pytest.mark.parametrize("loss", ["mse", "wmrb", "kl"])
@pytest.mark.parametrize("kinds", [("linear", "linear"), ("biased", "linear"), ("relu", "biased"), ("relu", "relu")])
def test_analytic_gradients_match_autograd_twin(loss, kinds):
    rng = np.random.default_rng(7)
    n_u, n_i, r, S, Fu, Fi = 13, 17, 4, 5, 6, 17
    Xu = sparse.random(n_u, Fu, density=0.5, random_state=1, format="csr", dtype=np.float64)
    Xi = np.eye(n_i)
    A = sparse.random(n_u, n_i, density=0.3, random_state=2, format="coo")
    rows, cols = A.row.astype(np.int64), A.col.astype(np.int64)
    order = np.lexsort((cols, rows))
    rows, cols = rows[order], cols[order]
    vals = rng.choice([-2.0, -1.0, 1.0, 3.0], size=rows.size)
    samp = np.stack([rng.choice(n_i, S, replace=False) for _ in range(n_u)])

    def params(kind, F):
        if kind == "relu":
            return {"W": rng.standard_normal((5 * r, r)) * 0.3, "Wr": rng.standard_normal((F, 5 * r)),
                    "br": rng.standard_normal((1, 5 * r)) * 0.1}
        p = {"W": rng.standard_normal((F, r)) * 0.5}
        if kind == "biased":
            p["b"] = rng.standard_normal((1, r)) * 0.1
        return p

    pu, pi = params(kinds[0], Fu), params(kinds[1], Fi)
    import torch
    ref = tw.train_step(loss, Xu, Xi, kinds[0], kinds[1], pu, pi, rows, cols, vals, samp, n_i, S, lr=0.1,
                        dtype=torch.float64)
    for step in (o.train_step_dense, o.train_step_sparse):
        got = step(loss, Xu, Xi, kinds[0], kinds[1], pu, pi, rows, cols, vals, samp, n_i, S, lr=0.1)
        np.testing.assert_allclose(got[0], ref[0], rtol=1e-10, atol=1e-12)
        for a, b in ((got[1], ref[1]), (got[2], ref[2]), (got[3], ref[3]), (got[4], ref[4])):
            assert a.keys() == b.keys()
            for k in a:
                np.testing.assert_allclose(a[k], b[k], rtol=1e-8, atol=1e-10, err_msg=k)


def test_adam_step1_table():
    c = G["adam1"]
    w = np.full(len(c["g"]), c["w"], np.float64)
    out = o.adam_step1(w, np.array(c["g"], np.float64), c["lr"])
    np.testing.assert_allclose(out, np.array(c["expected"]), rtol=1e-13)
    out32 = o.adam_step1(w.astype(np.float32), np.array(c["g"], np.float32), c["lr"])
    np.testing.assert_allclose(out32, np.array(c["expected"]), rtol=1e-6)
    assert out32[0] == np.float32(c["w"])  # g == 0 -> no update, exactly
    # closed form quoted in SURVEY A.6: lr*g/(|g| + eps/sqrt(1-b2))
    g = np.array(c["g"])
    np.testing.assert_allclose(out, c["w"] - c["lr"] * g / (np.abs(g) + 1e-7 / np.sqrt(1 - 0.999)), rtol=1e-12)


def test_topk_tie_order_and_grid_golden():
    assert o.topk_stable(np.array([1, 3, 3, 0, 3], np.float32), 3).tolist() == [1, 2, 4]
    c = G["grid_topk"]
    U = np.array(c["U_int"], np.float32) / c["scale"]
    V = np.array(c["V_int"], np.float32) / c["scale"]
    P = o.canonical_scores(U, V)
    assert np.array_equal(P, (U @ V.T).astype(np.float32))  # exact grid: any order gives the same bits
    assert o.retrieve_user_recs(P, k=c["k"]).tolist() == c["raw"]
    Ppos = np.where(P > 0, P, np.float32(0))
    assert o.topk_stable(Ppos, c["k"]).tolist() == c["clamped"]


def test_metrics_golden():
    c = G["metrics_4x5"]
    P, A, k = np.array(c["P"], np.float32), np.array(c["A"], np.float32), c["k"]
    assert o._topk_clamped(P, k).tolist() == c["topk_clamped"]
    np.testing.assert_array_equal(o.recall_at_k(P, A, k), np.array(c["recall_drop"], np.float32))
    keep = [np.inf if x == "inf" else x for x in c["recall_keep"]]
    np.testing.assert_array_equal(o.recall_at_k(P, A, k, preserve_rows=True), np.array(keep, np.float32))
    np.testing.assert_array_equal(o.precision_at_k(P, A, k), np.array(c["precision_drop"], np.float32))
    np.testing.assert_array_equal(o.precision_at_k(P, A, k, True), np.array(c["precision_keep"], np.float32))
    prec, rec = np.float32(0.75), np.float32(1.0)
    assert o.f1_at_k(P, A, k) == np.float32((2 * prec * rec) / (prec + rec))


def test_ndcg_hand_case():
    P = np.array([[0.3, 0.9, 0.1], [0.5, 0.5, 0.2]], np.float32)
    A = np.array([[2.0, 0.0, 1.0], [0.0, 0.0, 0.0]], np.float32)
    # row 0 ranking [1,0,2] -> gains [0,3,1]; dcg@2 = 0/log2(2) + 3/log2(3)
    np.testing.assert_allclose(o.dcg_at_k(P, A, 2), [3 / np.log2(3), 0.0], rtol=1e-6)
    np.testing.assert_allclose(o.idcg_at_k(P, A, 2), [3 + 1 / np.log2(3), 0.0], rtol=1e-6)
    out = o.ndcg_at_k(P, A, 2)
    assert out.shape == (1,)
    keep = o.ndcg_at_k(P, A, 2, preserve_rows=True)
    assert keep[1] == 0.0


def test_initializer_global_norm():
    W = o.normal_initializer(50, 8, np.random.default_rng(0))
    assert abs(np.linalg.norm(W.astype(np.float64)) - 1.0) < 1e-6
    assert np.all(o.uniform_initializer(10, 3, np.random.default_rng(0)) >= 0)


def test_generate_random_interaction_shape_consistency():
    (rows, cols, vals), A = o.generate_random_interaction(50, 100, density=0.05, random_state=3)
    assert A.shape == (50, 100)
    assert np.array_equal(A[rows, cols], vals) and np.count_nonzero(A) == vals.size
    assert np.all(np.diff(rows * 100 + cols) > 0)  # row-major sorted


def test_dense_and_sparse_fit_agree_short_trajectory():
    rng = np.random.default_rng(3)
    (rows, cols, vals), _ = o.generate_random_interaction(30, 40, density=0.1, random_state=5)
    U0 = o.uniform_initializer(30, 4, rng).astype(np.float64)
    V0 = o.uniform_initializer(40, 4, rng).astype(np.float64)
    samp = np.stack([rng.choice(40, 8, replace=False) for _ in range(30)])
    a = o.fit(3, "wmrb", np.eye(30), np.eye(40), "linear", "linear", {"W": U0}, {"W": V0}, rows, cols,
              vals.astype(np.float64), samp, 40, 8, lr=0.1, dense=True)
    b = o.fit(3, "wmrb", np.eye(30), np.eye(40), "linear", "linear", {"W": U0}, {"W": V0}, rows, cols,
              vals.astype(np.float64), samp, 40, 8, lr=0.1, dense=False)
    np.testing.assert_allclose(a[4], b[4], rtol=1e-9)


def test_stateful_adam_matches_the_keras_formula_in_the_shim():
    """oracle.adam_step (extension oracle) against the Keras-Adam stand-in used to run the reference source."""
    import os
    import sys
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden", "tf_shim"))
    try:
        import tensorflow as tf  # the shim
        assert tf.__version__.endswith("shim")
        rng = np.random.default_rng(5)
        w0 = rng.standard_normal((7, 3)).astype(np.float32)
        var = tf.Variable(w0)
        opt = tf.keras.optimizers.Adam(learning_rate=0.05)
        w, m, v = w0.copy(), np.zeros_like(w0), np.zeros_like(w0)
        for t in range(1, 5):
            g = rng.standard_normal((7, 3)).astype(np.float32)
            opt.apply_gradients([(torch.from_numpy(g), var)])
            w, m, v = o.adam_step(w, g, m, v, t, 0.05)
            np.testing.assert_allclose(var.detach().numpy(), w, rtol=1e-5, atol=1e-6)
            if t == 1:  # the first step of a stateful Adam is the reference's per-epoch fresh step
                np.testing.assert_allclose(w, o.adam_step1(w0, g, 0.05), rtol=2e-6, atol=2e-7)
    finally:
        sys.path.pop(0)
        for mod in ("tensorflow", "tensorflow_probability"):
            sys.modules.pop(mod, None)
