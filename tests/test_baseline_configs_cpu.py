"""CPU tests at the exact shapes of BASELINE.json configs[0] (C1) and configs[1] (C2): the oracle's sparse restatement
(what the GPU parity tests compare against at these shapes, ``test_gpu_baseline_configs.py``) agrees with its
reference-faithful DENSE statement (dense identity features, full ``U V^T``, gathers -- ``matrix_factorization.py:130-176``)
and with the torch-autograd twin of that dense graph, on the configurations' own synthetic inputs."""
import numpy as np
import pytest
import torch

import _baseline_configs as cfg
from oracle import autograd_twin as tw
from oracle import mf_oracle as o


@pytest.mark.parametrize("name", ["c1", "c2"])
def test_sparse_oracle_equals_dense_reference_statement(name):
    c = getattr(cfg, name)()
    n_u, n_i, S, loss, lr = c["n_u"], c["n_i"], c["S"] or None, c["loss"], c["lr"]
    rows, cols, samp = c["rows"], c["cols"], c["samp"]
    vals = c["vals"].astype(np.float64)
    pu, pi = {"W": c["U0"].astype(np.float64)}, {"W": c["V0"].astype(np.float64)}
    Xu, Xi = np.eye(n_u), np.eye(n_i)
    sp = o.train_step_sparse(loss, Xu, Xi, "linear", "linear", pu, pi, rows, cols, vals, samp, n_i, S, lr=lr)
    de = o.train_step_dense(loss, Xu, Xi, "linear", "linear", pu, pi, rows, cols, vals, samp, n_i, S, lr=lr)
    np.testing.assert_allclose(sp[0], de[0], rtol=1e-11, atol=1e-13)
    for a, b in ((sp[1]["W"], de[1]["W"]), (sp[2]["W"], de[2]["W"])):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12 * max(np.abs(b).max(), 1.0))
    # the autograd twin differentiates the dense graph exactly like GradientTape differentiates the reference's
    lvec, gu, gi, nu, ni = tw.train_step(loss, Xu, Xi, "linear", "linear", pu, pi, rows, cols, vals, samp, n_i, S, lr=lr, dtype=torch.float64)
    np.testing.assert_allclose(np.asarray(lvec), sp[0], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(np.asarray(gu["W"]), sp[1]["W"], rtol=1e-8, atol=1e-11 * max(np.abs(sp[1]["W"]).max(), 1.0))
    np.testing.assert_allclose(np.asarray(gi["W"]), sp[2]["W"], rtol=1e-8, atol=1e-11 * max(np.abs(sp[2]["W"]).max(), 1.0))


def test_c2_has_the_shape_baseline_names():
    c = cfg.c2()
    assert (c["n_u"], c["n_i"], c["r"], c["S"]) == (943, 1682, 32, 1682 // 5)
    assert c["all_rows"].size == 100_000 and np.all(c["vals"] >= 4)
    key = c["all_rows"] * c["n_i"] + c["all_cols"]
    assert np.all(np.diff(key) > 0)  # deduplicated, row-major sorted like utils.py:53-57 produces
    assert all(len(set(row)) == c["S"] for row in c["samp"][:50])  # sampled without replacement (utils.py:20)
