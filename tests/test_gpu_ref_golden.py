"""The CUDA path (through the reference-shaped plugin surface and the C ABI) against golden vectors produced by the
REFERENCE'S OWN SOURCE run over tests/golden/tf_shim (tests/golden/make_ref_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "ref_golden.json")))


def arr(x, dtype=np.float32):
    def conv(v):
        if isinstance(v, list):
            return [conv(t) for t in v]
        return float(v) if isinstance(v, str) else v
    return np.asarray(conv(x), dtype=dtype)


def cpu(t):
    return t.detach().cpu().numpy()


D = G["data"]
ROWS, COLS, VALS = arr(D["rows"], np.int64), arr(D["cols"], np.int64), arr(D["vals"])
SAMP = arr(D["samp"], np.int64)
N_U, N_I, R, S = D["n_users"], D["n_items"], D["r"], D["S"]


def close_update(got, want, lr, what):
    """see tests/test_ref_golden.py: fresh-Adam updates are sign-like in the gradient"""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, what
    diff = np.abs(got - want)
    assert (diff <= 2.0 * lr + 1e-6).all(), what
    assert (diff > 2e-5).mean() <= 0.01, f"{what}: {(diff > 2e-5).mean():.3%} of the entries differ"


def fixed_init(W):
    from teamoflow_b200.mf.initializer_graphs import Initializer

    class Fixed(Initializer):
        def initialize_weights(self, n_features, n_components):
            assert W.shape == (n_features, n_components)
            return torch.as_tensor(W, device="cuda")
    return Fixed()


@pytest.mark.parametrize("ci", range(len(G["fit"]["cases"])))
def test_fit_steps_match_reference_source(ci):
    from teamoflow_b200.mf import embedding_graphs as E, loss_graphs as Lg
    from teamoflow_b200.mf._tensors import SparseInteractions
    from teamoflow_b200.mf.matrix_factorization import MatrixFactorization
    case = G["fit"]["cases"][ci]
    emb = {"linear": E.LinearEmbedding, "biased": E.BiasedLinearEmbedding, "relu": E.ReLUEmbedding}
    loss = {"mse": Lg.MSELoss, "wmrb": Lg.WMRBLoss, "kl": Lg.KLDivergenceLoss}[case["loss"]]
    Xu = np.eye(N_U, dtype=np.float32) if case["features"] == "id" else arr(G["fit"]["Xu_feat"])
    Xi = np.eye(N_I, dtype=np.float32) if case["features"] == "id" else arr(G["fit"]["Xi_feat"])
    inter = SparseInteractions(np.stack([ROWS, COLS], 1), VALS, (N_U, N_I))
    for epochs, key in ((1, "after1"), (2, "after2")):
        model = MatrixFactorization(R, user_repr_graph=emb[case["user"]](), item_repr_graph=emb[case["item"]](), loss_graph=loss(),
                                    user_weight_graph=fixed_init(arr(case["Wu0"])), item_weight_graph=fixed_init(arr(case["Wi0"])),
                                    n_users=N_U, n_items=N_I, n_samples=S)
        model.random_ind = torch.as_tensor(SAMP, device="cuda")
        for side, kind in (("user", case["user"]), ("item", case["item"])):
            if kind == "relu":  # injected exactly like the generator injects them into the reference model
                setattr(model, f"{side}_relu_weight", torch.as_tensor(arr(case["relu"][f"{side[0]}_rw"]), device="cuda"))
                setattr(model, f"{side}_relu_bias", torch.as_tensor(arr(case["relu"][f"{side[0]}_rb"]), device="cuda"))
        model.fit(epochs, torch.as_tensor(Xu), torch.as_tensor(Xi), inter, lr=case["lr"], verbose=False)
        want = case[key]
        tag = f"{case['loss']}/{case['user']}/{case['item']}/{case['features']} epochs={epochs}"
        assert len(model.user_trainable) == len(want["user_trainable"]) and len(model.item_trainable) == len(want["item_trainable"])
        for got, w in zip(model.user_trainable, want["user_trainable"]):
            close_update(cpu(got), arr(w), case["lr"], tag + " user trainable")
        for got, w in zip(model.item_trainable, want["item_trainable"]):
            close_update(cpu(got), arr(w), case["lr"], tag + " item trainable")
        if case["user"] != "relu":
            close_update(cpu(model.user_embedding), arr(want["user_embedding"]), case["lr"] * (1 + np.abs(Xu).sum(1).max()), tag + " E_u")
        if case["item"] != "relu":
            close_update(cpu(model.item_embedding), arr(want["item_embedding"]), case["lr"] * (1 + np.abs(Xi).sum(1).max()), tag + " E_i")


def test_loss_graphs_match_reference_source():
    from teamoflow_b200.mf import loss_graphs as Lg
    from teamoflow_b200.mf._tensors import SparseInteractions
    from teamoflow_b200.mf.utils import gather_matrix_indices
    c = G["loss_graphs"]
    P = torch.as_tensor(arr(c["P"]), device="cuda")
    inter = SparseInteractions(np.stack([ROWS, COLS], 1), VALS, (N_U, N_I))
    serial = P[torch.as_tensor(ROWS, device="cuda"), torch.as_tensor(COLS, device="cuda")]
    sample_preds = gather_matrix_indices(P, torch.as_tensor(SAMP, device="cuda"))
    np.testing.assert_allclose(cpu(Lg.MSELoss().get_loss(tf_interactions=inter, predictions=P)), arr(c["mse"]), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(cpu(Lg.WMRBLoss().get_loss(tf_interactions=inter, tf_sample_predictions=sample_preds,
                                                          tf_prediction_serial=serial, n_items=N_I, n_samples=S)),
                               arr(c["wmrb"]), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(cpu(Lg.KLDivergenceLoss().get_loss(tf_prediction_serial=serial, tf_interactions=inter)).reshape(()),
                               arr(c["kl"]), rtol=1e-5, atol=1e-7)
    g = G["gather_matrix_indices"]
    got = gather_matrix_indices(torch.as_tensor(arr(g["input"]), device="cuda"), torch.as_tensor(arr(g["index"], np.int64), device="cuda"))
    assert np.array_equal(cpu(got), arr(g["out"]))


def test_embedding_graphs_match_reference_source():
    from teamoflow_b200.mf import embedding_graphs as E
    c = G["embeddings"]
    t = lambda k: torch.as_tensor(arr(c[k]), device="cuda")  # noqa: E731
    np.testing.assert_allclose(cpu(E.LinearEmbedding().get_repr(t("X"), t("W"))[0]), arr(c["linear"]), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(cpu(E.BiasedLinearEmbedding().get_repr(t("X"), t("W"), linear_bias=t("b"))[0]), arr(c["biased"]), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(cpu(E.ReLUEmbedding().get_repr(t("X"), t("W5"), relu_weight=t("Wr"), relu_bias=t("br"))[0]), arr(c["relu"]),
                               rtol=1e-5, atol=1e-5)


def test_evaluation_surface_matches_reference_source_bit_exactly():
    from teamoflow_b200.mf._engine import new_storage
    from teamoflow_b200.mf.matrix_factorization import MatrixFactorization
    e = G["evaluate"]
    U, V, A = arr(e["U"]), arr(e["V"]), arr(e["A"])
    m = MatrixFactorization(R)
    m.user_embedding = new_storage(N_U, R, torch.as_tensor(U, device="cuda"))[:, :R]
    m.item_embedding = new_storage(N_I, R, torch.as_tensor(V, device="cuda"))[:, :R]
    At = torch.as_tensor(A, device="cuda")
    assert np.array_equal(cpu(m.predict()), arr(e["predict"]))
    assert np.array_equal(cpu(m.predict(At)[1]), arr(e["predict_unobserved"]))
    assert np.array_equal(cpu(m.predict_ranks(At)), arr(e["predict_ranks"], np.int64))
    for k in (3, 5, 17):
        w = e[f"k{k}"]
        np.testing.assert_array_equal(cpu(m.recall_at_k(At, k=k)), arr(w["recall"]))
        np.testing.assert_array_equal(cpu(m.recall_at_k(At, k=k, preserve_rows=True)), arr(w["recall_keep"]))
        np.testing.assert_array_equal(cpu(m.precision_at_k(At, k=k)), arr(w["precision"]))
        np.testing.assert_array_equal(cpu(m.precision_at_k(At, k=k, preserve_rows=True)), arr(w["precision_keep"]))
        np.testing.assert_allclose(float(m.f1_at_k(At, k=k)), arr(w["f1"]), rtol=1e-6)
        np.testing.assert_allclose(float(m.f1_at_k(At, k=k, beta=2.0)), arr(w["f1_beta2"]), rtol=1e-6)
        np.testing.assert_allclose(cpu(m.dcg_at_k(At, k=k)), arr(w["dcg"]), rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(cpu(m.idcg_at_k(At, k=k)), arr(w["idcg"]), rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(cpu(m.ndcg_at_k(At, k=k)), arr(w["ndcg"]), rtol=4e-6, atol=1e-6)
        np.testing.assert_allclose(cpu(m.ndcg_at_k(At, k=k, preserve_rows=True)), arr(w["ndcg_keep"]), rtol=4e-6, atol=1e-6)
        assert np.array_equal(m.retrieve_user_recs(k=k), arr(w["recs_all"], np.int32))
        assert np.array_equal(m.retrieve_user_recs(user=2, k=k), arr(w["recs_user2"], np.int32))
    assert np.array_equal(m.retrieve_user_recs(user=9), arr(e["recs_user9_full"], np.int32))
    assert np.array_equal(m.retrieve_user_recs(), arr(e["recs_full"], np.int32))
