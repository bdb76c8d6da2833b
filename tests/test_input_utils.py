"""CPU tests of the host-side input ETL: same results as the reference's list/dense code path,
without densifying."""
import random

import numpy as np
import pandas as pd
import pytest
from scipy import sparse

from teamoflow_b200.mf import input_utils as iu


def _df():
    return pd.DataFrame({"User ID": [10, 10, 7, 7, 3, 10, 42], "Items": ["a", "b", "a", "c", "b", "c", "a"],
                         "rating": [5.0, 3.0, 4.0, 1.0, 2.0, 4.5, 3.5]})


def test_create_iterable_interaction_first_appearance_order():
    rows, n_u, n_i = iu.create_iterable_interaction(_df())
    assert (n_u, n_i) == (4, 3)
    assert [r[0] for r in rows] == [0, 0, 1, 1, 2, 0, 3]
    assert [r[1] for r in rows] == [0, 1, 0, 2, 1, 2, 0]


def test_mask_train_test_split_matches_reference_statement():
    rows, n_u, n_i = iu.create_iterable_interaction(_df())
    random.seed(3)
    ref_rows = [list(r) for r in rows]
    random.shuffle(ref_rows)  # the reference shuffles in place with python's random (ref:48-50)
    thr = int(0.8 * len(ref_rows))
    want_train = sparse.csr_matrix(([r[2] for r in ref_rows[:thr]], ([r[0] for r in ref_rows[:thr]], [r[1] for r in ref_rows[:thr]])),
                                   shape=(n_u, n_i))
    random.seed(3)
    train, test, tri, tei = iu.mask_train_test_split([list(r) for r in rows], n_u, n_i)
    assert (train != want_train).nnz == 0
    assert train.shape == test.shape == (n_u, n_i)
    assert train.nnz + test.nnz == 7 and len(tri) == thr and len(tei) == 7 - thr
    assert iu.test_sparse_transformation(train, tri)


def test_df_to_sparse_pipeline_shapes():
    random.seed(0)
    train, test = iu.df_to_sparse_pipeline(_df())
    assert train.shape == test.shape == (4, 3) and train.nnz + test.nnz == 7


@pytest.mark.gpu
def test_convert_family_agrees_without_densifying():
    rng = np.random.default_rng(0)
    A = np.where(rng.random((7, 9)) < 0.3, rng.integers(1, 6, (7, 9)), 0).astype(np.float32)
    want_rows, want_cols = np.nonzero(A)
    for src in (A, A.tolist(), pd.DataFrame(A), sparse.csr_matrix(A), sparse.coo_matrix(A)):
        sp = iu.convert_to_tf_sparse(src)
        idx = sp.indices.cpu().numpy()
        assert sp.dense_shape == (7, 9)
        assert np.array_equal(idx[:, 0], want_rows) and np.array_equal(idx[:, 1], want_cols)
        assert np.array_equal(sp.values.cpu().numpy(), A[want_rows, want_cols])
    t = iu.convert_to_tensor_constant(A.tolist())
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), A)
