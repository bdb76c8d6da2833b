"""Input conversion -- same function names and argument lists as the reference ``mf/input_utils.py`` (cited as
``ref:LINE``), re-done without the dense round trips: the reference builds every sparse tensor through a dense
ndarray (``ref:141-143``; the scipy branch calls ``.toarray()`` at ``ref:186``), which is impossible beyond
~10^5 x 10^5.  Host-side ETL (pandas / scipy) stays on the host; tensors land on the GPU once.
"""
import random

import numpy as np
import torch
from scipy import sparse

from ._tensors import SparseInteractions, as_features, as_interactions, to_device  # noqa: F401


def create_iterable_interaction(df):
    """ref:10-23 -- remap the 'User ID' / 'Items' columns to dense ids (order of first appearance) and return
    ``(rows as a list of [user, item, rating...], n_users, n_items)``.  Like the reference it mutates ``df``."""
    user_codes, users = _factorize(df['User ID'])
    item_codes, items = _factorize(df['Items'])
    df['User ID'] = user_codes
    df['Items'] = item_codes
    return df.values.tolist(), len(users), len(items)


def _factorize(col):
    import pandas as pd
    codes, uniques = pd.factorize(col, sort=False)  # == dict(enumerate(col.unique())) inverted, ref:11-15
    return codes, uniques


def mask_train_test_split(interactions, n_users, n_items, test_size=0.2, shuffle=True, return_indices=True):
    """ref:26-79 -- shuffle (python ``random``, in place, like the reference), split the rows at
    ``int((1 - test_size) * len)`` and build two same-shape CSR matrices (duplicates are summed by scipy)."""
    if shuffle:
        random.shuffle(interactions)
    train_size = 1.0 - test_size
    train_thresh = int(train_size * len(interactions))
    train, test = interactions[:train_thresh], interactions[train_thresh:]

    def _csr(part):
        arr = np.asarray(part, dtype=np.float64).reshape(-1, 3) if len(part) else np.zeros((0, 3))
        rows, cols, vals = arr[:, 0].astype(np.int64), arr[:, 1].astype(np.int64), arr[:, 2]
        return sparse.csr_matrix((vals, (rows, cols)), shape=(n_users, n_items)), rows, cols, vals

    train_sparse, tr_r, tr_c, tr_v = _csr(train)
    test_sparse, te_r, te_c, te_v = _csr(test)
    if return_indices:
        train_indices = list(zip(zip(tr_r.tolist(), tr_c.tolist()), tr_v.tolist()))
        test_indices = list(zip(zip(te_r.tolist(), te_c.tolist()), te_v.tolist()))
        return train_sparse, test_sparse, train_indices, test_indices
    return train_sparse, test_sparse


def test_sparse_transformation(sparse_interactions, li_indices):
    """ref:82-104 -- consistency check of a sparse matrix against its (index, value) list.  The reference
    densifies and, as written, can only ever return True (it appends False on a MATCH and tests ``not any``);
    this keeps that observable behaviour without the ``.toarray()``."""
    csr = sparse_interactions.tocsr()
    li_int = []
    for tup, val in li_indices:
        row, col = tup
        if csr[int(row), int(col)] == val:
            li_int.append(False)
    return not any(li_int)


test_sparse_transformation.__test__ = False  # not a pytest test despite the reference's name


def df_to_sparse_pipeline(df, test_size=0.2):
    """ref:107-130 -- DataFrame -> (train, test) CSR matrices of equal shape.  (Like the reference, the split
    itself is always 80/20: ``ref:120`` passes a literal 0.2.)"""
    li_df, n_users, n_items = create_iterable_interaction(df)
    train, test, train_indices, test_indices = mask_train_test_split(li_df, n_users, n_items, test_size=0.2,
                                                                     shuffle=True, return_indices=True)
    bool_train = test_sparse_transformation(train, train_indices)
    bool_test = test_sparse_transformation(test, test_indices)
    if bool_train and bool_test:
        return train, test
    print('Please check your input for errors.')
    return


def convert_np_to_tf_sparse(np_arr):
    """ref:133-153 -- dense ndarray -> sparse interactions (non-zeros in row-major order)."""
    return as_interactions(np.asarray(np_arr, dtype=np.float32))


def convert_tf_to_tf_sparse(tf_arr):
    """ref:156-161 -- dense tensor -> sparse interactions."""
    return as_interactions(tf_arr)


def convert_list_to_tf_sparse(li_arr):
    """ref:164-169"""
    return convert_np_to_tf_sparse(np.array(li_arr))


def convert_df_to_tf_sparse(df_arr):
    """ref:172-177"""
    return convert_np_to_tf_sparse(np.array(df_arr))


def convert_sp_sparse_to_tf_sparse(sp_arr):
    """ref:180-198 -- scipy sparse -> sparse interactions WITHOUT ``.toarray()``."""
    return as_interactions(sp_arr)


def convert_to_tf_sparse(arr):
    """list / ndarray / DataFrame / tensor / scipy sparse -> ``SparseInteractions`` (ref:201-220)."""
    if hasattr(arr, "values") and hasattr(arr, "columns"):  # pandas DataFrame
        return convert_df_to_tf_sparse(arr)
    return as_interactions(arr)


convert_to_sparse = convert_to_tf_sparse


def convert_to_tensor_constant(A):
    """array-like -> fp32 CUDA tensor (ref:223-242)."""
    if isinstance(A, torch.Tensor):
        return to_device(A, torch.float32)
    return to_device(np.asarray(A, dtype=np.float32), torch.float32)


def convert_to_tensor_trainable(self, arr):
    """ref:244-253 -- kept with the reference's (odd) signature: a module-level function that takes ``self``."""
    const_arr = self.convert_to_tensor_constant(arr)
    return const_arr.clone().requires_grad_(True)
