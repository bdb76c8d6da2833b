"""Utility functions -- same names/signatures as the reference ``mf/utils.py``."""
import numpy as np
import torch
from scipy import sparse

from .. import _abi
from ._tensors import SparseInteractions, device, to_device

# above this many (user, item) cells the reference's O(n_users * n_items) host loop is replaced
# by the device sampler (tmf_sample_items)
HOST_SAMPLER_MAX_CELLS = 1 << 26


def random_sampler(n_items, n_users, n_samples, replace=False, seed=None):
    """Sampled item ids per user, int64 ``[n_users, n_samples]`` (reference ``utils.py:8-22``).

    Small problems run the reference's own statement -- ``np.random.choice(replace=...)`` per user from
    numpy's global RNG -- so ``np.random.seed`` reproduces the reference's negatives bit for bit.
    Large problems (or an explicit ``seed``) use the device sampler: a keyed bijection of
    ``[0, n_items)`` per user, i.e. the first ``n_samples`` entries of a pseudo-random permutation.
    """
    if n_samples > n_items and not replace:
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")  # numpy's message
    if seed is None and (replace or n_users * n_items <= HOST_SAMPLER_MAX_CELLS):
        items_per_user = [np.random.choice(a=n_items, size=n_samples, replace=replace) for _ in range(n_users)]
        return to_device(np.array(items_per_user, dtype=np.int64).reshape(n_users, n_samples), torch.int64)
    if replace:
        raise NotImplementedError("device sampler draws without replacement only")
    out = torch.empty(n_users, n_samples, dtype=torch.int64, device=device())
    s = int(seed) if seed is not None else int(np.random.randint(0, 2 ** 62))
    _abi.call("tmf_sample_items", n_users, n_items, n_samples, s, _abi.ptr(out))
    return out


def generate_random_interaction(n_users, n_items, min_val=0.0, max_val=5.0, density=0.50):
    """Random interaction table (reference ``utils.py:25-59``): returns ``(sparse, dense)`` where the
    sparse part holds the row-major-sorted nonzeros."""
    p = sparse.random(n_users, n_items, density=density)
    p = (max_val - min_val) * p + min_val * p.ceil()
    random_arr = np.round(p.toarray())
    scipy_random_arr = sparse.csr_matrix(random_arr)
    A = to_device(random_arr, torch.float32)
    row, col = scipy_random_arr.nonzero()
    nonzero_ind = np.stack([row, col], axis=1).astype(np.int64).reshape(-1, 2)
    vals = np.asarray(scipy_random_arr[row, col]).ravel().astype(np.float32) if row.size else np.zeros(0, np.float32)
    return SparseInteractions(nonzero_ind, vals, (n_users, n_items)), A


def gather_matrix_indices(input_arr, index_arr):
    """``out[r, c] = input_arr[r, index_arr[r, c]]`` (reference ``utils.py:62-105``; == torch.gather dim=1)."""
    inp = to_device(input_arr, torch.float32)
    idx = to_device(index_arr, torch.int64)
    if inp.dim() != 2 or idx.dim() != 2 or inp.shape[0] != idx.shape[0]:
        raise ValueError("input_arr [R, C] and index_arr [R, K] must share the row count")
    out = torch.empty(idx.shape, dtype=torch.float32, device=inp.device)
    _abi.call("tmf_gather_rows2d", _abi.ptr(inp), inp.shape[0], inp.shape[1], _abi.ptr(idx), idx.shape[1], _abi.ptr(out))
    return out
