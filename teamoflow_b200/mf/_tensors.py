"""Device-side containers standing in for the reference's TensorFlow types.

* ``SparseInteractions``  <->  ``tf.sparse.SparseTensor`` (``indices`` [nnz,2] int64, ``values`` [nnz]
  fp32, ``dense_shape``) as consumed by ``loss_graphs.py:47,74`` and ``matrix_factorization.py:154``.
* ``FeatureMatrix``       <->  the dense feature tensors of ``embedding_graphs.py:38``; stored as CSR
  (or flagged identity, the ``tf.eye`` every reference example uses) so that ``X @ W`` is a row gather.

torch is used for allocation / host<->device copies only.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _abi

try:  # scipy is optional at run time
    from scipy import sparse as _sp
except Exception:  # pragma: no cover
    _sp = None


def device():
    _abi.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def pad4(n):
    return (int(n) + 3) // 4 * 4


def to_device(x, dtype):
    """numpy / list / torch (any device) -> contiguous CUDA tensor of ``dtype`` (pinned staging for numpy)."""
    dev = device()
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=dtype, non_blocking=True).contiguous()
    arr = np.ascontiguousarray(np.asarray(x))
    t = torch.from_numpy(arr)
    if t.dtype != dtype:
        t = t.to(dtype)
    if t.numel() > (1 << 16):
        t = t.pin_memory()
    return t.to(dev, non_blocking=True)


def build_transpose(keys_i32, n_keys):
    """Stable counting transpose on the device: returns ``(ptr[n_keys+1], perm[n])`` int32."""
    n = keys_i32.numel()
    dev = keys_i32.device
    ptr = torch.empty(n_keys + 1, dtype=torch.int32, device=dev)
    perm = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    ws_bytes = _abi.query("tmf_transpose_ws_bytes", n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _abi.call("tmf_transpose_build", _abi.ptr(keys_i32), n, n_keys, _abi.ptr(ptr), _abi.ptr(perm), _abi.ptr(ws), ws_bytes)
    return ptr, perm[:n]


class SparseInteractions:
    """COO interaction table in stored order, plus the sorted CSR view the kernels use.

    ``indices`` may be int64 (``tf.sparse.SparseTensor``) or int32 (half the host->device bytes); the attribute
    ``.indices`` is the reference's int64 ``[nnz, 2]`` tensor either way (materialised on first use)."""

    def __init__(self, indices, values, dense_shape):
        raw = indices if isinstance(indices, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(indices)))
        if raw.dtype not in (torch.int32, torch.int64):
            raw = raw.to(torch.int64)
        if not raw.is_cuda and raw.numel() > (1 << 16) and not raw.is_pinned():
            raw = raw.pin_memory()
        self._raw = raw.to(device(), non_blocking=True).reshape(-1, 2).contiguous()
        self._idx64 = self._raw if self._raw.dtype == torch.int64 else None
        self.values = to_device(values, torch.float32).reshape(-1)
        self.dense_shape = (int(dense_shape[0]), int(dense_shape[1]))
        if self._raw.shape[0] != self.values.shape[0]:
            raise ValueError("indices and values disagree on nnz")
        if self.values.numel() >= 2 ** 31:
            raise ValueError("nnz must be < 2^31")
        self._csr = None

    @property
    def indices(self):
        if self._idx64 is None:
            self._idx64 = self._raw.to(torch.int64)
        return self._idx64

    @property
    def shape(self):
        return self.dense_shape

    @property
    def nnz(self):
        return int(self.values.numel())

    def csr(self):
        """``(row_ptr, col_idx, vals, coo_rows, perm)``: row-major sorted int32 view; ``perm`` maps sorted
        position -> stored position (None when the stored order is already row-major, as
        ``utils.py:53-57`` / ``input_utils.py:145-151`` produce).  Ids outside ``dense_shape`` raise (the kernels
        would gather rows past the end of the embedding tables; the reference's gather_nd raises there too,
        matrix_factorization.py:154)."""
        if self._csr is None:
            n_u, n_i = self.dense_shape
            nnz = self.nnz
            dev = self.values.device
            rows = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)[:nnz]
            cols = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)[:nnz]
            flags = torch.empty(1, dtype=torch.int32, device=dev)
            _abi.call("tmf_coo_split", _abi.ptr(self._raw), self._raw.element_size(), nnz, n_u, n_i, _abi.ptr(rows), _abi.ptr(cols),
                      _abi.ptr(flags))
            f = int(flags.item())
            if f & 1:
                r64, c64 = self._raw[:, 0], self._raw[:, 1]
                raise ValueError(f"interaction indices out of range for dense_shape {self.dense_shape}: rows in "
                                 f"[{int(r64.min())}, {int(r64.max())}], cols in [{int(c64.min())}, {int(c64.max())}] "
                                 "(are the ids 1-based?)")
            perm = None
            vals = self.values
            if f & 2:  # stored order is not row-major: stable sort once, losses go back to stored order
                key = rows.to(torch.int64) * n_i + cols.to(torch.int64)
                key, perm = torch.sort(key, stable=True)
                rows, cols = rows[perm].contiguous(), cols[perm].contiguous()
                vals = vals[perm].contiguous()
            row_ptr = torch.empty(n_u + 1, dtype=torch.int32, device=dev)
            _abi.call("tmf_rowptr_from_sorted", _abi.ptr(rows), nnz, n_u, _abi.ptr(row_ptr))
            self._csr = (row_ptr, cols, vals, rows, perm)
        return self._csr

    def to_dense(self):
        n_u, n_i = self.dense_shape
        A = torch.zeros(n_u, n_i, dtype=torch.float32, device=self.values.device)
        A.index_put_((self.indices[:, 0], self.indices[:, 1]), self.values, accumulate=True)
        return A

    def __repr__(self):
        return f"SparseInteractions(shape={self.dense_shape}, nnz={self.nnz})"


def as_interactions(x, shape=None):
    """Anything interaction-like -> ``SparseInteractions`` without densifying sparse inputs
    (the reference's ``convert_to_tf_sparse`` always goes through ``.toarray()``, ``input_utils.py:186``)."""
    if isinstance(x, SparseInteractions):
        return x
    if _sp is not None and _sp.issparse(x):
        coo = x.tocsr().tocoo()  # csr round trip: duplicates summed, row-major order
        mask = coo.data != 0
        idx = np.stack([coo.row[mask], coo.col[mask]], axis=1).astype(np.int64)
        return SparseInteractions(idx, coo.data[mask].astype(np.float32), coo.shape)
    if isinstance(x, torch.Tensor) and x.layout != torch.strided:
        c = x.coalesce() if x.layout == torch.sparse_coo else x.to_sparse_coo().coalesce()
        return SparseInteractions(c.indices().t().contiguous(), c.values(), tuple(c.shape))
    if isinstance(x, (tuple, list)) and len(x) == 3 and getattr(x[0], "ndim", 0) == 2 and len(x[2]) == 2:
        return SparseInteractions(x[0], x[1], x[2])
    # dense array-like: nonzeros in row-major order (convert_np_to_tf_sparse, input_utils.py:133-153)
    A = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float32))
    A = A.to(device(), dtype=torch.float32)
    idx = torch.nonzero(A)
    return SparseInteractions(idx, A[idx[:, 0], idx[:, 1]], tuple(A.shape))


# a dense feature tensor with more than this fraction of non-zeros stays dense (X W on the tensor cores, tmf_gemm_tc); below
# it the CSR gather / segment-sum wins (north_star: "a dense tcgen05/TMA GEMM only when the feature matrices are dense")
DENSE_FEATURE_MIN_DENSITY = 0.25


class FeatureMatrix:
    """Feature matrix ``X [n, F]`` as CSR (``ptr``, ``idx``, ``val``), the identity, or -- when genuinely dense -- a dense
    fp32 ``[n, ceil4(F)]`` storage (``dense``)."""

    def __init__(self, n_rows, n_cols, ptr=None, idx=None, val=None, identity=False, dense=None):
        self.shape = (int(n_rows), int(n_cols))
        self.identity = bool(identity)
        self.ptr, self.idx, self.val = ptr, idx, val
        self.dense = dense
        self._t = None

    @classmethod
    def eye(cls, n):
        return cls(n, n, identity=True)

    @property
    def nnz(self):
        if self.dense is not None:
            return int((self.dense != 0).sum())
        return self.shape[0] if self.identity else int(self.idx.numel())

    def transpose(self):
        """CSR of ``X^T`` (segments = feature columns), built once with the stable device transpose."""
        if self._t is None:
            n, F = self.shape
            ptr_t, perm = build_transpose(self.idx, F)
            rows = torch.repeat_interleave(torch.arange(n, dtype=torch.int32, device=self.idx.device),
                                           (self.ptr[1:] - self.ptr[:-1]).to(torch.int64))
            self._t = (ptr_t, rows[perm.long()].contiguous(), perm.contiguous())
        return self._t

    def to_dense(self):
        n, F = self.shape
        dev = device()
        if self.identity:
            return torch.eye(n, dtype=torch.float32, device=dev)
        if self.dense is not None:
            return self.dense[:, :F].clone()
        out = torch.zeros(n, F, dtype=torch.float32, device=dev)
        rows = torch.repeat_interleave(torch.arange(n, device=dev), (self.ptr[1:] - self.ptr[:-1]).to(torch.int64))
        out.index_put_((rows, self.idx.long()), self.val, accumulate=True)
        return out


def _csr_to_feature(m):
    m = m.tocsr()
    m.sum_duplicates()
    m.sort_indices()
    n, F = m.shape
    if n == F and m.nnz == n and np.array_equal(m.indices, np.arange(n)) and np.all(m.data == 1):
        return FeatureMatrix.eye(n)
    return FeatureMatrix(n, F, to_device(m.indptr, torch.int32), to_device(m.indices, torch.int32),
                         to_device(m.data, torch.float32))


def as_features(x):
    """Dense tensor / ndarray / scipy sparse / torch sparse / ``FeatureMatrix`` -> ``FeatureMatrix``.
    A dense identity (``tf.eye`` in every reference example) is detected and never multiplied."""
    if isinstance(x, FeatureMatrix):
        return x
    if _sp is not None and _sp.issparse(x):
        return _csr_to_feature(x)
    if isinstance(x, torch.Tensor) and x.layout != torch.strided:
        c = x.to_sparse_csr()
        n, F = c.shape
        return FeatureMatrix(n, F, to_device(c.crow_indices(), torch.int32), to_device(c.col_indices(), torch.int32),
                             to_device(c.values(), torch.float32))
    X = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float32))
    if X.dim() != 2:
        raise ValueError("features must be 2-D [n, n_features]")
    X = X.to(device(), dtype=torch.float32)
    n, F = X.shape
    if n == F and bool((X == torch.eye(n, device=X.device)).all()):
        return FeatureMatrix.eye(n)
    nz = torch.nonzero(X)  # row-major order
    if nz.shape[0] > DENSE_FEATURE_MIN_DENSITY * n * F:
        st = torch.zeros(n, pad4(F), dtype=torch.float32, device=X.device)
        st[:, :F] = X
        return FeatureMatrix(n, F, dense=st)
    counts = torch.bincount(nz[:, 0], minlength=n)
    ptr = torch.zeros(n + 1, dtype=torch.int32, device=X.device)
    ptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
    return FeatureMatrix(n, F, ptr, nz[:, 1].to(torch.int32).contiguous(), X[nz[:, 0], nz[:, 1]].contiguous())
