"""Minimal TensorFlow stand-in over torch-CPU (TEST INFRASTRUCTURE ONLY, see ../README.md).

Implements exactly the entry points `/root/reference/src/teamoflow/mf/*.py` calls, with TensorFlow's
documented semantics.  Tensors are plain ``torch.Tensor`` objects (fp32 / int64 on the CPU); variables are
leaf tensors with ``requires_grad=True`` that the optimizer updates in place.
"""
from __future__ import annotations

import collections
import types

import numpy as _np
import torch as _t

__version__ = "0.0-shim"

float32, float64, int32, int64, bool = _t.float32, _t.float64, _t.int32, _t.int64, _t.bool  # noqa: A001
newaxis = None
Tensor = _t.Tensor


def _as(x, dtype=None):
    if isinstance(x, _t.Tensor):
        return x if dtype is None or x.dtype == dtype else x.to(dtype)
    if isinstance(x, _np.ndarray):
        out = _t.from_numpy(_np.ascontiguousarray(x))
    else:
        out = _t.as_tensor(x)
    if dtype is None and out.dtype == _t.float64 and not isinstance(x, _np.ndarray):
        dtype = _t.float32  # python floats become float32 like tf.constant(1.5)
    return out if dtype is None else out.to(dtype)


def constant(value, dtype=None, shape=None):
    out = _as(value, dtype).clone()
    return out if shape is None else out.reshape(shape)


def convert_to_tensor(value, dtype=None):
    return _as(value, dtype)


def Variable(initial_value, trainable=True, dtype=None):  # noqa: N802
    v = _as(initial_value, dtype).detach().clone()
    if trainable and v.is_floating_point():
        v.requires_grad_(True)
    return v


def zeros(shape, dtype=float32):
    return _t.zeros(tuple(shape), dtype=dtype)


def eye(n, dtype=float32):
    return _t.eye(int(n), dtype=dtype)


def range(start, limit=None, delta=1, dtype=None):  # noqa: A001
    if limit is None:
        start, limit = 0, start
    return _t.arange(start, limit, delta, dtype=dtype)


def shape(x):
    return _as(x).shape


def cast(x, dtype):
    return _as(x).to(dtype)


def transpose(x):
    x = _as(x)
    return x.permute(*reversed(builtins_range(x.dim())))


def builtins_range(n):
    import builtins
    return builtins.range(n)


def matmul(a, b):
    return _t.matmul(_as(a), _as(b))


def expand_dims(x, axis):
    return _as(x).unsqueeze(axis)


def concat(values, axis):
    return _t.cat([_as(v) for v in values], dim=axis)


def repeat(x, repeats, axis=None):
    x = _as(x)
    if axis is None:
        return _t.repeat_interleave(x.reshape(-1), int(repeats))
    return _t.repeat_interleave(x, int(repeats), dim=axis)


def gather(params, indices, axis=0):
    return _t.index_select(_as(params), axis, _as(indices).to(_t.int64).reshape(-1)).reshape(
        tuple(_as(indices).shape) + tuple(_as(params).shape[1:])) if axis == 0 else NotImplemented


def gather_nd(params, indices):
    params, indices = _as(params), _as(indices).to(_t.int64)
    n = indices.shape[-1]
    return params[tuple(indices[..., i] for i in builtins_range(n))]


def where(condition, x=None, y=None):
    condition = _as(condition)
    if x is None and y is None:
        return _t.nonzero(condition)  # [n, ndim] int64, row-major order like tf.where
    x = _as(x)
    y = _as(y, x.dtype if isinstance(x, _t.Tensor) else None)
    return _t.where(condition, x, y)


def boolean_mask(tensor, mask):
    return _as(tensor)[_as(mask).to(_t.bool)]


def greater(x, y):
    return _as(x) > y


def less_equal(x, y):
    return _as(x) <= y


def maximum(x, y):
    # tf.maximum: the gradient goes to x where x >= y (to both halves on ties in newer TF versions only for
    # tensor-tensor; with a constant y the x-branch receives it when x >= y) [TF-sem]
    x = _as(x)
    yv = _as(y, x.dtype)
    return _MaxGE.apply(x, yv)


class _MaxGE(_t.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        ctx.save_for_backward(x >= y)
        ctx.xshape, ctx.yshape = x.shape, y.shape
        return _t.maximum(x, y)

    @staticmethod
    def backward(ctx, g):
        (ge,) = ctx.saved_tensors
        gx = _t.where(ge, g, _t.zeros_like(g))
        gy = _t.where(ge, _t.zeros_like(g), g)
        return _unbroadcast(gx, ctx.xshape), _unbroadcast(gy, ctx.yshape)


def _unbroadcast(g, shp):
    while g.dim() > len(shp):
        g = g.sum(0)
    for i, s in enumerate(shp):
        if s == 1 and g.shape[i] != 1:
            g = g.sum(i, keepdim=True)
    return g


def square(x):
    x = _as(x)
    return x * x


def sqrt(x):
    return _t.sqrt(_as(x))


def pow(x, y):  # noqa: A001
    return _t.pow(_as(x), _as(y))


def reduce_sum(x, axis=None):
    x = _as(x)
    return x.sum() if axis is None else x.sum(dim=axis)


def reduce_mean(x, axis=None):
    x = _as(x)
    return x.mean() if axis is None else x.mean(dim=axis)


_TopK = collections.namedtuple("TopKV2", ["values", "indices"])


def _top_k(x, k=1, sorted=True):  # noqa: A002
    """tf.math.top_k: descending values, equal values in ascending index order [TF-sem]; indices int32."""
    x = _as(x)
    k = int(k)
    vals, idx = _t.sort(x, dim=-1, descending=True, stable=True)
    return _TopK(vals[..., :k], idx[..., :k].to(_t.int32))


def _count_nonzero(x, axis=None, dtype=int64):
    x = _as(x)
    nz = (x != 0)
    return (nz.sum() if axis is None else nz.sum(dim=axis)).to(dtype)


def _l2_normalize(x, axis=None, epsilon=1e-12):
    """x * rsqrt(max(sum(x**2), epsilon)); axis=None reduces over the whole tensor [TF-sem]."""
    x = _as(x)
    ss = (x * x).sum() if axis is None else (x * x).sum(dim=axis, keepdim=True)
    return x * _t.rsqrt(_t.clamp(ss, min=epsilon))


math = types.SimpleNamespace(
    top_k=_top_k, count_nonzero=_count_nonzero, l2_normalize=_l2_normalize,
    log=lambda x: _t.log(_as(x)), log1p=lambda x: _t.log1p(_as(x)), pow=pow, square=square, sqrt=sqrt,
    not_equal=lambda x, y: _as(x) != y, is_nan=lambda x: _t.isnan(_as(x)), greater=greater, less_equal=less_equal,
    maximum=maximum, reduce_sum=reduce_sum, reduce_mean=reduce_mean)


def _moments(x, axes):
    """(mean, population variance) [TF-sem]."""
    x = _as(x)
    dims = tuple(axes)
    mean = x.mean(dim=dims)
    var = ((x - x.mean(dim=dims, keepdim=True)) ** 2).mean(dim=dims)
    return mean, var


nn = types.SimpleNamespace(relu=lambda x: _t.relu(_as(x)), moments=_moments, top_k=_top_k)

random = types.SimpleNamespace(
    normal=lambda shape, mean=0.0, stddev=1.0, dtype=float32, seed=None: _t.randn(tuple(shape), dtype=dtype) * stddev + mean,
    uniform=lambda shape, minval=0.0, maxval=1.0, dtype=float32, seed=None: _t.rand(tuple(shape), dtype=dtype) * (maxval - minval) + minval)


class _SparseTensor:
    """tf.sparse.SparseTensor: indices [nnz, ndim] int64, values [nnz], dense_shape."""

    def __init__(self, indices, values, dense_shape):
        self.indices = _as(indices).to(_t.int64)
        self.values = _as(values)
        self.dense_shape = tuple(int(s) for s in dense_shape)

    @property
    def shape(self):
        return self.dense_shape


sparse = types.SimpleNamespace(SparseTensor=_SparseTensor)
SparseTensor = _SparseTensor


class GradientTape:
    """Eager tape: torch autograd already records; gradient() of a non-scalar target differentiates its SUM [TF-sem]."""

    def __init__(self, persistent=False):
        self.persistent = persistent

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def gradient(self, target, sources):
        single = isinstance(sources, _t.Tensor)
        srcs = [sources] if single else list(sources)
        grads = _t.autograd.grad(target.sum(), srcs, retain_graph=True, allow_unused=True)
        return grads[0] if single else list(grads)


class _Adam:
    """tf.keras.optimizers.Adam: beta_1 0.9, beta_2 0.999, epsilon 1e-7; per-variable step counter and zero-initialised
    moments; update  w -= lr * sqrt(1 - b2^t) / (1 - b1^t) * m / (sqrt(v) + eps)  [TF-sem]."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps = float(learning_rate), beta_1, beta_2, epsilon
        self.t = 0
        self.state = {}

    def apply_gradients(self, grads_and_vars):
        self.t += 1
        f = _t.float32
        alpha = (_t.tensor(self.lr, dtype=f) * _t.sqrt(_t.tensor(1.0, dtype=f) - _t.tensor(self.b2, dtype=f) ** self.t)
                 / (_t.tensor(1.0, dtype=f) - _t.tensor(self.b1, dtype=f) ** self.t))
        with _t.no_grad():
            for g, w in grads_and_vars:
                if g is None:
                    continue
                m, v = self.state.setdefault(id(w), (_t.zeros_like(w), _t.zeros_like(w)))
                m.mul_(self.b1).add_(g * (_t.tensor(1.0, dtype=f) - _t.tensor(self.b1, dtype=f)))
                v.mul_(self.b2).add_((g * g) * (_t.tensor(1.0, dtype=f) - _t.tensor(self.b2, dtype=f)))
                w.sub_(alpha * m / (_t.sqrt(v) + self.eps))


keras = types.SimpleNamespace(optimizers=types.SimpleNamespace(Adam=_Adam))
