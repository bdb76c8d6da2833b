"""Embedding graphs -- same classes/signatures as the reference ``mf/embedding_graphs.py``.

``get_repr(features, weights, aux_dim, relu_weight, relu_bias, linear_bias)`` returns
``(embedding, [trainables])`` exactly like the reference; ``features`` may be a dense tensor
(as in the reference), a scipy/torch sparse matrix or a ``FeatureMatrix``.  Inside
``MatrixFactorization.fit`` the classes act as descriptors selecting the fused kernels
(``_engine.Tower``); called directly they launch the same forward kernels.
"""
from abc import ABC, abstractmethod

import numpy as np
import torch

from .. import _abi
from ._engine import BIASED, LINEAR, RELU, Tower, new_storage, storage_of
from ._tensors import as_features


class Embeddings(ABC):
    """Abstract base class of embedding graphs (reference ``embedding_graphs.py:7-22``)."""
    kind = None

    @abstractmethod
    def get_repr(self, features, weights, aux_dim=None, relu_weight=None, relu_bias=None, linear_bias=None):
        pass


def _public(st, r):
    return st[:, :r]


class LinearEmbedding(Embeddings):
    """``X @ W`` (reference ``:30-38``)."""
    kind = LINEAR

    def get_repr(self, features, weights, aux_dim=None, relu_weight=None, relu_bias=None, linear_bias=None):
        X = as_features(features)
        r = weights.shape[1]
        t = Tower(LINEAR, X, r, storage_of(weights))
        return _public(t.forward(), r), [weights]


class BiasedLinearEmbedding(Embeddings):
    """``X @ W + b`` with a trainable ``b [1, r]`` created as zeros on first use (reference ``:45-58``)."""
    kind = BIASED

    def get_repr(self, features, weights, aux_dim=None, relu_weight=None, relu_bias=None, linear_bias=None):
        X = as_features(features)
        r = weights.shape[1]
        if linear_bias is None:
            linear_bias = _public(new_storage(1, r), r)
        t = Tower(BIASED, X, r, storage_of(weights), b=storage_of(linear_bias))
        return _public(t.forward(), r), [weights, linear_bias]


def new_relu_params(n_features, aux_dim, seed=None):
    """``relu_weight ~ N(0,1)`` un-normalised (reference ``:81``), ``relu_bias`` zeros (``:83``)."""
    Wr = new_storage(n_features, aux_dim)
    s = int(seed) if seed is not None else int(np.random.randint(0, 2 ** 62))
    _abi.call("tmf_fill_normal", _abi.ptr(Wr), n_features, aux_dim, Wr.shape[1], s)
    return _public(Wr, aux_dim), _public(new_storage(1, aux_dim), aux_dim)


class ReLUEmbedding(Embeddings):
    """``relu(X @ W_r + b_r) @ W`` -- one hidden layer of width ``5 * n_components`` (reference ``:66-87``)."""
    kind = RELU

    def get_repr(self, features, weights, aux_dim=None, relu_weight=None, relu_bias=None, linear_bias=None):
        X = as_features(features)
        n_features = X.shape[1]
        r = weights.shape[1]
        if aux_dim is None:
            aux_dim = 5 * r
        if weights.shape[0] != aux_dim:
            raise ValueError(f"ReLUEmbedding weights must be [aux_dim={aux_dim}, n_components], got {tuple(weights.shape)}")
        if relu_weight is None or relu_bias is None:
            rw, rb = new_relu_params(n_features, aux_dim)
            relu_weight = rw if relu_weight is None else relu_weight
            relu_bias = rb if relu_bias is None else relu_bias
        t = Tower(RELU, X, r, storage_of(weights), Wr=storage_of(relu_weight), br=storage_of(relu_bias))
        t.aux = aux_dim
        return _public(t.forward(), r), [weights, relu_weight, relu_bias]
