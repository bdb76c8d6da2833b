"""Weight initializers -- same classes/signatures as the reference ``mf/initializer_graphs.py``.

``initialize_weights(n_features, n_components)`` returns a trainable fp32 CUDA tensor ``[F, r]``:
iid N(0,1) (``:34``) or U[0,1) (``:51``) drawn by a Philox kernel, then L2-normalised over the WHOLE
tensor (``tf.math.l2_normalize`` with axis=None).  TensorFlow's RNG stream cannot be reproduced;
parity tests inject weights through a custom ``Initializer`` subclass, the one genuinely open
plugin point of the reference (``matrix_factorization.py:116-123``).
"""
from abc import ABC, abstractmethod

import numpy as np

from .. import _abi
from ._engine import new_storage, reduce_ws


class Initializer(ABC):
    """Abstract base class for weight initializers (reference ``initializer_graphs.py:7-19``)."""

    @abstractmethod
    def initialize_weights(self, n_features, n_components):
        """:return: torch CUDA tensor [n_features, n_components], fp32"""
        pass


def _seed(seed):
    return int(seed) if seed is not None else int(np.random.randint(0, 2 ** 62))


class NormalInitializer(Initializer):
    """Standard-normal entries, globally L2-normalised (reference ``:27-35``)."""

    def __init__(self, seed=None):
        self.seed = seed

    def initialize_weights(self, n_features, n_components):
        st = new_storage(n_features, n_components)
        _abi.call("tmf_fill_normal", _abi.ptr(st), n_features, n_components, st.shape[1], _seed(self.seed))
        _abi.call("tmf_l2_normalize_global", _abi.ptr(st), st.numel(), _abi.ptr(reduce_ws()))
        return st[:, :n_components]


class UniformInitializer(Initializer):
    """U[0,1) entries, globally L2-normalised (reference ``:43-52``)."""

    def __init__(self, seed=None):
        self.seed = seed

    def initialize_weights(self, n_features, n_components):
        st = new_storage(n_features, n_components)
        _abi.call("tmf_fill_uniform", _abi.ptr(st), n_features, n_components, st.shape[1], _seed(self.seed))
        _abi.call("tmf_l2_normalize_global", _abi.ptr(st), st.numel(), _abi.ptr(reduce_ws()))
        return st[:, :n_components]
