"""Prediction graphs -- same surface as the reference ``mf/predict_graphs.py``."""
from abc import ABC, abstractmethod

import torch

from .. import _abi
from ._engine import storage_of


class PredictionGraph(ABC):
    """Abstract base class of user-item scoring (reference ``predict_graphs.py:6-21``)."""

    @abstractmethod
    def get_prediction(self, user_embedding, item_embedding):
        """:return: tensor [n_users, n_items]"""
        pass


def dense_scores(user_embedding, item_embedding):
    """Full ``U V^T`` with the canonical score (fp64-accumulated, rounded to fp32)."""
    U, V = storage_of(user_embedding), storage_of(item_embedding)
    r = user_embedding.shape[1]
    if item_embedding.shape[1] != r:
        raise ValueError("user and item embeddings disagree on n_components")
    P = torch.empty(U.shape[0], V.shape[0], dtype=torch.float32, device=U.device)
    _abi.call("tmf_predict_dense", _abi.ptr(U), U.shape[0], _abi.ptr(V), V.shape[0], r, U.shape[1], _abi.ptr(P))
    return P


class DotProductPrediction(PredictionGraph):
    """``U @ V^T`` (reference ``predict_graphs.py:29-35``)."""

    def get_prediction(self, user_embedding, item_embedding):
        return dense_scores(user_embedding, item_embedding)
