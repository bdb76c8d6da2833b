"""Input conversion -- the entry points of the reference ``mf/input_utils.py`` that sit directly
in front of ``fit`` (``convert_to_tf_sparse`` family, :133-220, and ``convert_to_tensor_constant``,
:223-242), re-done without the dense round trip (``.toarray()`` at :186)."""
import numpy as np
import torch

from ._tensors import SparseInteractions, as_features, as_interactions, to_device  # noqa: F401


def convert_to_tf_sparse(arr):
    """list / ndarray / DataFrame / tensor / scipy sparse -> ``SparseInteractions`` (reference :201-220)."""
    if hasattr(arr, "values") and hasattr(arr, "columns"):  # pandas DataFrame
        arr = np.asarray(arr, dtype=np.float32)
    return as_interactions(arr)


convert_to_sparse = convert_to_tf_sparse


def convert_to_tensor_constant(A):
    """array-like -> fp32 CUDA tensor (reference :223-242)."""
    if isinstance(A, torch.Tensor):
        return to_device(A, torch.float32)
    return to_device(np.asarray(A, dtype=np.float32), torch.float32)
