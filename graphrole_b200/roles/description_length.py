"""Minimum-description-length costs used by RoleExtractor's model selection, on the GPU
(same definitions as graphrole/roles/description_length.py; SURVEY.md section 8f "next" #4).

The host functions keep the reference's signatures (NumPy in, floats out) and move float64
copies to the device; RoleExtractor's grid search calls the device forms directly on the
factors it already holds in HBM (roles/extract.py).
"""
from typing import Tuple

import numpy as np
import torch

from graphrole_b200 import _native
from graphrole_b200.types import FactorTuple, MatrixLike


def _device():
    return torch.device('cuda', torch.cuda.current_device())


def _to_device(a) -> torch.Tensor:
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if a.ndim != 2:
        a = a.reshape(1, -1)
    return torch.as_tensor(a, device=_device())


def get_description_length_costs(V: MatrixLike, model: FactorTuple) -> Tuple[float, float]:
    """(encoding cost, error cost) of representing V by the encoded factor pair `model`
    (description_length.py:8-29); the product G F is formed inside the cost kernel."""
    G_encoded, F_encoded = model
    V_orig = V.values if hasattr(V, 'values') else V
    G, F = _to_device(G_encoded), _to_device(F_encoded)
    return (encoding_cost_device(G, F), _native.mdl_error_cost(_to_device(V_orig), G, F))


def count_distinct_device(M: torch.Tensor) -> int:
    """np.unique(M).size on the device (sort + run count)."""
    quantizer = _native.Quantizer(M.numel(), M.device)
    try:
        return quantizer.bind(M).count_distinct()
    finally:
        quantizer.close()


def encoding_cost_from_counts(n_distinct_G: int, n_distinct_F: int, n_entries: int) -> float:
    """bits per entry (from the larger codebook of the two factors) x number of entries
    (description_length.py:32-41)."""
    return float(np.ceil(np.log2(max(n_distinct_G, n_distinct_F))) * n_entries)


def encoding_cost_device(G: torch.Tensor, F: torch.Tensor) -> float:
    return encoding_cost_from_counts(count_distinct_device(G), count_distinct_device(F),
                                     G.numel() + F.numel())


def get_encoding_cost(model: FactorTuple) -> float:
    G_encoded, F_encoded = model
    return encoding_cost_device(_to_device(G_encoded), _to_device(F_encoded))


def get_error_cost(V: np.ndarray, V_approx: np.ndarray) -> float:
    """Generalised KL divergence sum(v log(v / v') - v + v') over the entries with v != 0
    (section 2.3 of the RolX paper; description_length.py:44-61)."""
    return _native.mdl_kl(_to_device(V), _to_device(V_approx))
