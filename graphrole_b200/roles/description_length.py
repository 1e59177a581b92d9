"""Minimum-description-length costs used by RoleExtractor's model selection (host side;
same definitions as graphrole/roles/description_length.py; SURVEY.md section 8f "next" #4)."""
from typing import Tuple

import numpy as np

from graphrole_b200.types import FactorTuple, MatrixLike


def get_description_length_costs(V: MatrixLike, model: FactorTuple) -> Tuple[float, float]:
    """(encoding cost, error cost) of representing V by the encoded factor pair `model`."""
    G_encoded, F_encoded = model
    V_orig = V.values if hasattr(V, 'values') else V
    return get_encoding_cost(model), get_error_cost(V_orig, G_encoded @ F_encoded)


def get_encoding_cost(model: FactorTuple) -> float:
    """bits per entry (from the larger codebook of the two factors) x number of entries."""
    G_encoded, F_encoded = model
    codebook = max(np.unique(G_encoded).size, np.unique(F_encoded).size)
    return np.ceil(np.log2(codebook)) * (G_encoded.size + F_encoded.size)


def get_error_cost(V: np.ndarray, V_approx: np.ndarray) -> float:
    """Generalised KL divergence sum(v log(v / v') - v + v') over the entries with v != 0
    (section 2.3 of the RolX paper)."""
    v = np.asarray(V, dtype=float).ravel()
    v_approx = np.asarray(V_approx, dtype=float).ravel()
    nz = v != 0
    terms = np.zeros_like(v)
    with np.errstate(divide='ignore', invalid='ignore'):
        terms[nz] = v[nz] * np.log(v[nz] / v_approx[nz]) - v[nz] + v_approx[nz]
    return float(np.sum(terms))
