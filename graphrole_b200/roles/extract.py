"""RolX role extraction (host driver around hot path B).

API surface of graphrole/roles/extract.py.  The factorisation itself,
graphrole_b200.roles.factor.get_nmf_decomposition, runs on the GPU; the MDL grid search,
quantisation and cost bookkeeping stay on the host (SURVEY.md section 8, rows B7 and f#4).
"""
from typing import Dict, Optional, Tuple

import numpy as np
import pandas as pd

from graphrole_b200.roles.description_length import get_description_length_costs
from graphrole_b200.roles.factor import encode, get_nmf_decomposition
from graphrole_b200.types import DataFrameLike, FactorTuple, Node


class RoleExtractor:

    """ Assign node roles based on input features """

    N_ROLE_RANGE = (2, 8)
    N_BIT_RANGE = (1, 8)

    def __init__(
        self,
        n_roles: Optional[int] = None,
        n_role_range: Optional[Tuple[int, int]] = None,
        n_bit_range: Optional[Tuple[int, int]] = None,
    ) -> None:
        """
        :param n_roles: optional number of roles to select; default uses MDL model selection
        :param n_role_range: optional (min, max) roles for the model selection grid search
        :param n_bit_range: optional (min, max) bits for the model selection grid search
        """
        self.n_roles = n_roles
        self.min_roles, self.max_roles = n_role_range if n_role_range else self.N_ROLE_RANGE
        self.min_bits, self.max_bits = n_bit_range if n_bit_range else self.N_BIT_RANGE
        self.node_role_factor: Optional[pd.DataFrame] = None
        self.role_feature_factor: Optional[pd.DataFrame] = None

    @property
    def roles(self) -> Optional[Dict[Node, str]]:
        """{node: name of the role with the largest factor entry}"""
        if self.node_role_factor is None:
            return None
        return self.node_role_factor.idxmax(axis=1).to_dict()

    @property
    def role_percentage(self) -> Optional[DataFrameLike]:
        """Node-role factor with every row normalised to sum to one."""
        if self.node_role_factor is None:
            return None
        return self.node_role_factor.div(self.node_role_factor.sum(axis=1), axis=0)

    def extract_role_factors(self, features: pd.DataFrame) -> None:
        """Factor the node-feature frame into node-role and role-feature frames."""
        if self.n_roles:
            # n_roles * (n_nodes + n_features) factor entries -> about log2 of that many bits
            n_bits = int(np.log2(self.n_roles * min(features.shape)))
            node_role, role_feature = self._get_encoded_role_factors(
                features, self.n_roles, n_bits)
        else:
            node_role, role_feature = self._select_model(features)

        role_labels = [f'role_{i}' for i in range(node_role.shape[1])]
        self.node_role_factor = pd.DataFrame(node_role, index=features.index,
                                             columns=role_labels)
        self.role_feature_factor = pd.DataFrame(role_feature, index=role_labels,
                                                columns=features.columns)

    def explain(self):
        raise NotImplementedError('Role explanation ("sense making") is not yet implemented.')

    def _select_model(self, features: pd.DataFrame) -> FactorTuple:
        """Grid search over (n_roles, n_bits); the model with the smallest rescaled
        encoding + error description length wins."""
        bit_stop = self.max_bits + 1
        role_stop = min(min(features.shape), self.max_roles) + 1
        encoding_costs = np.full((role_stop, bit_stop), np.nan)
        error_costs = np.full((role_stop, bit_stop), np.nan)
        models: Dict[Tuple[int, int], FactorTuple] = {}

        for roles in range(self.min_roles, role_stop):
            for bits in range(self.min_bits, bit_stop):
                try:
                    model = self._get_encoded_role_factors(features, roles, bits)
                    costs = get_description_length_costs(features, model)
                except ValueError:
                    # more bins than samples to quantise: skip this grid cell
                    continue
                encoding_costs[roles, bits], error_costs[roles, bits] = costs
                models[(roles, bits)] = model

        total = self._rescale_costs(encoding_costs) + self._rescale_costs(error_costs)
        best_roles, best_bits = np.argwhere(total == np.nanmin(total))[0]
        return models[(int(best_roles), int(best_bits))]

    @staticmethod
    def _get_encoded_role_factors(features: pd.DataFrame, n_roles: int,
                                  n_bits: int) -> FactorTuple:
        """NMF of the feature values followed by n_bits quantisation of both factors."""
        n_bins = int(2 ** n_bits)
        G, F = get_nmf_decomposition(features.values, n_roles)
        return encode(G, n_bins), encode(F, n_bins)

    @staticmethod
    def _rescale_costs(costs: np.ndarray) -> np.ndarray:
        """Divide every row (fixed n_roles) by its NaN-skipping Euclidean norm so encoding and
        error costs are comparable."""
        norms = np.sqrt(np.nansum(np.square(costs), axis=1, keepdims=True))
        with np.errstate(invalid='ignore', divide='ignore'):
            return costs / norms
