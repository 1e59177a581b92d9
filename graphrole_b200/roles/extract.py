"""RolX role extraction (driver around hot path B).

API surface of graphrole/roles/extract.py.  Everything heavy stays in HBM: the feature matrix is
copied to the device once, every factorisation (gr_nmf_mu_f32), quantisation (gr_quantizer_*),
description-length cost (gr_mdl_error_cost_*) and the role epilogue (gr_roles_*) runs there, and
only scalars -- two costs per grid cell -- come back during model selection (SURVEY.md section 8,
rows B7 and f#4).  The NMF of a given n_roles does not depend on n_bits, so the grid factorises
once per n_roles instead of once per cell like roles/extract.py:121-133 (`refit_per_cell=True`
restores the reference's behaviour, which consumes NumPy's global random stream per cell).
"""
from typing import Dict, Optional, Tuple

import numpy as np
import pandas as pd
import torch

from graphrole_b200 import _native
from graphrole_b200.roles import factor
from graphrole_b200.roles.description_length import encoding_cost_from_counts
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
        frame = self.node_role_factor
        if self._device_factor_matches(frame):
            arg, _ = _native.roles(self._node_role_device, want_percentage=False)
            labels = frame.columns.to_numpy()[arg.cpu().numpy()]
            return dict(zip(frame.index, labels))
        return frame.idxmax(axis=1).to_dict()

    @property
    def role_percentage(self) -> Optional[DataFrameLike]:
        """Node-role factor with every row normalised to sum to one."""
        if self.node_role_factor is None:
            return None
        frame = self.node_role_factor
        if self._device_factor_matches(frame):
            _, pct = _native.roles(self._node_role_device, want_argmax=False)
            return pd.DataFrame(pct.cpu().numpy(), index=frame.index, columns=frame.columns)
        return frame.div(frame.sum(axis=1), axis=0)

    def _device_factor_matches(self, frame: pd.DataFrame) -> bool:
        """True while node_role_factor is the frame extract_role_factors produced (its float64
        device copy is then used by the role kernels); a frame assigned by hand is served by
        pandas."""
        return getattr(self, '_node_role_device', None) is not None and \
            getattr(self, '_node_role_frame_id', None) == id(frame)

    def extract_role_factors(self, features: pd.DataFrame,
                             refit_per_cell: bool = False) -> None:
        """Factor the node-feature frame into node-role and role-feature frames."""
        if self.n_roles:
            # n_roles * (n_nodes + n_features) factor entries -> about log2 of that many bits
            n_bits = int(np.log2(self.n_roles * min(features.shape)))
            node_role, role_feature = self._get_encoded_role_factors(
                features, self.n_roles, n_bits)
        else:
            node_role, role_feature = self._select_model(features, refit_per_cell)

        role_labels = [f'role_{i}' for i in range(node_role.shape[1])]
        self.node_role_factor = pd.DataFrame(node_role, index=features.index,
                                             columns=role_labels)
        self.role_feature_factor = pd.DataFrame(role_feature, index=role_labels,
                                                columns=features.columns)
        self._node_role_device = torch.as_tensor(
            np.ascontiguousarray(node_role, dtype=np.float64), device=_cuda_device())
        self._node_role_frame_id = id(self.node_role_factor)

    def explain(self):
        raise NotImplementedError('Role explanation ("sense making") is not yet implemented.')

    def _select_model(self, features: pd.DataFrame, refit_per_cell: bool = False) -> FactorTuple:
        """Grid search over (n_roles, n_bits); the model with the smallest rescaled
        encoding + error description length wins (roles/extract.py:98-142).  The grid itself
        (encoding_costs, error_costs of the last search) is kept in `grid_costs_`."""
        bit_stop = self.max_bits + 1
        role_stop = min(min(features.shape), self.max_roles) + 1
        encoding_costs = np.full((role_stop, bit_stop), np.nan)
        error_costs = np.full((role_stop, bit_stop), np.nan)

        grid = DeviceModelGrid(features.values)
        try:
            for roles in range(self.min_roles, role_stop):
                for bits in range(self.min_bits, bit_stop):
                    try:
                        costs = grid.costs(roles, bits, refit=refit_per_cell)
                    except _native.TooManyBinsError:
                        # more bins than samples to quantise: skip this grid cell
                        # (roles/extract.py:127-129)
                        continue
                    encoding_costs[roles, bits], error_costs[roles, bits] = costs
            total = self._rescale_costs(encoding_costs) + self._rescale_costs(error_costs)
            self.grid_costs_ = (encoding_costs, error_costs)
            if np.all(np.isnan(total)):
                raise ValueError('model selection found no (n_roles, n_bits) cell it could '
                                 'evaluate: every quantisation had more bins than samples')
            best_roles, best_bits = np.argwhere(total == np.nanmin(total))[0]
            return grid.model(int(best_roles), int(best_bits))
        finally:
            grid.close()

    @staticmethod
    def _get_encoded_role_factors(features: pd.DataFrame, n_roles: int,
                                  n_bits: int) -> FactorTuple:
        """NMF of the feature values followed by n_bits quantisation of both factors."""
        grid = DeviceModelGrid(features.values)
        try:
            return grid.model(n_roles, n_bits)
        finally:
            grid.close()

    @staticmethod
    def _rescale_costs(costs: np.ndarray) -> np.ndarray:
        """Divide every row (fixed n_roles) by its NaN-skipping Euclidean norm so encoding and
        error costs are comparable."""
        norms = np.sqrt(np.nansum(np.square(costs), axis=1, keepdims=True))
        with np.errstate(invalid='ignore', divide='ignore'):
            return costs / norms


def _cuda_device() -> torch.device:
    return torch.device('cuda', torch.cuda.current_device())


class DeviceModelGrid:
    """The (n_roles, n_bits) cells of RoleExtractor's model selection with everything in HBM.

    Holds the feature matrix V (float32) on the device; per n_roles one factorisation
    V ~ G F (cached: it does not depend on n_bits) and one bound quantiser per factor (the sort
    and prefix sums behind every n_bins tried on it); per cell the encoded factors, their
    codebook sizes and the description-length error cost.  Only scalars reach the host until
    `model()` fetches the winning cell's factors."""

    def __init__(self, V, device: Optional[torch.device] = None):
        if isinstance(V, torch.Tensor) and V.is_cuda:
            # features already in HBM (e.g. DeviceRecursiveFeatureExtractor's output)
            if V.dim() != 2:
                raise ValueError('the feature matrix must be 2-D')
            if not bool(torch.isfinite(V).all()):
                raise ValueError('Input X contains NaN or infinity.')
            if bool((V < 0).any()):
                raise ValueError('Negative values in data passed to NMF')
            self.device = V.device
            self._features = factor.FeatureMatrix(V)
        else:
            V = np.asarray(V)
            if V.ndim != 2:
                raise ValueError('the feature matrix must be 2-D')
            if not np.all(np.isfinite(V)):
                raise ValueError('Input X contains NaN or infinity.')
            if np.any(V < 0):
                raise ValueError('Negative values in data passed to NMF')
            self.device = device or _cuda_device()
            self._features = factor.FeatureMatrix(
                torch.as_tensor(np.ascontiguousarray(V, dtype=np.float32), device=self.device))
        # [n, f] view of a buffer whose rows are padded to a multiple of 4 columns, so that any
        # feature count runs the tensor-core kernels
        self.V = self._features.values
        self.n, self.f = self.V.shape
        self._factors: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}
        self._quantizers: Dict[int, Tuple[_native.Quantizer, _native.Quantizer]] = {}
        self.n_fits = 0
        # set `timed` before the first call to get seconds per phase in `timings_s` (every phase
        # is then bracketed by a device synchronisation)
        self.timed = False
        self.timings_s: Dict[str, float] = {}
        self.nmf_iterations: Dict[int, int] = {}

    def _phase(self, name: str, fn):
        if not self.timed:
            return fn()
        import time
        torch.cuda.synchronize(self.device)
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize(self.device)
        self.timings_s[name] = self.timings_s.get(name, 0.0) + time.perf_counter() - t0
        return out

    @classmethod
    def from_device(cls, V: torch.Tensor) -> 'DeviceModelGrid':
        return cls(V)

    def factors(self, n_roles: int, refit: bool = False):
        if n_roles > factor.MAX_ROLES:
            raise ValueError(f'n_roles = {n_roles}: the CUDA solver supports at most '
                             f'{factor.MAX_ROLES}')
        if refit or n_roles not in self._factors:
            W0, H0 = self._phase('nndsvda_init', lambda: factor.nndsvda_init(self.V, n_roles))
            W, H, n_iter, _ = self._phase('nmf_mu', lambda: self._features.nmf_mu(
                W0, H0, max_iter=factor.MAX_ITER, tol=factor.TOL))
            self.nmf_iterations[n_roles] = n_iter
            self.n_fits += 1
            self._drop_quantizers(n_roles)
            self._factors[n_roles] = (W, H)
        return self._factors[n_roles]

    def _drop_quantizers(self, n_roles: int):
        for q in self._quantizers.pop(n_roles, ()):
            q.close()

    def quantizers(self, n_roles: int, refit: bool = False):
        W, H = self.factors(n_roles, refit)
        if n_roles not in self._quantizers:
            self._quantizers[n_roles] = self._phase('quantizer_bind', lambda: (
                _native.Quantizer(W.numel(), self.device).bind(W),
                _native.Quantizer(H.numel(), self.device).bind(H)))
        return self._quantizers[n_roles]

    def encoded(self, n_roles: int, n_bits: int, refit: bool = False):
        """(G_encoded, F_encoded, codebook size of G, of F) on the device."""
        n_bins = int(2 ** n_bits)
        qG, qF = self.quantizers(n_roles, refit)
        G, info_g = self._phase('encode', lambda: qG.encode(n_bins))
        F, info_f = self._phase('encode', lambda: qF.encode(n_bins))
        return G, F, info_g['n_distinct'], info_f['n_distinct']

    def costs(self, n_roles: int, n_bits: int, refit: bool = False) -> Tuple[float, float]:
        """(encoding cost, error cost) of one grid cell (description_length.py:8-29)."""
        G, F, n_g, n_f = self.encoded(n_roles, n_bits, refit)
        return (encoding_cost_from_counts(n_g, n_f, G.numel() + F.numel()),
                self._phase('error_cost', lambda: _native.mdl_error_cost(self.V, G, F)))

    def model(self, n_roles: int, n_bits: int) -> FactorTuple:
        """The cell's encoded factors as float64 host arrays (what the reference returns)."""
        G, F, _, _ = self.encoded(n_roles, n_bits)
        return G.double().cpu().numpy(), F.double().cpu().numpy()

    def close(self):
        for n_roles in list(self._quantizers):
            self._drop_quantizers(n_roles)
        self._factors.clear()
