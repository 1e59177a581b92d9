"""Hot path B: the RolX non-negative matrix factorisation, on the GPU.

`get_nmf_decomposition(X, n_roles)` has the contract of graphrole/roles/factor.py:10-26, which
wraps sklearn NMF(solver='mu', init='nndsvda') (sklearn/decomposition/_nmf.py:726-888).  Here
the multiplicative-update loop is the CUDA library's gr_nmf_mu_f32 (include/graphrole_b200.h);
the NNDSVDa start is computed on the device with torch.linalg (not a hot path, SURVEY.md
section 8 row B2).
"""
import os
from ctypes import byref, c_double, c_int32, c_void_p
from typing import Optional, Tuple

import numpy as np
import torch

from graphrole_b200 import _native
from graphrole_b200.types import FactorTuple

# sklearn NMF defaults the reference relies on (_nmf.py:1533-1548)
MAX_ITER = 200
TOL = 1e-4
CHECK_EVERY = 10
MAX_ROLES = 32      # rank limit of the CUDA kernels (gr_nmf_create)
# which kernels the most recent nmf_mu call ran: 'tcgen05' or 'ffma'
last_path = None


def get_nmf_decomposition(X: np.ndarray, n_roles: int) -> FactorTuple:
    """
    Compute NMF decomposition X ~= G F
    :param X: matrix to factor (n_nodes x n_features, non-negative)
    :param n_roles: rank of decomposition
    :return: (G [n_nodes, n_roles], F [n_roles, n_features]) as float64 ndarrays
    """
    X = np.asarray(X)
    if X.ndim != 2:
        raise ValueError('X must be a 2-D array')
    if not np.all(np.isfinite(X)):
        # sklearn's input validation (check_array) refuses these before the solver starts
        raise ValueError('Input X contains NaN or infinity.')
    if np.any(X < 0):
        raise ValueError('Negative values in data passed to NMF')
    if n_roles > MAX_ROLES:
        raise ValueError(f'n_roles = {n_roles}: the CUDA solver supports at most {MAX_ROLES}')
    device = torch.device('cuda', torch.cuda.current_device())
    Xd = FeatureMatrix(torch.as_tensor(np.ascontiguousarray(X, dtype=np.float32), device=device))
    W0, H0 = nndsvda_init(Xd.values, n_roles)
    W, H, _, _ = Xd.nmf_mu(W0, H0, max_iter=MAX_ITER, tol=TOL)
    return W.double().cpu().numpy(), H.double().cpu().numpy()


class FeatureMatrix:
    """The float32 feature matrix in HBM, stored with its rows padded by zero columns to a
    multiple of 4 features -- what the tensor-core kernels need (16-byte row pitch for TMA).
    `values` is the [n, f] view, `padded` the [n, f4] buffer behind it.  A zero column of X is
    inert in the factorisation: with the matching column of H0 set to zero it stays zero (its
    update has numerator 0) and adds exact zeros to X H^T, H H^T and the residual, so the first
    f columns of the padded problem are the problem."""

    def __init__(self, V: torch.Tensor):
        if not (isinstance(V, torch.Tensor) and V.is_cuda and V.dim() == 2):
            raise ValueError('the feature matrix must be a 2-D CUDA tensor')
        n, f = V.shape
        f4 = -(-f // 4) * 4
        if f4 == f:
            self.padded = V.to(torch.float32).contiguous()
        else:
            self.padded = torch.zeros(n, f4, dtype=torch.float32, device=V.device)
            self.padded[:, :f] = V
        self.values = self.padded[:, :f]

    def nmf_mu(self, W0: torch.Tensor, H0: torch.Tensor, **kwargs):
        """nmf_mu on the padded matrix; returns (W, H [r, f], n_iter, error)."""
        f, f4 = self.values.shape[1], self.padded.shape[1]
        if f4 != f:
            H0 = torch.nn.functional.pad(H0, (0, f4 - f))
        W, H, n_iter, err = nmf_mu(self.padded, W0, H0, **kwargs)
        return W, (H if f4 == f else H[:, :f].contiguous()), n_iter, err


class NmfSolver:
    """A gr_nmf_t handle (workspaces, tensor maps) that can be reused across calls on matrices
    of the same shape -- `update` runs multiplicative updates IN PLACE on W and H."""

    def __init__(self, n: int, f: int, r: int, device=None):
        self.n, self.f, self.r = int(n), int(f), int(r)
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None \
            else torch.device(device)
        self._lib = _native.load()
        self._handle = c_void_p()
        with torch.cuda.device(self.device):
            _native.check(self._lib.gr_nmf_create(byref(self._handle), self.n, self.f, self.r,
                                                  self.device.index or 0), 'gr_nmf_create')
        self.last_path = None

    def update(self, X, W, H, max_iter=MAX_ITER, tol=TOL, check_every=CHECK_EVERY,
               use_tf32=True, want_error=True, stream=None) -> Tuple[int, float]:
        """Returns (n_iter, error); error is NaN when tol == 0 and want_error is False (no
        extra pass over X is made then)."""
        for name, t, shape in (('X', X, (self.n, self.f)), ('W', W, (self.n, self.r)),
                               ('H', H, (self.r, self.f))):
            if not (t.is_cuda and t.dtype == torch.float32 and tuple(t.shape) == shape):
                raise ValueError(f'{name} must be a float32 CUDA tensor of shape {shape}')
        if X.stride(1) != 1 or not W.is_contiguous() or not H.is_contiguous():
            raise ValueError('X needs unit column stride; W and H must be contiguous')
        n_iter, err = c_int32(0), c_double(float('nan'))
        err_ptr = byref(err) if (want_error or tol > 0) else None
        nvtx = os.environ.get('GR_NVTX') == '1'
        if nvtx:
            torch.cuda.nvtx.range_push(f'nmf mu n={self.n} f={self.f} r={self.r} <= {max_iter} it')
        with torch.cuda.device(self.device):
            _native.check(self._lib.gr_nmf_mu_f32(
                self._handle, c_void_p(X.data_ptr()), X.stride(0), c_void_p(W.data_ptr()),
                c_void_p(H.data_ptr()), max_iter, float(tol), check_every,
                1 if use_tf32 else 0, byref(n_iter), err_ptr, _native._stream_ptr(stream)),
                'gr_nmf_mu_f32')
        if nvtx:
            torch.cuda.nvtx.range_pop()
        self.last_path = 'tcgen05' if self._lib.gr_nmf_last_path(self._handle) else 'ffma'
        return int(n_iter.value), float(err.value)

    def close(self):
        if getattr(self, '_handle', None):
            self._lib.gr_nmf_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nmf_mu(X: torch.Tensor, W0: torch.Tensor, H0: torch.Tensor, max_iter: int = MAX_ITER,
           tol: float = TOL, check_every: int = CHECK_EVERY, use_tf32: bool = True,
           stream=None) -> Tuple[torch.Tensor, torch.Tensor, int, float]:
    """Multiplicative-update NMF from an explicit start (sklearn's init='custom').

    X [n, f], W0 [n, r], H0 [r, f]: float32 CUDA tensors.  Returns (W, H, n_iter, error) with
    error = ||X - W H||_F at the last convergence check (at the end when tol == 0).
    """
    for name, t in (('X', X), ('W0', W0), ('H0', H0)):
        if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2):
            raise ValueError(f'{name} must be a 2-D float32 CUDA tensor')
    n, f = X.shape
    r = W0.shape[1]
    if W0.shape != (n, r) or H0.shape != (r, f):
        raise ValueError('shapes must be X [n, f], W0 [n, r], H0 [r, f]')
    if X.stride(1) != 1:
        X = X.contiguous()
    W = W0.clone().contiguous()
    H = H0.clone().contiguous()
    solver = NmfSolver(n, f, r, X.device)
    try:
        n_iter, err = solver.update(X, W, H, max_iter, tol, check_every, use_tf32, True, stream)
        global last_path
        last_path = solver.last_path
    finally:
        solver.close()
    return W, H, n_iter, err


def nmf_error(X: torch.Tensor, W: torch.Tensor, H: torch.Tensor, use_tf32: bool = False) -> float:
    """||X - W H||_F evaluated by the library (dense residual, fp64 accumulation).  use_tf32:
    the tensor-core form the multiplicative-update loop uses for its convergence checks."""
    n, f = X.shape
    r = W.shape[1]
    lib = _native.load()
    handle = c_void_p()
    W = W.contiguous()
    H = H.contiguous()
    with torch.cuda.device(X.device):
        _native.check(lib.gr_nmf_create(byref(handle), n, f, r, X.device.index or 0),
                      'gr_nmf_create')
        try:
            err = c_double(0.0)
            fn = lib.gr_nmf_error_tf32 if use_tf32 else lib.gr_nmf_error_f32
            _native.check(fn(
                handle, c_void_p(X.data_ptr()), X.stride(0), c_void_p(W.data_ptr()),
                c_void_p(H.data_ptr()), byref(err), _native._stream_ptr(None)),
                'gr_nmf_error_tf32' if use_tf32 else 'gr_nmf_error_f32')
        finally:
            lib.gr_nmf_destroy(handle)
    return float(err.value)


def orthonormal_basis(Y: torch.Tensor) -> torch.Tensor:
    """Orthonormal basis of range(Y), Y [m, k] -- the `Q` of sklearn's range finder
    (sklearn/utils/extmath.py randomized_range_finder: linalg.qr(A @ Q, mode='economic')).

    Tall matrices (m >= 64 k: the n x k sketches of a feature matrix) take CholeskyQR2 with a
    float64 Gram matrix: two passes of [G = Y^T Y, Y <- Y chol(G)^-T], i.e. four streaming reads
    of Y instead of k Householder reflections over it (cuSOLVER's geqrf + orgqr took ~0.15 s per
    10 M x 18 sketch, 8 sketches per init: 1.65 s of a C5 initialisation, against 0.9 s for the
    whole multiplicative-update fit).  Any orthonormal basis of the same range gives the same
    projection B = Q^T A, hence the same U, S, V.  The Gram matrix squares the condition number,
    so the float64 Cholesky fails beyond cond(Y) ~ 1e8 (an exactly rank-deficient sketch, e.g.
    rank(X) < k); then -- and for short matrices -- Householder QR as before."""
    m, k = Y.shape
    if m < 64 * k:
        return torch.linalg.qr(Y)[0]
    Q = Y.double()
    for _ in range(2):
        G = Q.T @ Q
        L, info = torch.linalg.cholesky_ex(G)
        if int(info) != 0 or not bool(torch.isfinite(L).all()):
            return torch.linalg.qr(Y)[0]
        Q = Q @ torch.linalg.inv(L).T
    # a sketch whose Cholesky factor went through but lost the basis (cond^2 near 1 / eps)
    dev = (Q.T @ Q - torch.eye(k, dtype=Q.dtype, device=Q.device)).abs().max()
    if not bool(dev < 1e-6):
        return torch.linalg.qr(Y)[0]
    return Q.to(Y.dtype)


def nndsvda_init(X: torch.Tensor, n_components: int, eps: float = 1e-6,
                 n_oversamples: int = 10,
                 seed: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """NNDSVDa start (Boutsidis & Gallopoulos 2008; sklearn _nmf.py:214-366) on the device.

    The truncated SVD is a randomised subspace iteration (Halko et al. 2011) like sklearn's
    `_randomized_svd`; its Gaussian test matrix is drawn from NumPy's global RNG (or `seed`),
    so `np.random.seed` makes the start reproducible exactly as it does for the reference.
    """
    n, f = X.shape
    if n_components > min(n, f):
        raise ValueError("init = 'nndsvda' can only be used when "
                         'n_components <= min(n_samples, n_features)')
    rng = np.random.mtrand._rand if seed is None else np.random.RandomState(seed)
    k = min(n_components + n_oversamples, min(n, f))
    n_iter = 7 if n_components < 0.1 * min(n, f) else 4
    work = torch.float64 if n * f <= (1 << 26) else torch.float32
    A = X.to(work)
    Q = torch.from_numpy(rng.normal(size=(f, k))).to(X.device, work)
    Q = orthonormal_basis(A @ Q)
    for _ in range(n_iter):
        Q = orthonormal_basis(A.T @ Q)
        Q = orthonormal_basis(A @ Q)
    B = Q.T @ A                                   # k x f
    Ub, S, Vt = torch.linalg.svd(B.double(), full_matrices=False)
    U = (Q.double() @ Ub)[:, :n_components]
    S = S[:n_components]
    Vt = Vt[:n_components]
    # sign convention: largest-magnitude entry of every left vector is positive
    idx = U.abs().argmax(dim=0)
    signs = torch.sign(U[idx, torch.arange(n_components, device=U.device)])
    signs[signs == 0] = 1
    U = U * signs
    Vt = Vt * signs[:, None]

    W = torch.zeros_like(U)
    H = torch.zeros_like(Vt)
    W[:, 0] = S[0].sqrt() * U[:, 0].abs()
    H[0] = S[0].sqrt() * Vt[0].abs()
    for j in range(1, n_components):
        x, y = U[:, j], Vt[j]
        xp, yp = x.clamp(min=0), y.clamp(min=0)
        xn, yn = (-x).clamp(min=0), (-y).clamp(min=0)
        mp = xp.norm() * yp.norm()
        mn = xn.norm() * yn.norm()
        if mp > mn:
            u, v, sigma = xp / xp.norm(), yp / yp.norm(), mp
        else:
            u, v, sigma = xn / xn.norm(), yn / yn.norm(), mn
        scale = (S[j] * sigma).sqrt()
        W[:, j] = scale * u
        H[j] = scale * v
    W[W < eps] = 0
    H[H < eps] = 0
    avg = X.mean(dtype=torch.float64)       # fp64 accumulation without an fp64 copy of X
    W[W == 0] = avg
    H[H == 0] = avg
    return W.float().contiguous(), H.float().contiguous()


def encode(X: np.ndarray, n_bins: int) -> np.ndarray:
    """Quantise X to n_bins levels with a Lloyd-Max quantiser (1-D k-means on the entries), as
    graphrole/roles/factor.py:29-49 does with KMeans(n_clusters=n_bins, random_state=1); raises
    ValueError (TooManyBinsError) when n_bins exceeds the number of entries -- callers rely on
    that.  Runs on the GPU (gr_quantizer_*, csrc/rolx_epilogue.cu): same k-means++ stream, same
    stopping rules, labels equal scikit-learn's when X holds at least n_bins distinct values."""
    X = np.asarray(X, dtype=np.float64)
    if n_bins > X.size:
        raise _native.TooManyBinsError(f'n_samples={X.size} should be >= n_clusters={n_bins}.')
    device = torch.device('cuda', torch.cuda.current_device())
    flat = np.ascontiguousarray(X).reshape(-1, X.shape[-1] if X.ndim >= 1 and X.size else 1)
    quantizer = _native.Quantizer(X.size, device)
    try:
        out, _ = quantizer.bind(torch.as_tensor(flat, device=device)).encode(n_bins)
        return out.cpu().numpy().reshape(X.shape)
    finally:
        quantizer.close()
