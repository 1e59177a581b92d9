"""Hot path B on several GPUs: row-sharded NMF multiplicative updates.

Rows of X [n, f] and W [n, r] are split over the ranks (one process per GPU, torch.distributed),
H [r, f] is replicated.  The W update is row-local; the H update needs W^T X and W^T W summed over
all rows -- r f + r^2 floats, the one exchange step of an iteration:

    rank-local   gr_nmf_iteration_local_f32   W_local update, local W^T X and W^T W
    collective   all-reduce(sum) of [W^T X | W^T W]           (NCCL over NVLink; <= 68 KB)
    replicated   gr_nmf_update_h_f32          H update, identical on every rank

and, at every convergence check (sklearn's rule, _nmf.py:867-879), an all-reduce of the local
squared residual.  Every rank therefore takes the same decisions and holds the same H; the result
differs from the single-GPU run only by the order in which the per-CTA partial sums are added.

The compute steps are behind a small backend object so that the loop -- the part that is new
here -- can be exercised on CPU (gloo, world_size 2) with a float64 stand-in for the kernels that
lives in the tests; the product backend is the CUDA library and nothing else.
"""
import math
from ctypes import byref, c_double, c_void_p
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from graphrole_b200 import _native
from graphrole_b200.roles import factor


def row_shard(n: int, world: int, rank: int, align: int = 128) -> Tuple[int, int]:
    """Rows [lo, hi) of rank `rank`: equal shares rounded to whole 128-row blocks (the row
    blocks of the kernels), the remainder on the last rank."""
    per = -(-n // world)
    per = -(-per // align) * align
    lo = min(n, rank * per)
    hi = n if rank == world - 1 else min(n, lo + per)
    return lo, hi


class CudaNmfBackend:
    """The three compute steps on this rank's GPU (include/graphrole_b200.h)."""

    def __init__(self, n_local: int, f: int, r: int, device: torch.device):
        self.n, self.f, self.r, self.device = int(n_local), int(f), int(r), torch.device(device)
        self._lib = _native.load()
        self._handle = c_void_p()
        with torch.cuda.device(self.device):
            _native.check(self._lib.gr_nmf_create(byref(self._handle), self.n, self.f, self.r,
                                                  self.device.index or 0), 'gr_nmf_create')
        # [W^T X | W^T W]: one buffer, one all-reduce
        self.sums = torch.empty(self.r * self.f + self.r * self.r, dtype=torch.float32,
                                device=self.device)

    def local_iteration(self, X, W, H, use_tf32=True) -> torch.Tensor:
        wtx = self.sums.data_ptr()
        wtw = wtx + 4 * self.r * self.f
        with torch.cuda.device(self.device):
            _native.check(self._lib.gr_nmf_iteration_local_f32(
                self._handle, c_void_p(X.data_ptr()), X.stride(0), c_void_p(W.data_ptr()),
                c_void_p(H.data_ptr()), 1 if use_tf32 else 0, c_void_p(wtx), c_void_p(wtw),
                _native._stream_ptr(None)), 'gr_nmf_iteration_local_f32')
        return self.sums

    def update_h(self, sums, H) -> None:
        wtx = sums.data_ptr()
        with torch.cuda.device(self.device):
            _native.check(self._lib.gr_nmf_update_h_f32(
                self._handle, c_void_p(wtx), c_void_p(wtx + 4 * self.r * self.f),
                c_void_p(H.data_ptr()), _native._stream_ptr(None)), 'gr_nmf_update_h_f32')

    def error_sq(self, X, W, H, use_tf32=True) -> float:
        err = c_double(0.0)
        tc = use_tf32 and bool(self._lib.gr_nmf_takes_tensor_cores(
            self._handle, c_void_p(X.data_ptr()), X.stride(0)))
        fn = self._lib.gr_nmf_error_tf32 if tc else self._lib.gr_nmf_error_f32
        with torch.cuda.device(self.device):
            _native.check(fn(self._handle, c_void_p(X.data_ptr()), X.stride(0),
                             c_void_p(W.data_ptr()), c_void_p(H.data_ptr()), byref(err),
                             _native._stream_ptr(None)), 'gr_nmf_error')
        return float(err.value) ** 2

    @property
    def last_path(self) -> str:
        return 'tcgen05' if self._lib.gr_nmf_last_path(self._handle) else 'ffma'

    def close(self):
        if getattr(self, '_handle', None):
            self._lib.gr_nmf_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _all_ranks_agree(ok: bool, group, device) -> bool:
    """True when `ok` holds on every rank (all-reduce MIN); every rank gets the same answer, so a
    failure on one rank becomes an exception on all of them instead of a hang in the next
    collective."""
    if not dist.is_initialized():
        return ok
    if dist.get_backend(group) != 'nccl':
        device = torch.device('cpu')
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(flag.item())


class RowShardedNmf:
    """Multiplicative updates on this rank's rows; `fit` mirrors gr_nmf_mu_f32's loop.
    Construction is collective: every rank must hold at least one row and get its workspaces,
    or all ranks raise."""

    def __init__(self, n_local: int, f: int, r: int, device=None, group=None, backend=None):
        self.n, self.f, self.r = int(n_local), int(f), int(r)
        self.group = group
        device = torch.device('cpu') if backend is not None and device is None else (
            torch.device('cuda', torch.cuda.current_device()) if device is None
            else torch.device(device))
        if not _all_ranks_agree(self.n >= 1, group, device):
            raise ValueError('row-sharded NMF: a rank holds no rows (fewer row blocks than ranks)')
        error = None
        if backend is None:
            try:
                backend = CudaNmfBackend(self.n, self.f, self.r, device)
            except Exception as exc:           # e.g. out of memory on this rank only
                error = exc
        if not _all_ranks_agree(error is None, group, device):
            if backend is not None and hasattr(backend, 'close'):
                backend.close()
            raise RuntimeError(f'row-sharded NMF: set-up failed on a rank ({error!r})')
        self.backend = backend
        self.allreduce_bytes_per_iteration = 4 * (self.r * self.f + self.r * self.r)

    def _sum_over_ranks(self, value: float, like: torch.Tensor) -> float:
        t = torch.tensor([value], dtype=torch.float64, device=like.device)
        if dist.is_initialized():
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return float(t.item())

    def fit(self, X, W, H, max_iter: int = factor.MAX_ITER, tol: float = factor.TOL,
            check_every: int = factor.CHECK_EVERY, use_tf32: bool = True) -> Tuple[int, float]:
        """In place on W (this rank's rows) and H (replicated; must be equal on every rank).
        Returns (n_iter, ||X - W H||_F over ALL rows at the last check; NaN when tol == 0)."""
        b = self.backend
        error_at_init = previous = error = float('nan')
        if tol > 0:
            error_at_init = math.sqrt(self._sum_over_ranks(b.error_sq(X, W, H, use_tf32), H))
            previous = error = error_at_init
        n_iter = 0
        for it in range(1, max_iter + 1):
            sums = b.local_iteration(X, W, H, use_tf32)
            if dist.is_initialized():
                dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
            b.update_h(sums, H)
            n_iter = it
            if tol > 0 and it % check_every == 0:
                error = math.sqrt(self._sum_over_ranks(b.error_sq(X, W, H, use_tf32), H))
                if (previous - error) / error_at_init < tol:      # _nmf.py:877
                    break
                previous = error
        return n_iter, error

    def close(self):
        close = getattr(self.backend, 'close', None)
        if close:
            close()


def nmf_mu_row_sharded(X_local: torch.Tensor, W0_local: torch.Tensor, H0: torch.Tensor,
                       max_iter: int = factor.MAX_ITER, tol: float = factor.TOL,
                       check_every: int = factor.CHECK_EVERY, use_tf32: bool = True,
                       group=None) -> Tuple[torch.Tensor, torch.Tensor, int, float]:
    """factor.nmf_mu for a row shard: X_local [n_local, f], W0_local [n_local, r], H0 [r, f] (the
    same on every rank).  Ranks that are not a multiple of 4 run on zero-padded factors, as in
    the single-GPU library call.  Returns (W_local, H, n_iter, global error)."""
    for name, t in (('X_local', X_local), ('W0_local', W0_local), ('H0', H0)):
        if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2):
            raise ValueError(f'{name} must be a 2-D float32 CUDA tensor')
    n, f = X_local.shape
    r = W0_local.shape[1]
    if W0_local.shape != (n, r) or H0.shape != (r, f):
        raise ValueError('shapes must be X [n, f], W0 [n, r], H0 [r, f]')
    if r > factor.MAX_ROLES:
        raise ValueError(f'n_roles = {r}: the CUDA solver supports at most {factor.MAX_ROLES}')
    r4 = -(-r // 4) * 4 if use_tf32 else r
    W = torch.zeros(n, r4, dtype=torch.float32, device=X_local.device)
    W[:, :r] = W0_local
    H = torch.zeros(r4, f, dtype=torch.float32, device=X_local.device)
    H[:r] = H0
    X = X_local if X_local.stride(1) == 1 else X_local.contiguous()
    solver = RowShardedNmf(n, f, r4, X.device, group)
    try:
        n_iter, err = solver.fit(X, W, H, max_iter, tol, check_every, use_tf32)
    finally:
        solver.close()
    return W[:, :r].contiguous(), H[:r].contiguous(), n_iter, err
