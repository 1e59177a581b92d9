"""ctypes binding of libgraphrole_b200.so (C-ABI declared in include/graphrole_b200.h).

This module is deliberately thin: argument marshalling and error translation only.  There is
no CPU fallback anywhere in the package -- if the shared library is missing or the device is
not sm_100, calls raise.
"""
import ctypes
import os
from ctypes import (POINTER, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_uint32,
                    c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libgraphrole_b200.so')

GR_OK = 0
GR_ERR_INVALID_ARGUMENT = 1
GR_ERR_INVALID_GRAPH = 2
GR_ERR_CUDA = 3
GR_ERR_UNSUPPORTED_DEVICE = 4
GR_ERR_OUT_OF_MEMORY = 5
GR_CSR_VALIDATE = 1
GR_CSR_HOT_HINTS = 2

# every symbol include/graphrole_b200.h declares: (restype, argtypes)
SIGNATURES = {
    'gr_last_error': (c_char_p, []),
    'gr_version': (c_char_p, []),
    'gr_kernel_launch_count': (c_int64, []),
    'gr_csr_create': (c_int, [POINTER(c_void_p), c_int64, c_int64, c_int64, c_void_p, c_void_p,
                              c_int, c_int]),
    'gr_csr_destroy': (c_int, [c_void_p]),
    'gr_csr_tune_hot_rows': (c_int, [c_void_p, c_int64]),
    'gr_csr_info': (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64),
                            POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]),
    'gr_refex_aggregate_f32': (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int64, c_int64,
                                       c_void_p, c_void_p, c_int64, c_void_p]),
    'gr_refex_aggregate_bcast_f32': (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int64,
                                             c_int64, c_void_p, POINTER(c_void_p), c_int32,
                                             c_int64, c_void_p]),
    'gr_peer_alloc': (c_int, [POINTER(c_void_p), c_int64, c_int, c_void_p]),
    'gr_peer_open': (c_int, [POINTER(c_void_p), c_void_p, c_int]),
    'gr_peer_close': (c_int, [c_void_p, c_int]),
    'gr_peer_free': (c_int, [c_void_p, c_int]),
    'gr_peer_barrier': (c_int, [POINTER(c_void_p), c_int32, c_int32, c_int64, c_double,
                                c_void_p]),
    'gr_peer_flag_words': (c_int64, []),
    'gr_peer_barrier_status': (c_int, [c_void_p, POINTER(c_int64)]),
    'gr_refex_levels_host_f32': (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32,
                                         c_void_p, c_void_p]),
    'gr_refex_levels_host_sharded_f32': (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int32,
                                                 c_int64, POINTER(c_void_p), POINTER(c_void_p),
                                                 POINTER(c_void_p), c_int32, c_int32,
                                                 POINTER(c_int64), c_void_p, c_void_p]),
    'gr_nmf_create': (c_int, [POINTER(c_void_p), c_int64, c_int32, c_int32, c_int]),
    'gr_nmf_destroy': (c_int, [c_void_p]),
    'gr_nmf_mu_f32': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_double,
                              c_int32, c_int32, POINTER(c_int32), POINTER(c_double), c_void_p]),
    'gr_nmf_last_path': (c_int, [c_void_p]),
    'gr_nmf_takes_tensor_cores': (c_int, [c_void_p, c_void_p, c_int64]),
    'gr_pruner_create': (c_int, [POINTER(c_void_p), c_int64, c_int]),
    'gr_pruner_destroy': (c_int, [c_void_p]),
    'gr_prune_bin_f32': (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_double, c_void_p,
                                 c_int64, c_void_p]),
    'gr_prune_bin_f64': (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_double, c_void_p,
                                 c_int64, c_void_p]),
    'gr_prune_pairwise_gap_i32': (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_void_p,
                                          c_void_p]),
    'gr_level0_features_f64': (c_int, [c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int32,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                       c_void_p]),
    'gr_nmf_error_f32': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                 POINTER(c_double), c_void_p]),
    'gr_nmf_error_tf32': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                  POINTER(c_double), c_void_p]),
    'gr_nmf_iteration_local_f32': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int32,
                                           c_void_p, c_void_p, c_void_p]),
    'gr_nmf_update_h_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'gr_quantizer_create': (c_int, [POINTER(c_void_p), c_int64, c_int]),
    'gr_quantizer_destroy': (c_int, [c_void_p]),
    'gr_quantizer_bind_f32': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    'gr_quantizer_bind_f64': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    'gr_quantizer_encode_f32': (c_int, [c_void_p, c_int32, c_uint32, c_int32, c_double, c_void_p,
                                        c_int64, POINTER(c_double), POINTER(c_int32),
                                        POINTER(c_int64), c_void_p]),
    'gr_quantizer_encode_f64': (c_int, [c_void_p, c_int32, c_uint32, c_int32, c_double, c_void_p,
                                        c_int64, POINTER(c_double), POINTER(c_int32),
                                        POINTER(c_int64), c_void_p]),
    'gr_quantizer_count_distinct': (c_int, [c_void_p, POINTER(c_int64), c_void_p]),
    'gr_mdl_error_cost_f32': (c_int, [c_void_p, c_int64, c_int32, c_int64, c_void_p, c_int64,
                                      c_void_p, c_int64, c_int32, POINTER(c_double), c_int,
                                      c_void_p]),
    'gr_mdl_error_cost_f64': (c_int, [c_void_p, c_int64, c_int32, c_int64, c_void_p, c_int64,
                                      c_void_p, c_int64, c_int32, POINTER(c_double), c_int,
                                      c_void_p]),
    'gr_mdl_kl_f64': (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int64, c_int64,
                              POINTER(c_double), c_int, c_void_p]),
    'gr_roles_f32': (c_int, [c_void_p, c_int64, c_int32, c_int64, c_void_p, c_void_p, c_int64,
                             c_int, c_void_p]),
    'gr_roles_f64': (c_int, [c_void_p, c_int64, c_int32, c_int64, c_void_p, c_void_p, c_int64,
                             c_int, c_void_p]),
    'gr_numpy_random_sample': (c_int, [c_uint32, c_int64, POINTER(c_double)]),
    'gr_numpy_choice_uniform': (c_int, [c_uint32, c_int64, POINTER(c_int64)]),
}


class NativeLibraryError(RuntimeError):
    """The CUDA library is missing, failed to load, or a call into it failed."""

    def __init__(self, message, code=None):
        super().__init__(message)
        self.code = code


class TooManyBinsError(ValueError):
    """More quantisation bins than matrix entries: the ValueError scikit-learn's KMeans raises for
    n_samples < n_clusters, which RoleExtractor's grid search expects and skips
    (graphrole/roles/extract.py:127-129).  A class of its own so that the grid does not swallow
    unrelated argument errors."""


_lib = None


def load():
    """Load libgraphrole_b200.so and bind every declared symbol.  Raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; '
            f'g.build()"` (nvcc, sm_100a). graphrole_b200 has no CPU fallback.')
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as exc:
        raise NativeLibraryError(f'cannot load {LIB_PATH}: {exc}') from exc
    for name, (restype, argtypes) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise NativeLibraryError(f'{LIB_PATH} does not export {name}') from exc
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(code, what):
    """Translate a non-zero status into the exception class the reference's callers expect."""
    if code == GR_OK:
        return
    msg = (load().gr_last_error() or b'').decode('utf-8', 'replace')
    text = f'{what}: {msg} (status {code})'
    if code in (GR_ERR_INVALID_ARGUMENT, GR_ERR_INVALID_GRAPH):
        raise ValueError(text)
    if code == GR_ERR_OUT_OF_MEMORY:
        raise MemoryError(text)
    raise NativeLibraryError(text, code)


def launch_count():
    return int(load().gr_kernel_launch_count())


def version():
    return load().gr_version().decode()


def _stream_ptr(stream):
    """cudaStream_t of a torch stream (None = torch's current stream)."""
    import torch
    if stream is None:
        stream = torch.cuda.current_stream()
    return c_void_p(stream.cuda_stream)


class CsrHandle:
    """Owner of a gr_csr_t*.  Keeps the rowptr/colidx tensors alive (the library does not copy)."""

    def __init__(self, rowptr, colidx, n_cols=None, validate=True, hot_hints=True):
        import torch
        lib = load()
        if not (rowptr.is_cuda and colidx.is_cuda):
            raise ValueError('rowptr/colidx must be CUDA tensors')
        if rowptr.dtype != torch.int64 or colidx.dtype != torch.int32:
            raise ValueError('rowptr must be int64 and colidx int32')
        self.rowptr = rowptr.contiguous()
        self.colidx = colidx.contiguous()
        self.n_rows = self.rowptr.numel() - 1
        self.n_cols = int(self.n_rows if n_cols is None else n_cols)
        self.nnz = self.colidx.numel()
        self.device = self.rowptr.device
        handle = c_void_p()
        # the library reads rowptr with a blocking copy on the legacy stream: make sure the
        # producer stream is done
        torch.cuda.current_stream(self.device).synchronize()
        check(lib.gr_csr_create(byref(handle), self.n_rows, self.n_cols, self.nnz,
                                c_void_p(self.rowptr.data_ptr()),
                                c_void_p(self.colidx.data_ptr()),
                                self.device.index or 0,
                                (GR_CSR_VALIDATE if validate else 0)
                                | (GR_CSR_HOT_HINTS if hot_hints else 0)),
              'gr_csr_create')
        self._handle = handle

    def info(self):
        vals = [c_int64() for _ in range(6)]
        check(load().gr_csr_info(self._handle, *[byref(v) for v in vals]), 'gr_csr_info')
        keys = ('n_rows', 'n_cols', 'nnz', 'n_hub_rows', 'n_hub_segments', 'n_hot_rows')
        return {k: int(v.value) for k, v in zip(keys, vals)}

    def tune_hot_rows(self, row_bytes):
        """Re-tag the hot rows for feature rows of `row_bytes` bytes (cache hint only)."""
        check(load().gr_csr_tune_hot_rows(self._handle, int(row_bytes)), 'gr_csr_tune_hot_rows')

    def aggregate(self, X, out=None, row_lo=0, row_hi=None, stream=None):
        """One recursion level.  X: [n_cols, d] fp32 (row stride free, unit column stride).

        Returns `out` of shape [row_hi - row_lo, 2*d]: columns [0, d) = sum, [d, 2d) = mean,
        the reference's agg-major order (graphrole/features/extract.py:158-162).
        """
        import torch
        if X.dtype != torch.float32 or not X.is_cuda or X.dim() != 2:
            raise ValueError('X must be a 2-D float32 CUDA tensor')
        if X.stride(1) != 1 and X.shape[1] > 1:
            raise ValueError('X must have unit column stride')
        if X.shape[0] != self.n_cols:
            raise ValueError(f'X has {X.shape[0]} rows, graph addresses {self.n_cols}')
        d = X.shape[1]
        row_hi = self.n_rows if row_hi is None else row_hi
        rows = row_hi - row_lo
        if out is None:
            out = torch.empty((rows, 2 * d), dtype=torch.float32, device=X.device)
        if out.shape != (rows, 2 * d) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError(f'out must be a contiguous float32 [{rows}, {2 * d}] tensor')
        ldx = X.stride(0) if X.shape[0] > 1 else max(d, X.stride(0))
        # out pointers are addressed by handle-local row number: rebase to row 0
        base = out.data_ptr() - row_lo * 2 * d * 4
        check(load().gr_refex_aggregate_f32(
            self._handle, c_void_p(X.data_ptr()), ldx, d, row_lo, row_hi,
            c_void_p(base), c_void_p(base + d * 4), 2 * d, _stream_ptr(stream)),
            'gr_refex_aggregate_f32')
        return out

    def aggregate_into(self, X, out_sum, out_mean, row_lo=0, row_hi=None, stream=None):
        """Low-level form: separate [rows, d] outputs sharing one row stride (either may be
        None).  Used by the sharded path to write mean rows straight into the next level's
        full input matrix."""
        import torch
        if X.dtype != torch.float32 or not X.is_cuda or X.dim() != 2 or X.stride(1) != 1:
            raise ValueError('X must be a 2-D float32 CUDA tensor with unit column stride')
        if X.shape[0] != self.n_cols:
            raise ValueError(f'X has {X.shape[0]} rows, graph addresses {self.n_cols}')
        d = X.shape[1]
        row_hi = self.n_rows if row_hi is None else row_hi
        rows = row_hi - row_lo
        ldo = None
        ptrs = []
        for t in (out_sum, out_mean):
            if t is None:
                ptrs.append(c_void_p(0))
                continue
            if (t.dtype != torch.float32 or not t.is_cuda or tuple(t.shape) != (rows, d)
                    or (d > 1 and t.stride(1) != 1)):
                raise ValueError(f'outputs must be float32 CUDA [{rows}, {d}] row-major')
            ld = t.stride(0) if rows > 1 else d
            if ldo is not None and ld != ldo:
                raise ValueError('out_sum and out_mean must share one row stride')
            ldo = ld
            ptrs.append(c_void_p(t.data_ptr() - row_lo * ld * 4))
        if ldo is None:
            raise ValueError('at least one output is required')
        check(load().gr_refex_aggregate_f32(
            self._handle, c_void_p(X.data_ptr()), X.stride(0) if X.shape[0] > 1 else d, d,
            row_lo, row_hi, ptrs[0], ptrs[1], ldo, _stream_ptr(stream)),
            'gr_refex_aggregate_f32')

    def aggregate_bcast(self, X, out_sum, replica_ptrs, ldo, row_offset, stream=None):
        """Fused level + exchange: mean rows of this handle's rows go to every replica.

        replica_ptrs: base addresses (ints) of the [n_cols, ldo] fp32 replicas of the next
        input matrix -- own and peer-mapped; row_offset = global row number of handle row 0.
        out_sum: optional local [n_rows, d] tensor with row stride ldo."""
        import torch
        if X.dtype != torch.float32 or not X.is_cuda or X.dim() != 2 or X.stride(1) != 1:
            raise ValueError('X must be a 2-D float32 CUDA tensor with unit column stride')
        if X.shape[0] != self.n_cols:
            raise ValueError(f'X has {X.shape[0]} rows, graph addresses {self.n_cols}')
        d = X.shape[1]
        sum_ptr = c_void_p(0)
        if out_sum is not None:
            if (out_sum.dtype != torch.float32 or tuple(out_sum.shape) != (self.n_rows, d)
                    or out_sum.stride(1) != 1 or (self.n_rows > 1 and out_sum.stride(0) != ldo)):
                raise ValueError(f'out_sum must be float32 [{self.n_rows}, {d}], row stride {ldo}')
            sum_ptr = c_void_p(out_sum.data_ptr())
        arr = (c_void_p * len(replica_ptrs))(*[c_void_p(int(p) + row_offset * ldo * 4)
                                                for p in replica_ptrs])
        check(load().gr_refex_aggregate_bcast_f32(
            self._handle, c_void_p(X.data_ptr()), X.stride(0) if X.shape[0] > 1 else d, d,
            0, self.n_rows, sum_ptr, arr, len(replica_ptrs), ldo, _stream_ptr(stream)),
            'gr_refex_aggregate_bcast_f32')

    def levels_host(self, X_host, levels, recurse_on='mean', out_host=None, stream=None):
        """Host-buffer entry point: H2D of X, `levels` recursion levels, D2H of every level."""
        import torch
        if X_host.dtype != torch.float32 or X_host.is_cuda or X_host.dim() != 2:
            raise ValueError('X_host must be a 2-D float32 CPU tensor')
        if X_host.shape[0] != self.n_cols or X_host.stride(1) != 1:
            raise ValueError('X_host must be [n_cols, d] with unit column stride')
        d = X_host.shape[1]
        if out_host is None:
            out_host = torch.empty((levels, self.n_rows, 2 * d), dtype=torch.float32,
                                   pin_memory=True)
        if tuple(out_host.shape) != (levels, self.n_rows, 2 * d) or not out_host.is_contiguous():
            raise ValueError('out_host must be contiguous [levels, n_rows, 2*d]')
        with torch.cuda.device(self.device):
            check(load().gr_refex_levels_host_f32(
                self._handle, c_void_p(X_host.data_ptr()), X_host.stride(0), d, levels,
                {'sum': 0, 'mean': 1}[recurse_on], c_void_p(out_host.data_ptr()),
                _stream_ptr(stream)), 'gr_refex_levels_host_f32')
        return out_host

    def levels_host_sharded(self, X_host, col0, d, levels, row_offset, replicas_even,
                            replicas_odd, flag_ptrs, rank, epoch, out_host, stream=None):
        """gr_refex_levels_host_sharded_f32.  X_host: pinned float32 [n_cols, >= col0 + d];
        out_host: [levels, 2, n_rows, d] (sum rows, then mean rows).  Returns the advanced barrier
        epoch."""
        import torch
        if X_host.dtype != torch.float32 or X_host.is_cuda or X_host.dim() != 2 \
                or X_host.shape[0] != self.n_cols or X_host.stride(1) != 1 \
                or X_host.shape[1] < col0 + d:
            raise ValueError('X_host must be a float32 CPU [n_cols, >= col0 + d] tensor')
        if tuple(out_host.shape) != (levels, 2, self.n_rows, d) or not out_host.is_contiguous() \
                or out_host.dtype != torch.float32:
            raise ValueError('out_host must be contiguous float32 [levels, 2, n_rows, d]')
        k = len(replicas_even)
        ev = (c_void_p * k)(*[c_void_p(int(p)) for p in replicas_even])
        od = (c_void_p * k)(*[c_void_p(int(p)) for p in replicas_odd])
        fl = (c_void_p * k)(*[c_void_p(int(p)) for p in flag_ptrs])
        ep = c_int64(int(epoch))
        with torch.cuda.device(self.device):
            check(load().gr_refex_levels_host_sharded_f32(
                self._handle, c_void_p(X_host.data_ptr() + col0 * 4), X_host.stride(0), d, levels,
                row_offset, ev, od, fl, k, rank, byref(ep), c_void_p(out_host.data_ptr()),
                _stream_ptr(stream)), 'gr_refex_levels_host_sharded_f32')
        return int(ep.value)

    def close(self):
        if getattr(self, '_handle', None):
            load().gr_csr_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PeerBuffer:
    """A device buffer other processes on the box can map (gr_peer_alloc).  `handle` is the
    64-byte token to ship to the peers; PeerBuffer.open(token) maps a peer's buffer."""

    def __init__(self, nbytes, device_index, _ptr=None, _owned=True):
        self.nbytes = int(nbytes)
        self.device_index = int(device_index)
        self._owned = _owned
        if _ptr is not None:
            self.ptr = _ptr
            self.handle = None
            return
        ptr = c_void_p()
        token = ctypes.create_string_buffer(64)
        check(load().gr_peer_alloc(byref(ptr), self.nbytes, self.device_index, token),
              'gr_peer_alloc')
        self.ptr = int(ptr.value)
        self.handle = bytes(token.raw)

    @classmethod
    def open(cls, token, nbytes, device_index):
        ptr = c_void_p()
        buf = ctypes.create_string_buffer(bytes(token), 64)
        check(load().gr_peer_open(byref(ptr), buf, int(device_index)), 'gr_peer_open')
        return cls(nbytes, device_index, _ptr=int(ptr.value), _owned=False)

    @property
    def __cuda_array_interface__(self):
        return {'shape': (self.nbytes,), 'typestr': '|u1', 'data': (self.ptr, False),
                'version': 3, 'strides': None}

    def tensor(self, shape, dtype):
        """torch view of the buffer (keeps this object alive through the array interface)."""
        import torch
        flat = torch.as_tensor(self, device=torch.device('cuda', self.device_index))
        count = 1
        for s in shape:
            count *= int(s)
        itemsize = torch.empty((), dtype=dtype).element_size()
        return flat[:count * itemsize].view(dtype).view(*shape)

    def close(self):
        if getattr(self, 'ptr', None):
            lib = load()
            if self._owned:
                lib.gr_peer_free(c_void_p(self.ptr), self.device_index)
            else:
                lib.gr_peer_close(c_void_p(self.ptr), self.device_index)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def peer_flag_words():
    return int(load().gr_peer_flag_words())


def peer_barrier(flag_ptrs, rank, epoch, timeout_s=20.0, stream=None):
    arr = (c_void_p * len(flag_ptrs))(*[c_void_p(int(p)) for p in flag_ptrs])
    check(load().gr_peer_barrier(arr, len(flag_ptrs), rank, epoch, timeout_s, _stream_ptr(stream)),
          'gr_peer_barrier')


def peer_barrier_timed_out(own_flag_ptr):
    v = c_int64()
    check(load().gr_peer_barrier_status(c_void_p(int(own_flag_ptr)), byref(v)),
          'gr_peer_barrier_status')
    return int(v.value)


class Pruner:
    """Owner of a gr_pruner_t*: vertical log binning of feature columns and the pairwise
    Chebyshev gaps of the binned columns, for matrices of `n_rows` rows on one device."""

    def __init__(self, n_rows, device):
        import torch
        device = torch.device(device)
        if device.type != 'cuda':
            raise NativeLibraryError('the pruning kernels run on CUDA devices only')
        self.device = device if device.index is not None else \
            torch.device('cuda', torch.cuda.current_device())
        self.n_rows = int(n_rows)
        handle = c_void_p()
        check(load().gr_pruner_create(byref(handle), self.n_rows, self.device.index),
              'gr_pruner_create')
        self._handle = handle

    def bin_columns(self, X, frac=0.5, out=None, stream=None):
        """X: [n_rows, d] float32 / float64 CUDA tensor, unit column stride.  Returns the binned
        columns as a COLUMN-major int32 tensor of shape [d, n_rows]."""
        import torch
        if not X.is_cuda or X.dim() != 2 or X.dtype not in (torch.float32, torch.float64):
            raise ValueError('X must be a 2-D float32/float64 CUDA tensor')
        if X.shape[0] != self.n_rows:
            raise ValueError(f'X has {X.shape[0]} rows, the pruner was created for {self.n_rows}')
        d = X.shape[1]
        if d > 1 and X.stride(1) != 1:
            raise ValueError('X must have unit column stride')
        if out is None:
            out = torch.empty((d, self.n_rows), dtype=torch.int32, device=X.device)
        if tuple(out.shape) != (d, self.n_rows) or out.dtype != torch.int32 \
                or not out.is_contiguous():
            raise ValueError(f'out must be a contiguous int32 [{d}, {self.n_rows}] tensor')
        fn = load().gr_prune_bin_f32 if X.dtype == torch.float32 else load().gr_prune_bin_f64
        ldx = X.stride(0) if X.shape[0] > 1 else max(d, 1)
        with torch.cuda.device(self.device):
            check(fn(self._handle, c_void_p(X.data_ptr()), ldx, d, float(frac),
                     c_void_p(out.data_ptr()), self.n_rows, _stream_ptr(stream)), 'gr_prune_bin')
        return out

    def pairwise_gaps(self, bins, stream=None):
        """bins: [d, n_rows] int32 (bin_columns output).  Returns int32 [d, d]: max over rows of
        |bins[i] - bins[j]|."""
        import torch
        if not bins.is_cuda or bins.dtype != torch.int32 or bins.dim() != 2 \
                or not bins.is_contiguous() or bins.shape[1] != self.n_rows:
            raise ValueError(f'bins must be a contiguous int32 CUDA [d, {self.n_rows}] tensor')
        d = bins.shape[0]
        gap = torch.empty((d, d), dtype=torch.int32, device=bins.device)
        with torch.cuda.device(self.device):
            check(load().gr_prune_pairwise_gap_i32(self._handle, c_void_p(bins.data_ptr()),
                                                   self.n_rows, d, c_void_p(gap.data_ptr()),
                                                   _stream_ptr(stream)),
                  'gr_prune_pairwise_gap_i32')
        return gap

    def close(self):
        if getattr(self, '_handle', None):
            load().gr_pruner_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def level0_features(rowptr, colidx, weights=None, directed=False, stream=None):
    """gr_level0_features_f64 on CUDA tensors.  Returns a dict of float64 [n] tensors:
    out_weight, in_weight (directed only), diag, internal, external."""
    import torch
    if not (rowptr.is_cuda and colidx.is_cuda):
        raise NativeLibraryError('the level-0 kernels run on CUDA tensors only (no CPU fallback)')
    if rowptr.dtype != torch.int64 or colidx.dtype != torch.int32:
        raise ValueError('rowptr must be int64 and colidx int32')
    rowptr, colidx = rowptr.contiguous(), colidx.contiguous()
    n, nnz = rowptr.numel() - 1, colidx.numel()
    dev = rowptr.device
    if weights is not None:
        weights = weights.to(dev, torch.float64).contiguous()
        if weights.numel() != nnz:
            raise ValueError('weights must have one entry per arc')
    out = {k: torch.empty(n, dtype=torch.float64, device=dev)
           for k in ('out_weight', 'diag', 'internal', 'external')}
    if directed:
        out['in_weight'] = torch.empty(n, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(load().gr_level0_features_f64(
            n, nnz, c_void_p(rowptr.data_ptr()), c_void_p(colidx.data_ptr() if nnz else 0),
            c_void_p(weights.data_ptr() if weights is not None else 0), int(bool(directed)),
            c_void_p(out['out_weight'].data_ptr()),
            c_void_p(out['in_weight'].data_ptr() if directed else 0),
            c_void_p(out['diag'].data_ptr()), c_void_p(out['internal'].data_ptr()),
            c_void_p(out['external'].data_ptr()), dev.index or 0, _stream_ptr(stream)),
            'gr_level0_features_f64')
    return out


class Quantizer:
    """Owner of a gr_quantizer_t*: the Lloyd-Max quantiser of graphrole/roles/factor.py:29-49 on
    the device.  bind() a float32 / float64 CUDA matrix once (sort + prefix sums), then encode()
    it with as many bin counts as needed."""

    KMEANS_SEED = 1          # factor.py:41 KMeans(random_state=1)
    KMEANS_TOL = 1e-4        # sklearn defaults the reference relies on
    KMEANS_MAX_ITER = 300

    def __init__(self, capacity, device):
        import torch
        device = torch.device(device)
        if device.type != 'cuda':
            raise NativeLibraryError('the quantiser runs on CUDA devices only (no CPU fallback)')
        self.device = device if device.index is not None else \
            torch.device('cuda', torch.cuda.current_device())
        self.capacity = int(capacity)
        handle = c_void_p()
        check(load().gr_quantizer_create(byref(handle), self.capacity, self.device.index),
              'gr_quantizer_create')
        self._handle = handle
        self._bound = None

    def bind(self, X, stream=None):
        import torch
        if not X.is_cuda or X.dim() != 2 or X.dtype not in (torch.float32, torch.float64):
            raise ValueError('X must be a 2-D float32/float64 CUDA tensor')
        if X.shape[1] > 1 and X.stride(1) != 1:
            raise ValueError('X must have unit column stride')
        fn = load().gr_quantizer_bind_f32 if X.dtype == torch.float32 else \
            load().gr_quantizer_bind_f64
        ld = X.stride(0) if X.shape[0] > 1 else X.shape[1]
        with torch.cuda.device(self.device):
            check(fn(self._handle, c_void_p(X.data_ptr()), X.shape[0], X.shape[1], ld,
                     _stream_ptr(stream)), 'gr_quantizer_bind')
        self._bound = X          # keep the matrix alive: the handle reads it again
        return self

    def encode(self, n_bins, out=None, seed=KMEANS_SEED, tol=KMEANS_TOL,
               max_iter=KMEANS_MAX_ITER, stream=None):
        """Returns (encoded matrix, info) with info = dict(centers, n_iter, n_distinct)."""
        import numpy as np
        import torch
        X = self._bound
        if X is None:
            raise ValueError('bind() a matrix first')
        if n_bins > X.numel():
            raise TooManyBinsError(f'n_samples={X.numel()} should be >= n_clusters={n_bins}.')
        if out is None:
            out = torch.empty(X.shape, dtype=X.dtype, device=X.device)
        if out.shape != X.shape or out.dtype != X.dtype or not out.is_contiguous():
            raise ValueError('out must be a contiguous tensor of the bound matrix\'s shape/dtype')
        fn = load().gr_quantizer_encode_f32 if X.dtype == torch.float32 else \
            load().gr_quantizer_encode_f64
        centers = (c_double * int(n_bins))()
        n_iter, n_distinct = c_int32(0), c_int64(0)
        with torch.cuda.device(self.device):
            check(fn(self._handle, int(n_bins), int(seed), int(max_iter), float(tol),
                     c_void_p(out.data_ptr()), X.shape[1], centers, byref(n_iter),
                     byref(n_distinct), _stream_ptr(stream)), 'gr_quantizer_encode')
        return out, {'centers': np.array(centers[:]), 'n_iter': int(n_iter.value),
                     'n_distinct': int(n_distinct.value)}

    def count_distinct(self, stream=None):
        v = c_int64(0)
        check(load().gr_quantizer_count_distinct(self._handle, byref(v), _stream_ptr(stream)),
              'gr_quantizer_count_distinct')
        return int(v.value)

    def close(self):
        if getattr(self, '_handle', None):
            load().gr_quantizer_destroy(self._handle)
            self._handle = None
            self._bound = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _matrix_args(t, name):
    import torch
    if not t.is_cuda or t.dim() != 2 or t.dtype not in (torch.float32, torch.float64):
        raise ValueError(f'{name} must be a 2-D float32/float64 CUDA tensor')
    if t.shape[1] > 1 and t.stride(1) != 1:
        raise ValueError(f'{name} must have unit column stride')
    return c_void_p(t.data_ptr()), (t.stride(0) if t.shape[0] > 1 else t.shape[1])


def mdl_error_cost(V, G, F, stream=None):
    """sum over V != 0 of v log(v / a) - v + a with a = G @ F (description_length.py:19, :44-61).
    V [n, f], G [n, r], F [r, f]: CUDA tensors of one dtype (float32 or float64)."""
    import torch
    if not (V.dtype == G.dtype == F.dtype):
        raise ValueError('V, G and F must share one dtype')
    n, f = V.shape
    r = G.shape[1]
    if G.shape != (n, r) or F.shape != (r, f):
        raise ValueError('shapes must be V [n, f], G [n, r], F [r, f]')
    (pv, ldv), (pg, ldg), (pf, ldf) = (_matrix_args(t, k) for t, k in
                                       ((V, 'V'), (G, 'G'), (F, 'F')))
    fn = load().gr_mdl_error_cost_f32 if V.dtype == torch.float32 else \
        load().gr_mdl_error_cost_f64
    cost = c_double(0.0)
    check(fn(pv, n, f, ldv, pg, ldg, pf, ldf, r, byref(cost), V.device.index or 0,
             _stream_ptr(stream)), 'gr_mdl_error_cost')
    return float(cost.value)


def mdl_kl(V, V_approx, stream=None):
    """The same cost from an explicit approximation (get_error_cost(V, V_approx)); float64."""
    import torch
    if V.dtype != torch.float64 or V_approx.dtype != torch.float64 or V.shape != V_approx.shape:
        raise ValueError('V and V_approx must be float64 CUDA tensors of one shape')
    (pv, ldv), (pa, lda) = _matrix_args(V, 'V'), _matrix_args(V_approx, 'V_approx')
    cost = c_double(0.0)
    check(load().gr_mdl_kl_f64(pv, pa, V.shape[0], V.shape[1], ldv, lda, byref(cost),
                               V.device.index or 0, _stream_ptr(stream)), 'gr_mdl_kl_f64')
    return float(cost.value)


def roles(W, want_argmax=True, want_percentage=True, stream=None):
    """(argmax int32 [n] or None, row-normalised W or None) of a node-role factor on the device."""
    import torch
    pw, ldw = _matrix_args(W, 'W')
    n, r = W.shape
    arg = torch.empty(n, dtype=torch.int32, device=W.device) if want_argmax else None
    pct = torch.empty((n, r), dtype=W.dtype, device=W.device) if want_percentage else None
    fn = load().gr_roles_f32 if W.dtype == torch.float32 else load().gr_roles_f64
    check(fn(pw, n, r, ldw, c_void_p(arg.data_ptr() if arg is not None else 0),
             c_void_p(pct.data_ptr() if pct is not None else 0), r, W.device.index or 0,
             _stream_ptr(stream)), 'gr_roles')
    return arg, pct


def numpy_random_sample(seed, count):
    """RandomState(seed).random_sample(count) as the library generates it (host only)."""
    buf = (c_double * int(count))()
    check(load().gr_numpy_random_sample(int(seed), int(count), buf), 'gr_numpy_random_sample')
    return list(buf)


def numpy_choice_uniform(seed, n):
    """RandomState(seed).choice(n, p=np.full(n, 1 / n)) as the library computes it (host only)."""
    v = c_int64(0)
    check(load().gr_numpy_choice_uniform(int(seed), int(n), byref(v)), 'gr_numpy_choice_uniform')
    return int(v.value)
