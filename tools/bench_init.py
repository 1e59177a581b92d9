"""NNDSVDa start at C5 size (X 10M x 512): seconds per call, CholeskyQR2 sketches (default) vs
Householder QR (torch.linalg.qr, what round 1 ran), and where the time goes."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from graphrole_b200.roles import factor


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, out


def main():
    n, f = int(os.environ.get('N', 10_000_000)), 512
    dev = torch.device('cuda', 0)
    gen = torch.Generator(device=dev).manual_seed(0)
    X = torch.rand(n, 8, device=dev, generator=gen).square_() @ torch.rand(8, f, device=dev, generator=gen)
    X += 0.05 * torch.rand(n, f, device=dev, generator=gen)
    chol = factor.orthonormal_basis
    for r in (2, 8):
        np.random.seed(0)
        timed(lambda: factor.nndsvda_init(X, r))
        np.random.seed(0)
        t_c, (Wc, Hc) = timed(lambda: factor.nndsvda_init(X, r))
        factor.orthonormal_basis = lambda Y: torch.linalg.qr(Y)[0]
        np.random.seed(0)
        t_h, (Wh, Hh) = timed(lambda: factor.nndsvda_init(X, r))
        factor.orthonormal_basis = chol
        k = r + 10
        Q = torch.randn(f, k, device=dev)
        t_aq, Y = timed(lambda: X @ Q)
        t_atq, _ = timed(lambda: X.T @ Y)
        t_cq, _ = timed(lambda: chol(Y))
        t_hq, _ = timed(lambda: torch.linalg.qr(Y))
        print(json.dumps({'r': r, 'init_s_cholesky_qr2': round(t_c, 4), 'init_s_householder': round(t_h, 4),
                          # NNDSVDa is discontinuous at 0 (entries below 1e-6 become mean(X)): count the
                          # entries whose sign rounding noise flipped instead of a max difference
                          'frac_W_entries_differing': float(
                              ((Wc - Wh).abs() > 1e-4 * Wh.abs().max()).float().mean()),
                          'max_rel_diff_H': float((Hc - Hh).abs().max() / Hh.abs().max()),
                          'A@Q_s': round(t_aq, 4), 'A.T@Y_s': round(t_atq, 4),
                          'basis_cholesky_qr2_s': round(t_cq, 4), 'basis_householder_s': round(t_hq, 4)}),
              flush=True)


if __name__ == '__main__':
    main()
