#!/usr/bin/env python
"""Development check of the tcgen05 NMF kernel against the FFMA kernels (GPU box only)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from graphrole_b200 import _native
from graphrole_b200.roles import factor

if os.environ.get('GR_EXP_LIB'):     # a kernel variant built by tools/build_variant.sh
    _native.LIB_PATH = os.path.abspath(os.environ['GR_EXP_LIB'])

shapes = [(64, 128, 32), (64, 128, 8), (128, 128, 32), (200, 128, 5), (1000, 512, 32),
          (64 * 148 * 3 + 17, 512, 32), (5000, 96, 12), (4096, 768, 16), (300, 64, 4)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]
iters = int(os.environ.get('ITERS', '1'))
for n, f, r in shapes:
    g = torch.Generator(device='cuda').manual_seed(0)
    X = torch.rand(n, f, device='cuda', generator=g)
    W0 = torch.rand(n, r, device='cuda', generator=g) + 0.1
    H0 = torch.rand(r, f, device='cuda', generator=g) + 0.1
    Wf, Hf, _, ef = factor.nmf_mu(X, W0, H0, max_iter=iters, tol=0, use_tf32=False)
    pf = factor.last_path
    t0 = time.time()
    Wt, Ht, _, et = factor.nmf_mu(X, W0, H0, max_iter=iters, tol=0, use_tf32=True)
    torch.cuda.synchronize()
    pt = factor.last_path
    dw = float((Wt - Wf).abs().max() / Wf.abs().max())
    dh = float((Ht - Hf).abs().max() / Hf.abs().max())
    print(f'n={n} f={f} r={r} paths={pf}/{pt} relW={dw:.3e} relH={dh:.3e} err={ef:.6f}/{et:.6f} '
          f'({time.time() - t0:.2f}s)', flush=True)
    if dw > 5e-2 or dh > 5e-2:
        bad = (Wt - Wf).abs().argmax()
        print('   worst W at row', int(bad) // r, 'col', int(bad) % r,
              'tc', float(Wt.flatten()[bad]), 'ffma', float(Wf.flatten()[bad]))
        bad = (Ht - Hf).abs().argmax()
        print('   worst H at role', int(bad) // f, 'col', int(bad) % f,
              'tc', float(Ht.flatten()[bad]), 'ffma', float(Hf.flatten()[bad]))
