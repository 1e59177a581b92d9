"""Lloyd-Max quantiser (gr_quantizer_*) on a C5-sized node-role factor (10 M x 8 = 80 M entries):
ms per bind and per encode for 2 .. 256 bins."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from graphrole_b200 import _native


def main():
    n, r = int(os.environ.get('N', 10_000_000)), int(os.environ.get('R', 8))
    dev = torch.device('cuda', 0)
    gen = torch.Generator(device=dev).manual_seed(1)
    W = torch.rand(n, r, device=dev, generator=gen) ** 2

    def wall(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3, out

    q = _native.Quantizer(W.numel(), dev)
    ms, _ = wall(lambda: q.bind(W))
    print(json.dumps({'what': 'bind (centre, sort, prefix sums)', 'entries': W.numel(), 'ms': round(ms, 2)}))
    for bits in range(1, 9):
        ms, (G, info) = wall(lambda: q.encode(2 ** bits))
        print(json.dumps({'bins': 2 ** bits, 'ms': round(ms, 2), 'lloyd_iterations': info['n_iter'],
                          'distinct': info['n_distinct'],
                          'checksum': float(G.double().sum())}), flush=True)
    q.close()


if __name__ == '__main__':
    main()
