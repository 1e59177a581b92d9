"""Round-2 experiment: what a column-group shard of path A costs on one GPU.

Aggregation is column-separable (S[:, c] = A X[:, c]): a rank that owns d/C of the columns needs
no exchange between levels at all.  Its work is the whole graph at d/C columns, so its level time
is measurable on ONE GPU.  Sweeps d_local, the hot-row budget, rows per warp and row ranges
(for C x R hybrids) on the C3 graph."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from graphrole_b200 import _native  # noqa: E402
from graphrole_b200.graph.generators import barabasi_albert_csr, erdos_renyi_csr  # noqa: E402
from graphrole_b200.shard import nnz_balanced_ranges  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=10_000_000)
    ap.add_argument('--m', type=int, default=20)
    ap.add_argument('--family', default='ba')
    ap.add_argument('--widths', default='8,16,32,64')
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    g = (barabasi_albert_csr(args.n, args.m, seed=0, device=dev) if args.family == 'ba'
         else erdos_renyi_csr(args.n, args.m, seed=0, device=dev))
    n = g.n
    X = torch.rand(n, 64, device=dev)
    rows = []
    for dl in [int(w) for w in args.widths.split(',')]:
        Xl = X[:, :dl].contiguous()
        out = torch.empty((n, 2 * dl), device=dev)
        for hot_bytes in sorted({256, dl * 4}):
            for budget in ('48', '24', '72'):
                if budget != '48' and hot_bytes == 256 and dl != 64:
                    continue
                os.environ['GR_CSR_HOT_BUDGET_MB'] = budget
                h = _native.CsrHandle(g.rowptr, g.colidx, validate=False)
                h._handle and h.tune_hot_rows(hot_bytes)
                for rpw in (4, 8, 16, 31):
                    os.environ['GR_REFEX_ROWS_PER_WARP'] = str(rpw)
                    ms = timed(lambda: h.aggregate(Xl, out=out))
                    rows.append({'d_local': dl, 'hot_row_bytes': hot_bytes, 'budget_mb': budget,
                                 'hot_rows': h.info()['n_hot_rows'], 'rows_per_warp': rpw,
                                 'rows': 'all', 'ms': round(ms, 3)})
                    print(json.dumps(rows[-1]), flush=True)
                h.close()
        os.environ['GR_CSR_HOT_BUDGET_MB'] = '48'
        os.environ['GR_REFEX_ROWS_PER_WARP'] = '16'
        # C x R hybrids: first / last arc-balanced row range at this width
        h = _native.CsrHandle(g.rowptr, g.colidx, validate=False)
        h.tune_hot_rows(dl * 4)
        for R in (2, 4):
            ranges = nnz_balanced_ranges(g.rowptr, R)
            for which in (0, R - 1):
                lo, hi = ranges[which]
                o = out[:hi - lo]
                for rpw in (4, 16):
                    os.environ['GR_REFEX_ROWS_PER_WARP'] = str(rpw)
                    ms = timed(lambda: h.aggregate(Xl, out=o, row_lo=lo, row_hi=hi))
                    rows.append({'d_local': dl, 'R': R, 'range': which, 'rows': hi - lo,
                                 'rows_per_warp': rpw, 'ms': round(ms, 3)})
                    print(json.dumps(rows[-1]), flush=True)
        h.close()
        del Xl, out
    os.environ.pop('GR_REFEX_ROWS_PER_WARP', None)


if __name__ == '__main__':
    main()
