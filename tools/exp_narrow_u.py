"""Round-2 experiment: one vs two gathers per lane in flight for narrow rows (column-group shards),
C3 graph, a quarter of the rows (what one rank of a 2 x 4 grid does) and all rows."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from graphrole_b200 import _native  # noqa: E402
from graphrole_b200.graph.generators import barabasi_albert_csr  # noqa: E402
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
    ap.add_argument('--widths', default='32,16,8')
    ap.add_argument('--one', action='store_true', help='a single d = 32 launch (for ncu)')
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    g = barabasi_albert_csr(10_000_000, 20, seed=0, device=dev)
    X = torch.rand(g.n, 64, device=dev)
    ranges = nnz_balanced_ranges(g.rowptr, 4)
    for dl in [int(w) for w in args.widths.split(',')]:
        Xl = X[:, :dl]                       # strided view, like a column group reads X0
        Xc = Xl.contiguous()
        h = _native.CsrHandle(g.rowptr, g.colidx, validate=False)
        h.tune_hot_rows(dl * 4)
        out = torch.empty((g.n, 2 * dl), device=dev)
        if args.one:
            h.aggregate(Xc, out=out)
            torch.cuda.synchronize()
            return
        for u in ('1', '2'):
            os.environ['GR_REFEX_U'] = u
            row = {'d_local': dl, 'U': int(u),
                   'all_rows_ms': round(timed(lambda: h.aggregate(Xc, out=out)), 3),
                   'all_rows_strided_ms': round(timed(lambda: h.aggregate(Xl, out=out)), 3)}
            for which in (0, 3):
                lo, hi = ranges[which]
                o = out[:hi - lo]
                row[f'quarter{which}_ms'] = round(
                    timed(lambda: h.aggregate(Xc, out=o, row_lo=lo, row_hi=hi)), 3)
            print(json.dumps(row), flush=True)
        h.close()


if __name__ == '__main__':
    main()
