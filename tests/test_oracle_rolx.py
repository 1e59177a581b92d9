"""The RolX-epilogue oracle (oracle/rolx_oracle.py) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py::rolx_cases -> graphrole.roles.factor.encode,
graphrole.roles.description_length), and the NumPy RandomState stream the CUDA quantiser's
seeding draws from (host-only entry points of the library: no GPU needed)."""
import os

import numpy as np
import pytest

from oracle import rolx_oracle as oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def rolx_cases():
    return np.load(os.path.join(GOLDEN, 'rolx_cases.npz'))


def test_encode_matches_the_reference(rolx_cases):
    z = rolx_cases
    for name in z['encode_names']:
        X = z[f'{name}__X']
        for k in z[f'{name}__bins']:
            got = oracle.encode(X, int(k))
            np.testing.assert_allclose(got, z[f'{name}__enc{k}'], rtol=1e-12, atol=1e-12,
                                       err_msg=f'{name} bins={k}')


def test_encode_refuses_more_bins_than_entries():
    with pytest.raises(ValueError, match='should be >= n_clusters'):
        oracle.encode(np.random.rand(2, 2), 16)


def test_description_length_costs_match_the_reference(rolx_cases):
    z = rolx_cases
    V = z['dl__V']
    for name in z['dl_names']:
        G, F = z[f'dl__{name}__G'], z[f'dl__{name}__F']
        enc, err = z[f'dl__{name}__costs']
        assert oracle.encoding_cost(G, F) == enc
        assert oracle.error_cost(V, G @ F) == pytest.approx(err, rel=1e-12)


def test_grid_row_matches_the_reference(rolx_cases):
    z = rolx_cases
    V, G, F, want = z['grid__V'], z['grid__G'], z['grid__F'], z['grid__costs']
    for bits in range(1, 9):
        try:
            Ge, Fe = oracle.encode(G, 2 ** bits), oracle.encode(F, 2 ** bits)
        except ValueError:
            assert np.isnan(want[bits]).all()
            continue
        assert oracle.encoding_cost(Ge, Fe) == want[bits, 0]
        assert oracle.error_cost(V, Ge @ Fe) == pytest.approx(want[bits, 1], rel=1e-10)


def test_rescale_and_select():
    costs = np.array([[np.nan, np.nan], [3.0, 4.0], [np.nan, 2.0]])
    got = oracle.rescale_costs(costs)
    np.testing.assert_allclose(got[1], [0.6, 0.8])
    assert oracle.select_model(costs, costs) == (1, 0)


def test_library_random_stream_is_numpys():
    """mt19937.h: RandomState(seed).random_sample and .choice(n, p=uniform)."""
    from graphrole_b200 import _native
    for seed in (0, 1, 7, 2 ** 31 + 5):
        np.testing.assert_array_equal(np.array(_native.numpy_random_sample(seed, 1500)),
                                      np.random.RandomState(seed).random_sample(1500))
    for n in (1, 2, 3, 10, 999, 1000, 65536, 600_001):
        for seed in range(12):
            w = np.ones(n)
            assert _native.numpy_choice_uniform(seed, n) == \
                np.random.RandomState(seed).choice(n, p=w / w.sum())


def test_numpy_pairwise_sum_restatement_is_bit_exact():
    """The order in which the quantiser sums its data mean on the GPU is NumPy's; this pins the
    restatement of that order to the installed NumPy (np.add.reduce, ndarray.mean on the
    [n, 1] layout KMeans sees) bit for bit, ragged sizes included."""
    rng = np.random.RandomState(0)
    for n in (1, 5, 8, 9, 100, 128, 129, 1000, 4097, 20000, 80000, 123457, 300001):
        a = rng.rand(n) ** 2
        want = np.add.reduce(a)
        assert oracle.numpy_pairwise_sum(a) == want, n
        assert a.reshape(-1, 1).mean(axis=0)[0] == want / n, n


def _seeds_by_the_range_rule(x, k, rs):
    """NumPy model of the seeding in csrc/rolx_epilogue.cu: a candidate is evaluated, and the
    chosen centre applied, only on the SORTED range strictly between the chosen centres next to
    it (none when its value is a centre already, everything when a neighbour is within the
    rounding-noise guard); the sampling still sums the distances in the original order."""
    import bisect
    n = x.size
    order = np.argsort(x, kind='stable')
    xs = x[order]
    guard = 1e-14 * max(xs[0] ** 2, xs[-1] ** 2)
    trials = 2 + int(np.log(k))

    def sqd(c, v):
        return np.maximum((-2.0 * (c * v) + c * c) + v * v, 0)

    idx = np.full(k, -1, dtype=np.int64)
    idx[0] = rs.choice(n, p=np.ones(n) / n)
    closest = sqd(x[idx[0]], x)
    centres = [x[idx[0]]]
    for c in range(1, k):
        pot = closest.sum()
        cand = np.searchsorted(np.cumsum(closest), rs.uniform(size=trials) * pot)
        np.clip(cand, None, n - 1, out=cand)
        pots, ranges = [], []
        for j in cand:
            cv = x[j]
            pos = bisect.bisect_left(centres, cv)
            if pos < len(centres) and centres[pos] == cv:
                ranges.append((0, 0))
                pots.append(pot)
                continue
            c_lo = centres[pos - 1] if pos > 0 else None
            c_hi = centres[pos] if pos < len(centres) else None
            everything = (c_lo is not None and (cv - c_lo) ** 2 <= guard) or \
                         (c_hi is not None and (c_hi - cv) ** 2 <= guard)
            lo = 0 if (everything or c_lo is None) else int(np.searchsorted(xs, c_lo, 'right'))
            hi = n if (everything or c_hi is None) else int(np.searchsorted(xs, c_hi, 'left'))
            ranges.append((lo, hi))
            cl = closest[order[lo:hi]]
            pots.append(pot - np.sum(cl - np.minimum(cl, sqd(cv, xs[lo:hi]))))
        best = int(np.argmin(pots))
        lo, hi = ranges[best]
        cv = x[cand[best]]
        sel = order[lo:hi]
        closest[sel] = np.minimum(closest[sel], sqd(cv, xs[lo:hi]))
        idx[c] = cand[best]
        bisect.insort(centres, cv)
    return idx


@pytest.mark.parametrize('kind', ['uniform', 'integers', 'skewed', 'near_duplicates',
                                  'two_decimals', 'zero_mass'])
def test_range_rule_of_the_incremental_seeding_reproduces_kmeans_plusplus(kind):
    """The rule the CUDA seeding relies on -- a new centre only changes distances between the
    chosen centres next to it -- gives the centres of the full evaluation (the restatement of
    sklearn's _kmeans_plusplus above), sklearn's rounded distance formula included.  Compared by
    value: among candidates of EQUAL value the full evaluation's argmin is decided by BLAS
    rounding noise in NumPy itself.  Cases with fewer distinct values than centres are the
    degenerate regime the library handles separately."""
    rng = np.random.RandomState(len(kind))
    for n in (50, 500, 4000):
        v = {'uniform': lambda: rng.rand(n),
             'integers': lambda: rng.randint(0, 40, n).astype(float),
             'skewed': lambda: rng.rand(n) ** 4,
             'near_duplicates': lambda: np.concatenate([rng.rand(n // 2) * 1e-9 + 0.5,
                                                        rng.rand(n - n // 2)]),
             'two_decimals': lambda: np.round(rng.rand(n), 2),
             'zero_mass': lambda: np.abs(rng.randn(n)) * (rng.rand(n) < 0.3)}[kind]()
        x = v - v.mean()
        for k in (2, 4, 16, 32):
            if k > min(n, np.unique(v).size):
                continue
            want = oracle.kmeans_plusplus_1d(x, k, np.random.RandomState(1 + n))
            got = _seeds_by_the_range_rule(x, k, np.random.RandomState(1 + n))
            np.testing.assert_array_equal(x[got], x[want], err_msg=f'{kind} n={n} k={k}')
