"""Parity of the RolX epilogue on the GPU (csrc/rolx_epilogue.cu through the C-ABI) against the
pinned oracle, the reference's golden vectors and scikit-learn itself.

Stated tolerances: quantised values within 1e-9 of the reference's (identical labels, centres
equal up to fp64 summation order); description-length costs within 1e-9 relative in fp64
storage and 1e-5 relative in fp32 storage (the grid's device-resident form)."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

from graphrole_b200 import RoleExtractor, _native
from graphrole_b200.roles import description_length as dl
from graphrole_b200.roles import factor
from graphrole_b200.roles.extract import DeviceModelGrid
from oracle import rolx_oracle as oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def rolx_cases():
    return np.load(os.path.join(GOLDEN, 'rolx_cases.npz'))


def dev(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device='cuda:0')


def test_encode_matches_the_reference_golden(rolx_cases):
    z = rolx_cases
    before = _native.launch_count()
    for name in z['encode_names']:
        X = z[f'{name}__X']
        for k in z[f'{name}__bins']:
            got = factor.encode(X, int(k))
            assert got.shape == X.shape
            np.testing.assert_allclose(got, z[f'{name}__enc{k}'], rtol=1e-9, atol=1e-9,
                                       err_msg=f'{name} bins={k}')
    assert _native.launch_count() > before


def test_one_bind_serves_every_bin_count_and_reports_sklearn_centres(rolx_cases):
    X = rolx_cases['exp400x5__X']
    q = _native.Quantizer(X.size, 'cuda:0').bind(dev(X))
    for k in (2, 7, 16, 100, 256):
        out, info = q.encode(k)
        labels, centres, n_iter = oracle.kmeans_1d(X.ravel(), k)
        np.testing.assert_allclose(info['centers'], centres, rtol=1e-9, atol=1e-12)
        assert info['n_iter'] == n_iter
        np.testing.assert_allclose(out.cpu().numpy().ravel(), centres[labels], rtol=1e-9)
        assert info['n_distinct'] == np.unique(centres[labels]).size
    q.close()


@pytest.mark.parametrize('n,k', [(200_000, 16), (1_000_000, 64), (3_000_000, 256)])
def test_encode_equals_sklearn_kmeans_beyond_fixture_sizes(n, k):
    """scikit-learn itself (third-party owner of the arithmetic) on sizes it still finishes in
    seconds: identical labels, centres to 1e-9."""
    cluster = pytest.importorskip('sklearn.cluster')
    rng = np.random.RandomState(k)
    X = (rng.gamma(2.0, size=n) * rng.rand(n)).reshape(-1, 4)
    km = cluster.KMeans(n_clusters=k, random_state=1).fit(X.reshape(-1, 1))
    want = km.cluster_centers_.ravel()[km.labels_].reshape(X.shape)
    got = factor.encode(X, k)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize('kind', ['near_duplicates', 'ties', 'heavy_zero_mass', 'grid'])
def test_seeding_ranges_on_awkward_data(kind):
    """The k-means++ seeding evaluates a candidate only on the sorted range between the chosen
    centres next to it (csrc/rolx_epilogue.cu).  Data that stress that rule: values within
    rounding noise of each other (a candidate next to a centre closer than the guard takes the
    whole range), ties (a candidate equal to a chosen centre takes none) and a large mass of
    equal values -- against the NumPy restatement of scikit-learn's KMeans, 32 and 64 bins (the
    range form starts with the 20th centre).
    'grid': values rounded to three decimals.  There the k-means++ seeds are grid values, their
    midpoints are data values, and scikit-learn assigns those exactly-equidistant points by the
    last bit of the data mean -- with a mean from an ordinary parallel reduction the centres were
    1e-3 off after the first Lloyd step (same seeds).  The library therefore sums the mean in
    NumPy's pairwise order (numpy_leaf_sums_kernel + the tree on the host), which makes this
    case equal too."""
    rng = np.random.RandomState(11)
    n = 20_000
    if kind == 'near_duplicates':
        X = np.concatenate([0.5 + 1e-9 * rng.rand(n // 2), rng.rand(n - n // 2)])
    elif kind == 'ties':
        X = (rng.rand(1000) ** 2)[rng.randint(0, 1000, size=n)]
    elif kind == 'grid':
        X = np.round(rng.rand(n) ** 2, 3)
    else:
        X = np.abs(rng.randn(n)) * (rng.rand(n) < 0.3)
    X = rng.permutation(X).reshape(-1, 4)
    for k in ((8, 32, 64) if kind == 'grid' else (32, 64)):
        np.testing.assert_allclose(factor.encode(X, k), oracle.encode(X, k), rtol=1e-9, atol=1e-12)


def test_float32_storage_path_matches_oracle():
    """The device-resident grid quantises the fp32 NMF factors in place."""
    rng = np.random.RandomState(3)
    X32 = (rng.rand(5000, 8) ** 3).astype(np.float32)
    q = _native.Quantizer(X32.size, 'cuda:0').bind(dev(X32, torch.float32))
    for k in (2, 16, 256):
        out, info = q.encode(k)
        want = oracle.encode(X32.astype(np.float64), k)
        assert out.dtype == torch.float32
        np.testing.assert_allclose(out.cpu().numpy(), want, rtol=2e-7)
        assert info['n_distinct'] == np.unique(want.astype(np.float32)).size
    q.close()


def test_strided_input_and_vector():
    rng = np.random.RandomState(4)
    wide = dev(rng.rand(300, 10))
    view = wide[:, 2:7]
    q = _native.Quantizer(view.numel(), 'cuda:0').bind(view)
    out, _ = q.encode(8)
    np.testing.assert_allclose(out.cpu().numpy(), oracle.encode(view.cpu().numpy(), 8), rtol=1e-9)
    q.close()
    v = rng.rand(37)
    np.testing.assert_allclose(factor.encode(v, 5), oracle.encode(v, 5), rtol=1e-9)


def test_fewer_distinct_values_than_bins_is_exact():
    """Below n_bins distinct values scikit-learn's own result depends on np.argpartition's order of
    equal keys (and can leave the data range); here every distinct value is a level."""
    rng = np.random.RandomState(5)
    X = rng.randint(0, 6, size=(100, 3)).astype(np.float64)
    got = factor.encode(X, 256)
    np.testing.assert_array_equal(got, X)
    q = _native.Quantizer(X.size, 'cuda:0').bind(dev(X))
    _, info = q.encode(64)
    assert info['n_distinct'] == 6 and info['n_iter'] == 0
    assert q.count_distinct() == 6
    q.close()
    const = np.full((4, 4), 2.5)
    np.testing.assert_array_equal(factor.encode(const, 2), const)


def test_encode_bins_and_value_error():
    """tests/test_roles/test_factor.py:27-31 and the ValueError the grid relies on
    (roles/extract.py:127-129)."""
    X = np.random.RandomState(0).rand(20, 30)
    for n_bins in range(1, 8):
        assert len(np.unique(factor.encode(X, n_bins))) <= n_bins
    with pytest.raises(ValueError, match='should be >= n_clusters'):
        factor.encode(np.random.rand(2, 2), 16)
    q = _native.Quantizer(16, 'cuda:0').bind(dev(np.random.rand(2, 2)))
    with pytest.raises(_native.TooManyBinsError):
        q.encode(16)
    with pytest.raises(ValueError):
        q.encode(2048)          # library limit, a different error class than "too many bins"
    q.close()


def test_description_length_costs(rolx_cases, roles_cases):
    G = np.array([[0.0, 1.0], [1.0, 2.0], [3.0, 0.0]])
    F = np.array([[1.0, 0.0, 2.0], [0.0, 1.0, 1.0]])
    assert dl.get_encoding_cost((G, F)) == 2 * (6 + 6)      # 4 distinct values -> 2 bits
    V = np.random.RandomState(0).rand(5, 4)
    assert dl.get_error_cost(V, V) == pytest.approx(0.0, abs=1e-12)
    assert dl.get_error_cost(V, V * 1.5) == pytest.approx(oracle.error_cost(V, V * 1.5), rel=1e-12)
    z = rolx_cases
    for name in z['dl_names']:
        model = (z[f'dl__{name}__G'], z[f'dl__{name}__F'])
        enc, err = dl.get_description_length_costs(z['dl__V'], model)
        assert enc == z[f'dl__{name}__costs'][0]
        assert err == pytest.approx(z[f'dl__{name}__costs'][1], rel=1e-9)
    X = np.array(roles_cases['X'])
    for row in roles_cases['encoded']:
        model = (np.array(row['G']), np.array(row['F']))
        enc, err = dl.get_description_length_costs(pd.DataFrame(X), model)
        assert enc == pytest.approx(row['encoding_cost'])
        assert err == pytest.approx(row['error_cost'], rel=1e-9)


def test_error_cost_float32_storage_and_zero_mask():
    rng = np.random.RandomState(6)
    V = rng.rand(3000, 40)
    V[rng.rand(*V.shape) < 0.2] = 0.0
    G, F = rng.rand(3000, 7) + 0.05, rng.rand(7, 40) + 0.05
    want = oracle.error_cost(V, G @ F)
    assert _native.mdl_error_cost(dev(V), dev(G), dev(F)) == pytest.approx(want, rel=1e-9)
    V32, G32, F32 = (a.astype(np.float32) for a in (V, G, F))
    want32 = oracle.error_cost(V32, G32.astype(np.float64) @ F32.astype(np.float64))
    got32 = _native.mdl_error_cost(*(dev(a, torch.float32) for a in (V32, G32, F32)))
    assert got32 == pytest.approx(want32, rel=1e-9)          # fp64 arithmetic on fp32 storage
    assert got32 == pytest.approx(want, rel=1e-5)


def test_grid_row_matches_the_reference(rolx_cases):
    """One row of the model-selection grid with the factors held fixed (roles/extract.py:121-133):
    quantiser + both costs for bits 1..8, NaN where the reference skips the cell."""
    z = rolx_cases
    V, G, F, want = z['grid__V'], z['grid__G'], z['grid__F'], z['grid__costs']
    qG = _native.Quantizer(G.size, 'cuda:0').bind(dev(G))
    qF = _native.Quantizer(F.size, 'cuda:0').bind(dev(F))
    for bits in range(1, 9):
        try:
            Ge, ig = qG.encode(2 ** bits)
            Fe, i_f = qF.encode(2 ** bits)
        except _native.TooManyBinsError:
            assert np.isnan(want[bits]).all()
            continue
        enc = dl.encoding_cost_from_counts(ig['n_distinct'], i_f['n_distinct'], G.size + F.size)
        assert enc == want[bits, 0]
        assert _native.mdl_error_cost(dev(V), Ge, Fe) == pytest.approx(want[bits, 1], rel=1e-9)
    qG.close()
    qF.close()


def test_select_model_on_the_reference_test_input(rolx_cases):
    """tests/test_roles/test_extract.py:81-88: seeded 20 x 30 uniform data.  The reference's test
    expects 2 roles; with the scikit-learn installed here the unmodified reference selects 8
    (recorded in the fixture by make_golden.py) -- and so must this path.  The cost grid is
    compared cell by cell: encoding costs exactly, error costs within 5 % and their median
    relative difference below 1e-3 (the NMF behind each row is a TF32/fp32 run from an NNDSVDa
    start with its own random stream; where a factor entry sits next to a quantisation boundary
    a coarse codebook assigns it differently).  One NMF per
    n_roles instead of one per cell."""
    z = rolx_cases
    feats = pd.DataFrame(z['select__X'])
    rx = RoleExtractor()
    rx.extract_role_factors(feats)
    assert rx.node_role_factor.shape == (20, int(z['select__n_roles']))
    assert rx.role_feature_factor.shape == (int(z['select__n_roles']), 30)
    enc, err = rx.grid_costs_
    want_enc, want_err = z['select__encoding_costs'], z['select__error_costs']
    assert np.array_equal(np.isnan(enc), np.isnan(want_enc))
    ok = ~np.isnan(want_enc)
    np.testing.assert_array_equal(enc[ok], want_enc[ok])
    np.testing.assert_allclose(err[ok], want_err[ok], rtol=5e-2)
    assert np.median(np.abs(err[ok] - want_err[ok]) / want_err[ok]) < 1e-3
    grid = DeviceModelGrid(feats.values)
    for bits in (1, 2, 3):
        grid.costs(3, bits)
    assert grid.n_fits == 1
    grid.costs(3, 2, refit=True)
    assert grid.n_fits == 2
    grid.close()


def test_grid_refuses_what_the_solver_cannot_do():
    feats = pd.DataFrame(np.random.RandomState(0).rand(50, 40))
    with pytest.raises(ValueError, match='at most 32'):
        RoleExtractor(n_role_range=(33, 34), n_bit_range=(1, 2)).extract_role_factors(feats)
    bad = feats.copy()
    bad.iloc[3, 4] = np.nan
    with pytest.raises(ValueError, match='NaN'):
        RoleExtractor(n_roles=2).extract_role_factors(bad)


def test_roles_and_percentages_on_device():
    rng = np.random.RandomState(8)
    W = rng.rand(10_000, 5)
    W[7] = [0.2, 0.9, 0.9, 0.1, 0.9]            # first maximum wins, like idxmax
    for dtype in (torch.float64, torch.float32):
        arg, pct = _native.roles(dev(W, dtype))
        Wd = W.astype(np.float32 if dtype == torch.float32 else np.float64)
        np.testing.assert_array_equal(arg.cpu().numpy(), Wd.argmax(axis=1))
        np.testing.assert_allclose(pct.cpu().numpy(), Wd / Wd.sum(axis=1, keepdims=True),
                                   rtol=1e-6 if dtype == torch.float32 else 1e-14)
    feats = pd.DataFrame(rng.rand(60, 12), index=[f'n{i}' for i in range(60)])
    rx = RoleExtractor(n_roles=3)
    rx.extract_role_factors(feats)
    assert rx.roles == rx.node_role_factor.idxmax(axis=1).to_dict()
    np.testing.assert_allclose(rx.role_percentage.values,
                               rx.node_role_factor.div(rx.node_role_factor.sum(axis=1), axis=0).values)
