"""Pins oracle/nmf_oracle.py (CPU restatement of hot path B) against scikit-learn, the
third-party owner of the arithmetic: committed golden vectors and the installed package."""
import numpy as np
import pytest

from oracle import nmf_oracle as oracle

CASES = [('rand20x30', 2), ('rand20x30', 4), ('rand20x30', 7), ('rand300x64', 4),
         ('rand300x64', 8), ('rand300x64', 16), ('planted500x48', 4), ('planted500x48', 8)]


@pytest.mark.parametrize('name,r', CASES)
def test_mu_loop_matches_golden(nmf_cases, name, r):
    z = nmf_cases
    X, W0, H0 = z[f'{name}__X'], z[f'{name}__r{r}__W0'], z[f'{name}__r{r}__H0']
    W, H, n_iter = oracle.fit_multiplicative_update(X, W0, H0, max_iter=50, tol=0)
    assert n_iter == 50
    np.testing.assert_allclose(W, z[f'{name}__r{r}__W50'], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(H, z[f'{name}__r{r}__H50'], rtol=1e-10, atol=1e-13)
    W, H, n_iter = oracle.fit_multiplicative_update(X, W0, H0, max_iter=200, tol=1e-4)
    assert n_iter == int(z[f'{name}__r{r}__n_iter'])
    np.testing.assert_allclose(W, z[f'{name}__r{r}__Wconv'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(H, z[f'{name}__r{r}__Hconv'], rtol=1e-9, atol=1e-12)
    assert oracle.frobenius_error(X, W, H) == pytest.approx(float(z[f'{name}__r{r}__err']),
                                                             rel=1e-10)


def test_reference_wrapper_equals_seeded_sklearn_path(nmf_cases):
    """graphrole.roles.factor.get_nmf_decomposition under np.random.seed == MU from the
    NNDSVDa start drawn under the same seed (factor.py:19-25)."""
    z = nmf_cases
    for name, r in CASES:
        np.testing.assert_allclose(z[f'{name}__r{r}__Gref'], z[f'{name}__r{r}__Wconv'],
                                   rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(z[f'{name}__r{r}__Fref'], z[f'{name}__r{r}__Hconv'],
                                   rtol=1e-9, atol=1e-12)


def test_mu_loop_matches_installed_sklearn():
    sk = pytest.importorskip('sklearn.decomposition._nmf')
    rng = np.random.RandomState(5)
    X = rng.rand(120, 40)
    W0, H0 = rng.rand(120, 6) + 0.1, rng.rand(6, 40) + 0.1
    W_ref, H_ref, it_ref = sk._fit_multiplicative_update(X, W0.copy(), H0.copy(), 'frobenius',
                                                         max_iter=200, tol=1e-4)
    W, H, it = oracle.fit_multiplicative_update(X, W0, H0)
    assert it == it_ref
    np.testing.assert_allclose(W, W_ref, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(H, H_ref, rtol=1e-9, atol=1e-12)
    assert oracle.EPSILON == sk.EPSILON


def test_zero_denominator_uses_epsilon():
    X = np.array([[1.0, 0.0], [0.0, 2.0]])
    W0 = np.array([[1.0, 0.0], [0.0, 0.0]])
    H0 = np.array([[1.0, 0.0], [0.0, 0.0]])
    W, H, _ = oracle.fit_multiplicative_update(X, W0, H0, max_iter=3, tol=0)
    assert np.isfinite(W).all() and np.isfinite(H).all()


def test_nndsvda_from_exact_svd_matches_sklearn_on_well_separated_spectrum():
    sk = pytest.importorskip('sklearn.decomposition._nmf')
    rng = np.random.RandomState(0)
    # rank-4 with well separated singular values: the randomised SVD is then accurate
    X = (rng.rand(60, 4) * np.array([8.0, 4.0, 2.0, 1.0])) @ rng.rand(4, 25)
    np.random.seed(3)
    W_ref, H_ref = sk._initialize_nmf(X, 3, init='nndsvda')
    U, S, Vt = oracle.truncated_svd_exact(X, 3)
    W, H = oracle.nndsvda_from_svd(X, U, S, Vt)
    np.testing.assert_allclose(W, W_ref, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(H, H_ref, rtol=1e-6, atol=1e-8)


def test_zero_padding_is_a_fixed_point_of_the_updates():
    """What lets every rank and every feature count run the tensor-core kernels: a zero column of
    W with the matching zero row of H (rank padded to a multiple of 4 inside the library), and a
    zero column of X with the matching zero column of H0 (feature count padded by
    roles.factor.FeatureMatrix), stay zero under the multiplicative updates and leave the real
    entries what they were."""
    rng = np.random.RandomState(2)
    n, f, r = 200, 30, 5
    X = rng.rand(n, 4) @ rng.rand(4, f) + 0.05 * rng.rand(n, f)
    W0, H0 = rng.rand(n, r) + 0.1, rng.rand(r, f) + 0.1
    W, H, it = oracle.fit_multiplicative_update(X, W0, H0, max_iter=60, tol=0)
    # rank 5 -> 8
    Wp, Hp, _ = oracle.fit_multiplicative_update(
        X, np.pad(W0, ((0, 0), (0, 3))), np.pad(H0, ((0, 3), (0, 0))), max_iter=60, tol=0)
    assert not Wp[:, r:].any() and not Hp[r:].any()
    np.testing.assert_allclose(Wp[:, :r], W, rtol=1e-10)
    np.testing.assert_allclose(Hp[:r], H, rtol=1e-10)
    # 30 features -> 32
    Wq, Hq, _ = oracle.fit_multiplicative_update(
        np.pad(X, ((0, 0), (0, 2))), W0, np.pad(H0, ((0, 0), (0, 2))), max_iter=60, tol=0)
    assert not Hq[:, f:].any()
    np.testing.assert_allclose(Wq, W, rtol=1e-10)
    np.testing.assert_allclose(Hq[:, :f], H, rtol=1e-10)
    assert oracle.frobenius_error(np.pad(X, ((0, 0), (0, 2))), Wq, Hq) == pytest.approx(
        oracle.frobenius_error(X, W, H), rel=1e-12)


def _tf32(a):
    """float32 values rounded to TF32 (10 mantissa bits, nearest, ties away like cvt.rna)."""
    u = np.asarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x1000) & 0xFFFFE000).astype(np.uint32)
    return u.view(np.float32).astype(np.float64)


@pytest.mark.parametrize('n,f,r', [(1, 4, 4), (127, 32, 4), (300, 64, 4), (1000, 132, 20),
                                   (5003, 96, 12), (257, 512, 32)])
def test_tf32_operand_rounding_moves_the_error_within_the_stated_bound(n, f, r):
    """A NumPy model of the tensor-core convergence check (W and H rounded to TF32, everything
    else exact) against the float64 error: the difference stays inside the bound the GPU test
    states (tests/test_nmf_gpu.py::tf32_error_bound: 1e-5 relative + 3 dn / sqrt(r (n + f)) +
    dn^2 / (2 err) with dn = 2^-11 ||W H||_F), on random factors and on a fitted model."""
    rng = np.random.RandomState(n % 89 + f + r)
    X, W, H = (rng.rand(n, f).astype(np.float32).astype(np.float64),
               rng.rand(n, r).astype(np.float32).astype(np.float64),
               rng.rand(r, f).astype(np.float32).astype(np.float64))
    Wf, Hf, _ = oracle.fit_multiplicative_update(X, W, H, max_iter=30, tol=0)
    for Wc, Hc in ((W, H), (Wf.astype(np.float32).astype(np.float64),
                            Hf.astype(np.float32).astype(np.float64))):
        want = oracle.frobenius_error(X, Wc, Hc)
        got = oracle.frobenius_error(X, _tf32(Wc), _tf32(Hc))
        dn = 2.0 ** -11 * np.linalg.norm(Wc @ Hc)
        bound = 1e-5 * want + 3 * dn / np.sqrt(r * (n + f)) + dn * dn / (2 * want)
        assert abs(got - want) <= bound, (abs(got - want), bound)
