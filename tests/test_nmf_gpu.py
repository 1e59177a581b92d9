"""Parity of hot path B (NMF multiplicative updates) on the GPU against scikit-learn.

The reference's tests pin shapes and non-negativity only (tests/test_roles/test_factor.py:17-25);
numerical parity is therefore anchored on scikit-learn -- the owner of the arithmetic -- with a
SHARED start (init='custom' semantics): committed golden vectors (tests/golden/nmf_cases.npz),
the pinned NumPy oracle, and the installed sklearn run live.

Stated tolerances (fp32 storage on the GPU vs float64 in sklearn):
  FFMA path (use_tf32=False)   factors after a fixed 50 iterations within 2e-3 relative to the
                               factor's max entry; reconstruction error within 1e-4 relative
  TF32 path (use_tf32=True)    factors within 2e-2 relative to max entry on the small golden
                               cases; on the benchmarked f = 512 pair kernel within 1e-2, error
                               within 1e-3, node roles (argmax) identical on >= 99 % of nodes
  stopping iteration           identical, or off by one convergence check (10 iterations) when
                               sklearn's criterion is within rounding of the threshold
"""
import numpy as np
import pytest
import torch

from graphrole_b200 import RoleExtractor
from graphrole_b200.roles import factor
from oracle import nmf_oracle as oracle

pytestmark = pytest.mark.gpu

CASES = [('rand20x30', 2), ('rand20x30', 4), ('rand20x30', 7), ('rand300x64', 4),
         ('rand300x64', 8), ('rand300x64', 16), ('planted500x48', 4), ('planted500x48', 8)]


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device='cuda:0')


def run_mu(X, W0, H0, use_tf32, **kw):
    """The multiplicative-update loop on the path under test.  TF32: through FeatureMatrix, like
    the product (feature counts that are not a multiple of 4 are zero-padded), and the tcgen05
    kernel must actually have run -- every golden case, whatever its rank and width."""
    if not use_tf32:
        return factor.nmf_mu(dev(X), dev(W0), dev(H0), use_tf32=False, **kw)
    out = factor.FeatureMatrix(dev(X)).nmf_mu(dev(W0), dev(H0), use_tf32=True, **kw)
    assert factor.last_path == 'tcgen05'
    return out


def rel_to_max(got, ref):
    return float(np.abs(got - ref).max() / np.abs(ref).max())


@pytest.mark.parametrize('use_tf32,tol_f,tol_e', [(False, 2e-3, 1e-4), (True, 2e-2, 1e-3)])
@pytest.mark.parametrize('name,r', CASES)
def test_fixed_iterations_match_sklearn_golden(nmf_cases, name, r, use_tf32, tol_f, tol_e):
    z = nmf_cases
    X, W0, H0 = z[f'{name}__X'], z[f'{name}__r{r}__W0'], z[f'{name}__r{r}__H0']
    W, H, n_iter, err = run_mu(X, W0, H0, use_tf32, max_iter=50, tol=0)
    assert n_iter == 50
    W, H = W.cpu().numpy().astype(np.float64), H.cpu().numpy().astype(np.float64)
    assert (W >= 0).all() and (H >= 0).all()
    assert rel_to_max(W, z[f'{name}__r{r}__W50']) < tol_f
    assert rel_to_max(H, z[f'{name}__r{r}__H50']) < tol_f
    ref_err = oracle.frobenius_error(X, z[f'{name}__r{r}__W50'], z[f'{name}__r{r}__H50'])
    assert err == pytest.approx(ref_err, rel=tol_e)
    assert err == pytest.approx(oracle.frobenius_error(X, W, H), rel=1e-4)


@pytest.mark.parametrize('use_tf32,tol_f', [(False, 5e-3), (True, 2e-2)])
@pytest.mark.parametrize('name,r', CASES)
def test_convergence_loop_matches_sklearn(nmf_cases, name, r, use_tf32, tol_f):
    """tol = 1e-4, max_iter = 200, check every 10 iterations (sklearn NMF defaults) -- on the FFMA
    kernels and on the DEFAULT product path (tcgen05, TF32 contractions)."""
    z = nmf_cases
    X, W0, H0 = z[f'{name}__X'], z[f'{name}__r{r}__W0'], z[f'{name}__r{r}__H0']
    W, H, n_iter, err = run_mu(X, W0, H0, use_tf32)
    ref_iter = int(z[f'{name}__r{r}__n_iter'])
    assert abs(n_iter - ref_iter) <= 10 and n_iter % 10 == 0
    ref_err = float(z[f'{name}__r{r}__err'])
    assert err == pytest.approx(ref_err, rel=2e-3)
    if n_iter == ref_iter:
        Wn, Hn = W.cpu().numpy().astype(np.float64), H.cpu().numpy().astype(np.float64)
        Wr, Hr = z[f'{name}__r{r}__Wconv'], z[f'{name}__r{r}__Hconv']
        # what the factors are FOR -- the reconstruction -- agrees to 1e-2 of its norm on every case
        assert np.linalg.norm(Wn @ Hn - Wr @ Hr) / np.linalg.norm(Wr @ Hr) < 1e-2
        # the factors themselves: uniform noise has no identifiable low-rank structure, so after
        # 100+ iterations TF32 rounding moves them along the flat directions of the loss (measured:
        # up to 4.5e-2 of the largest entry on rand300x64, r = 4, with the error still equal to
        # 2e-3); they are compared where they are identifiable (planted data) or exact (fp32 path)
        if not use_tf32 or name.startswith('planted'):
            assert rel_to_max(Wn, Wr) < tol_f
            assert rel_to_max(Hn, Hr) < tol_f


@pytest.mark.parametrize('r', [4, 8, 16, 32])
@pytest.mark.parametrize('kind', ['planted', 'uniform'])
def test_benchmarked_pair_kernel_matches_sklearn_directly(kind, r):
    """The kernel bench.py times on C5 -- nmf_fused_tc_kernel<2> (f > 128: column-split CTA pairs)
    -- against scikit-learn's _fit_multiplicative_update itself (float64, shared W0 / H0), at
    f = 512 with several row blocks per CTA and a ragged tail (n = 64 * 148 * 3 + 5), 50
    iterations.  Stated tolerance: factors within 1e-2 of the factor's
    largest entry, reconstruction error within 1e-3 relative, and the role of a node (argmax of its
    row of W) identical on >= 99 % of the nodes whose two best roles differ by more than 1 %."""
    sk = pytest.importorskip('sklearn.decomposition._nmf')
    rng = np.random.RandomState(100 + r)
    n, f = 64 * 148 * 3 + 5, 512
    if kind == 'planted':
        X = rng.rand(n, r) ** 2 @ rng.rand(r, f) + 0.02 * rng.rand(n, f)
    else:
        X = rng.rand(n, f)
    W0, H0 = rng.rand(n, r) + 0.1, rng.rand(r, f) + 0.1
    X32, W32, H32 = (a.astype(np.float32) for a in (X, W0, H0))     # the bytes the GPU reads
    W_sk, H_sk, _ = sk._fit_multiplicative_update(
        X32.astype(np.float64), W32.astype(np.float64), H32.astype(np.float64), 'frobenius',
        max_iter=50, tol=0)
    W, H, n_iter, err = factor.nmf_mu(dev(X32), dev(W32), dev(H32), max_iter=50, tol=0,
                                      use_tf32=True)
    assert factor.last_path == 'tcgen05' and n_iter == 50
    W, H = W.cpu().numpy().astype(np.float64), H.cpu().numpy().astype(np.float64)
    ew, eh = rel_to_max(W, W_sk), rel_to_max(H, H_sk)
    ref_err = oracle.frobenius_error(X32.astype(np.float64), W_sk, H_sk)
    top2 = np.sort(W_sk, axis=1)[:, -2:]
    decided = (top2[:, 1] - top2[:, 0]) > 0.01 * top2[:, 1]
    agree = float((W.argmax(1) == W_sk.argmax(1))[decided].mean()) if decided.any() else 1.0
    print(f'pair kernel vs sklearn [{kind}, r={r}]: W {ew:.2e}  H {eh:.2e}  err rel '
          f'{abs(err - ref_err) / ref_err:.2e}  argmax agreement {agree:.4f} on '
          f'{int(decided.sum())} decided nodes')
    assert ew < 1e-2 and eh < 1e-2
    assert err == pytest.approx(ref_err, rel=1e-3)
    assert agree >= 0.99


def test_against_oracle_and_live_sklearn_medium():
    sk = pytest.importorskip('sklearn.decomposition._nmf')
    rng = np.random.RandomState(7)
    n, f, r = 5000, 96, 12
    X = rng.rand(n, 8) @ rng.rand(8, f) + 0.05 * rng.rand(n, f)
    W0, H0 = rng.rand(n, r) + 0.1, rng.rand(r, f) + 0.1
    W_sk, H_sk, _ = sk._fit_multiplicative_update(X, W0.copy(), H0.copy(), 'frobenius',
                                                  max_iter=30, tol=0)
    W_or, H_or, _ = oracle.fit_multiplicative_update(X, W0, H0, max_iter=30, tol=0)
    np.testing.assert_allclose(W_or, W_sk, rtol=1e-9)
    W, H, _, err = factor.nmf_mu(dev(X), dev(W0), dev(H0), max_iter=30, tol=0, use_tf32=False)
    assert rel_to_max(W.cpu().numpy(), W_sk) < 2e-3
    assert rel_to_max(H.cpu().numpy(), H_sk) < 2e-3
    assert err == pytest.approx(oracle.frobenius_error(X, W_sk, H_sk), rel=1e-4)
    assert factor.nmf_error(dev(X), dev(W_sk), dev(H_sk)) == pytest.approx(
        oracle.frobenius_error(X, W_sk, H_sk), rel=1e-5)


@pytest.mark.parametrize('n,f,r', [(1, 4, 1), (257, 512, 32), (1000, 132, 5), (5003, 96, 12),
                                   (64 * 148 * 3 + 5, 512, 4), (300, 130, 8)])
def test_error_pass_kernels_match_the_oracle(n, f, r, monkeypatch):
    """||X - W H||_F (sklearn _nmf.py:114-127, dense branch): the register-tiled kernel (f % 4 ==
    0) and the general thread-per-column kernel against the float64 oracle, ragged row / column
    tails included; f = 130 takes the general kernel by itself."""
    rng = np.random.RandomState(n % 97 + f + r)
    X, W, H = rng.rand(n, f), rng.rand(n, r), rng.rand(r, f)
    X32, W32, H32 = (a.astype(np.float32) for a in (X, W, H))
    want = oracle.frobenius_error(X32.astype(np.float64), W32.astype(np.float64),
                                  H32.astype(np.float64))
    got = factor.nmf_error(dev(X32), dev(W32), dev(H32))
    assert got == pytest.approx(want, rel=1e-5)
    monkeypatch.setenv('GR_NMF_ERROR_SIMPLE', '1')
    simple = factor.nmf_error(dev(X32), dev(W32), dev(H32))
    assert simple == pytest.approx(want, rel=1e-5)
    assert got == pytest.approx(simple, rel=2e-6)


def tf32_error_bound(want, W, H):
    """How far the TF32 rounding of the operands may move ||X - W H||_F.  Every term of W H is
    perturbed by <= 2^-11 relative and without bias: ||delta|| <= dn = 2^-11 ||W H||_F, and
    ||d + delta|| - ||d|| ~ (d . delta) / ||d|| + ||delta||^2 / (2 ||d||).  d . delta is a sum
    of zero-mean terms of which r (n + f) are independent (one rounding per entry of W and of H:
    the error of one entry of H moves a whole column of W H the same way).  1e-5 relative on top
    (fp32 residuals, accumulation order)."""
    W, H = np.asarray(W, dtype=np.float64), np.asarray(H, dtype=np.float64)
    dn = 2.0 ** -11 * np.linalg.norm(W @ H)
    independent = W.shape[1] * (W.shape[0] + H.shape[1])
    return 1e-5 * want + 3 * dn / np.sqrt(independent) + dn * dn / (2 * want)


@pytest.mark.parametrize('n,f,r', [(1, 4, 4), (127, 32, 4), (129, 36, 8), (257, 512, 32),
                                   (1000, 132, 20), (5003, 96, 12), (4096, 768, 16),
                                   (3000, 1024, 32), (777, 260, 28), (300, 64, 4),
                                   (128 * 148 * 2 + 131, 512, 8)])
def test_tensor_core_error_pass_matches_the_oracle(n, f, r):
    """gr_nmf_error_tf32 -- W H on tcgen05 (M = 128 row blocks x 64-column stages), residual and
    squares in fp32, fp64 sums: the convergence check of the default (TF32) product path.
    Against the float64 oracle (sklearn _nmf.py:114-127) evaluated on the same fp32 inputs,
    within tf32_error_bound: 1e-5 ... 2e-5 relative from 10^5 entries on (measured on C5: 2e-8
    against the fp32 FFMA pass), looser only on toy sizes and on (near-)exact fits, where the
    operand rounding is all that is left of the residual.  Shapes: ragged row blocks (n % 128),
    ragged 32-column boxes, 64-column stages and 256-column spans (f = 36, 132, 260), one to
    four spans, several blocks per CTA."""
    rng = np.random.RandomState(n % 89 + f + r)
    X, W, H = rng.rand(n, f), rng.rand(n, r), rng.rand(r, f)
    X32, W32, H32 = (a.astype(np.float32) for a in (X, W, H))
    want = oracle.frobenius_error(X32.astype(np.float64), W32.astype(np.float64),
                                  H32.astype(np.float64))
    got = factor.nmf_error(dev(X32), dev(W32), dev(H32), use_tf32=True)
    assert abs(got - want) <= tf32_error_bound(want, W32, H32)
    if n * f >= 100_000:
        assert got == pytest.approx(want, rel=2e-5)
    # a fitted model (small residual: the regime of the stopping rule)
    Wd, Hd, _, _ = factor.nmf_mu(dev(X32), dev(W32), dev(H32), max_iter=30, tol=0, use_tf32=True)
    Wn, Hn = Wd.double().cpu().numpy(), Hd.double().cpu().numpy()
    want = oracle.frobenius_error(X32.astype(np.float64), Wn, Hn)
    got = factor.nmf_error(dev(X32), Wd, Hd, use_tf32=True)
    assert abs(got - want) <= tf32_error_bound(want, Wn, Hn)
    if n * f >= 100_000:
        assert got == pytest.approx(want, rel=2e-5)
    # (the fp32 FFMA pass: on an exact fit all that is left of the residual is fp32 rounding)
    assert factor.nmf_error(dev(X32), Wd, Hd) == pytest.approx(
        want, rel=1e-5, abs=1e-6 * np.linalg.norm(X32))


def test_tensor_core_error_pass_refuses_other_shapes():
    X, W, H = torch.rand(50, 130, device='cuda:0'), torch.rand(50, 5, device='cuda:0'), \
        torch.rand(5, 130, device='cuda:0')
    with pytest.raises(ValueError):
        factor.nmf_error(X, W, H, use_tf32=True)
    assert factor.nmf_error(X, W, H) > 0


def test_config5_10m_x_512_at_size():
    """BASELINE.json configs[4] at its own size: X = 10 M x 512 (5.1e9 entries: every index past
    2^32), r = 8.  scikit-learn cannot run here in test time, so the checks are the ones that do
    not depend on size: the tcgen05 pair kernel against the fp32 FFMA kernels from the same
    start (3 iterations), non-negativity, and the multiplicative update's monotone decrease of
    ||X - W H||_F (Lee & Seung) -- evaluated by the error pass, itself checked against a float64
    evaluation on a row sample."""
    free_gpu, _ = torch.cuda.mem_get_info(0)
    if free_gpu < 60e9:
        pytest.skip('needs ~45 GB of HBM')
    n, f, r = 10_000_000, 512, 8
    gen = torch.Generator(device='cuda:0').manual_seed(5)
    X = torch.rand(n, 8, device='cuda:0', generator=gen).square_() @ \
        torch.rand(8, f, device='cuda:0', generator=gen)
    X += 0.05 * torch.rand(n, f, device='cuda:0', generator=gen)
    W0 = torch.rand(n, r, device='cuda:0', generator=gen) + 0.1
    H0 = torch.rand(r, f, device='cuda:0', generator=gen) + 0.1
    e0 = factor.nmf_error(X, W0, H0)
    Wt, Ht, _, et = factor.nmf_mu(X, W0, H0, max_iter=3, tol=0, use_tf32=True)
    assert factor.last_path == 'tcgen05'
    Wf, Hf, _, ef = factor.nmf_mu(X, W0, H0, max_iter=3, tol=0, use_tf32=False)
    assert bool((Wt >= 0).all()) and bool((Ht >= 0).all())
    assert float((Wt - Wf).abs().max() / Wf.abs().max()) < 5e-3
    assert float((Ht - Hf).abs().max() / Hf.abs().max()) < 5e-3
    assert et == pytest.approx(ef, rel=1e-3)
    assert et < e0
    W6, H6, _, e6 = factor.nmf_mu(X, Wt, Ht, max_iter=3, tol=0, use_tf32=True)
    assert e6 < et
    # the error pass itself, on the last 50 000 rows (indices past 2^32 in the flattened matrix)
    tail = slice(n - 50_000, n)
    want = torch.linalg.norm(X[tail].double() - W6[tail].double() @ H6.double()).item()
    got = factor.nmf_error(X[tail].contiguous(), W6[tail].contiguous(), H6)
    assert got == pytest.approx(want, rel=1e-6)
    got = factor.nmf_error(X[tail].contiguous(), W6[tail].contiguous(), H6, use_tf32=True)
    assert got == pytest.approx(want, rel=1e-5)
    # both forms on the whole matrix (the tensor-core pass is what the loop's checks run)
    assert factor.nmf_error(X, W6, H6, use_tf32=True) == pytest.approx(
        factor.nmf_error(X, W6, H6), rel=1e-5)


def test_tensor_core_path_is_taken_and_matches_ffma():
    """Shapes the tcgen05 kernel takes (f % 4 == 0, f <= 1024; any rank up to 32 -- ranks that are
    not a multiple of 4 run on zero-padded factors inside the library) must actually run it -- no
    silent fallback -- and agree with the FFMA kernels; other shapes report 'ffma'."""
    rng = np.random.RandomState(3)
    for n, f, r, expect in [(3000, 512, 32, 'tcgen05'), (64 * 148 * 2 + 5, 128, 8, 'tcgen05'),
                            (1000, 96, 12, 'tcgen05'), (500, 704, 16, 'tcgen05'),
                            (500, 768, 16, 'tcgen05'), (9000, 1024, 32, 'tcgen05'),
                            (700, 132, 4, 'tcgen05'), (64 * 74 * 3 + 1, 256, 32, 'tcgen05'),
                            (500, 64, 5, 'tcgen05'), (400, 128, 2, 'tcgen05'),
                            (400, 96, 7, 'tcgen05'), (64 * 148 + 3, 512, 29, 'tcgen05'),
                            (257, 36, 3, 'tcgen05'), (200, 8, 1, 'tcgen05'),
                            (500, 1028, 16, 'ffma'), (500, 30, 4, 'ffma')]:
        X = rng.rand(n, f)
        W0, H0 = rng.rand(n, r) + 0.1, rng.rand(r, f) + 0.1
        Wt, Ht, _, _ = factor.nmf_mu(dev(X), dev(W0), dev(H0), max_iter=5, tol=0, use_tf32=True)
        assert factor.last_path == expect, (n, f, r, factor.last_path)
        Wf, Hf, _, _ = factor.nmf_mu(dev(X), dev(W0), dev(H0), max_iter=5, tol=0, use_tf32=False)
        assert factor.last_path == 'ffma'
        assert rel_to_max(Wt.cpu().numpy(), Wf.cpu().numpy()) < 5e-3
        assert rel_to_max(Ht.cpu().numpy(), Hf.cpu().numpy()) < 5e-3


@pytest.mark.parametrize('n,f,r', [(600, 30, 5), (300, 7, 3), (1000, 513, 6), (64, 33, 2)])
def test_any_feature_count_runs_the_tensor_core_path(n, f, r):
    """factor.FeatureMatrix stores the features with zero columns up to a multiple of 4 (the
    16-byte row pitch TMA needs): the padded problem's first f columns are the problem -- same
    factors as the unpadded FFMA run to the TF32 tolerance, H's padding stays exactly zero, same
    error as the unpadded matrix."""
    rng = np.random.RandomState(f)
    X = rng.rand(n, f)
    W0, H0 = rng.rand(n, r) + 0.1, rng.rand(r, f) + 0.1
    fm = factor.FeatureMatrix(dev(X))
    assert fm.padded.shape[1] % 4 == 0 and fm.values.shape == (n, f)
    assert bool((fm.padded[:, f:] == 0).all()) and bool((fm.values == dev(X)).all())
    Wt, Ht, it_t, err_t = fm.nmf_mu(dev(W0), dev(H0), max_iter=20, tol=0)
    assert factor.last_path == 'tcgen05' and Ht.shape == (r, f) and it_t == 20
    Wf, Hf, _, err_f = factor.nmf_mu(dev(X), dev(W0), dev(H0), max_iter=20, tol=0, use_tf32=False)
    assert factor.last_path == 'ffma'
    assert rel_to_max(Wt.cpu().numpy(), Wf.cpu().numpy()) < 1e-2
    assert rel_to_max(Ht.cpu().numpy(), Hf.cpu().numpy()) < 1e-2
    assert err_t == pytest.approx(err_f, rel=1e-3)
    assert err_t == pytest.approx(
        oracle.frobenius_error(X, Wt.double().cpu().numpy(), Ht.double().cpu().numpy()), rel=1e-4)


def test_row_sharded_loop_on_one_rank_is_the_library_loop():
    """roles/sharded.py::RowShardedNmf with a single shard (no process group): the split
    iteration (gr_nmf_iteration_local_f32 + gr_nmf_update_h_f32) adds the per-CTA partials in the
    same order as gr_nmf_mu_f32, so factors, stopping iteration and error are bit-identical --
    on the tcgen05 path and on the FFMA path, padded ranks included."""
    from graphrole_b200.roles.sharded import nmf_mu_row_sharded
    rng = np.random.RandomState(11)
    for n, f, r, use_tf32 in [(64 * 148 + 77, 512, 8, True), (3000, 96, 5, True),
                              (2000, 70, 6, False)]:
        X = (rng.rand(n, 6) ** 2) @ rng.rand(6, f) + 0.05 * rng.rand(n, f)
        W0, H0 = rng.rand(n, r) + 0.1, rng.rand(r, f) + 0.1
        Ws, Hs, it_s, err_s = nmf_mu_row_sharded(dev(X), dev(W0), dev(H0), use_tf32=use_tf32)
        Wl, Hl, it_l, err_l = factor.nmf_mu(dev(X), dev(W0), dev(H0), use_tf32=use_tf32)
        assert factor.last_path == ('tcgen05' if use_tf32 else 'ffma')
        assert it_s == it_l and err_s == err_l
        assert torch.equal(Ws, Wl) and torch.equal(Hs, Hl)


@pytest.mark.parametrize('shards', [2, 3])
def test_row_sharded_iterations_match_unsharded(shards):
    """The exchange step emulated on one GPU: every shard runs gr_nmf_iteration_local_f32 on its
    rows, the [W^T X | W^T W] sums are added (what the all-reduce does) and every shard applies
    gr_nmf_update_h_f32 -- every shard holds the same H bit for bit, and the factors equal the
    unsharded run up to the order of the fp32 partial sums.  On the tcgen05 path W is held at TF32
    precision, so a last-bit difference in H can flip the rounding of an entry of W by one TF32
    step (2^-11 of its value) and the two TF32 trajectories separate at that level: stated 1e-2
    of the largest entry for W, 2e-3 for H, 1e-4 for the error (measured 1.7e-3 / < 1e-4 here,
    3.4e-3 / 7e-4 / 1.3e-5 after 40 iterations on 2..8 GPUs; the FFMA path: 2e-6).  The real
    two-GPU run: test_two_gpu_row_sharded_nmf."""
    from graphrole_b200.roles.sharded import CudaNmfBackend, row_shard
    rng = np.random.RandomState(shards)
    n, f, r = 128 * 40 + 19, 256, 8
    X = (rng.rand(n, 6) ** 2) @ rng.rand(6, f) + 0.05 * rng.rand(n, f)
    W0, H0 = rng.rand(n, r) + 0.1, rng.rand(r, f) + 0.1
    Xd = dev(X)
    spans = [row_shard(n, shards, s) for s in range(shards)]
    Ws = [dev(W0[lo:hi]) for lo, hi in spans]
    Hs = [dev(H0) for _ in spans]
    backends = [CudaNmfBackend(hi - lo, f, r, Xd.device) for lo, hi in spans]
    for _ in range(20):
        sums = [b.local_iteration(Xd[lo:hi], W, H).clone()
                for b, (lo, hi), W, H in zip(backends, spans, Ws, Hs)]
        total = torch.stack(sums).sum(dim=0)
        for b, H in zip(backends, Hs):
            b.update_h(total, H)
    assert all(b.last_path == 'tcgen05' for b in backends)
    err_sq = sum(b.error_sq(Xd[lo:hi], W, H) for b, (lo, hi), W, H in zip(backends, spans, Ws, Hs))
    for b in backends:
        b.close()
    Wu, Hu, _, err_u = factor.nmf_mu(Xd, dev(W0), dev(H0), max_iter=20, tol=0)
    for H in Hs[1:]:
        assert torch.equal(H, Hs[0])
    assert rel_to_max(torch.cat(Ws).cpu().numpy(), Wu.cpu().numpy()) < 1e-2
    assert rel_to_max(Hs[0].cpu().numpy(), Hu.cpu().numpy()) < 2e-3
    assert err_sq ** 0.5 == pytest.approx(err_u, rel=1e-4)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs on one box')
def test_two_gpu_row_sharded_nmf():
    """torchrun, one rank per GPU, NCCL all-reduce (tools/check_sharded_nmf.py)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run(
        [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
         '--master-addr', '127.0.0.1', '--master-port', '29741',
         os.path.join(root, 'tools', 'check_sharded_nmf.py'), '--rows', '300000'],
        capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and 'SHARDED NMF CHECK OK' in res.stdout, res.stdout + res.stderr


def test_zero_denominators_and_edge_shapes():
    X = np.array([[1.0, 0.0, 0.0], [0.0, 2.0, 0.0], [0.0, 0.0, 0.0]])
    W0 = np.array([[1.0, 0.0], [0.0, 0.0], [0.0, 0.0]])
    H0 = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, 0.0]])
    W, H, _, _ = factor.nmf_mu(dev(X), dev(W0), dev(H0), max_iter=5, tol=0, use_tf32=False)
    Wr, Hr, _ = oracle.fit_multiplicative_update(X, W0, H0, max_iter=5, tol=0)
    assert torch.isfinite(W).all() and torch.isfinite(H).all()
    np.testing.assert_allclose(W.cpu().numpy(), Wr, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(H.cpu().numpy(), Hr, rtol=1e-4, atol=1e-7)
    # single row / single feature / rank 1, odd sizes
    rng = np.random.RandomState(1)
    for n, f, r in [(1, 5, 1), (7, 1, 1), (65, 33, 3), (257, 513, 5), (1000, 7, 7)]:
        X = rng.rand(n, f)
        W0, H0 = rng.rand(n, r) + 0.1, rng.rand(r, f) + 0.1
        W, H, _, _ = factor.nmf_mu(dev(X), dev(W0), dev(H0), max_iter=10, tol=0, use_tf32=False)
        Wr, Hr, _ = oracle.fit_multiplicative_update(X, W0, H0, max_iter=10, tol=0)
        assert rel_to_max(W.cpu().numpy(), Wr) < 1e-3
        assert rel_to_max(H.cpu().numpy(), Hr) < 1e-3
    with pytest.raises(ValueError):
        factor.nmf_mu(dev(rng.rand(50, 40)), dev(rng.rand(50, 33)), dev(rng.rand(33, 40)))


def test_get_nmf_decomposition_contract():
    """The reference's own test (tests/test_roles/test_factor.py:17-25): shapes, non-negative."""
    np.random.seed(0)
    X = np.random.rand(20, 30)
    for n_roles in range(2, 8):
        G, F = factor.get_nmf_decomposition(X, n_roles)
        assert G.shape == (20, n_roles) and F.shape == (n_roles, 30)
        assert (G >= 0).all() and (F >= 0).all()
        assert G.dtype == np.float64
    # quality: reconstruction error no worse than sklearn's from its own NNDSVDa start (+2 %)
    from sklearn.decomposition import NMF
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for n_roles in (3, 6):
            np.random.seed(1)
            G, F = factor.get_nmf_decomposition(X, n_roles)
            np.random.seed(1)
            model = NMF(n_components=n_roles, solver='mu', init='nndsvda')
            Gs = model.fit_transform(X)
            ours = np.linalg.norm(X - G @ F)
            theirs = np.linalg.norm(X - Gs @ model.components_)
            assert ours <= theirs * 1.02


def test_nndsvda_init_matches_sklearn_on_separated_spectrum():
    sk = pytest.importorskip('sklearn.decomposition._nmf')
    rng = np.random.RandomState(0)
    X = (rng.rand(400, 4) * np.array([8.0, 4.0, 2.0, 1.0])) @ rng.rand(4, 60)
    np.random.seed(3)
    W_ref, H_ref = sk._initialize_nmf(X, 3, init='nndsvda')
    np.random.seed(3)
    W, H = factor.nndsvda_init(dev(X), 3)
    np.testing.assert_allclose(W.cpu().numpy(), W_ref, rtol=2e-4, atol=1e-5)
    np.testing.assert_allclose(H.cpu().numpy(), H_ref, rtol=2e-4, atol=1e-5)


def test_nndsvda_init_tall_matrix_same_start_as_householder_qr(monkeypatch):
    """The sketches of a tall feature matrix are orthonormalised by CholeskyQR2
    (factor.orthonormal_basis); the start it leads to equals the one from Householder QR -- any
    orthonormal basis of the same range gives the same projection -- and sklearn's."""
    sk = pytest.importorskip('sklearn.decomposition._nmf')
    rng = np.random.RandomState(1)
    X = (rng.rand(6000, 5) ** 2 * np.array([9.0, 5.0, 3.0, 2.0, 1.0])) @ rng.rand(5, 40) + \
        0.01 * rng.rand(6000, 40)
    np.random.seed(7)
    W, H = factor.nndsvda_init(dev(X), 4)
    calls = []
    real_qr = torch.linalg.qr
    monkeypatch.setattr(factor, 'orthonormal_basis', lambda Y: (calls.append(1), real_qr(Y)[0])[1])
    np.random.seed(7)
    Wh, Hh = factor.nndsvda_init(dev(X), 4)
    assert calls
    np.testing.assert_allclose(W.cpu().numpy(), Wh.cpu().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(H.cpu().numpy(), Hh.cpu().numpy(), rtol=1e-4, atol=1e-5)
    np.random.seed(7)
    W_ref, H_ref = sk._initialize_nmf(X, 4, init='nndsvda')
    np.testing.assert_allclose(W.cpu().numpy(), W_ref, rtol=5e-4, atol=2e-5)
    np.testing.assert_allclose(H.cpu().numpy(), H_ref, rtol=5e-4, atol=2e-5)


def test_role_extractor_end_to_end(refex_cases):
    """RoleExtractor surface on the karate features (tests/test_roles/test_extract.py:38-75)."""
    from helpers import frame_from_json
    feats = frame_from_json(refex_cases['karate']['features'])
    rx = RoleExtractor(n_roles=3)
    rx.extract_role_factors(feats)
    assert rx.node_role_factor.shape == (34, 3)
    assert rx.role_feature_factor.shape == (3, feats.shape[1])
    assert list(rx.node_role_factor.columns) == ['role_0', 'role_1', 'role_2']
    assert set(rx.roles.keys()) == set(feats.index)
    np.testing.assert_allclose(rx.role_percentage.sum(axis=1).values, 1.0)
    rx = RoleExtractor(n_roles=None, n_role_range=(2, 4), n_bit_range=(2, 4))
    rx.extract_role_factors(feats)
    assert 2 <= rx.node_role_factor.shape[1] <= 4
