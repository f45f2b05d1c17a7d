"""The AEVB step with the wide layers on tcgen05 (precision 'bf16x3' and 'bf16') against the
fp64 oracle.  Stated tolerances (north star): bf16x3 -> the fp32 tier, 1e-4 relative on the
per-datapoint bound / SGVB and |d| <= 1e-4*max(|ref|, 0.1*||ref||_inf) on gradients (a hi+lo
bf16 pair carries 16 mantissa bits, i.e. 2^-17 = 7.6e-6 relative representation error per
operand, so entries far below the tensor's max norm are held to 1e-5 of that norm); bf16 ->
1e-2 relative on the bound (gradients within 3e-2 of the tensor's max norm)."""
import numpy as np
import pytest

from oracle import vaeb_oracle as O
from tests.util import assert_close_tensor

pytestmark = pytest.mark.gpu


def _model(x, H, Z, M, L, est, params, precision, continuous=False):
    import vaeb_b200
    return vaeb_b200.VAEB(x, continuous, H, Z, M, L, 0.01, est == "LA", False, params, precision=precision)


def _rand_params(D, H, Z, seed, scale):
    rng = np.random.RandomState(seed)
    return [rng.normal(0, scale, s).astype(np.float32) for s in O.param_shapes(D, H, Z, False)]


def _check(x, H, Z, M, L, est, params, precision, seed, idx=1, continuous=False):
    rng = np.random.RandomState(seed)
    eps = rng.normal(size=(L, M, Z)).astype(np.float32)
    m = _model(x, H, Z, M, L, est, params, precision, continuous)
    o = O.OracleVAEB(x, continuous, H, Z, M, L=L, estimator=est, params=params, dtype=np.float64)
    xb = x[idx * M:(idx + 1) * M]
    sg_ref, rows_ref, g_ref = o.grads(xb, eps)
    sg, rows, g = m.gradients(index=idx, eps=eps)
    if precision == "bf16x3":
        assert sg == pytest.approx(sg_ref, rel=1e-4)
        np.testing.assert_allclose(rows, rows_ref, rtol=1e-4)
        for a, b, n in zip(g, g_ref, O.param_names(continuous)):
            assert_close_tensor(a, b, 1e-4, floor=0.1, name="grad " + n)
    else:
        assert sg == pytest.approx(sg_ref, rel=1e-2)
        np.testing.assert_allclose(rows, rows_ref, rtol=1e-2)
        for a, b, n in zip(g, g_ref, O.param_names(continuous)):
            assert_close_tensor(a, b, 3e-2, floor=1.0, name="grad " + n)
    # the same numbers through the host-staged path (x mirrored per call) and through validate
    sg2, rows2, g2 = m.gradients(x=xb, eps=eps)
    np.testing.assert_array_equal(rows, rows2)
    for a, b in zip(g, g2):
        np.testing.assert_array_equal(a, b)
    v, vr = m.validate(xb, eps=eps, per_row=True)
    # (bf16 tier at large batch: validate runs the latent layers in fp32, the training step on the tensor cores in bf16)
    np.testing.assert_allclose(vr, rows, rtol=1e-6 if precision == "bf16x3" else 1e-3)
    ret = m.update(idx, eps=eps)
    assert float(ret) == pytest.approx(sg_ref / M, rel=1e-4 if precision == "bf16x3" else 1e-2)
    m.close()


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("est,L", [("LB", 1), ("LA", 2)])
def test_tc_step_small_ragged(precision, est, L):
    D, H, Z, M = 37, 29, 3, 11
    x = np.random.RandomState(5).uniform(size=(3 * M, D)).astype(np.float32)
    _check(x, H, Z, M, L, est, _rand_params(D, H, Z, 7, 0.3), precision, 11)


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
def test_tc_step_c2_full_width(precision):
    x = O.synthetic_mnist(300)
    _check(x, 500, 20, 100, 1, "LB", _rand_params(784, 500, 20, 3, 0.05), precision, 22, idx=2)


def test_tc_step_c2_init_weights_bf16x3():
    x = O.synthetic_mnist(200)
    _check(x, 500, 20, 100, 1, "LB", O.init_params(784, 500, 20, False), "bf16x3", 23)


def test_tc_step_large_batch_bf16x3():
    # several 128-row tiles, BN = 128 path (rows >= 1024), L = 1
    x = O.synthetic_mnist(2304)
    _check(x, 500, 20, 1152, 1, "LB", _rand_params(784, 500, 20, 4, 0.05), "bf16x3", 24, idx=1)


@pytest.mark.parametrize("precision,est,Z,M", [("fp32", "LA", 10, 1100), ("fp32", "LB", 20, 1024), ("fp32", "LB", 3, 1030),
                                              ("bf16x3", "LA", 20, 1280), ("bf16", "LB", 2, 1152)])
def test_large_batch_latent_kernels(precision, est, Z, M):
    """rows >= 1024 switches the latent layers (enc2 + reparam + dec1, their backward, the thin weight
    gradients) to the 64-row tilings of kernels_latent.cu: ragged row counts, both estimators, Z in {2,3,10,20}."""
    import vaeb_b200
    x = O.synthetic_mnist(2 * M)
    params = _rand_params(784, 500, Z, 5, 0.05)
    rng = np.random.RandomState(31)
    eps = rng.normal(size=(1, M, Z)).astype(np.float32)
    m = vaeb_b200.VAEB(x, False, 500, Z, M, 1, 0.01, est == "LA", False, params, precision=precision)
    o = O.OracleVAEB(x, False, 500, Z, M, L=1, estimator=est, params=params, dtype=np.float64)
    sg_ref, rows_ref, g_ref = o.grads(x[M:2 * M], eps)
    sg, rows, g = m.gradients(index=1, eps=eps)
    tol, gtol, floor = (1e-2, 3e-2, 1.0) if precision == "bf16" else (1e-4, 1e-4, 0.1)
    assert sg == pytest.approx(sg_ref, rel=tol)
    np.testing.assert_allclose(rows, rows_ref, rtol=tol)
    for a, b, n in zip(g, g_ref, O.param_names(False)):
        assert_close_tensor(a, b, gtol, floor=floor, name="grad " + n)
    m.close()


def test_large_batch_philox_rows_match_small_batch_rows():
    """Philox eps is keyed by the row: the first rows of a 1152-row minibatch (64-row tilings) get the same
    noise, hence the same per-row bound, as the same rows in a 384-row minibatch (small-batch kernels)."""
    import vaeb_b200
    x = O.synthetic_mnist(1152)
    params = _rand_params(784, 500, 20, 6, 0.05)
    big = vaeb_b200.VAEB(x, False, 500, 20, 1152, 1, 0.01, False, False, params)
    small = vaeb_b200.VAEB(x[:384], False, 500, 20, 384, 1, 0.01, False, False, params)
    _, rows_b, _ = big.gradients(index=0)
    _, rows_s, _ = small.gradients(index=0)
    np.testing.assert_allclose(rows_b[:384], rows_s, rtol=2e-5)
    big.close(); small.close()


def test_tc_training_tracks_fp32_training():
    import vaeb_b200
    x = O.synthetic_mnist(1000)
    out = {}
    for prec in ("fp32", "bf16x3", "bf16"):
        m = vaeb_b200.VAEB(x, False, 500, 20, 100, 1, 0.01, False, False, precision=prec)
        lb = [float(np.mean(m.update_many(np.arange(10)))) for _ in range(3)]
        out[prec] = lb
        m.close()
    assert out["fp32"][-1] > out["fp32"][0]
    np.testing.assert_allclose(out["bf16x3"], out["fp32"], rtol=2e-3)
    np.testing.assert_allclose(out["bf16"], out["fp32"], rtol=2e-2)


@pytest.mark.parametrize("precision,M,est", [("bf16x3", 100, "LB"), ("bf16x3", 4096, "LB"), ("bf16x3", 1152, "LA"),
                                              ("bf16", 1152, "LB"), ("bf16", 100, "LB")])
def test_tc_step_gaussian_decoder(precision, M, est):
    """Gaussian decoder on the tensor cores: the output layer is ONE GEMM over the interleaved
    columns [W2|W6]' (VAEB.py:257-263), log-density and both deltas in its epilogue (:304-307), dgrad over K = 2D, the
    [W2|W6]' weight gradient de-interleaved by the slice reduction.  C1's shape (Frey 560-200-2), M = 100 (layer
    launches) and large batches (every contraction on tcgen05, one weight-gradient launch up to 4096 rows)."""
    D, H, Z = 560, 200, 2
    rng = np.random.RandomState(77)
    x = np.clip(rng.normal(0.5, 0.2, (2 * M, D)), 0.01, 0.99).astype(np.float32)
    params = [rng.normal(0, 0.05, s).astype(np.float32) for s in O.param_shapes(D, H, Z, True)]
    _check(x, H, Z, M, 1, est, params, precision, 31, idx=1, continuous=True)


def test_tc_gaussian_updates_match_fp32_path():
    """Updates at the C1 shape, M = 2048: the tensor-core Gaussian path (bf16x3) lands where the fp32 kernels do -- the
    bounds of three consecutive updates, and the first Adagrad step wherever it is well conditioned (it is
    lr * g / (|g| + 1e-6): entries with |g| ~ 0 keep only a sign that rounding decides; tests/test_gpu_parity.py)."""
    import vaeb_b200
    D, H, Z, M = 560, 200, 2, 2048
    rng = np.random.RandomState(78)
    x = np.clip(rng.normal(0.5, 0.2, (2 * M, D)), 0.01, 0.99).astype(np.float32)
    params = [rng.normal(0, 0.05, s).astype(np.float32) for s in O.param_shapes(D, H, Z, True)]
    eps = rng.normal(size=(1, M, Z)).astype(np.float32)
    out = {}
    for prec in ("fp32", "bf16x3"):
        m = vaeb_b200.VAEB(x, True, H, Z, M, 1, 0.01, False, False, params, precision=prec)
        _, _, g = m.gradients(index=0, eps=eps)
        b0 = float(m.update(0, eps=eps))
        after = m.get_params()
        bounds = [b0] + [float(m.update(i % 2, eps=eps)) for i in (1, 2)]
        out[prec] = (bounds, after, g)
        m.close()
    np.testing.assert_allclose(out["bf16x3"][0], out["fp32"][0], rtol=2e-4)
    for a, b, q0, gr in zip(out["bf16x3"][1], out["fp32"][1], params, out["fp32"][2]):
        well = np.abs(gr) > 1e-2 * np.abs(gr).max()
        np.testing.assert_allclose((a - q0)[well], (b - q0)[well], rtol=5e-3, atol=1e-7)
