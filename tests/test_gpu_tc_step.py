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


def _model(x, H, Z, M, L, est, params, precision):
    import vaeb_b200
    return vaeb_b200.VAEB(x, False, H, Z, M, L, 0.01, est == "LA", False, params, precision=precision)


def _rand_params(D, H, Z, seed, scale):
    rng = np.random.RandomState(seed)
    return [rng.normal(0, scale, s).astype(np.float32) for s in O.param_shapes(D, H, Z, False)]


def _check(x, H, Z, M, L, est, params, precision, seed, idx=1):
    rng = np.random.RandomState(seed)
    eps = rng.normal(size=(L, M, Z)).astype(np.float32)
    m = _model(x, H, Z, M, L, est, params, precision)
    o = O.OracleVAEB(x, False, H, Z, M, L=L, estimator=est, params=params, dtype=np.float64)
    xb = x[idx * M:(idx + 1) * M]
    sg_ref, rows_ref, g_ref = o.grads(xb, eps)
    sg, rows, g = m.gradients(index=idx, eps=eps)
    if precision == "bf16x3":
        assert sg == pytest.approx(sg_ref, rel=1e-4)
        np.testing.assert_allclose(rows, rows_ref, rtol=1e-4)
        for a, b, n in zip(g, g_ref, O.param_names(False)):
            assert_close_tensor(a, b, 1e-4, floor=0.1, name="grad " + n)
    else:
        assert sg == pytest.approx(sg_ref, rel=1e-2)
        np.testing.assert_allclose(rows, rows_ref, rtol=1e-2)
        for a, b, n in zip(g, g_ref, O.param_names(False)):
            assert_close_tensor(a, b, 3e-2, floor=1.0, name="grad " + n)
    # the same numbers through the host-staged path (x mirrored per call) and through validate
    sg2, rows2, g2 = m.gradients(x=xb, eps=eps)
    np.testing.assert_array_equal(rows, rows2)
    for a, b in zip(g, g2):
        np.testing.assert_array_equal(a, b)
    v, vr = m.validate(xb, eps=eps, per_row=True)
    np.testing.assert_allclose(vr, rows, rtol=1e-6)
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


def test_tc_rejects_gaussian_decoder():
    import vaeb_b200
    with pytest.raises(ValueError, match="Bernoulli"):
        vaeb_b200.VAEB(np.zeros((8, 6), np.float32), True, 4, 2, 4, 1, 0.01, False, False, precision="bf16")


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
def test_c3_full_size_persistent_kernels_match_oracle(precision):
    """BASELINE config C3 at its full per-GPU size (M = 16384 rows, 784-500-20): at this size the activation layers
    run in the persistent tcgen05 kernel (128 x 256 tiles, two TMEM accumulators) and no fp32 activations are kept.
    Bound, per-row bounds and every gradient tensor against the fp64 oracle; the bound is also the sum of its rows
    (size-independent property)."""
    import vaeb_b200
    M, Z = 16384, 20
    x = O.synthetic_mnist(M)
    params = _rand_params(784, 500, Z, 7, 0.05)
    eps = np.random.RandomState(41).normal(size=(1, M, Z)).astype(np.float32)
    m = vaeb_b200.VAEB(x, False, 500, Z, M, 1, 0.01, False, False, params, precision=precision)
    o = O.OracleVAEB(x, False, 500, Z, M, L=1, estimator="LB", params=params, dtype=np.float64)
    sg_ref, rows_ref, g_ref = o.grads(x, eps)
    sg, rows, g = m.gradients(index=0, eps=eps)
    tol, gtol, floor = (1e-2, 3e-2, 1.0) if precision == "bf16" else (1e-4, 1e-4, 0.1)
    assert sg == pytest.approx(sg_ref, rel=tol)
    assert float(np.sum(rows, dtype=np.float64)) == pytest.approx(sg, rel=1e-5)
    np.testing.assert_allclose(rows, rows_ref, rtol=tol)
    for a, b, n in zip(g, g_ref, O.param_names(False)):
        assert_close_tensor(a, b, gtol, floor=floor, name="grad " + n)
    # one update through the same kernels moves the parameters like the oracle's Adagrad step
    before = float(m.update(0, eps=eps))
    assert before == pytest.approx(o.update(0, eps), rel=tol)
    m.close()
