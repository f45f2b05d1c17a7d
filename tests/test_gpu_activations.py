"""Sigmoid and ReLU hidden layers (the alternatives to tanh that the reference's report compares,
Report/replication/replic.tex:73-82) on the fp32 per-layer kernels, against the fp64 oracle
(whose activations are pinned against torch.autograd in tests/test_oracle.py): gradients, the bound, two updates,
validate, reconstruct and the importance-sampled log p(x).  Tolerances: the fp32 tier of tests/test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import vaeb_oracle as O
from tests.util import assert_close_tensor

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _params(D, H, Z, continuous, seed, scale):
    rng = np.random.RandomState(seed)
    return [rng.normal(0, scale, s).astype(np.float32) for s in O.param_shapes(D, H, Z, continuous)]


@pytest.mark.parametrize("act", ["sigmoid", "relu"])
@pytest.mark.parametrize("continuous,est,L,D,H,Z,M", [
    (False, "LB", 1, 784, 500, 20, 100),      # C2 shape
    (True, "LB", 1, 560, 200, 2, 100),        # C1 shape, Gaussian decoder
    (False, "LA", 2, 40, 24, 3, 17),          # L > 1, ragged
    (False, "LB", 1, 784, 500, 20, 1152),     # the 64-row large-batch latent kernels
])
def test_step_with_other_activations(act, continuous, est, L, D, H, Z, M):
    import vaeb_b200
    x = O.synthetic_mnist(2 * M, seed=5, D=D)
    if continuous:
        x = np.clip(np.random.RandomState(3).normal(0.5, 0.2, (2 * M, D)), 0.01, 0.99).astype(np.float32)
    params = _params(D, H, Z, continuous, 21, 0.05 if D > 100 else 0.2)
    eps = np.random.RandomState(9).normal(size=(L, M, Z)).astype(np.float32)
    m = vaeb_b200.VAEB(x, continuous, H, Z, M, L, 0.01, est == "LA", False, params, activation=act)
    o = O.OracleVAEB(x, continuous, H, Z, M, L=L, estimator=est, params=params, dtype=np.float64, activation=act)
    xb = x[M:2 * M]
    sg_ref, rows_ref, g_ref = o.grads(xb, eps)
    sg, rows, g = m.gradients(index=1, eps=eps)
    assert sg == pytest.approx(sg_ref, rel=RTOL)
    np.testing.assert_allclose(rows, rows_ref, rtol=RTOL)
    for a, b, n in zip(g, g_ref, O.param_names(continuous)):
        assert_close_tensor(a, b, RTOL, name="%s grad %s" % (act, n))
    # two updates: the returned bounds, then the parameters where the first Adagrad steps are well conditioned
    p0 = [q.copy() for q in o.params]
    for idx in (1, 0):
        ret_ref = o.update(idx, eps)
        ret = m.update(idx, eps=eps)
        assert float(ret) == pytest.approx(ret_ref, rel=RTOL)
    for a, b, q0, gr, n in zip(m.get_params(), o.params, p0, g_ref, O.param_names(continuous)):
        well = np.abs(gr) > 1e-2 * np.abs(gr).max()
        np.testing.assert_allclose((a - q0)[well], (b - q0)[well], rtol=5e-3, atol=1e-7, err_msg="%s step %s" % (act, n))
    # validate on a ragged set with the updated parameters
    xv = x[:M - 3]
    ev = np.random.RandomState(4).normal(size=(L, len(xv), Z)).astype(np.float32)
    sgv_ref, rowsv_ref = o.validate(xv, ev)
    sgv, rowsv = m.validate(xv, eps=ev, per_row=True)
    assert sgv == pytest.approx(sgv_ref, rel=2e-4)
    np.testing.assert_allclose(rowsv, rowsv_ref, rtol=2e-4)
    m.close()


@pytest.mark.parametrize("act", ["sigmoid", "relu"])
def test_log_px_with_other_activations(act):
    import vaeb_b200
    D, H, Z, n, L = 784, 500, 20, 24, 7
    x = O.synthetic_mnist(n, seed=8)
    params = _params(D, H, Z, False, 33, 0.05)
    eps = np.random.RandomState(2).normal(size=(n, L, Z)).astype(np.float32)
    m = vaeb_b200.VAEB(x, False, H, Z, n, 1, 0.01, False, False, params, activation=act)
    p64 = O.as_dict([q.astype(np.float64) for q in params], False)
    # the estimator's specification (oracle is_log_px) with the model's activation
    xd = x.astype(np.float64)
    h = O.hidden_act(xd @ p64["W3"] + p64["b3"], act)
    mu, ls = h @ p64["W4"] + p64["b4"], h @ p64["W5"] + p64["b5"]
    logw = np.empty((n, L))
    for l in range(L):
        e = eps[:, l, :].astype(np.float64)
        z = mu + np.exp(0.5 * ls) * e
        a = O.hidden_act(z @ p64["W1"] + p64["b1"], act) @ p64["W2"] + p64["b2"]
        lp = (xd * a - O.softplus(a)).sum(1)
        logw[:, l] = lp + (-0.5 * O.LOG2PI - 0.5 * z ** 2).sum(1) - (-0.5 * O.LOG2PI - 0.5 * ls - 0.5 * e ** 2).sum(1)
    mx = logw.max(1, keepdims=True)
    ref = (mx[:, 0] + np.log(np.exp(logw - mx).sum(1))) - np.log(L)
    got = m.log_px(x, L=L, eps=eps)
    np.testing.assert_allclose(got, ref, rtol=RTOL)
    m.close()


def test_other_activations_need_the_fp32_path():
    import vaeb_b200
    x = O.synthetic_mnist(64, seed=1)
    with pytest.raises(Exception):
        vaeb_b200.VAEB(x, False, 500, 20, 32, 1, 0.01, False, False, precision="bf16x3", activation="relu")
    with pytest.raises(ValueError):
        vaeb_b200.VAEB(x, False, 500, 20, 32, 1, 0.01, False, False, activation="gelu")
