"""Deeper encoders (Report/replication/replic.tex:46-57: "Increasing the depth of the encoder", 2-4 hidden layers of 500
units on MNIST).  The reference repository holds no code for the experiment (VAEB.py builds one hidden
layer), so the capability is specified by the oracle (tests/test_oracle.py pins its backward against autograd): extra
H x H layers between the first hidden layer and the heads, parameters W3_k, b3_k appended after the reference's list.
fp32 per-layer kernels; tolerances of tests/test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import vaeb_oracle as O
from tests.util import assert_close_tensor

pytestmark = pytest.mark.gpu
RTOL = 1e-4


@pytest.mark.parametrize("depth,act,continuous,est,L,D,H,Z,M", [
    (2, "tanh", False, "LB", 1, 784, 500, 10, 100),      # the report's set-up: MNIST, 500 units, Nz = 10
    (3, "tanh", False, "LB", 1, 784, 500, 10, 100),
    (4, "relu", False, "LA", 2, 40, 24, 3, 17),
    (2, "tanh", True, "LB", 1, 560, 200, 2, 100),
    (2, "tanh", False, "LB", 1, 784, 500, 10, 1152),     # large-batch latent kernels behind a deep encoder
])
def test_step_with_deeper_encoder(depth, act, continuous, est, L, D, H, Z, M):
    import vaeb_b200
    if continuous:
        x = np.clip(np.random.RandomState(3).normal(0.5, 0.2, (2 * M, D)), 0.01, 0.99).astype(np.float32)
    else:
        x = O.synthetic_mnist(2 * M, seed=6, D=D)
    rng = np.random.RandomState(41)
    params = [rng.normal(0, 0.05 if D > 100 else 0.2, s).astype(np.float32)
              for s in O.param_shapes(D, H, Z, continuous, depth)]
    eps = rng.normal(size=(L, M, Z)).astype(np.float32)
    m = vaeb_b200.VAEB(x, continuous, H, Z, M, L, 0.01, est == "LA", False, params, activation=act, encoder_layers=depth)
    assert len(m.params) == len(params) and m.params[-1].name == "b3_%d" % depth
    o = O.OracleVAEB(x, continuous, H, Z, M, L=L, estimator=est, params=params, dtype=np.float64, activation=act)
    xb = x[M:2 * M]
    sg_ref, rows_ref, g_ref = o.grads(xb, eps)
    sg, rows, g = m.gradients(index=1, eps=eps)
    assert sg == pytest.approx(sg_ref, rel=RTOL)
    np.testing.assert_allclose(rows, rows_ref, rtol=RTOL)
    names = O.param_names(continuous, depth)
    assert len(g) == len(names)
    for a, b, n in zip(g, g_ref, names):
        assert_close_tensor(a, b, RTOL, name="depth %d grad %s" % (depth, n))
    p0 = [q.copy() for q in o.params]
    for idx in (1, 0):
        assert float(m.update(idx, eps=eps)) == pytest.approx(o.update(idx, eps), rel=RTOL)
    for a, b, q0, gr, n in zip(m.get_params(), o.params, p0, g_ref, names):
        well = np.abs(gr) > 1e-2 * np.abs(gr).max()
        np.testing.assert_allclose((a - q0)[well], (b - q0)[well], rtol=5e-3, atol=5e-5, err_msg="step %s" % n)   # atol: 0.5 % of one lr-sized step (two steps may cancel)
    xv = x[:M - 3]
    ev = rng.normal(size=(L, len(xv), Z)).astype(np.float32)
    sgv_ref, rowsv_ref = o.validate(xv, ev)
    sgv, rowsv = m.validate(xv, eps=ev, per_row=True)
    assert sgv == pytest.approx(sgv_ref, rel=2e-4)
    np.testing.assert_allclose(rowsv, rowsv_ref, rtol=2e-4)
    m.close()


def test_deep_encoder_log_px_save_load_and_default_init(tmp_path):
    import vaeb_b200
    D, H, Z, n, L = 784, 500, 10, 16, 5
    x = O.synthetic_mnist(n, seed=8)
    m = vaeb_b200.VAEB(x, False, H, Z, n, 1, 0.01, False, False, encoder_layers=3)      # the library's own initialisation
    p = m.get_params()
    assert [a.shape for a in p] == [tuple(s) for s in O.param_shapes(D, H, Z, False, 3)]
    assert abs(float(np.std(p[-4])) - 0.01) < 1e-3 and not p[-1].any()                  # W3_2 ~ N(0, 0.01^2), b3_3 = 0
    eps = np.random.RandomState(2).normal(size=(n, L, Z)).astype(np.float32)
    ref, _ = O.is_log_px([q.astype(np.float64) for q in p], x.astype(np.float64), eps.astype(np.float64), False)
    np.testing.assert_allclose(m.log_px(x, L=L, eps=eps), ref, rtol=RTOL)
    f = str(tmp_path / "deep.mdl")
    m.save(f)
    m2, _ = vaeb_b200.VAEB.load(f, data=(x, x))
    assert m2.encoder_layers == 3
    for a, b in zip(m2.get_params(), p):
        np.testing.assert_array_equal(a, b)
    m.close(); m2.close()


def test_deep_encoder_needs_the_fp32_path():
    import vaeb_b200
    x = O.synthetic_mnist(64, seed=1)
    with pytest.raises(Exception):
        vaeb_b200.VAEB(x, False, 500, 20, 32, 1, 0.01, False, False, precision="bf16x3", encoder_layers=2)
