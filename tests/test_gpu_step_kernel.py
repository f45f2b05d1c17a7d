"""The tensor-core single-launch step kernel (vaeb_b200/csrc/step_tc.cu): the update of VAEB.py:408-415 at M <= 128,
against the fp64 oracle, for the configurations the other files do not pin it on:
  * it IS the kernel that runs (vaeb_step_kernel) and an update is one launch;
  * C1 = BASELINE configs[0] at its size (Frey Face 560-200-2, Gaussian decoder: [W2|W6] as one layer of virtual
    columns, two-pass dgrad), L^B and L^A, three consecutive updates, every parameter and accumulator;
  * ragged shapes (minibatch not a multiple of 8 / 16, H not a multiple of 64, odd Z, a single row);
  * the overflow hand-over of the Gaussian decoder: a step whose deltas leave the fp16 operand range is finished by
    the fp32 kernel, with the bounds and parameters the oracle gives;
  * update_many == a sequence of update() calls, bit for bit (Philox noise).
Tolerances: bounds 1e-4 relative; tensors |d| <= 1e-4 * max(|ref|, 0.05 ||ref||_inf) (the fp32 tier)."""
import numpy as np
import pytest

from oracle import vaeb_oracle as O
from tests.util import assert_close_tensor, frey_trained_params

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _rand_params(D, H, Z, cont, seed, scale):
    rng = np.random.RandomState(seed)
    return [rng.normal(0, scale, s).astype(np.float32) for s in O.param_shapes(D, H, Z, cont)]


def _model(x, cont, H, Z, M, est, params, **kw):
    import vaeb_b200
    return vaeb_b200.VAEB(x, cont, H, Z, M, 1, 0.01, est == "LA", False, params, **kw)


def _three_updates(x, cont, H, Z, M, est, params, seed):
    m = _model(x, cont, H, Z, M, est, params)
    assert m.step_kernel_name().startswith("step_tc_kernel"), m.step_kernel_name()
    o = O.OracleVAEB(x, cont, H, Z, M, L=1, estimator=est, params=params, dtype=np.float64)
    rng = np.random.RandomState(seed)
    names = O.param_names(cont)
    well = [np.ones(p.shape, bool) for p in params]
    p0 = [p.astype(np.float64) for p in params]
    for step in range(3):
        eps = rng.normal(size=(1, M, Z)).astype(np.float32)
        idx = step % (x.shape[0] // M)
        _, _, g_ref = o.grads(x[idx * M:(idx + 1) * M], eps)
        for w, g in zip(well, g_ref):
            w &= np.abs(g) > 1e-3 * np.abs(g).max()
        l0 = m.launch_count()
        got, ref = float(m.update(idx, eps=eps)), o.update(idx, eps)
        assert got == pytest.approx(ref, rel=RTOL), (est, step)
        if step:
            assert m.launch_count() - l0 == 1          # (the first call also builds the operand mirrors)
    # Adagrad normalises a step to +-lr whatever |g| is: compare where the gradient is determined at every step
    for a, b, q, w, n in zip(m.get_params(), o.params, p0, well, names):
        np.testing.assert_allclose((a - q)[w], (b - q)[w], rtol=2e-3, atol=2e-6, err_msg="step " + n)
    for a, b, w, n in zip(m._get_buffer(1), o.ada, well, names):
        # ADA = sum of g^2: at the mask's edge (|g| = 1e-3 ||g||_inf) the gradient tolerance 1e-4 * 0.05 ||g||_inf is
        # 5e-3 of the entry, i.e. 1e-2 of its square
        np.testing.assert_allclose(a[w], b[w], rtol=1e-2, err_msg="ADA " + n)
    m.close()


@pytest.mark.parametrize("est", ["LB", "LA"])
def test_step_kernel_c1_frey_gaussian_decoder(est):
    x = O.synthetic_frey(300)
    _three_updates(x, True, 200, 2, 100, est, _rand_params(560, 200, 2, True, 5, 0.05), 31)


def test_step_kernel_c1_gradient_of_one_update():
    """After ONE update from zero accumulators ADA = g^2: the gradient the kernel used, tensor by tensor."""
    x = O.synthetic_frey(200)
    params = _rand_params(560, 200, 2, True, 6, 0.08)
    m = _model(x, True, 200, 2, 100, "LB", params)
    o = O.OracleVAEB(x, True, 200, 2, 100, L=1, params=params, dtype=np.float64)
    eps = np.random.RandomState(7).normal(size=(1, 100, 2)).astype(np.float32)
    _, _, g_ref = o.grads(x[100:200], eps)
    assert float(m.update(1, eps=eps)) == pytest.approx(o.update(1, eps), rel=RTOL)
    for a, g, n in zip(m._get_buffer(1), g_ref, O.param_names(True)):
        assert_close_tensor(np.sqrt(a.astype(np.float64)), np.abs(g), RTOL, name="sqrt(ada) " + n)
    m.close()


@pytest.mark.parametrize("cont,D,H,Z,M", [(False, 64, 64, 1, 1), (False, 200, 68, 3, 37), (True, 72, 100, 5, 128),
                                           (False, 1024, 512, 20, 100), (True, 512, 256, 7, 99)])
def test_step_kernel_ragged_shapes(cont, D, H, Z, M):
    rng = np.random.RandomState(D + H)
    x = rng.uniform(size=(2 * M, D)).astype(np.float32)
    params = _rand_params(D, H, Z, cont, 9, 0.1)
    m = _model(x, cont, H, Z, M, "LB", params)
    assert m.step_kernel_name().startswith("step_tc_kernel")
    o = O.OracleVAEB(x, cont, H, Z, M, L=1, params=params, dtype=np.float64)
    for step in range(2):
        eps = rng.normal(size=(1, M, Z)).astype(np.float32)
        _, _, g_ref = o.grads(x[step * M:(step + 1) * M], eps)
        ada0 = [a.astype(np.float64) for a in m._get_buffer(1)]
        assert float(m.update(step, eps=eps)) == pytest.approx(o.update(step, eps), rel=RTOL)
        # the gradient of THIS step: ADA grows by g^2 (the oracle's own second gradient is taken at the oracle's
        # parameters, which can differ by 2 lr wherever a first-step gradient was ~0 -- hence only step 0 is pinned)
        if step == 0:
            for a, a0, g, n in zip(m._get_buffer(1), ada0, g_ref, O.param_names(cont)):
                assert_close_tensor(np.sqrt(np.maximum(a.astype(np.float64) - a0, 0)), np.abs(g), RTOL, name="|g| " + n)
    m.close()


def test_step_kernel_gaussian_overflow_is_finished_by_the_fp32_kernel():
    """The reference's trained Frey weights at lr = 0.01 explode at the second update (bound -2.5e7, deltas > 1e8): the
    launch stops before that update's first parameter write and the fp32 kernel takes the step."""
    x = O.synthetic_frey(300)
    params = frey_trained_params()
    m = _model(x[:200], True, 200, 2, 100, "LB", params)
    assert m.step_kernel_name().startswith("step_tc_kernel")
    o = O.OracleVAEB(x[:200], True, 200, 2, 100, params=params)
    rng = np.random.RandomState(1)
    for i in range(2):
        eps = rng.normal(size=(1, 100, 2)).astype(np.float32)
        assert float(m.update(i, eps=eps)) == pytest.approx(o.update(i, eps), rel=RTOL)
    eps = rng.normal(size=(1, 100, 2)).astype(np.float32)
    assert float(m.validate(x[200:], eps=eps)) == pytest.approx(o.validate(x[200:], eps)[0], rel=RTOL)
    # and inside ONE launch of many updates: same bounds as update by update (Philox noise on both sides)
    m1 = _model(x[:200], True, 200, 2, 100, "LB", params, seed=3)
    m2 = _model(x[:200], True, 200, 2, 100, "LB", params, seed=3)
    order = np.array([0, 1, 0, 1], np.int32)
    a = m1.update_many(order)
    b = np.array([float(m2.update(int(i))) for i in order], np.float32)
    np.testing.assert_allclose(a, b, rtol=1e-5)
    m.close(); m1.close(); m2.close()


@pytest.mark.parametrize("cont", [False, True])
def test_step_kernel_update_many_equals_updates(cont):
    D, H, Z, M = (560, 200, 2, 100) if cont else (784, 500, 20, 100)
    x = O.synthetic_frey(600) if cont else O.synthetic_mnist(600)
    params = _rand_params(D, H, Z, cont, 11, 0.05)
    m1 = _model(x, cont, H, Z, M, "LB", params, seed=5)
    m2 = _model(x, cont, H, Z, M, "LB", params, seed=5)
    order = np.array([3, 0, 5, 1, 1, 4, 2], np.int32)
    a = m1.update_many(order)
    b = np.array([float(m2.update(int(i))) for i in order], np.float32)
    np.testing.assert_array_equal(a, b)
    for p, q in zip(m1.get_params(), m2.get_params()):
        np.testing.assert_array_equal(p, q)
    m1.close(); m2.close()
