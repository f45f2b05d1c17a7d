"""Pins the CPU oracle: known-answer value from the reference, torch.autograd (fp64)
re-derivation of every objective, finite differences, RNG restatements."""
import math

import numpy as np
import pytest
import torch

from oracle import vaeb_oracle as O

torch.set_default_dtype(torch.float64)


def _rand_problem(seed, M, D, H, Z, L, continuous, scale=0.3):
    rng = np.random.RandomState(seed)
    params = [rng.normal(0, scale, s) for s in O.param_shapes(D, H, Z, continuous)]
    x = rng.uniform(size=(M, D))
    eps = rng.normal(size=(L, M, Z))
    return params, x, eps


def _torch_objective(tp, x, eps, continuous, estimator):
    """Written from the reference's Theano expressions (VAEB.py:245-346), not from the oracle."""
    names = O.param_names(continuous)
    p = dict(zip(names, tp))
    x = torch.as_tensor(x)
    h = torch.tanh(x @ p["W3"] + p["b3"])
    mu = h @ p["W4"] + p["b4"]
    ls = h @ p["W5"] + p["b5"]
    L = eps.shape[0]
    sgvb = 0
    rows = 0
    for l in range(L):
        e = torch.as_tensor(eps[l])
        z = mu + torch.exp(0.5 * ls) * e
        hd = torch.tanh(z @ p["W1"] + p["b1"])
        y = torch.sigmoid(hd @ p["W2"] + p["b2"])
        if continuous:
            lv = hd @ p["W6"] + p["b6"]
            lp = (-0.5 * math.log(2 * math.pi) - 0.5 * lv - 0.5 * (x - y) ** 2 / torch.exp(lv)).sum(1)
        else:
            lp = (x * torch.log(y) + (1 - x) * torch.log(1 - y)).sum(1)
        if estimator == "LA":
            prior = (-0.5 * math.log(2 * math.pi) - 0.5 * z ** 2).sum(1)
            logq = (-0.5 * math.log(2 * math.pi) - 0.5 * ls - 0.5 * (z - mu) ** 2 / torch.exp(ls)).sum(1)
            rows = rows + (lp + prior - logq)
        else:
            rows = rows + lp
    rows = rows / L
    if estimator == "LB":
        rows = rows + 0.5 * (1 + ls - mu ** 2 - torch.exp(ls)).sum(1)
    return rows


def test_known_answer_logpdf_main():
    # degenerate-vae/logpdf.py:119-123 -- the only known-answer input in the reference
    Y = np.array([[0, 0, 1], [0, 0, 1]], dtype=np.float64)
    P = np.array([[0.01, 0.01, 0.99], [0.01, 0.01, 0.99]])
    assert O.lpdf_bernoulli(Y, P) == pytest.approx(6 * math.log(0.99 + 1e-7), rel=1e-13)
    assert O.lpdf_bernoulli(Y, P) == pytest.approx(-0.0603014090604, rel=1e-9)


@pytest.mark.parametrize("continuous", [False, True])
@pytest.mark.parametrize("estimator", ["LB", "LA"])
@pytest.mark.parametrize("L", [1, 3])
def test_grads_match_autograd(continuous, estimator, L):
    params, x, eps = _rand_problem(1, 7, 11, 9, 4, L, continuous)
    out = O.elbo_and_grads(params, x, eps, continuous, estimator)
    tp = [torch.tensor(q, requires_grad=True) for q in params]
    rows = _torch_objective(tp, x, eps, continuous, estimator)
    crit = rows.sum() - 0.5 * sum((q ** 2).sum() for q in tp)
    crit.backward()
    assert out.sgvb == pytest.approx(rows.sum().item(), rel=1e-12)
    np.testing.assert_allclose(out.per_row, rows.detach().numpy(), rtol=1e-11, atol=1e-12)
    for g, q, n in zip(out.grads, tp, O.param_names(continuous)):
        np.testing.assert_allclose(g, q.grad.numpy(), rtol=1e-9, atol=1e-11, err_msg=n)


def test_grads_finite_differences():
    params, x, eps = _rand_problem(2, 5, 6, 5, 3, 2, True)
    out = O.elbo_and_grads(params, x, eps, True, "LB")
    rng = np.random.RandomState(0)
    for ti, q in enumerate(params):
        for _ in range(3):
            idx = tuple(rng.randint(0, s) for s in q.shape)
            h = 1e-6
            old = q[idx]
            q[idx] = old + h
            fp = O.elbo_and_grads(params, x, eps, True, "LB", want_grads=False).sgvb - 0.5 * sum((t ** 2).sum() for t in params)
            q[idx] = old - h
            fm = O.elbo_and_grads(params, x, eps, True, "LB", want_grads=False).sgvb - 0.5 * sum((t ** 2).sum() for t in params)
            q[idx] = old
            assert out.grads[ti][idx] == pytest.approx((fp - fm) / (2 * h), rel=2e-6, abs=1e-7)


def test_bernoulli_softplus_form_equals_log_sigmoid_form():
    rng = np.random.RandomState(3)
    a = rng.normal(0, 4, (5, 8)); x = rng.uniform(size=(5, 8))
    y = 1 / (1 + np.exp(-a))
    ref = (x * np.log(y) + (1 - x) * np.log(1 - y)).sum(1)
    np.testing.assert_allclose(O.log_px_given_z(x, a, None, False), ref, rtol=1e-12)


def test_adagrad_rule():
    # VAEB.py:438-442
    p = [np.array([1.0, -2.0])]; ada = [np.array([0.5, 0.0])]; g = [np.array([0.3, -0.4])]
    O.adagrad_update(p, ada, g, lr=0.01)
    np.testing.assert_allclose(ada[0], [0.5 + 0.09, 0.16])
    np.testing.assert_allclose(p[0], [1.0 + 0.01 * 0.3 / (math.sqrt(0.59) + 1e-6),
                                      -2.0 + 0.01 * -0.4 / (0.4 + 1e-6)])


def test_update_returns_pre_update_value_and_matches_autograd_step():
    M, D, H, Z = 6, 10, 8, 3
    params, x, eps = _rand_problem(4, 2 * M, D, H, Z, 1, False)
    eps = eps[:, :M]
    m = O.OracleVAEB(x, False, H, Z, M, params=params)
    before = [q.copy() for q in m.params]
    val = m.update(1, eps)
    tp = [torch.tensor(q, requires_grad=True) for q in before]
    rows = _torch_objective(tp, x[M:2 * M], eps, False, "LB")
    (rows.sum() - 0.5 * sum((q ** 2).sum() for q in tp)).backward()
    assert val == pytest.approx(rows.sum().item() / M, rel=1e-12)
    for q_new, q_old, t in zip(m.params, before, tp):
        g = t.grad.numpy()
        np.testing.assert_allclose(q_new, q_old + 0.01 * g / (np.sqrt(g * g) + 1e-6), rtol=1e-10)


def test_fullbayes_variant_matches_autograd():
    # VAEBfullbayes.py:139-145,183-184: mean objective, no prior, -lr*1e-6*p^2
    M, D, H, Z = 5, 9, 7, 3
    params, x, eps = _rand_problem(5, M, D, H, Z, 1, True)
    m = O.OracleVAEB(x, True, H, Z, M, params=params, variant="fullbayes")
    before = [q.copy() for q in m.params]
    val = m.update(0, eps)
    tp = [torch.tensor(q, requires_grad=True) for q in before]
    obj = _torch_objective(tp, x, eps, True, "LB").mean()
    obj.backward()
    assert val == pytest.approx(obj.item(), rel=1e-12)
    for q_new, q_old, t in zip(m.params, before, tp):
        g = t.grad.numpy()
        np.testing.assert_allclose(q_new, q_old + 0.01 * g / (np.sqrt(g * g) + 1e-6) - 0.01 * 1e-6 * q_old ** 2,
                                   rtol=1e-9, atol=1e-14)


def test_fvb_faithful_only_prior_terms_train():
    # data term has no gradient w.r.t. the variational parameters
    M, D, H, Z = 4, 8, 6, 2
    params, x, eps = _rand_problem(6, M, D, H, Z, 1, False)
    m = O.OracleVAEB(x, False, H, Z, M, params=params, estimator="FVB")
    sgvb, _, grads = m.grads(x, eps)
    tf = [torch.tensor(q, requires_grad=True) for q in m.fvp]
    tp = [torch.tensor(q) for q in params]
    rows = _torch_objective(tp, x, eps, False, "LB")
    tprior = sum(0.5 * torch.sum(1 + torch.log(tf[i + 1] ** 2) - tf[i] ** 2 - tf[i + 1] ** 2) for i in range(0, len(tf), 2))
    val = M * rows.sum() + tprior
    (val - 0.5 * sum((q ** 2).sum() for q in tf)).backward()
    assert sgvb == pytest.approx(val.item(), rel=1e-12)
    for g, t in zip(grads, tf):
        np.testing.assert_allclose(g, t.grad.numpy(), rtol=1e-10)
    with pytest.raises(ValueError):
        O.OracleVAEB(x, False, H, Z, M, L=2, params=params, estimator="FVB").grads(x, np.zeros((2, M, Z)))


def test_fvb_sampled_matches_autograd():
    M, D, H, Z = 4, 8, 6, 2
    params, x, eps = _rand_problem(7, M, D, H, Z, 1, True)
    m = O.OracleVAEB(x, True, H, Z, M, params=params, estimator="FVB_SAMPLED")
    rng = np.random.RandomState(1)
    for i in range(1, len(m.fvp), 2):
        m.fvp[i] = rng.normal(0, 0.05, m.fvp[i].shape)      # mixed-sign sigmas exercise |sigma|
    zeta = [rng.normal(size=q.shape) for q in params]
    sgvb, _, grads = m.grads(x, eps, zeta)
    tf = [torch.tensor(q, requires_grad=True) for q in m.fvp]
    theta = [tf[2 * i] + torch.sqrt(tf[2 * i + 1] ** 2) * torch.tensor(zeta[i]) for i in range(len(params))]
    rows = _torch_objective(theta, x, eps, True, "LB")
    tprior = sum(0.5 * torch.sum(1 + torch.log(tf[i + 1] ** 2) - tf[i] ** 2 - tf[i + 1] ** 2) for i in range(0, len(tf), 2))
    val = M * rows.sum() + tprior
    (val - 0.5 * sum((q ** 2).sum() for q in tf)).backward()
    assert sgvb == pytest.approx(val.item(), rel=1e-12)
    for g, t in zip(grads, tf):
        np.testing.assert_allclose(g, t.grad.numpy(), rtol=1e-9, atol=1e-10)


def test_is_estimator_definition_and_bound():
    N, D, H, Z, L = 5, 9, 7, 3, 64
    params, x, _ = _rand_problem(8, N, D, H, Z, 1, False)
    eps = np.random.RandomState(2).normal(size=(N, L, Z))
    logp, logw = O.is_log_px(params, x, eps, False)
    # each log w equals the getLA integrand (VAEB.py:319-327) for that sample
    la_rows = O.elbo_and_grads(params, x, eps[:, 3, :][None], False, "LA", want_grads=False).per_row
    np.testing.assert_allclose(logw[:, 3], la_rows, rtol=1e-12)
    ref = torch.logsumexp(torch.tensor(logw), 1).numpy() - math.log(L)
    np.testing.assert_allclose(logp, ref, rtol=1e-12)
    assert np.all(logp >= logw.mean(1) - 1e-12)   # Jensen: IS estimate >= mean log-weight


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    z = O.philox4x32_10(np.zeros((1, 4), np.uint32), np.zeros((1, 2), np.uint32))[0]
    assert [hex(v) for v in z] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    f = np.full((1, 4), 0xFFFFFFFF, np.uint32)
    r = O.philox4x32_10(f, np.full((1, 2), 0xFFFFFFFF, np.uint32))[0]
    assert [hex(v) for v in r] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    pi = O.philox4x32_10(np.array([[0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]], np.uint32),
                         np.array([[0xa4093822, 0x299f31d0]], np.uint32))[0]
    assert [hex(v) for v in pi] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_philox_normal_moments_and_determinism():
    a = O.philox_normal(10, 0, 3, 200001)
    b = O.philox_normal(10, 0, 3, 1000)
    np.testing.assert_array_equal(a[:1000], b)
    assert abs(a.mean()) < 0.01 and abs(a.std() - 1) < 0.01
    assert not np.array_equal(a[:1000], O.philox_normal(10, 0, 4, 1000))
    assert not np.array_equal(a[:1000], O.philox_normal(10, 1, 3, 1000))


def test_init_params_draw_order():
    # VAEB.py:58-109: two discarded draws, then W3,W4,W5,W1,W2,(W6); biases zero
    D, H, Z = 12, 7, 3
    ps = O.init_params(D, H, Z, True)
    r = np.random.RandomState(10)
    r.normal(0, 0.01, (D, H)); r.normal(0, 0.01, (H, Z))
    exp = [r.normal(0, 0.01, s).astype(np.float32) for s in [(D, H), (H, Z), (H, Z), (Z, H), (H, D), (H, D)]]
    for a, b in zip(ps[:6], exp):
        np.testing.assert_array_equal(a, b)
    assert all(np.all(q == 0) for q in ps[6:]) and all(q.dtype == np.float32 for q in ps)


def test_ae_primitives():
    rng = np.random.RandomState(0)
    x = rng.normal(size=(4, 5)); Ws = [rng.normal(size=(5, 6)), rng.normal(size=(6, 3))]
    bs = [rng.normal(size=6), rng.normal(size=3)]
    np.testing.assert_allclose(O.construct_mlp(x, Ws, bs), np.tanh(np.tanh(x @ Ws[0] + bs[0]) @ Ws[1] + bs[1]))
    th = [rng.normal(size=(3, 2))]
    assert O.normal_prior(th, 2.0) == pytest.approx(-0.5 * (np.sum(th[0] ** 2) / 2 + 6 * math.log(4 * math.pi)))
    mu0, mu1 = rng.normal(size=4), rng.normal(size=4)
    s0, s1 = rng.uniform(0.5, 2, 4), rng.uniform(0.5, 2, 4)
    kl = torch.distributions.kl_divergence(
        torch.distributions.Normal(torch.tensor(mu0), torch.tensor(s0).sqrt()),
        torch.distributions.Normal(torch.tensor(mu1), torch.tensor(s1).sqrt())).sum().item()
    assert O.gauss_dkl(mu0, s0, mu1, s1) == pytest.approx(kl, rel=1e-12)
    Y = rng.normal(size=(3, 4)); mu = rng.normal(size=(3, 4)); l2 = rng.normal(size=(3, 4))
    ref = torch.distributions.Normal(torch.tensor(mu), torch.tensor(np.exp(0.5 * l2))).log_prob(torch.tensor(Y)).sum().item()
    assert O.lpdf_indep_normal(Y, mu, l2) == pytest.approx(ref, rel=1e-12)


@pytest.mark.parametrize("act", ["sigmoid", "relu"])
@pytest.mark.parametrize("continuous", [False, True])
@pytest.mark.parametrize("est", ["LB", "LA"])
def test_other_hidden_activations_match_autograd(act, continuous, est):
    """Sigmoid / ReLU hidden layers (Report/replication/replic.tex:73-82): the oracle's hand-derived backward against
    torch.autograd in float64."""
    torch = pytest.importorskip("torch")
    rng = np.random.RandomState(0)
    D, H, Z, M, L = 12, 9, 3, 7, 2
    params = [rng.normal(0, 0.3, s) for s in O.param_shapes(D, H, Z, continuous)]
    x = rng.uniform(0, 1, (M, D))
    eps = rng.normal(size=(L, M, Z))
    out = O.elbo_and_grads(params, x, eps, continuous, est, True, act=act)
    tp = [torch.tensor(p, requires_grad=True) for p in params]
    P = dict(zip(O.param_names(continuous), tp))
    f = {"sigmoid": torch.sigmoid, "relu": torch.relu}[act]
    xt = torch.tensor(x)
    he = f(xt @ P["W3"] + P["b3"])
    mu, ls = he @ P["W4"] + P["b4"], he @ P["W5"] + P["b5"]
    tot = 0
    for l in range(L):
        e = torch.tensor(eps[l])
        z = mu + torch.exp(0.5 * ls) * e
        hd = f(z @ P["W1"] + P["b1"])
        a = hd @ P["W2"] + P["b2"]
        if continuous:
            lv = hd @ P["W6"] + P["b6"]
            lp = (-0.5 * O.LOG2PI - 0.5 * lv - 0.5 * (xt - torch.sigmoid(a)) ** 2 / torch.exp(lv)).sum(1)
        else:
            lp = (xt * a - torch.nn.functional.softplus(a)).sum(1)
        if est == "LA":
            pr = (-0.5 * O.LOG2PI - 0.5 * z ** 2).sum(1)
            lq = (-0.5 * O.LOG2PI - 0.5 * ls - 0.5 * (z - mu) ** 2 / torch.exp(ls)).sum(1)
            tot = tot + (lp + pr - lq).sum() / L
        else:
            tot = tot + lp.sum() / L
    if est == "LB":
        tot = tot + 0.5 * (1 + ls - mu ** 2 - torch.exp(ls)).sum()
    (tot - 0.5 * sum((p ** 2).sum() for p in tp)).backward()
    assert abs(out.sgvb - float(tot.detach())) < 1e-9
    for g, t in zip(out.grads, tp):
        np.testing.assert_allclose(g, t.grad.numpy(), atol=1e-9)


@pytest.mark.parametrize("depth", [2, 3, 4])
@pytest.mark.parametrize("continuous,est", [(False, "LB"), (True, "LA")])
def test_deeper_encoders_match_autograd(depth, continuous, est):
    """Deeper encoders (Report/replication/replic.tex:46-57: extra H x H hidden layers before the heads; the reference holds
    no code for them): the oracle's hand-derived backward against torch.autograd in float64."""
    rng = np.random.RandomState(1)
    D, H, Z, M, L = 12, 9, 3, 7, 2
    params = [rng.normal(0, 0.3, s) for s in O.param_shapes(D, H, Z, continuous, depth)]
    x = rng.uniform(0, 1, (M, D))
    eps = rng.normal(size=(L, M, Z))
    out = O.elbo_and_grads(params, x, eps, continuous, est, True)
    tp = [torch.tensor(p, requires_grad=True) for p in params]
    P = dict(zip(O.param_names(continuous, depth), tp))
    xt = torch.tensor(x)
    he = torch.tanh(xt @ P["W3"] + P["b3"])
    for k in range(2, depth + 1):
        he = torch.tanh(he @ P["W3_%d" % k] + P["b3_%d" % k])
    mu, ls = he @ P["W4"] + P["b4"], he @ P["W5"] + P["b5"]
    tot = 0
    for l in range(L):
        e = torch.tensor(eps[l])
        z = mu + torch.exp(0.5 * ls) * e
        hd = torch.tanh(z @ P["W1"] + P["b1"])
        a = hd @ P["W2"] + P["b2"]
        if continuous:
            lv = hd @ P["W6"] + P["b6"]
            lp = (-0.5 * O.LOG2PI - 0.5 * lv - 0.5 * (xt - torch.sigmoid(a)) ** 2 / torch.exp(lv)).sum(1)
        else:
            lp = (xt * a - torch.nn.functional.softplus(a)).sum(1)
        if est == "LA":
            pr = (-0.5 * O.LOG2PI - 0.5 * z ** 2).sum(1)
            lq = (-0.5 * O.LOG2PI - 0.5 * ls - 0.5 * (z - mu) ** 2 / torch.exp(ls)).sum(1)
            tot = tot + (lp + pr - lq).sum() / L
        else:
            tot = tot + lp.sum() / L
    if est == "LB":
        tot = tot + 0.5 * (1 + ls - mu ** 2 - torch.exp(ls)).sum()
    (tot - 0.5 * sum((p ** 2).sum() for p in tp)).backward()
    assert len(out.grads) == len(tp)
    assert abs(out.sgvb - float(tot.detach())) < 1e-9
    for g, t in zip(out.grads, tp):
        np.testing.assert_allclose(g, t.grad.numpy(), atol=1e-9)
