"""CPU oracle for the AEVB hot path of budzianowski/VAEB -- TEST INFRASTRUCTURE ONLY.

This file is a plain-numpy restatement of the reference's arithmetic.  It is the
checker for the CUDA path, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm may
import it.  Nothing under ``vaeb_b200/`` imports this module.

PINNED BY THE REFERENCE'S OWN SOURCE: the reference ships no tests, golden input/output
pairs or known-answer vectors for this path and is Python 2 +
Theano, which cannot run in this image -- so tests/golden/make_reference_golden.py
executes the reference's files where they lie (VAEB.py, VAEBfullbayes.py,
degenerate-vae/{logpdf,mlp,infalg}.py) on a lazy-graph Theano stand-in
(tests/golden/theano_shim.py, torch CPU float64) and commits inputs, the noise each call
drew and all outputs as tests/golden/ref_*.npz.  tests/test_reference_golden.py holds
this oracle to those vectors (<= 1e-9 on every bound, parameter and accumulator: both
decoders x both estimators x L in {1,2}, full-VB, the fullbayes variant, AdaDelta, the
reference initialisation, the shipped trained Frey weights).  Unpinned remains only
Theano's own arithmetic under the calls the stand-in substitutes, and the capabilities
the reference does not have (IS estimator, sampled-weights full-VB: specified here).
Further checks:
  * the one known-answer input in the reference, ``degenerate-vae/logpdf.py:119-123``
    (6*ln(0.99+1e-7), now also produced by the reference's own ``bernoulli``);
  * an independent ``torch.autograd`` (CPU, fp64) re-derivation of every objective
    (tests/test_oracle.py) and central finite differences;
  * the trained fp32 Frey weights shipped in ``reconstruction_res/*.mdl`` as
    realistic-scale parameters (tests/golden/, extracted by
    tests/golden/make_golden.py).

Every function cites the reference ``file:line`` it restates (paths relative to the
reference root).  Arithmetic runs in the dtype of the inputs (float64 for the
oracle proper, float32 for the timed CPU baseline).

Conventions (VAEB.py): x[M,D]; parameters in the reference's list order
``[W3,W4,W5,W1,W2,(W6),b3,b4,b5,b1,b2,(b6)]`` (VAEB.py:111-115), weights stored
``[in,out]`` row-major, ``log_sigma`` is a log-VARIANCE (VAEB.py:45,343).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

LOG2PI = math.log(2.0 * math.pi)

PARAM_NAMES_DISCRETE = ["W3", "W4", "W5", "W1", "W2", "b3", "b4", "b5", "b1", "b2"]
PARAM_NAMES_CONTINUOUS = ["W3", "W4", "W5", "W1", "W2", "W6", "b3", "b4", "b5", "b1", "b2", "b6"]


def param_names(continuous, encoder_layers=1):
    """Order of ``self.params`` -- VAEB.py:111-115.  Deeper encoders (Report/replication/replic.tex:46-57; no code in the
    reference): the extra layers' W3_k, then b3_k, k = 2.., follow the reference's list."""
    names = list(PARAM_NAMES_CONTINUOUS if continuous else PARAM_NAMES_DISCRETE)
    names += ["W3_%d" % k for k in range(2, encoder_layers + 1)]
    names += ["b3_%d" % k for k in range(2, encoder_layers + 1)]
    return names


def param_shapes(D, H, Z, continuous, encoder_layers=1):
    """Shapes behind VAEB.py:58-109."""
    s = {"W3": (D, H), "W4": (H, Z), "W5": (H, Z), "W1": (Z, H), "W2": (H, D),
         "b3": (H,), "b4": (Z,), "b5": (Z,), "b1": (H,), "b2": (D,)}
    if continuous:
        s["W6"] = (H, D)
        s["b6"] = (D,)
    for k in range(2, encoder_layers + 1):
        s["W3_%d" % k] = (H, H)
        s["b3_%d" % k] = (H,)
    return [s[n] for n in param_names(continuous, encoder_layers)]


def as_dict(params, continuous):
    base = len(PARAM_NAMES_CONTINUOUS if continuous else PARAM_NAMES_DISCRETE)
    layers = 1 + (len(params) - base) // 2          # the list's length says how deep the encoder is
    return dict(zip(param_names(continuous, layers), params))


def encoder_depth(p):
    return 1 + sum(1 for n in p if n.startswith("W3_"))


# --------------------------------------------------------------------------------------
# a1  initialisation -- VAEB.py:50-125
# --------------------------------------------------------------------------------------
def init_params(D, H, Z, continuous, dtype=np.float32, prng=None, sigma_init=0.01):
    """``initialize_params`` (VAEB.py:50-115).  ``RandomState(10)`` and sigma 0.01 are
    forced by the constructor regardless of its arguments (VAEB.py:148-149).  Draw order:
    W3,W4 are drawn and then overwritten by a duplicated block (VAEB.py:58-67), so the
    stream is W3',W4',W3,W4,W5,W1,W2,(W6); biases are zeros (VAEB.py:53)."""
    prng = np.random.RandomState(10) if prng is None else prng

    def init_w(din, dout):
        return prng.normal(0, sigma_init, (din, dout)).astype(dtype)

    init_w(D, H)  # discarded W3 (VAEB.py:58)
    init_w(H, Z)  # discarded W4 (VAEB.py:64)
    p = {}
    p["W3"] = init_w(D, H)
    p["W4"] = init_w(H, Z)
    p["W5"] = init_w(H, Z)
    p["W1"] = init_w(Z, H)
    p["W2"] = init_w(H, D)
    if continuous:
        p["W6"] = init_w(H, D)
    for n, shp in zip(param_names(continuous), param_shapes(D, H, Z, continuous)):
        if n.startswith("b"):
            p[n] = np.zeros(shp, dtype=dtype)
    return [p[n] for n in param_names(continuous)]


def init_full_variational(params, sigma_vb_init=1e-3):
    """VAEB.py:117-125: interleaved ``[mu0, sigma0, mu1, sigma1, ...]`` with
    ``mu = param`` and ``sigma = 1e-3 * ones`` (fullVBSigmaInit, VAEB.py:146)."""
    out = []
    for p in params:
        out.append(np.array(p, copy=True))
        out.append(np.full_like(p, sigma_vb_init))
    return out


# --------------------------------------------------------------------------------------
# Theano RandomStreams emulation (best effort; recalled third-party behaviour)
# --------------------------------------------------------------------------------------
class TheanoRandomStreams:
    """``T.shared_randomstreams.RandomStreams(seed)`` as used at VAEB.py:158,42.
    ``gen_seedgen = RandomState(seed)``; every ``srng.normal`` node owns
    ``RandomState(gen_seedgen.randint(2**30))``; each call of the compiled function draws
    ``rs.normal(0, 1, size)`` in fp64 and casts to floatX.  ``update`` and ``validate``
    share the node states, so draws interleave in call order."""

    def __init__(self, seed=10, n_nodes=1):
        gen = np.random.RandomState(seed)
        self.nodes = [np.random.RandomState(int(gen.randint(2 ** 30))) for _ in range(n_nodes)]

    def normal(self, node, shape, dtype=np.float32):
        return self.nodes[node].normal(0.0, 1.0, size=shape).astype(dtype)

    def draw(self, M, Z, dtype=np.float32):
        """eps[L,M,Z] for one function call (one draw per reparam node, loop order)."""
        return np.stack([self.normal(l, (M, Z), dtype) for l in range(len(self.nodes))])


# --------------------------------------------------------------------------------------
# Philox4x32-10 + Box-Muller: the on-device eps generator, restated bit-exactly for uint32
# (new; not in the reference -- SURVEY 7 "RNG equivalence")
# --------------------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(ctr, key):
    """ctr: uint32[...,4], key: uint32[...,2] -> uint32[...,4] (Salmon et al. 2011)."""
    c = [np.asarray(ctr[..., i], dtype=np.uint32).copy() for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32).copy()
    k1 = np.asarray(key[..., 1], dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PHILOX_M0 * c[0].astype(np.uint64)
            p1 = _PHILOX_M1 * c[2].astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = p0.astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = p1.astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = (k0 + _PHILOX_W0).astype(np.uint32)
            k1 = (k1 + _PHILOX_W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def philox_normal(seed, stream, step, n_elems, sample=0):
    """fp32 N(0,1) for flat elements 0..n_elems-1 of (stream, step, sample).

    Counter layout (shared with vaeb_b200/csrc/philox.cuh): group g = elem // 4;
    ctr = (g_lo, g_hi | stream << 24, sample, step); key = (seed_lo, seed_hi).  The four
    outputs map to uniforms u = ((r >> 8) + 0.5) * 2**-24 in (0,1); Box-Muller pairs
    (u0,u1) -> (n0,n1), (u2,u3) -> (n2,n3); element e takes n[e % 4]."""
    n_groups = (n_elems + 3) // 4
    g = np.arange(n_groups, dtype=np.uint64)
    ctr = np.empty((n_groups, 4), dtype=np.uint32)
    ctr[:, 0] = (g & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr[:, 1] = ((g >> np.uint64(32)).astype(np.uint32) & np.uint32(0x00FFFFFF)) | np.uint32((stream & 0xFF) << 24)
    ctr[:, 2] = np.uint32(sample & 0xFFFFFFFF)
    ctr[:, 3] = np.uint32(step & 0xFFFFFFFF)
    key = np.empty((n_groups, 2), dtype=np.uint32)
    key[:, 0] = np.uint32(seed & 0xFFFFFFFF)
    key[:, 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    r = philox4x32_10(ctr, key)
    u = ((r >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)
    rad0 = np.sqrt(np.float32(-2.0) * np.log(u[:, 0]))
    rad1 = np.sqrt(np.float32(-2.0) * np.log(u[:, 2]))
    th0 = np.float32(2.0 * math.pi) * u[:, 1]
    th1 = np.float32(2.0 * math.pi) * u[:, 3]
    n = np.stack([rad0 * np.cos(th0), rad0 * np.sin(th0), rad1 * np.cos(th1), rad1 * np.sin(th1)], axis=-1)
    return n.reshape(-1)[:n_elems].astype(np.float32)


# --------------------------------------------------------------------------------------
# forward pieces
# --------------------------------------------------------------------------------------
def softplus(a):
    """log(1+exp(a)), the stable form Theano's rewrite of log(sigmoid) produces."""
    return np.maximum(a, 0) + np.log1p(np.exp(-np.abs(a)))


def sigmoid(a):
    return 0.5 * (np.tanh(0.5 * a) + 1.0)


def hidden_act(a, act="tanh"):
    """Hidden-layer activation: tanh in the reference (VAEB.py:246,254); sigmoid / ReLU are the alternatives its
    report compares (Report/replication/replic.tex:73-82)."""
    if act == "tanh":
        return np.tanh(a)
    if act == "sigmoid":
        return sigmoid(a)
    if act == "relu":
        return np.maximum(a, 0)
    raise ValueError("unknown activation %r" % (act,))


def hidden_act_grad(h, act="tanh"):
    """Derivative of the activation in terms of its VALUE h."""
    if act == "tanh":
        return 1.0 - h ** 2
    if act == "sigmoid":
        return h * (1.0 - h)
    if act == "relu":
        return (h > 0).astype(h.dtype)
    raise ValueError("unknown activation %r" % (act,))


def encoder_hiddens(p, x, act="tanh"):
    """Activations of every encoder hidden layer (one in VAEB.py:246; deeper encoders: replic.tex:46-57)."""
    hs = [hidden_act(x @ p["W3"] + p["b3"], act)]
    for k in range(2, encoder_depth(p) + 1):
        hs.append(hidden_act(hs[-1] @ p["W3_%d" % k] + p["b3_%d" % k], act))
    return hs


def encoder(p, x, act="tanh"):
    """VAEB.py:245-251."""
    h = encoder_hiddens(p, x, act)[-1]
    mu = h @ p["W4"] + p["b4"]
    ls = h @ p["W5"] + p["b5"]
    return h, mu, ls


def reparam(mu, ls, eps):
    """VAEB.py:41-47: z = mu + exp(0.5*log_sigma)*eps."""
    return mu + np.exp(0.5 * ls) * eps


def decoder(p, z, continuous, act="tanh"):
    """VAEB.py:253-265.  Returns (h, a, lv): a is the PRE-sigmoid activation
    (y = sigmoid(a)); lv = h.W6+b6 only for the continuous model."""
    h = hidden_act(z @ p["W1"] + p["b1"], act)
    a = h @ p["W2"] + p["b2"]
    lv = (h @ p["W6"] + p["b6"]) if continuous else None
    return h, a, lv


def log_px_given_z(x, a, lv, continuous):
    """``posterior_log_prob`` VAEB.py:302-313, per row.  Bernoulli:
    -binary_crossentropy(sigmoid(a), x).sum(1) = sum_d x*a - softplus(a)."""
    if continuous:
        mu_x = sigmoid(a)
        return (-0.5 * LOG2PI - 0.5 * lv - 0.5 * (x - mu_x) ** 2 / np.exp(lv)).sum(axis=1)
    return (x * a - softplus(a)).sum(axis=1)


def kl_rows(mu, ls):
    """VAEB.py:343 (negative KL, per row)."""
    return 0.5 * (1.0 + ls - mu ** 2 - np.exp(ls)).sum(axis=1)


def theta_prior(fvp):
    """``thetaPrior`` VAEB.py:359-363 over the interleaved variational list."""
    tot = 0.0
    for i in range(0, len(fvp), 2):
        mu_t, sg_t = fvp[i], fvp[i + 1]
        tot = tot + 0.5 * np.sum(1.0 + np.log(sg_t ** 2) - mu_t ** 2 - sg_t ** 2)
    return tot


@dataclass
class StepOut:
    sgvb: float          # the SGVB scalar the compiled function works with (sum semantics)
    per_row: np.ndarray  # per-datapoint ELBO terms [M] (data part only)
    grads: list          # d(train_criterion)/d(param), reference order (None for eval)


def _zeros_like_list(params):
    return [np.zeros_like(q) for q in params]


def elbo_and_grads(params, x, eps, continuous, estimator="LB", want_grads=True,
                   prior_scale=1.0, row_weight=None, act="tanh"):
    """The symbolic graph of ``getGradient`` (VAEB.py:370-399) for the LB (VAEB.py:332-346)
    and LA (VAEB.py:315-330) estimators with the weight prior of VAEB.py:386-390.

    eps: [L,M,Z] injected noise.  Returns SGVB (sum over the minibatch, mean over L),
    the per-row values and the gradient of ``SGVB - 0.5*prior_scale*sum(p**2)``.
    ``row_weight`` (scalar) scales the data objective (used by the variants)."""
    p = as_dict(params, continuous)
    dt = x.dtype
    L, M, Z = eps.shape
    w = dt.type(1.0 if row_weight is None else row_weight)
    hs = encoder_hiddens(p, x, act)
    h_e = hs[-1]
    mu, ls = h_e @ p["W4"] + p["b4"], h_e @ p["W5"] + p["b5"]
    sd = np.exp(0.5 * ls)
    per_row = np.zeros(M, dtype=dt)
    g = {n: np.zeros_like(q) for n, q in p.items()}
    d_mu = np.zeros_like(mu)
    d_ls = np.zeros_like(ls)
    invL = dt.type(1.0 / L)
    for l in range(L):
        e = eps[l]
        z = mu + sd * e
        h_d, a, lv = decoder(p, z, continuous, act)
        lp = log_px_given_z(x, a, lv, continuous)
        if estimator == "LA":
            prior = (-0.5 * LOG2PI - 0.5 * z ** 2).sum(axis=1)                       # VAEB.py:322
            logq = (-0.5 * LOG2PI - 0.5 * ls - 0.5 * (z - mu) ** 2 / np.exp(ls)).sum(axis=1)  # :324-325
            per_row += invL * (lp + prior - logq)
        else:
            per_row += invL * lp
        if not want_grads:
            continue
        s = w * invL
        if continuous:
            mu_x = sigmoid(a)
            r = (x - mu_x) * np.exp(-lv)
            d_a = s * r * mu_x * (1.0 - mu_x)
            d_lv = s * (-0.5 + 0.5 * (x - mu_x) * r)
        else:
            d_a = s * (x - sigmoid(a))
            d_lv = None
        g["W2"] += h_d.T @ d_a
        g["b2"] += d_a.sum(axis=0)
        d_h = d_a @ p["W2"].T
        if continuous:
            g["W6"] += h_d.T @ d_lv
            g["b6"] += d_lv.sum(axis=0)
            d_h += d_lv @ p["W6"].T
        d_a1 = d_h * hidden_act_grad(h_d, act)
        g["W1"] += z.T @ d_a1
        g["b1"] += d_a1.sum(axis=0)
        d_z = d_a1 @ p["W1"].T
        if estimator == "LA":
            d_z = d_z - s * z           # d prior / dz
            # logQ = sum(-.5log2pi - .5 ls - .5 eps^2) once z-mu = sd*eps is substituted:
            # the (z-mu)^2/exp(ls) term has zero net gradient, so -logQ gives d_ls += .5
            d_ls += s * 0.5
        d_mu += d_z
        d_ls += d_z * (0.5 * sd * e)
    if estimator == "LB":
        per_row += kl_rows(mu, ls)
        if want_grads:
            d_mu += w * (-mu)
            d_ls += w * 0.5 * (1.0 - np.exp(ls))
    sgvb = per_row.sum()
    if not want_grads:
        return StepOut(float(sgvb), per_row, None)
    g["W4"] = h_e.T @ d_mu
    g["b4"] = d_mu.sum(axis=0)
    g["W5"] = h_e.T @ d_ls
    g["b5"] = d_ls.sum(axis=0)
    d_he = d_mu @ p["W4"].T + d_ls @ p["W5"].T
    d_a3 = d_he * hidden_act_grad(h_e, act)
    depth = len(hs)
    for k in range(depth, 1, -1):                   # back through the extra encoder hidden layers
        g["W3_%d" % k] = hs[k - 2].T @ d_a3
        g["b3_%d" % k] = d_a3.sum(axis=0)
        d_a3 = (d_a3 @ p["W3_%d" % k].T) * hidden_act_grad(hs[k - 2], act)
    g["W3"] = x.T @ d_a3
    g["b3"] = d_a3.sum(axis=0)
    ps = dt.type(prior_scale)
    grads = [g[n] - ps * p[n] for n in param_names(continuous, depth)]   # VAEB.py:389-390
    return StepOut(float(sgvb), per_row, grads)


def adagrad_update(params, ada, grads, lr, eps=1e-6, p2_coeff=0.0):
    """``getUpdates`` VAEB.py:426-444 (ascent): acc = ada + g^2; p += lr*g/(sqrt(acc)+eps).
    ``p2_coeff`` adds the VAEBfullbayes.py:183-184 term ``- lr*eps*p**2`` (p2_coeff = eps).
    Identical rule in degenerate-vae/infalg.py:148-164.  In place."""
    for q, a, g in zip(params, ada, grads):
        dt = q.dtype.type
        a += g * g
        upd = dt(lr) * g / (np.sqrt(a) + dt(eps))
        if p2_coeff:
            upd = upd - dt(lr) * dt(p2_coeff) * q * q
        q += upd


def adadelta_update(params, g_ac, dx_ac, grads, rho=0.95, eps=1e-6):
    """``getAdaDeltaUpdates`` VAEB.py:449-469 (the alternative to Adagrad that the reference keeps
    commented out at VAEB.py:404): g_ac = rho*g_ac + (1-rho)*g^2; dx = sqrt(dx_ac+eps)*g/sqrt(g_ac+eps);
    x += dx; dx_ac = rho*dx_ac + (1-rho)*dx^2.  In place."""
    for q, a, d, g in zip(params, g_ac, dx_ac, grads):
        dt = q.dtype.type
        a *= dt(rho)
        a += dt(1.0 - rho) * g * g
        dx = np.sqrt(d + dt(eps)) * g / np.sqrt(a + dt(eps))
        q += dx
        d *= dt(rho)
        d += dt(1.0 - rho) * dx * dx


# --------------------------------------------------------------------------------------
# model objects mirroring the compiled functions
# --------------------------------------------------------------------------------------
class OracleVAEB:
    """State + ``update``/``validate`` of VAEB.py:408-422 with injected eps.

    estimator: "LB" | "LA" | "FVB" (reference-faithful) | "FVB_SAMPLED" (new
    semantics named by the north star; parity unpinned by the reference).
    variant: "vaeb" | "fullbayes" (VAEBfullbayes.py:121-201: mean objective, no weight
    prior in the criterion, extra -lr*1e-6*p**2 in the update)."""

    def __init__(self, x_train, continuous, H, Z, batch_size, L=1, lr=0.01, estimator="LB",
                 params=None, dtype=np.float64, variant="vaeb", optimizer="adagrad", rho=0.95, activation="tanh"):
        self.activation = activation
        self.x = np.asarray(x_train, dtype=dtype)
        self.N, self.D = self.x.shape
        self.continuous, self.H, self.Z = continuous, H, Z
        self.M, self.L, self.lr = batch_size, L, lr
        self.estimator, self.variant, self.dtype = estimator, variant, dtype
        self.optimizer, self.rho = optimizer, rho
        if params is None:
            params = init_params(self.D, H, Z, continuous, dtype=np.float32)
        self.params = [np.array(q, dtype=dtype, copy=True) for q in params]
        if estimator.startswith("FVB"):
            self.fvp = init_full_variational(self.params)          # VAEB.py:117-125
            self.ada = _zeros_like_list(self.fvp)
        else:
            self.fvp = None
            self.ada = _zeros_like_list(self.params)               # VAEB.py:178-182
            self.dx_ac = _zeros_like_list(self.params)             # VAEB.py:457 (AdaDelta only)

    # -- objective -------------------------------------------------------------------
    def _objective(self, x, eps, want_grads, zeta=None):
        est = self.estimator
        if self.variant == "fullbayes":
            # VAEBfullbayes.py:139-145: mean of (KL + logp), no prior term
            M = x.shape[0]
            out = elbo_and_grads(self.params, x, eps, self.continuous, "LB", want_grads,
                                 prior_scale=0.0, row_weight=1.0 / M)
            return out.sgvb / M, out.per_row, out.grads
        if est in ("LB", "LA"):
            out = elbo_and_grads(self.params, x, eps, self.continuous, est, want_grads, act=self.activation)
            return out.sgvb, out.per_row, out.grads
        # ---- full VB, VAEB.py:349-367 + :391-393, :399 ----
        if self.L != 1:
            raise ValueError("getFVBL clobbers `mu` at VAEB.py:361: undefined for L>1")
        M = x.shape[0]
        if est == "FVB":
            theta = self.params                                     # frozen MAP params (F5)
        else:
            theta = [self.fvp[2 * i] + np.abs(self.fvp[2 * i + 1]) * zeta[i]   # VAEB.py:127-129
                     for i in range(len(self.params))]
        out = elbo_and_grads(theta, x, eps, self.continuous, "LB",
                             want_grads and est == "FVB_SAMPLED", prior_scale=0.0, row_weight=M)
        sgvb = M * out.sgvb + float(theta_prior(self.fvp))          # VAEB.py:364
        grads = None
        if want_grads:
            grads = []
            for i in range(len(self.params)):
                mu_t, sg_t = self.fvp[2 * i], self.fvp[2 * i + 1]
                g_mu = -2.0 * mu_t                                   # thetaPrior + (-.5 mu^2) term
                g_sg = 1.0 / sg_t - 2.0 * sg_t
                if est == "FVB_SAMPLED":
                    g_mu = g_mu + out.grads[i]
                    g_sg = g_sg + out.grads[i] * zeta[i] * np.sign(sg_t)
                grads += [g_mu, g_sg]
        return sgvb, out.per_row, grads

    def grads(self, x, eps, zeta=None):
        return self._objective(np.asarray(x, self.dtype), np.asarray(eps, self.dtype), True, zeta)

    def update(self, index, eps, zeta=None):
        """``update(index)`` VAEB.py:408-415: returns SGVB/batch_size computed with the
        PRE-update parameters; applies Adagrad."""
        xb = self.x[index * self.M:(index + 1) * self.M]
        sgvb, _, grads = self._objective(xb, np.asarray(eps, self.dtype), True, zeta)
        if self.variant == "fullbayes":
            adagrad_update(self.params, self.ada, grads, self.lr, 1e-6, p2_coeff=1e-6)
            return sgvb                                              # VAEBfullbayes.py:153
        target = self.fvp if self.fvp is not None else self.params
        if self.optimizer == "adadelta":
            adadelta_update(target, self.ada, self.dx_ac, grads, self.rho, 1e-6)
        else:
            adagrad_update(target, self.ada, grads, self.lr, 1e-6)
        return sgvb / self.M

    def validate(self, x, eps, zeta=None):
        """``validate(x)`` VAEB.py:418-422: SGVB (a sum; the caller divides, :582)."""
        sgvb, per_row, _ = self._objective(np.asarray(x, self.dtype), np.asarray(eps, self.dtype), False, zeta)
        return sgvb, per_row


# --------------------------------------------------------------------------------------
# a19  importance-sampled marginal log-likelihood (new)
# --------------------------------------------------------------------------------------
def is_log_px(params, x, eps, continuous):
    """log p_hat(x_i) = logsumexp_l(log w_il) - log L with
    log w = log p(x|z) + log p(z) - log q(z|x): the ``getLA`` integrand (VAEB.py:319-327).
    eps: [N,L,Z].  Returns (logp[N], logw[N,L])."""
    p = as_dict(params, continuous)
    _, mu, ls = encoder(p, x)
    N, L, Z = eps.shape
    logw = np.empty((N, L), dtype=x.dtype)
    sd = np.exp(0.5 * ls)
    for l in range(L):
        e = eps[:, l, :]
        z = mu + sd * e
        _, a, lv = decoder(p, z, continuous)
        lp = log_px_given_z(x, a, lv, continuous)
        prior = (-0.5 * LOG2PI - 0.5 * z ** 2).sum(axis=1)
        logq = (-0.5 * LOG2PI - 0.5 * ls - 0.5 * e ** 2).sum(axis=1)
        logw[:, l] = lp + prior - logq
    m = logw.max(axis=1)
    logp = m + np.log(np.exp(logw - m[:, None]).sum(axis=1)) - math.log(L)
    return logp, logw


# --------------------------------------------------------------------------------------
# reconstruct -- VAEB.py:267-300, deterministic part only
# --------------------------------------------------------------------------------------
def reconstruct_mean(params, x, eps, continuous):
    """Decoder output averaged over n samples (eps[n,M,Z]); n == 0 decodes ``mu``
    (VAEB.py:269-292).  Returns y (discrete) or (y_mu, y_log_sigma) (continuous); the
    reference's final ``multivariate_normal`` draw (VAEB.py:295-297) is host RNG."""
    p = as_dict(params, continuous)
    _, mu, ls = encoder(p, x)
    zs = [mu] if eps is None or len(eps) == 0 else [reparam(mu, ls, e) for e in eps]
    acc_y, acc_lv = 0.0, 0.0
    for z in zs:
        _, a, lv = decoder(p, z, continuous)
        acc_y = acc_y + sigmoid(a)
        if continuous:
            acc_lv = acc_lv + lv
    n = len(zs)
    return (acc_y / n, acc_lv / n) if continuous else acc_y / n


# --------------------------------------------------------------------------------------
# AE-side primitives named by the north star (degenerate-vae/{mlp,logpdf,infalg}.py)
# --------------------------------------------------------------------------------------
def construct_mlp(x, Ws, bs, f=np.tanh):
    """``ConstructMLP`` degenerate-vae/mlp.py:66-74: f on EVERY layer incl. the last."""
    h = x
    for W, b in zip(Ws, bs):
        h = f(h @ W + b)
    return h


def normal_prior(theta, s2):
    """``ConstructNormalPrior`` degenerate-vae/mlp.py:87-91."""
    return -0.5 * sum(np.sum(q ** 2 / s2 + math.log(2 * math.pi * s2)) for q in theta)


def gauss_dkl(mu0, s20, mu1, s21):
    """``GaussDKL`` degenerate-vae/mlp.py:157-159."""
    return 0.5 * np.sum(s20 / s21 + (mu1 - mu0) ** 2 / s21 - 1.0 + np.log(s21) - np.log(s20))


def lpdf_bernoulli(Y, P):
    """degenerate-vae/logpdf.py:85-86 (note the +1e-7 inside both logs)."""
    return np.sum(Y * np.log(P + 1e-7) + (1 - Y) * np.log(1.0 - P + 1e-7))


def lpdf_indep_normal(Y, mu, logs2):
    """degenerate-vae/logpdf.py:112-114."""
    return -0.5 * np.sum(LOG2PI + logs2 + (Y - mu) ** 2 / np.exp(logs2))


def out_to_probs(Hm, W, b):
    """degenerate-vae/logpdf.py:46-47."""
    return sigmoid(Hm @ W + b)


def out_to_real(Hm, W, b):
    """degenerate-vae/logpdf.py:72-73."""
    return Hm @ W + b


# --------------------------------------------------------------------------------------
# AE baselines: degenerate-vae/ae.py:41-117 and vanilla-ae/ae.py:45-104, one hidden layer
# per side as LearnFreyFace / LearnMNIST build them.  Parameters in the VAEB list order
# [W3,W4,W5,W1,W2,(W6),b3,b4,b5,b1,b2,(b6)] with W3=Wenc0, W4=Wz, W1=Wdec0, W2=Wout|Wmu, W6=Wlogs2 (W5, b5 unused).
# --------------------------------------------------------------------------------------
def ae_objective_and_grads(params, x, continuous, kind, s2=1.0):
    """Returns (reported value, gradient of logjoint w.r.t. params) for one minibatch x.
    kind 'degenerate': Z = Hz.Wz+bz; loglik = logpdf.bernoulli (1e-7 inside both logs) | indep_normal;
      logjoint = loglik + NormalPrior(theta, s2) + NormalPrior([Z], 1); reported loglik / M   (ae.py:46-87)
    kind 'vanilla': Z = tanh(Hz.Wz+bz); se = sum((X - sigmoid(.))^2); logjoint = -se + NormalPrior(theta, s2);
      reported se / M                                                                        (vanilla-ae/ae.py:50-83)"""
    p = as_dict(params, continuous)
    M = x.shape[0]
    h_e = np.tanh(x @ p["W3"] + p["b3"])
    pre = h_e @ p["W4"] + p["b4"]
    z = np.tanh(pre) if kind == "vanilla" else pre
    h_d = np.tanh(z @ p["W1"] + p["b1"])
    a = h_d @ p["W2"] + p["b2"]
    P = sigmoid(a)
    Q = sigmoid(-a)                                     # 1 - P without cancellation
    g = {n: np.zeros_like(q) for n, q in p.items()}
    d_lv = None
    if kind == "vanilla":
        assert not continuous
        value = ((x - P) ** 2).sum() / M
        d_a = 2.0 * (x - P) * P * Q                     # d(-se)/da
    elif continuous:
        lv = h_d @ p["W6"] + p["b6"]
        r = (x - P) * np.exp(-lv)
        value = (-0.5 * (LOG2PI + lv + (x - P) * r)).sum() / M          # logpdf.py:112-114
        d_a = r * P * Q
        d_lv = -0.5 + 0.5 * (x - P) * r
    else:
        e = 1e-7
        value = (x * np.log(P + e) + (1.0 - x) * np.log(Q + e)).sum() / M   # logpdf.py:85-86
        d_a = (x / (P + e) - (1.0 - x) / (Q + e)) * P * Q
    g["W2"] = h_d.T @ d_a
    g["b2"] = d_a.sum(axis=0)
    d_h = d_a @ p["W2"].T
    if d_lv is not None:
        g["W6"] = h_d.T @ d_lv
        g["b6"] = d_lv.sum(axis=0)
        d_h += d_lv @ p["W6"].T
    d_a1 = d_h * (1.0 - h_d ** 2)
    g["W1"] = z.T @ d_a1
    g["b1"] = d_a1.sum(axis=0)
    d_z = d_a1 @ p["W1"].T
    d_pre = d_z * (1.0 - z ** 2) if kind == "vanilla" else d_z - z      # tanh'  |  NormalPrior([Z], 1.0)
    g["W4"] = h_e.T @ d_pre
    g["b4"] = d_pre.sum(axis=0)
    d_a3 = (d_pre @ p["W4"].T) * (1.0 - h_e ** 2)
    g["W3"] = x.T @ d_a3
    g["b3"] = d_a3.sum(axis=0)
    names = param_names(continuous)
    grads = [g[n] - p[n] / s2 for n in names]           # mlp.py:87-91: -0.5*sum(p^2/s2 + log(2 pi s2))
    for i, n in enumerate(names):
        if n in ("W5", "b5"):
            grads[i] = np.zeros_like(p[n])              # not part of theta in ae.py
    return float(value), grads


class OracleAE:
    """`train(idx)` of ConstructAE with infalg.AdaGrad(eta) (infalg.py:148-164): gather rows, ascend logjoint."""

    def __init__(self, x_train, continuous, params, kind="degenerate", s2=1.0, eta=0.01, dtype=np.float64):
        self.x = np.asarray(x_train, dtype=dtype)
        self.continuous, self.kind, self.s2, self.eta = continuous, kind, s2, eta
        self.params = [np.array(q, dtype=dtype, copy=True) for q in params]
        self.ada = _zeros_like_list(self.params)

    def train(self, idx):
        value, grads = ae_objective_and_grads(self.params, self.x[np.asarray(idx)], self.continuous, self.kind, self.s2)
        adagrad_update(self.params, self.ada, grads, self.eta, 1e-6)
        return value

    def forward(self, x, what="reconstruct"):
        p = as_dict(self.params, self.continuous)
        if what == "decode":
            z = np.asarray(x, self.x.dtype)
        else:
            pre = np.tanh(np.asarray(x, self.x.dtype) @ p["W3"] + p["b3"]) @ p["W4"] + p["b4"]
            z = np.tanh(pre) if self.kind == "vanilla" else pre
            if what == "encode":
                return z
        return sigmoid(np.tanh(z @ p["W1"] + p["b1"]) @ p["W2"] + p["b2"])


# --------------------------------------------------------------------------------------
# synthetic data of the reference's shapes -- shared by tests and bench
# --------------------------------------------------------------------------------------
def synthetic_mnist(n, seed=15485863, D=784):
    rng = np.random.RandomState(seed)
    x = rng.uniform(size=(n, D)).astype(np.float32)
    x *= (rng.uniform(size=(n, D)) < 0.19)
    return x


def synthetic_frey(n=1965, seed=15485863, D=560):
    rng = np.random.RandomState(seed)
    return np.clip(0.5 + 0.2 * rng.normal(size=(n, D)), 0, 1).astype(np.float32)
