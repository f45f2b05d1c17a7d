"""Host-side mirror of the reference's autoencoder baselines over the C-ABI:

  * ``degenerate-vae/ae.py:41-117``  ConstructAE -- point-estimate "degenerate VAE": Z = Hz.Wz + bz with a N(0,1)
    prior on Z, ``logpdf.bernoulli`` (1e-7 inside both logs) or ``logpdf.indep_normal`` outputs;
  * ``vanilla-ae/ae.py:45-104``      ConstructAE -- Z = tanh(Hz.Wz + bz), squared error of sigmoid outputs;

both with ``mlp.ConstructNormalPrior(theta, s2)`` weight decay and ``infalg.AdaGrad(eta)``, trained on minibatches
GATHERED by index (``givens = {X: Xtr[idx]}``).  ``ConstructAE`` keeps the reference's signature and return tuple
``(train, reconstruct, encode, decode, theta)``; the extra keyword ``kind`` picks the variant (the reference has one
module per variant).  One hidden layer per side, tanh, as ``LearnFreyFace`` / ``LearnMNIST`` build them
(``ae.py:137-139,182-183``).  Every number comes from the device; there is no CPU path here."""
from __future__ import annotations

import ctypes as C

import numpy as np
import numpy.random as rnd

from . import _lib
from .model import SharedParam, _f32, _ptr


class AdaGrad(object):
    """``infalg.AdaGrad`` (degenerate-vae/infalg.py:141-186): carries the learning rate eta."""

    def __init__(self, eta):
        if not eta > 0:
            raise ValueError('eta must be greater than zero and less than one.')      # infalg.py:173
        self.eta = eta

    def name(self):
        return 'AdaGrad'


def WeightMatrix(Din, Dout, name=None):
    """mlp.py:48-49: N(0, 0.01^2) from the GLOBAL numpy RNG."""
    return rnd.normal(0.0, 0.01, size=(Din, Dout)).astype(np.float32)


def BiasVector(D, name=None):
    """mlp.py:36-37: biases are N(0, 0.01^2) too (unlike VAEB.py's zeros)."""
    return rnd.normal(0.0, 0.01, size=(D)).astype(np.float32)


def rmse(X, Xpr):
    """ae.py:120-121."""
    return np.sqrt(np.mean(np.sum((X - Xpr) ** 2, 1)))


def mse(X, Xpr):
    """vanilla-ae/ae.py:120-121."""
    return np.mean(np.sum((X - Xpr) ** 2, 1))


class _AEModel(object):
    """Owns the device handle; parameters live in the VAEB layout (W3=Wenc, W4=Wz, W1=Wdec, W2=Wout|Wmu, W6=Wlogs2)."""

    def __init__(self, Xtr, H, Dz, cont, kind, s2, eta, device):
        self._lib = _lib.load()
        self.kind = _lib.AE_VANILLA if kind == "vanilla" else _lib.AE_DEGENERATE
        x = _f32(Xtr)
        self.N, self.D = x.shape
        cfg = _lib.Config(input_dim=self.D, hidden_units=H, latent_size=Dz, batch_size=1, L=1, continuous=int(cont),
                          estimator=_lib.EST_LB, variant=_lib.VARIANT_VAEB, precision=_lib.PREC_FP32, device=device,
                          learning_rate=eta, adagrad_eps=1e-6, prior_scale=1.0 / s2, sigma_vb_init=1e-3, seed=10)
        self._h = C.c_void_p()
        _lib.check(self._lib.vaeb_create(C.byref(cfg), C.byref(self._h)))
        self._names = ["W3", "W4", "W5", "W1", "W2"] + (["W6"] if cont else []) + ["b3", "b4", "b5", "b1", "b2"] + \
                      (["b6"] if cont else [])
        shapes = {"W3": (self.D, H), "W4": (H, Dz), "W5": (H, Dz), "W1": (Dz, H), "W2": (H, self.D), "W6": (H, self.D),
                  "b3": (H,), "b4": (Dz,), "b5": (Dz,), "b1": (H,), "b2": (self.D,), "b6": (self.D,)}
        self._shapes = [shapes[n] for n in self._names]
        _lib.check(self._lib.vaeb_upload_data(self._h, _ptr(x), self.N))

    # SharedParam plumbing (same protocol as model.VAEB)
    def _get_buffer(self, which):
        out = [np.empty(s, np.float32) for s in self._shapes]
        ptrs = (C.c_void_p * len(out))(*[a.ctypes.data for a in out])
        _lib.check(self._lib.vaeb_get_tensors(self._h, which, ptrs))
        return out

    def _set_buffer(self, which, values):
        vals = [_f32(v).reshape(s) for v, s in zip(values, self._shapes)]
        ptrs = (C.c_void_p * len(vals))(*[a.ctypes.data for a in vals])
        _lib.check(self._lib.vaeb_set_tensors(self._h, which, ptrs))

    def train(self, idx):
        ia = np.ascontiguousarray(idx, dtype=np.int32)
        out = C.c_float()
        _lib.check(self._lib.vaeb_ae_train(self._h, self.kind, _ptr(ia), len(ia), C.byref(out)))
        return np.asarray(out.value, dtype=np.float32)

    def forward(self, what, a, width_out):
        xa = _f32(a)
        xa = xa.reshape(-1, xa.shape[-1])
        out = np.empty((xa.shape[0], width_out), np.float32)
        _lib.check(self._lib.vaeb_ae_forward(self._h, self.kind, what, _ptr(xa), xa.shape[0], _ptr(out)))
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.vaeb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ConstructAE(Xtr, Denc=[500], Dz=20, Ddec=[500], f="tanh", s2=1.0, inf=AdaGrad(0.01), otype='binary',
                kind="degenerate", device=0):
    """degenerate-vae/ae.py:41-117 (kind='degenerate') / vanilla-ae/ae.py:45-104 (kind='vanilla').

    Xtr: ndarray or anything with get_value().  Returns (train, reconstruct, encode, decode, theta) where
    theta lists the parameters in the reference's order -- Wenc + benc + [Wz, bz] + Wdec + bdec + [Wout, bout]
    (binary) or + [Wmu, Wlogs2, bmu, blogs2] (cont) -- as objects with get_value()/set_value()."""
    if len(Denc) != 1 or len(Ddec) != 1 or Denc[0] != Ddec[0]:
        raise ValueError("one hidden layer of equal width per side (as LearnFreyFace / LearnMNIST build it)")
    if f not in ("tanh", np.tanh):
        raise ValueError("f: tanh only")
    if kind not in ("degenerate", "vanilla"):
        raise ValueError("kind must be 'degenerate' or 'vanilla'")
    if kind == "vanilla":
        otype = 'binary'                                   # vanilla-ae/ae.py has sigmoid outputs only
    if otype not in ('binary', 'cont'):
        raise ValueError('otype currently only supports binary.')        # ae.py:74 (sic)
    X = np.asarray(Xtr.get_value() if hasattr(Xtr, "get_value") else Xtr)
    Dobs, H = X.shape[1], Denc[0]
    cont = otype == 'cont'
    # the reference's draw order from the global numpy RNG (ae.py:49-71)
    vals = {}
    vals["W3"], vals["b3"] = WeightMatrix(Dobs, H), BiasVector(H)
    vals["W4"], vals["b4"] = WeightMatrix(H, Dz), BiasVector(Dz)
    vals["W1"], vals["b1"] = WeightMatrix(Dz, H), BiasVector(H)
    if cont:
        vals["W2"], vals["W6"] = WeightMatrix(H, Dobs), WeightMatrix(H, Dobs)
        vals["b2"], vals["b6"] = BiasVector(Dobs), BiasVector(Dobs)
    else:
        vals["W2"], vals["b2"] = WeightMatrix(H, Dobs), BiasVector(Dobs)
    m = _AEModel(X, H, Dz, cont, kind, s2, inf.eta, device)
    vals["W5"], vals["b5"] = np.zeros((H, Dz), np.float32), np.zeros(Dz, np.float32)      # not part of an AE
    m._set_buffer(_lib.BUF_PARAMS, [vals[n] for n in m._names])
    order = ["W3", "b3", "W4", "b4", "W1", "b1"] + (["W2", "W6", "b2", "b6"] if cont else ["W2", "b2"])
    theta = [SharedParam(m, _lib.BUF_PARAMS, m._names.index(n), n, m._shapes[m._names.index(n)]) for n in order]

    def train(idx):
        return m.train(idx)

    def reconstruct(Xin):
        return m.forward(0, Xin, Dobs)

    def encode(Xin):
        return m.forward(1, Xin, Dz)

    def decode(Zin):
        return m.forward(2, Zin, Dobs)

    train.model = m                                        # keeps the handle alive / lets callers close() it
    return train, reconstruct, encode, decode, theta


def LearnAE(X, epochs=100, Dz=20, Ntr=1500, H=200, otype='cont', kind="degenerate", batch_size=100, verbose=True):
    """The epoch loop shared by LearnFreyFace / LearnMNIST (ae.py:127-166,174-210): a fresh permutation of the row
    indices every epoch, ragged last minibatch, learning curve of train() values, train/test rmse."""
    Xtr, Xte = X[:Ntr], X[Ntr:]
    train, reconstruct, encode, decode, theta = ConstructAE(Xtr, Denc=[H], Dz=Dz, Ddec=[H], otype=otype, kind=kind,
                                                            inf=AdaGrad(0.01))
    curve = []
    for i in range(epochs):
        idx = rnd.permutation(np.arange(Ntr)).astype(np.int32)
        lb = 0
        while lb < Ntr:
            ub = min(lb + batch_size, Ntr)
            curve.append(float(train(idx[lb:ub])))
            lb = ub
        if verbose:
            print('Epoch ' + str(i) + ('. mse = ' if kind == "vanilla" else '. mean loglik = ') + str(curve[-1]))
    rm = (rmse(Xtr, reconstruct(Xtr)), rmse(Xte, reconstruct(Xte)) if len(Xte) else float("nan"))
    if verbose:
        print('training rmse = ' + str(rm[0]))
        print('testing rmse = ' + str(rm[1]))
    return reconstruct, encode, decode, curve, rm
