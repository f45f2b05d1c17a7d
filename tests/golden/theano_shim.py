"""A minimal stand-in for the subset of Theano that the reference's hot path uses, so that the
reference's OWN source (/root/reference/VAEB.py, VAEBfullbayes.py, degenerate-vae/{mlp,logpdf,
infalg}.py) can be executed in this container to produce golden vectors (tests/golden/
make_reference_golden.py).  Test infrastructure only -- nothing under vaeb_b200/ imports it.

What it is: a lazy expression graph (like Theano's) evaluated with torch on the CPU; `T.grad` is
`torch.autograd.grad` of the evaluated cost w.r.t. the shared-variable leaves; `theano.function`
evaluates outputs and `updates` against the pre-update values and then assigns, honouring
`givens`.  `RandomStreams(seed).normal(shape)` follows Theano's scheme (a seed generator
RandomState(seed); every normal() node owns RandomState(gen.randint(2**30)); every function call
draws normal(0,1,shape) in fp64 and casts to floatX) and LOGS every draw, so a golden fixture
records the noise the reference consumed and the parity tests inject exactly that.

What it is not: Theano.  Semantics of the ops used (dot, tanh, sigmoid, exp, log, sqrt, pow, sum,
binary_crossentropy(o, t) = -(t log o + (1-t) log(1-o)), subtensor slices, ones/zeros_like, eye)
are the documented ones; Theano's graph optimiser (e.g. the softplus rewrite of
log(sigmoid)) is not reproduced -- fixtures are generated in float64 where that does not matter.
"""
from __future__ import annotations

import sys
import types

import numpy as np
import torch

_DT = {"float32": torch.float32, "float64": torch.float64}


class _Config(object):
    floatX = "float64"


config = _Config()


def _tdtype():
    return _DT[config.floatX]


class Var(object):
    """A node of the lazy graph.  `fn(env)` computes its value (torch tensor / int / tuple)."""
    __array_ufunc__ = None       # numpy scalars defer to our reflected operators
    __array_priority__ = 1000

    def __init__(self, fn, name=None):
        self._fn = fn
        self.name = name

    # -- evaluation ------------------------------------------------------------------
    def _value(self, env):
        k = id(self)
        if k in env["givens"]:
            return _ev(env["givens"][k], env)
        memo = env["memo"]
        if k not in memo:
            memo[k] = self._fn(env)
        return memo[k]

    def eval(self, inputs_to_values=None):
        env = _new_env()
        for k, v in (inputs_to_values or {}).items():
            env["memo"][id(k)] = _to_t(v)
        return _to_np(self._value(env))

    # -- operators -------------------------------------------------------------------
    def __add__(self, o): return _bin(self, o, lambda a, b: a + b)
    def __radd__(self, o): return _bin(o, self, lambda a, b: a + b)
    def __sub__(self, o): return _bin(self, o, lambda a, b: a - b)
    def __rsub__(self, o): return _bin(o, self, lambda a, b: a - b)
    def __mul__(self, o): return _bin(self, o, lambda a, b: a * b)
    def __rmul__(self, o): return _bin(o, self, lambda a, b: a * b)
    def __truediv__(self, o): return _bin(self, o, lambda a, b: a / b)
    def __rtruediv__(self, o): return _bin(o, self, lambda a, b: a / b)
    __div__, __rdiv__ = __truediv__, __rtruediv__
    def __pow__(self, o): return _bin(self, o, lambda a, b: a ** b)
    def __neg__(self): return Var(lambda env: -self._value(env))

    def __getitem__(self, idx):
        def fn(env):
            v = self._value(env)
            if isinstance(idx, slice):
                lo = None if idx.start is None else int(_ev(idx.start, env))
                hi = None if idx.stop is None else int(_ev(idx.stop, env))
                return v[lo:hi]
            if isinstance(idx, tuple):
                return v[idx]
            i = _ev(idx, env)
            if isinstance(i, torch.Tensor) and i.ndim >= 1:          # Xtr[idx] with an ivector: gather of rows
                return v[i.long()]
            return v[int(i)]
        return Var(fn)

    def sum(self, axis=None, keepdims=False):
        return sum(self, axis=axis, keepdims=keepdims)

    def mean(self, axis=None, keepdims=False):
        return mean(self, axis=axis, keepdims=keepdims)

    @property
    def shape(self):
        return Var(lambda env: tuple(self._value(env).shape))

    @property
    def T(self):
        return Var(lambda env: self._value(env).t())

    def dimshuffle(self, *pattern):
        def fn(env):
            v = self._value(env)
            keep = [p for p in pattern if p != 'x']
            v = v.permute(*keep) if len(keep) > 1 else v
            for i, p in enumerate(pattern):
                if p == 'x':
                    v = v.unsqueeze(i)
            return v
        return Var(fn)


def _new_env():
    return {"memo": {}, "givens": {}, "leaves": {}}


def _to_t(v):
    if isinstance(v, torch.Tensor):
        return v
    if isinstance(v, (int, np.integer)):
        return int(v)
    if isinstance(v, (float, np.floating)):
        return float(v)
    a = np.asarray(v)
    if a.dtype.kind == "f":
        return torch.as_tensor(a.astype(config.floatX))
    return torch.as_tensor(a)


def _to_np(v):
    if isinstance(v, torch.Tensor):
        return v.detach().numpy().copy()
    if isinstance(v, (list, tuple)):
        return type(v)(_to_np(u) for u in v)
    return v


def _ev(x, env):
    if isinstance(x, Var):
        return x._value(env)
    return _to_t(x)


def _bin(a, b, f):
    return Var(lambda env: f(_ev(a, env), _ev(b, env)))


def _un(a, f):
    return Var(lambda env: f(_as_tensor(_ev(a, env))))


def _as_tensor(v):
    return v if isinstance(v, torch.Tensor) else torch.as_tensor(v, dtype=_tdtype())


# ---- shared variables -------------------------------------------------------------------
class SharedVariable(Var):
    def __init__(self, value, name=None, borrow=False):
        Var.__init__(self, None, name)
        self.set_value(value)

    def get_value(self, borrow=False):
        return self._np

    def set_value(self, value, borrow=False):
        a = np.array(value)
        self._np = a.astype(config.floatX) if a.dtype.kind == "f" else a

    def _value(self, env):
        k = id(self)
        if k in env["givens"]:
            return _ev(env["givens"][k], env)
        if k not in env["leaves"]:
            t = torch.tensor(self._np)
            if t.is_floating_point():
                t.requires_grad_(True)
            env["leaves"][k] = t
        return env["leaves"][k]


def shared(value, name=None, borrow=False, **kw):
    return SharedVariable(value, name=name)


# ---- theano.function ------------------------------------------------------------------------
class Function(object):
    def __init__(self, inputs, outputs, updates=None, givens=None, allow_input_downcast=False, **kw):
        self.inputs, self.outputs = list(inputs), outputs
        self.updates = list(updates.items()) if isinstance(updates, dict) else list(updates or [])
        self.givens = list(givens.items()) if isinstance(givens, dict) else list(givens or [])

    def __call__(self, *args):
        assert len(args) == len(self.inputs)
        env = _new_env()
        for v, a in zip(self.inputs, args):
            env["memo"][id(v)] = _to_t(a)
        for k, e in self.givens:
            env["givens"][id(k)] = e
        outs = self.outputs if isinstance(self.outputs, (list, tuple)) else [self.outputs]
        vals = [_ev(o, env) for o in outs]
        new = [(s, _ev(e, env)) for s, e in self.updates]      # all against the pre-update values
        for s, v in new:
            s.set_value(_to_np(v))
        res = [np.asarray(_to_np(v)) for v in vals]
        return res if isinstance(self.outputs, (list, tuple)) else res[0]


def function(inputs=(), outputs=None, updates=None, givens=None, **kw):
    return Function(inputs, outputs, updates, givens, **kw)


# ---- theano.tensor ---------------------------------------------------------------------------
def _input(name=None):
    def fn(env):
        raise RuntimeError("symbolic input %r has no value" % (name,))
    return Var(fn, name)


def matrix(name=None, dtype=None): return _input(name)
def vector(name=None, dtype=None): return _input(name)
def iscalar(name=None): return _input(name)
def lscalar(name=None): return _input(name)
def ivector(name=None): return _input(name)
def lvector(name=None): return _input(name)
def scalar(name=None, dtype=None): return _input(name)


def dot(a, b): return _bin(a, b, lambda x, y: _as_tensor(x) @ _as_tensor(y))
def tanh(a): return _un(a, torch.tanh)
def exp(a): return _un(a, torch.exp)
def log(a): return _un(a, torch.log)
def sqrt(a): return _un(a, torch.sqrt)
def sqr(a): return _un(a, lambda v: v * v)
def abs_(a): return _un(a, torch.abs)
def pow(a, b): return _bin(a, b, lambda x, y: x ** y)   # noqa: A001
def ones_like(a): return _un(a, torch.ones_like)
def zeros_like(a): return _un(a, torch.zeros_like)
def zeros(shape, dtype=None): return Var(lambda env: torch.zeros(tuple(_ev(shape, env)), dtype=_tdtype()))
def ones(shape, dtype=None): return Var(lambda env: torch.ones(tuple(_ev(shape, env)), dtype=_tdtype()))
def eye(n): return Var(lambda env: torch.eye(int(_ev(n, env)), dtype=_tdtype()))
def cast(a, dtype): return _un(a, lambda v: v.to(_DT.get(str(dtype), v.dtype)))
def as_tensor_variable(a): return a if isinstance(a, Var) else Var(lambda env: _to_t(a))
def concatenate(xs, axis=0): return Var(lambda env: torch.cat([_as_tensor(_ev(x, env)) for x in xs], dim=axis))


def sum(a, axis=None, keepdims=False):   # noqa: A001
    def fn(env):
        v = _as_tensor(_ev(a, env))
        return v.sum() if axis is None else v.sum(dim=axis, keepdim=keepdims)
    return Var(fn)


def mean(a, axis=None, keepdims=False):
    def fn(env):
        v = _as_tensor(_ev(a, env))
        return v.mean() if axis is None else v.mean(dim=axis, keepdim=keepdims)
    return Var(fn)


def grad(cost, wrt, **kw):
    single = not isinstance(wrt, (list, tuple))
    wl = [wrt] if single else list(wrt)
    holder = Var(None)

    def all_grads(env):
        c = _ev(cost, env)
        leaves = [w._value(env) for w in wl]
        gs = torch.autograd.grad(c, leaves, retain_graph=True, allow_unused=True)
        return [torch.zeros_like(l) if g is None else g.detach() for g, l in zip(gs, leaves)]
    holder._fn = all_grads
    outs = [Var((lambda i: (lambda env: holder._value(env)[i]))(i)) for i in range(len(wl))]
    return outs[0] if single else outs


class _NNet(object):
    @staticmethod
    def sigmoid(a): return _un(a, torch.sigmoid)

    @staticmethod
    def softplus(a): return _un(a, torch.nn.functional.softplus)

    @staticmethod
    def softmax(a): return _un(a, lambda v: torch.softmax(v, dim=-1))

    @staticmethod
    def binary_crossentropy(output, target):
        return _bin(output, target, lambda o, t: -(t * torch.log(o) + (1.0 - t) * torch.log(1.0 - o)))


class RandomStreams(object):
    """theano.tensor.shared_randomstreams.RandomStreams: see the module docstring.  `draws` is
    the log [(node index, float64 array as drawn)] in evaluation order."""

    def __init__(self, seed=None):
        self.gen_seedgen = np.random.RandomState(seed)
        self.nodes = []
        self.draws = []

    def normal(self, size=None, avg=0.0, std=1.0, ndim=None, dtype=None):
        rs = np.random.RandomState(int(self.gen_seedgen.randint(2 ** 30)))
        node = len(self.nodes)
        self.nodes.append(rs)

        def fn(env):
            shp = tuple(int(s) for s in _ev(size, env))
            d = rs.normal(avg, std, size=shp)
            self.draws.append((node, d.copy()))
            return torch.as_tensor(d.astype(config.floatX))
        return Var(fn)


def install(floatX="float64"):
    """Registers the shim as `theano`, `theano.tensor`, ... in sys.modules (plus inert stand-ins for
    matplotlib.pyplot / VAEBImage, which the reference imports but the hot path never calls)."""
    config.floatX = floatX
    me = sys.modules[__name__]
    th = types.ModuleType("theano")
    T = types.ModuleType("theano.tensor")
    for n in ("matrix", "vector", "iscalar", "lscalar", "ivector", "lvector", "scalar", "dot", "tanh", "exp", "log",
              "sqrt", "sqr", "abs_", "pow", "ones_like", "zeros_like", "zeros", "ones", "eye", "cast", "sum", "mean",
              "grad", "as_tensor_variable", "concatenate"):
        setattr(T, n, getattr(me, n))
    T.nnet = _NNet
    srs = types.ModuleType("theano.tensor.shared_randomstreams")
    srs.RandomStreams = RandomStreams
    T.shared_randomstreams = srs
    th.tensor, th.config, th.shared, th.function, th.grad = T, config, shared, function, grad
    sys.modules["theano"] = th
    sys.modules["theano.tensor"] = T
    sys.modules["theano.tensor.nnet"] = _NNet
    sys.modules["theano.tensor.shared_randomstreams"] = srs
    for stub in ("matplotlib", "matplotlib.pyplot", "VAEBImage"):
        if stub not in sys.modules:
            m = types.ModuleType(stub)
            m.plot = lambda *a, **k: None
            sys.modules[stub] = m
    return th
