"""`.mdl` / `.trc` result layout of the reference.

`.mdl` (VAEB.py:189-203): sequential pickles in one file -- n_hidden_units, n_latent,
continuous, learning_rate, batch_size, prng, sigmaInit, L, [genericEstimator], then every
parameter in the order [W3,W4,W5,W1,W2,(W6),b3,b4,b5,b1,b2,(b6)].  The shipped files are
Python-2 protocol-0 pickles whose parameters are Theano `CudaNdarraySharedVariable` objects;
they are read with a restricted unpickler that only ever constructs numpy arrays/dtypes and
inert stubs.  `reconstruction_res/*.mdl` carry the 9-object header that `VAEB.load`
(VAEB.py:206-242) expects, `full_vb_res/*.mdl` the 8-object header that the current
`VAEB.save` writes; both are accepted.  We write the 9-object header so that `load` works.

`.trc` (VAEB.py:568-570,583-585,591-593): header `num_samples,L,Lvalid`, then one line per
epoch written TWICE, floats as Python 2 printed them (12 significant digits)."""
from __future__ import annotations

import pickle

import numpy as np


class _Stub(object):
    """Inert stand-in for any class the restricted unpickler does not allow."""

    def __init__(self, *args, **kwargs):
        self._args = args
        self._state = None

    def __setstate__(self, state):
        self._state = state


def _stub_factory(*args, **kwargs):
    return _Stub(*args, **kwargs)


class _RestrictedUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module in ("numpy.core.multiarray", "numpy._core.multiarray") and name in ("_reconstruct", "scalar"):
            try:                                   # numpy >= 2
                import numpy._core.multiarray as m
            except ImportError:                    # numpy 1.x
                import numpy.core.multiarray as m
            return getattr(m, name)
        if module == "numpy" and name in ("ndarray", "dtype"):
            return getattr(np, name)
        if module == "_codecs" and name == "encode":     # protocol-2 bytes payloads of ndarrays
            import _codecs
            return _codecs.encode
        if module in ("copy_reg", "copyreg") and name == "_reconstructor":
            return lambda cls, base, state: cls() if isinstance(cls, type) else _Stub()
        if module in ("__builtin__", "builtins") and name == "object":
            return object
        # theano.*, numpy.random.*: never instantiate the real thing
        return type(str(name), (_Stub,), {"__module__": module}) if name[:1].isupper() else _stub_factory


def _find_array(obj, depth=0):
    """First ndarray reachable from a stubbed Theano shared variable."""
    if isinstance(obj, np.ndarray):
        return obj
    if depth > 8:
        return None
    kids = []
    if isinstance(obj, _Stub):
        kids = [obj._state, obj._args]
    elif isinstance(obj, dict):
        # the value lives under container -> storage
        kids = [obj.get("container"), obj.get("storage")] + [v for k, v in obj.items() if k not in ("container", "storage")]
    elif isinstance(obj, (list, tuple)):
        kids = list(obj)
    for k in kids:
        r = _find_array(k, depth + 1)
        if r is not None:
            return r
    return None


def _load_all(file_name):
    objs = []
    with open(file_name, "rb") as f:
        while True:
            try:
                objs.append(_RestrictedUnpickler(f, encoding="latin1").load())
            except EOFError:
                break
    return objs


def read_mdl(file_name):
    """Returns (header dict, [float32 arrays in reference order])."""
    objs = _load_all(file_name)
    if len(objs) < 8 + 10:
        raise ValueError("%s: not a VAEB .mdl file (%d objects)" % (file_name, len(objs)))
    keys = ["n_hidden_units", "n_latent", "continuous", "learning_rate", "batch_size", "prng", "sigmaInit", "L"]
    header = dict(zip(keys, objs[:8]))
    rest = objs[8:]
    header["genericEstimator"] = False
    if isinstance(rest[0], (bool, np.bool_)):        # 9-object header (VAEB.py:218)
        header["genericEstimator"] = bool(rest[0])
        rest = rest[1:]
    params = []
    for o in rest:
        a = _find_array(o)
        if a is None:
            raise ValueError("%s: parameter object without an array" % file_name)
        params.append(np.ascontiguousarray(a, dtype=np.float32))
    n_expected = 12 if header["continuous"] else 10
    extra = len(params) - n_expected                 # deeper encoders append (W3_k, b3_k) pairs after the reference's list
    if extra < 0 or extra % 2 or extra > 6:
        raise ValueError("%s: expected %d parameter tensors (+ 2 per extra encoder layer), found %d"
                         % (file_name, n_expected, len(params)))
    header["encoder_layers"] = 1 + extra // 2
    header["prng"] = None      # the constructor re-seeds RandomState(10) regardless (VAEB.py:148)
    return header, params


def write_mdl(file_name, header, params):
    with open(file_name, "wb") as f:
        for k in ["n_hidden_units", "n_latent", "continuous", "learning_rate", "batch_size", "prng", "sigmaInit", "L",
                  "genericEstimator"]:
            pickle.dump(header[k], f, protocol=2)
        for p in params:
            pickle.dump(np.asarray(p, dtype=np.float32), f, protocol=2)


def read_pkl_list(file_name):
    """`modelFrey.pkl` layout (freyFace.py:139-145): one pickled list of 12 ndarrays."""
    with open(file_name, "rb") as f:
        obj = _RestrictedUnpickler(f, encoding="latin1").load()
    return [np.ascontiguousarray(_find_array(o), dtype=np.float32) for o in obj]


def py2_float(v):
    """str(float) as Python 2 printed it: 12 significant digits."""
    s = "%.12g" % float(v)
    if "." not in s and "e" not in s and "n" not in s and "i" not in s:
        s += ".0"
    return s


def trace_header(trace_file):
    with open(trace_file, "w") as f:
        f.write("num_samples,L,Lvalid\n")


def trace_line(trace_file, num_samples, lb, lb_valid):
    with open(trace_file, "a") as f:
        f.write("{0},{1},{2}\n".format(num_samples, py2_float(lb), py2_float(lb_valid)))
