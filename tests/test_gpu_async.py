"""Byte-valued host minibatches (vaeb_update_host_async_u8): the reference's data sets are 8-bit images stored as
pixel / 256 (mnist.pkl.gz, freyfaces.pkl; VAEB.py:230-239, 544-555).  Sending the bytes and expanding them on the device
must be indistinguishable from sending the float32 array the reference's loader holds."""
import numpy as np
import pytest

from oracle import vaeb_oracle as O

pytestmark = pytest.mark.gpu


def _bytes(n, D, seed):
    rng = np.random.RandomState(seed)
    return (rng.randint(0, 256, (n, D)) * (rng.uniform(size=(n, D)) < 0.25)).astype(np.uint8)


@pytest.mark.parametrize("precision,M,H,Z,continuous,D", [
    ("fp32", 100, 500, 20, False, 784),          # the single-launch step kernel behind the staging ring
    ("fp32", 100, 200, 2, True, 560),            # Frey: Gaussian decoder
    ("bf16x3", 1152, 500, 20, False, 784),       # large-batch tensor-core layers + tail
    ("fp32", 37, 24, 3, False, 45),              # ragged: D * rows not a multiple of 16 -> scalar tail of the expansion
])
def test_u8_minibatches_equal_float_minibatches(precision, M, H, Z, continuous, D):
    import vaeb_b200
    n_batches = 6
    xb = vaeb_b200.pinned_empty((n_batches * M, D), np.uint8)
    xb[:] = _bytes(n_batches * M, D, 11)
    if continuous:
        xb[:] = np.clip(xb, 3, 250)
    xf = vaeb_b200.pinned_empty(xb.shape, np.float32)
    xf[:] = xb.astype(np.float32) * np.float32(1.0 / 256.0)       # what the reference's pickles hold
    rng = np.random.RandomState(2)
    params = [rng.normal(0, 0.05, s).astype(np.float32) for s in O.param_shapes(D, H, Z, continuous)]
    ms = [vaeb_b200.VAEB(xf[:M], continuous, H, Z, M, 1, 0.01, False, False, params, precision=precision, seed=5)
          for _ in range(2)]
    for i in range(n_batches):                                     # more minibatches than one staging group
        ms[0].update_host_async(xf[i * M:(i + 1) * M])
        ms[1].update_host_async(xb[i * M:(i + 1) * M])
    a, b = ms[0].collect(), ms[1].collect()
    assert len(a) == n_batches and np.isfinite(a).all()
    np.testing.assert_array_equal(a, b)
    for p, q in zip(ms[0].get_params(), ms[1].get_params()):
        np.testing.assert_array_equal(p, q)
    for m in ms:
        m.close()


def test_u8_scale_and_argument_checks():
    import vaeb_b200
    M, D = 64, 48
    xb = vaeb_b200.pinned_empty((M, D), np.uint8)
    xb[:] = _bytes(M, D, 3)
    xf = vaeb_b200.pinned_empty((M, D), np.float32)
    xf[:] = xb.astype(np.float32) * np.float32(1.0 / 255.0)       # any positive scale: one fp32 product per element
    m1 = vaeb_b200.VAEB(xf, False, 16, 2, M, 1, 0.01, False, False, seed=1)
    m2 = vaeb_b200.VAEB(xf, False, 16, 2, M, 1, 0.01, False, False, seed=1)
    m1.update_host_async(xf)
    m2.update_host_async(xb, scale=1.0 / 255.0)
    np.testing.assert_array_equal(m1.collect(), m2.collect())
    with pytest.raises(Exception):
        m2.update_host_async(xb.copy())                           # pageable memory
    with pytest.raises(Exception):
        m2.update_host_async(xb, scale=0.0)
    with pytest.raises(ValueError):
        m2.update_host_async(xb[:, :40])                           # wrong width / not contiguous
    m1.close(); m2.close()
