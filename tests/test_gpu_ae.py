"""AE baselines through the C-ABI against golden vectors produced by the reference's own
degenerate-vae/ae.py and vanilla-ae/ae.py (tests/golden/make_reference_golden.py), and against the oracle at the
reference's own shapes."""
import numpy as np
import pytest

from oracle import vaeb_oracle as O
from tests.util import assert_close_tensor, load_golden

pytestmark = pytest.mark.gpu

ORDER_B = ["W3", "b3", "W4", "b4", "W1", "b1", "W2", "b2"]
ORDER_C = ["W3", "b3", "W4", "b4", "W1", "b1", "W2", "W6", "b2", "b6"]


def _sub(g, prefix):
    return {k[len(prefix) + 2:]: v for k, v in g.items() if k.startswith(prefix + "__")}


@pytest.mark.parametrize("tag,kind,otype", [("deg_binary", "degenerate", "binary"), ("deg_cont", "degenerate", "cont"),
                                            ("vanilla", "vanilla", "binary")])
def test_cuda_matches_reference_ae_baselines(tag, kind, otype):
    from vaeb_b200 import ae
    c = _sub(load_golden("ref_ae_baselines.npz"), tag)
    order = ORDER_C if otype == "cont" else ORDER_B
    H, Dz = c["init_W4"].shape
    train, reconstruct, encode, decode, theta = ae.ConstructAE(c["x"], Denc=[H], Dz=Dz, Ddec=[H], otype=otype, kind=kind)
    assert [t.name for t in theta] == order
    for t, n in zip(theta, order):
        t.set_value(c["init_" + n])
    rets = [float(train(row[row >= 0])) for row in c["idx"]]
    np.testing.assert_allclose(rets, c["train_returns"], rtol=1e-4)
    for t, n in zip(theta, order):
        assert_close_tensor(t.get_value(), c["final_" + n], 1e-4, 0.05, n)
    x6 = c["x"][:6]
    np.testing.assert_allclose(reconstruct(x6), c["reconstruct"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(encode(x6), c["encode"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(decode(c["encode"]), c["decode"], rtol=1e-4, atol=1e-6)
    train.model.close()


@pytest.mark.parametrize("kind,otype,D,H,Dz", [("degenerate", "cont", 560, 200, 2), ("degenerate", "binary", 784, 500, 20),
                                               ("vanilla", "binary", 784, 500, 5)])
def test_cuda_ae_matches_oracle_at_reference_shapes(kind, otype, D, H, Dz):
    """LearnFreyFace (560-200-Dz, cont) and LearnMNIST (784-500-Dz, binary) shapes, batch 100 gathered by a
    permutation (ae.py:146-152), the reference's own N(0, 0.01^2) initialisation incl. the biases."""
    from vaeb_b200 import ae
    cont = otype == "cont"
    x = (O.synthetic_frey(400) if cont else O.synthetic_mnist(400)).astype(np.float32)
    np.random.seed(5)
    train, reconstruct, encode, decode, theta = ae.ConstructAE(x, Denc=[H], Dz=Dz, Ddec=[H], otype=otype, kind=kind)
    m = train.model
    p0 = [p.astype(np.float64) for p in m._get_buffer(0)]
    assert np.abs(p0[m._names.index("b3")]).max() > 0            # mlp.BiasVector: biases are drawn, not zero
    o = O.OracleAE(x.astype(np.float64), cont, p0, kind=kind)
    rng = np.random.RandomState(9)
    for _ in range(3):
        idx = rng.permutation(400)[:100].astype(np.int32)
        assert float(train(idx)) == pytest.approx(o.train(idx), rel=1e-4)
    for a, b, n in zip(m._get_buffer(0), o.params, m._names):
        # AdaGrad normalises each step to ~eta whatever |g|: an entry whose gradient is a sum of cancelling terms
        # (fp32 relative error ~1e-3 of the entry) moves by up to 1e-3*eta per update differently
        err = np.abs(a.astype(np.float64) - b)
        assert np.all(err <= 1e-4 * np.maximum(np.abs(b), 0.05 * np.abs(b).max()) + 1e-3 * 0.01 * 3), (n, err.max())
    np.testing.assert_allclose(reconstruct(x[:50]), o.forward(x[:50].astype(np.float64)), rtol=1e-4, atol=1e-6)
    m.close()


def test_ae_learn_loop_improves_and_rejects_bad_arguments():
    from vaeb_b200 import ae
    x = O.synthetic_frey(300).astype(np.float32)
    np.random.seed(3)
    reconstruct, encode, decode, curve, rm = ae.LearnAE(x, epochs=3, Dz=2, Ntr=250, H=64, otype='cont', verbose=False)
    assert len(curve) == 3 * 3 and curve[-1] > curve[0]           # 250 rows -> 100 + 100 + 50 per epoch; loglik rises
    assert encode(x[:7]).shape == (7, 2) and decode(encode(x[:7])).shape == (7, 560)
    with pytest.raises(ValueError):
        ae.ConstructAE(x, Denc=[64, 32], Dz=2, Ddec=[64])
    with pytest.raises(ValueError):
        ae.AdaGrad(-1.0)
