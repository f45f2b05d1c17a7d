"""Host logic of the manifold renderer (freyFace.py:346-369, VAEBImage.py) -- no GPU: the decoder is a stand-in."""
import numpy as np
import pytest
import scipy.stats

from vaeb_b200 import manifold


class _FakeModel:
    def __init__(self, D, continuous, Z=2):
        self.input_size, self.continuous, self.n_latent = D, continuous, Z

    def decode(self, z):
        y = np.tile(1.0 / (1.0 + np.exp(-z.sum(axis=1, keepdims=True))), (1, self.input_size)).astype(np.float32)
        return (y, np.full_like(y, -2.0)) if self.continuous else y


def test_grid_matches_the_reference_quantiles():
    z = manifold.grid_points()
    ref = np.asarray([[scipy.stats.norm.ppf((ii + 0.9) / 10.), scipy.stats.norm.ppf((jj + 0.9) / 10.)]
                      for ii in range(10) for jj in range(10)])       # freyFace.py:350-352
    np.testing.assert_allclose(z, ref, atol=1e-6)


@pytest.mark.parametrize("D,continuous,tile", [(784, False, (28, 28)), (560, True, (28, 20))])
def test_render_tiles_rows_of_ii_and_columns_of_jj(D, continuous, tile):
    faces, tiled = manifold.render(_FakeModel(D, continuous), n=10)
    assert faces.shape == (100, D)
    assert tiled.shape == (10 * tile[0], 10 * tile[1])
    # tile (ii, jj) holds face ii*10 + jj (VAEBImage.py:25-41: hstack over jj, vstack over ii)
    np.testing.assert_array_equal(tiled[3 * tile[0]:4 * tile[0], 7 * tile[1]:8 * tile[1]], manifold.to_image(faces[37]))
    s1, _ = manifold.render(_FakeModel(D, continuous), sample=True, rng=np.random.RandomState(0))
    assert (np.abs(s1 - faces).max() > 0) == continuous


def test_render_needs_two_latent_dimensions(tmp_path):
    with pytest.raises(ValueError, match="2-d latent"):
        manifold.render(_FakeModel(784, False, Z=3))
    _, tiled = manifold.render(_FakeModel(784, False), n=2)
    manifold.save_pgm(tiled, str(tmp_path / "m.pgm"))
    raw = open(tmp_path / "m.pgm", "rb").read()
    assert raw.startswith(b"P5 56 56 255\n") and len(raw) == len(b"P5 56 56 255\n") + 56 * 56


def test_save_image_orientation_and_inversion(tmp_path):
    """VAEBImage.save_image (VAEBImage.py:13-21): (1 - x)*255, MNIST rows in C order, Frey rows in F order rotated
    by -90 degrees; PGM without PIL, JPG through PIL like the reference."""
    import numpy as np
    from vaeb_b200 import manifold
    x = np.zeros(784, np.float32); x[3] = 1.0                  # pixel (row 0, col 3)
    a = manifold.save_image(x, str(tmp_path / "m.pgm"))
    assert a.shape == (28, 28) and a[0, 3] == 0 and a[0, 0] == 255
    raw = open(tmp_path / "m.pgm", "rb").read()
    assert raw.startswith(b"P5 28 28 255\n") and len(raw) == len(b"P5 28 28 255\n") + 784
    f = np.zeros(560, np.float32); f[1] = 1.0                  # F order: (row 1, col 0) of the 20 x 28 array
    b = manifold.save_image(f, str(tmp_path / "f.pgm"))
    assert b.shape == (28, 20) and (b == 0).sum() == 1
    try:
        import PIL  # noqa: F401
    except ImportError:
        return
    manifold.save_image(x, str(tmp_path / "m.jpg"))
    from PIL import Image
    assert Image.open(tmp_path / "m.jpg").size == (28, 28)


def test_reconstruction_mse_host_logic():
    """reconstruction.MSE = mean_i ||reconstruct(x_i) - x_i||^2 (reconstruction.py:9-18) -- host arithmetic only."""
    import numpy as np
    from vaeb_b200 import reconstruction as R

    class Fake(object):
        input_size, n_latent, continuous = 6, 2, False

        def reconstruct(self, x, n, eps=None):
            return x + 0.5

    x = np.arange(18, dtype=np.float32).reshape(3, 6)
    assert R.MSE(Fake(), x, 0) == 6 * 0.25
