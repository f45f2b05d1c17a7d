#!/usr/bin/env python
"""Drop-in for the reference's `VAEBfullbayes.py` script (hard-coded Frey, n_latent=10,
2000 epochs, VAEBfullbayes.py:203-244).  `--synthetic` / `--n_epochs N` are extensions."""
import sys

from vaeb_b200.fullbayes import VAE, main  # noqa: F401

if __name__ == '__main__':
    n_epochs = int(sys.argv[sys.argv.index('--n_epochs') + 1]) if '--n_epochs' in sys.argv else 2000
    main(n_epochs=n_epochs, synthetic='--synthetic' in sys.argv)
