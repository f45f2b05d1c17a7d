"""Small driver for ncu captures: a few update() steps of config c2 (MNIST 784-500-20, M=100)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vaeb_b200
from vaeb_b200.data import synthetic_mnist

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 100
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
x = synthetic_mnist(max(4 * M, 2000))
m = vaeb_b200.VAEB(x, False, 500, 20, M, 1, 0.01, False, False, precision=prec)
out = m.update_many(np.arange(steps) % (len(x) // M))
print("bound", out)
m.close()
