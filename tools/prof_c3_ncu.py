import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vaeb_b200
from vaeb_b200.data import synthetic_mnist
M = 16384
x = synthetic_mnist(M * 2)
m = vaeb_b200.VAEB(x, False, 500, 20, M, 1, 0.01, False, False, precision=sys.argv[1] if len(sys.argv) > 1 else "bf16")
print(m.update_many(np.arange(3) % 2))
m.close()
