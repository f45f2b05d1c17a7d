"""A few large-batch updates for an ncu capture of the chain kernel: python tools/ncu_chain.py bf16 16384"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vaeb_b200
from vaeb_b200.data import synthetic_mnist
prec, M = sys.argv[1], int(sys.argv[2])
x = synthetic_mnist(M * 2)
m = vaeb_b200.VAEB(x, False, 500, 20, M, 1, 0.01, False, False, precision=prec)
for i in range(3):
    m.update_many(np.array([i % 2]))
m.close()
