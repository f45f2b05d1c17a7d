"""tcgen05 / TMA building blocks on the B200: the generic bf16 GEMM in all four operand-major
combinations, ragged shapes, every UMMA N used by the step kernels."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bf16_round(a):
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def _run(M, N, K, a_mn, b_mn, bn, seed=0):
    from vaeb_b200 import _lib
    lib = _lib.load()
    rng = np.random.RandomState(seed)
    A = rng.normal(size=(M, K)).astype(np.float32)
    B = rng.normal(size=(K, N)).astype(np.float32)
    Cm = np.empty((M, N), np.float32)
    _lib.check(lib.vaeb_tc_gemm_test(0, M, N, K, a_mn, b_mn, bn, A.ctypes.data_as(C.c_void_p),
                                     B.ctypes.data_as(C.c_void_p), Cm.ctypes.data_as(C.c_void_p)))
    ref = _bf16_round(A).astype(np.float64) @ _bf16_round(B).astype(np.float64)
    return Cm, ref


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("bn", [64, 128, 256])
def test_tc_gemm_majors(a_mn, b_mn, bn):
    Cm, ref = _run(256, 512, 192, a_mn, b_mn, bn)
    np.testing.assert_allclose(Cm, ref, rtol=2e-5, atol=2e-4)


@pytest.mark.parametrize("M,N,K", [(104, 504, 784), (128, 784, 504), (504, 784, 104), (8, 8, 8), (136, 72, 520)])
def test_tc_gemm_ragged_shapes(M, N, K):
    # the AEVB shapes: M=100->104, H=500->504, D=784; TMA zero-fills the out-of-bounds part of a box
    for a_mn, b_mn in ((0, 1), (0, 0), (1, 1)):
        Cm, ref = _run(M, N, K, a_mn, b_mn, 64, seed=M + N)
        np.testing.assert_allclose(Cm, ref, rtol=2e-5, atol=5e-4)


def test_tc_gemm_small_n_tile():
    Cm, ref = _run(128, 96, 256, 0, 0, 32)
    np.testing.assert_allclose(Cm, ref, rtol=2e-5, atol=2e-4)
