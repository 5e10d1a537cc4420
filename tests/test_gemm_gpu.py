"""GPU unit test of the GEMM tiers (fp32 SIMT, fp16 tcgen05, fp16x3 tcgen05) against float64 NumPy."""
import ctypes

import numpy as np
import pytest

import april_asr_b200 as A
import common as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    m = A.Model(C.TINY_MODEL, C.product_lib())
    yield m
    m.close()


def run(model, tier, a, w, b):
    M, K = a.shape
    N = w.shape[0]
    out = np.zeros((M, N), np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    rc = model.lib.c.aab_debug_gemm(model.handle, tier, M, N, K, a.ctypes.data_as(fp), w.ctypes.data_as(fp), b.ctypes.data_as(fp), out.ctypes.data_as(fp))
    assert rc == 0
    return out


# (256, 512, 2048), (64, 4096, 512), (1000, 512, 2304): few tiles and a long K, the engine cuts K into ranges (split-K, CTA pairs and single
# CTAs) and sums the partial tiles in a second kernel
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (200, 64, 64), (1000, 512, 2304), (777, 264, 144), (4096, 4096, 512), (49664 // 8, 2048, 512),
                                   (256, 512, 2048), (64, 4096, 512), (300, 1024, 1024)])
def test_gemm_tiers(model, M, N, K):
    rs = np.random.RandomState(M + N + K)
    a = (rs.randn(M, K) * rs.choice([0.01, 1.0, 30.0], size=(M, 1))).astype(np.float32)
    w = (rs.randn(N, K) / np.sqrt(K)).astype(np.float32)
    b = rs.randn(N).astype(np.float32)
    ref = a.astype(np.float64) @ w.astype(np.float64).T + b
    scale = np.abs(a).astype(np.float64) @ np.abs(w).astype(np.float64).T + np.abs(b)     # per-element magnitude of the sum
    for tier, tol in ((0, 2e-6), (3, 2e-6), (1, 2e-3)):
        out = run(model, tier, a, w, b)
        err = float((np.abs(out - ref) / scale).max())
        print("M=%d N=%d K=%d tier %d: max err / sum|a||w| = %.3g" % (M, N, K, tier, err))
        assert err < tol, (tier, err)
