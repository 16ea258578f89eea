"""tcgen05 / TMEM GEMM vs the CPU oracle (kept in its own file: see tests/test_gpu_kernels.py)."""
import numpy as np
import pytest

from oracle import forward as ofw
from test_gpu_kernels import _qgemm_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('bits', [2, 3, 4])
@pytest.mark.parametrize('M', [33, 128, 129, 300, 2048])
@pytest.mark.parametrize('symmetric', [True, False])
def test_qgemm_tc_vs_oracle(bits, M, symmetric):
    from gpu_util import run_qgemm
    shapes = [(128, 128), (256, 1024), (384, 640)] if M < 2048 else [(512, 1024)]
    for (N, K) in shapes:
        codes, scales, zeros, X, bias, want = _qgemm_case(bits, N, K, M, symmetric, bits * 10 + M)
        z, _ = run_qgemm(codes, scales, zeros, bits, X, path=2, bias=bias, symmetric=symmetric)
        assert not np.isnan(z.astype(np.float32)).any()
        err = ofw.rel_err(z, want)
        assert err < 3e-4, (bits, M, symmetric, N, K, err)


def test_qgemm_tc_small_M_and_ragged_N():
    from gpu_util import run_qgemm
    for (N, K, M) in [(48, 256, 5), (144, 384, 40), (4096, 4096, 64)]:
        codes, scales, zeros, X, bias, want = _qgemm_case(2, N, K, M, False, N + M)
        z, _ = run_qgemm(codes, scales, zeros, 2, X, path=2, bias=bias, symmetric=False)
        assert ofw.rel_err(z, want) < 3e-4, (N, K, M)


def test_tc_and_skinny_agree():
    from gpu_util import run_qgemm
    codes, scales, zeros, X, bias, want = _qgemm_case(2, 512, 1024, 24, True, 99)
    a, _ = run_qgemm(codes, scales, zeros, 2, X, path=1, bias=bias, symmetric=True)
    b, _ = run_qgemm(codes, scales, zeros, 2, X, path=2, bias=bias, symmetric=True)
    assert ofw.rel_err(a, b) < 2e-4
