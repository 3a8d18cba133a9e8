"""GPU: the tcgen05 3xTF32 GEMM kernel (TMA + TMEM) against an fp64 reference, all three kinds."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[(64, 1), (128, 1), (128, 0), (160, 1)],
                ids=["tile128x64", "tile128x128-paired", "tile128x128", "tile128x160"])
def lib(request):
    """Every tile variant of the kernel (the step picks per launch group by a cost model; B200SAC_TC_BN forces a width,
    B200SAC_TC_PAIR128=0 the unpaired 128-wide flavour)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import os
    from distributed_sac_b200 import _lib
    os.environ["B200SAC_TC_BN"] = str(request.param[0])
    os.environ["B200SAC_TC_PAIR128"] = str(request.param[1])
    yield _lib.load()
    os.environ.pop("B200SAC_TC_BN", None)
    os.environ.pop("B200SAC_TC_PAIR128", None)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def run_tc(lib, mode, A, B, M, N, K, bias=None, mask=None, relu=0, want_c2=False):
    from distributed_sac_b200 import _lib
    Cm = torch.full((M, N), float("nan"), device="cuda")
    C2 = torch.full((M,), float("nan"), device="cuda") if want_c2 else None
    _lib.check(lib.b200sac_tc_gemm_test(mode, M, N, K, _p(A), A.stride(0), _p(B), B.stride(0), _p(bias), _p(mask),
                                        mask.stride(0) if mask is not None else 0, _p(Cm), Cm.stride(0), _p(C2), relu, None))
    return Cm, C2


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


# fp32-class accuracy: 3xTF32 drops only lo*lo (~2^-22 relative per product); the tensor core's
# truncating fp32 accumulation adds a bias ~ (#MMAs per accumulator) * 2^-24 (rotating accumulators: up to 4 pairs for
# 64-wide tiles, 2 pairs or 3 mains + 1 cross for 128-wide ones, 2 mains + 1 cross for 160-wide ones)
TOL = 5e-6

SHAPES = [(256, 256, 256), (512, 256, 256), (128, 64, 32), (1024, 400, 400), (1280, 400, 400), (200, 72, 40),
          (96, 48, 64), (33, 40, 36)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_fwd(lib, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.1
    b = torch.randn(N, device="cuda", generator=g)
    out, _ = run_tc(lib, 0, A, W, M, N, K, bias=b, relu=1)
    ref = torch.relu(A.double() @ W.double().T + b.double())
    assert torch.isfinite(out).all()
    assert rel(out, ref) < TOL, rel(out, ref)
    out2, _ = run_tc(lib, 0, A, W, M, N, K)            # no bias, no relu
    assert rel(out2, A.double() @ W.double().T) < TOL


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_dgrad(lib, M, N, K):
    """C[M][N] = (dY[M][K] @ W[K][N]) * (h > 0)"""
    g = torch.Generator(device="cuda").manual_seed(M + N * 3)
    dY = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(K, N, device="cuda", generator=g) * 0.1
    h = torch.relu(torch.randn(M, N, device="cuda", generator=g))
    out, _ = run_tc(lib, 1, dY, W, M, N, K, mask=h)
    ref = (dY.double() @ W.double()) * (h > 0)
    assert torch.isfinite(out).all()
    assert rel(out, ref) < TOL, rel(out, ref)
    out2, _ = run_tc(lib, 1, dY, W, M, N, K)
    assert rel(out2, dY.double() @ W.double()) < TOL


@pytest.mark.parametrize("M,N,K", [(256, 256, 256), (400, 400, 1024), (400, 400, 1280), (72, 40, 200), (40, 72, 96),
                                   (256, 256, 512), (64, 32, 32)])
def test_wgrad(lib, M, N, K):
    """C[M][N] = dY[K][M]^T @ X[K][N];  C2[M] = column sums of dY  (M = out features, K = batch)"""
    g = torch.Generator(device="cuda").manual_seed(M * 5 + K)
    dY = torch.randn(K, M, device="cuda", generator=g)
    X = torch.randn(K, N, device="cuda", generator=g)
    out, c2 = run_tc(lib, 2, dY, X, M, N, K, want_c2=True)
    ref = dY.double().T @ X.double()
    assert torch.isfinite(out).all() and torch.isfinite(c2).all()
    assert rel(out, ref) < TOL, rel(out, ref)
    assert rel(c2, dY.double().sum(0)) < 1e-6, rel(c2, dY.double().sum(0))


def test_strided_views(lib):
    """Operands that are row ranges / column-offset views of larger buffers (as the step uses them)."""
    g = torch.Generator(device="cuda").manual_seed(3)
    big = torch.randn(512, 256, device="cuda", generator=g)
    A = big[256:]                                           # rows B..2B-1 of the actor activations
    W = torch.randn(256, 256, device="cuda", generator=g)
    out, _ = run_tc(lib, 0, A, W, 256, 256, 256)
    assert rel(out, A.double() @ W.double().T) < TOL
