"""GPU: the two fp32 FFMA GEMM engines of the step (32x32-tile kernel, thin backward kernel) against an fp64 product.

Reference math: nn.Linear forward and its autograd (LunarLander_Distributed_SAC/src/model.py:41-44,119-125).
Tolerance: fp32 accumulation of K <= 1280 products, relative L2 <= 2e-6 (written here, checked below)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 2e-6


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from distributed_sac_b200 import _lib
    torch.zeros(1, device="cuda")
    return _lib.load()


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def run(lib, engine, mode, A, B, M, N, K, ldc=None, bias=None, mask=None, relu=0, want_c2=False):
    from distributed_sac_b200 import _lib
    ldc = ldc or N
    Cm = torch.full((M, ldc), float("nan"), device="cuda")
    C2 = torch.full((M,), float("nan"), device="cuda") if want_c2 else None
    _lib.check(lib.b200sac_gemm_test(engine, mode, M, N, K, _p(A), A.stride(0), _p(B), B.stride(0), _p(bias), _p(mask),
                                     mask.stride(0) if mask is not None else 0, _p(Cm), ldc, _p(C2), relu, None))
    return Cm, C2


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _padded(rows, cols, pitch, g):
    """[rows][pitch] buffer whose first `cols` columns are data and whose padding is NaN (must never leak into a result)."""
    t = torch.full((rows, pitch), float("nan"), device="cuda")
    t[:, :cols] = torch.randn(rows, cols, device="cuda", generator=g)
    return t


# (M, N, K, pitch of the N-wide operand): LunarLander critic/actor input layers, padded and unpadded, ragged sizes
THIN = [(256, 10, 256, 12), (256, 8, 256, 8), (400, 16, 1280, 16), (64, 3, 100, 3), (37, 13, 1000, 16), (256, 10, 700, 10),
        (1, 1, 5, 4)]


@pytest.mark.parametrize("engine", [0, 2], ids=["tile", "thin"])
@pytest.mark.parametrize("M,N,K,ld", THIN)
def test_thin_wgrad(lib, engine, M, N, K, ld):
    """dW[m][n] = sum_k dZ[k][m] X[k][n], db[m] = sum_k dZ[k][m]; the output keeps the operand's pitch (padded weights)."""
    g = torch.Generator(device="cuda").manual_seed(M + 31 * N + K)
    dZ = torch.randn(K, M, device="cuda", generator=g)
    X = _padded(K, N, ld, g)
    out, c2 = run(lib, engine, 2, dZ, X, M, N, K, ldc=ld, want_c2=True)
    ref = dZ.double().T @ X[:, :N].double()
    assert rel(out[:, :N], ref) <= TOL
    assert torch.isnan(out[:, N:]).all()                      # padding columns are not written
    assert rel(c2, dZ.double().sum(0)) <= TOL


@pytest.mark.parametrize("engine", [0, 2], ids=["tile", "thin"])
@pytest.mark.parametrize("M,N,K,ld", THIN)
@pytest.mark.parametrize("masked", [False, True])
def test_thin_dgrad(lib, engine, M, N, K, ld, masked):
    """dX[m][n] = (sum_k dZ[m][k] W[k][n]) * [mask > 0]."""
    g = torch.Generator(device="cuda").manual_seed(M + 17 * N + K + int(masked))
    dZ = torch.randn(M, K, device="cuda", generator=g)
    W = _padded(K, N, ld, g)
    mask = torch.randn(M, ld, device="cuda", generator=g) if masked else None
    out, _ = run(lib, engine, 1, dZ, W, M, N, K, ldc=ld, mask=mask)
    ref = dZ.double() @ W[:, :N].double()
    if masked:
        ref = ref * (mask[:, :N] > 0)
    assert rel(out[:, :N], ref) <= TOL
    assert torch.isnan(out[:, N:]).all()


@pytest.mark.parametrize("M,N,K", [(512, 256, 8), (256, 256, 10), (1024, 400, 43), (1280, 400, 53), (256, 256, 256), (33, 40, 36),
                                   (2560, 400, 400), (100, 7, 300)])
def test_tile_fwd(lib, M, N, K):
    """h = relu(X W^T + b) with the first-layer operands at their 4-float padded pitch."""
    g = torch.Generator(device="cuda").manual_seed(M * 3 + N + K)
    ld = (K + 3) & ~3
    X, W = _padded(M, K, ld, g), _padded(N, K, ld, g)
    W[:, :K] *= 0.1
    b = torch.randn(N, device="cuda", generator=g)
    out, _ = run(lib, 0, 0, X, W, M, N, K, bias=b, relu=1)
    assert rel(out, torch.relu(X[:, :K].double() @ W[:, :K].double().T + b.double())) <= TOL


def test_thin_engine_rejects_wide_or_forward_problems(lib):
    A = torch.zeros(64, 64, device="cuda")
    out = torch.zeros(64, 64, device="cuda")
    assert lib.b200sac_gemm_test(2, 1, 64, 17, 64, _p(A), 64, _p(A), 64, None, None, 0, _p(out), 64, None, 0, None) < 0
    assert lib.b200sac_gemm_test(2, 0, 64, 8, 64, _p(A), 64, _p(A), 64, None, None, 0, _p(out), 64, None, 0, None) < 0
    assert lib.b200sac_gemm_test(7, 1, 64, 8, 64, _p(A), 64, _p(A), 64, None, None, 0, _p(out), 64, None, 0, None) < 0
