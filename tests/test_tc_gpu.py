"""GPU parity for the tcgen05 bf16 tensor-core GEMM (mlb_gemm_bf16_tc) in all operand
major-ness combinations / epilogues, vs float64 products of the SAME bf16-rounded inputs
(so the only difference is fp32 accumulation order: tolerance rel-L2 1e-5 for fp32 output,
4e-3 for bf16 output = bf16 rounding of the result)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _tc(mlb, A, B, M, N, K, a_mn, b_mn, epi, splitk=1, bias=None, C0=None):
    from madrona_learn_b200._lib import c_int, call, ptr
    Ad = A.to(DEV).to(torch.bfloat16).contiguous()
    Bd = B.to(DEV).to(torch.bfloat16).contiguous()
    if epi == 1:
        C = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
    else:
        C = torch.zeros(M, N, dtype=torch.float32, device=DEV) if C0 is None else C0.to(DEV).clone()
    bd = None if bias is None else bias.to(DEV)
    call('mlb_gemm_bf16_tc', ptr(Ad), ptr(Bd), ptr(C), ptr(bd), c_int(M), c_int(N), c_int(K),
         c_int(Ad.shape[1]), c_int(Bd.shape[1]), c_int(N), c_int(a_mn), c_int(b_mn), c_int(epi), c_int(splitk))
    torch.cuda.synchronize()
    return C.float().cpu().double().numpy()


def _ref(A, B, a_mn, b_mn):
    a = A.to(torch.bfloat16).double()
    b = B.to(torch.bfloat16).double()
    a = a.t() if a_mn else a            # -> [M, K]
    b = b if b_mn else b.t()            # -> [K, N]
    return (a @ b).numpy()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize('M,N,K', [(128, 64, 64), (128, 256, 64), (256, 256, 256), (300, 128, 192),
                                   (1000, 64, 256), (8192, 256, 64), (4096, 512, 512), (128, 64, 40)])
def test_tc_gemm_kmajor_fwd(mlb, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)                      # [N, K] K-major (W^T)
    bias = torch.randn(N, generator=g)
    got = _tc(mlb, A, B, M, N, K, 0, 0, 0, bias=bias)
    assert _rel(got, _ref(A, B, 0, 0) + bias.double().numpy()) < 1e-5
    got = _tc(mlb, A, B, M, N, K, 0, 0, 1)
    assert _rel(got, _ref(A, B, 0, 0)) < 4e-3


@pytest.mark.parametrize('M,N,K,splitk', [(64, 256, 1024, 1), (256, 256, 4096, 8), (256, 64, 3000, 4),
                                          (128, 128, 640, 2), (512, 512, 2048, 4)])
def test_tc_gemm_mnmajor_dw(mlb, M, N, K, splitk):
    """dW = X^T dZ: A stored [K, M], B stored [K, N], split-K atomics into a pre-set C."""
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(K, M, generator=g)
    B = torch.randn(K, N, generator=g)
    C0 = torch.randn(M, N, generator=g)
    got = _tc(mlb, A, B, M, N, K, 1, 1, 2, splitk=splitk, C0=C0)
    assert _rel(got, _ref(A, B, 1, 1) + C0.double().numpy()) < 1e-5


@pytest.mark.parametrize('a_mn,b_mn', [(1, 0), (0, 1)])
def test_tc_gemm_mixed_major(mlb, a_mn, b_mn):
    M, N, K = 256, 128, 320
    g = torch.Generator().manual_seed(7)
    A = torch.randn((K, M) if a_mn else (M, K), generator=g)
    B = torch.randn((K, N) if b_mn else (N, K), generator=g)
    got = _tc(mlb, A, B, M, N, K, a_mn, b_mn, 0)
    assert _rel(got, _ref(A, B, a_mn, b_mn)) < 1e-5


def test_cast_kernels(mlb):
    from madrona_learn_b200._lib import c_int, c_ll, call, ptr
    x = torch.randn(1003, device=DEV)
    y = torch.empty(1003, dtype=torch.bfloat16, device=DEV)
    call('mlb_cast_f32_bf16', ptr(x), ptr(y), c_ll(1003))
    assert torch.equal(y, x.to(torch.bfloat16))
    W = torch.randn(70, 45, device=DEV)
    Wt = torch.zeros(45, 72, dtype=torch.bfloat16, device=DEV)
    Wc = torch.zeros(70, 48, dtype=torch.bfloat16, device=DEV)
    call('mlb_cast_weight_bf16', ptr(W), ptr(Wt), ptr(Wc), c_int(70), c_int(45), c_int(45), c_int(72), c_int(48))
    assert torch.equal(Wt[:, :70], W.t().to(torch.bfloat16))
    assert torch.equal(Wc[:, :45], W.to(torch.bfloat16))
