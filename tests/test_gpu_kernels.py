"""Single-kernel parity: the fp32 SIMT engine and the tcgen05 engine (bf16x3 split / bf16) against fp64 torch ops."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

PRECS = [(0, "simt", 1e-5), (1, "split", 1e-4), (2, "bf16", 1.5e-2)]  # relative to max|ref|; K up to 18432


def _lib():
    from ipoke_b200 import _lib
    return _lib


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 256, 128), (200, 48, 72), (1000, 144, 200), (4096, 2048, 2048)])
@pytest.mark.parametrize("prec,name,tol", PRECS)
def test_gemm(M, N, K, prec, name, tol):
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    out = torch.full((M, N), float("nan"), device="cuda")
    L.check(L.lib().ipk_test_gemm(A.data_ptr(), W.data_ptr(), out.data_ptr(), M, N, K, prec, None), "ipk_test_gemm")
    ref = (A.double() @ W.double().t())
    err = (out.double() - ref).abs().max().item()
    assert err < tol * max(1.0, ref.abs().max().item()), f"{name} gemm {M}x{N}x{K}: max err {err}"


@pytest.mark.parametrize("Fr,H,W,Cin,Cout", [(2, 8, 8, 32, 64), (4, 8, 8, 2048, 32), (3, 16, 16, 64, 128), (2, 32, 32, 128, 64),
                                              (1, 64, 64, 64, 64), (1, 128, 128, 64, 64), (5, 8, 8, 64, 256)])
@pytest.mark.parametrize("prec,name,tol", PRECS)
def test_conv3x3(Fr, H, W, Cin, Cout, prec, name, tol):
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(H * Cin + Cout)
    x = torch.randn(Fr, H, W, Cin, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (Cin * 9) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g)
    out = torch.full((Fr, H, W, Cout), float("nan"), device="cuda")
    L.check(L.lib().ipk_test_conv3x3(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), Fr, H, W, Cin, Cout, prec, None), "conv3x3")
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), padding=1).permute(0, 2, 3, 1)
    err = (out.double() - ref).abs().max().item()
    assert err < tol * max(1.0, ref.abs().max().item()), f"{name} conv3x3: max err {err}"


@pytest.mark.parametrize("Fr,H,W,Cin,Cout", [(2, 8, 8, 64, 64), (2, 16, 16, 256, 128), (1, 64, 64, 128, 64)])
@pytest.mark.parametrize("prec,name,tol", PRECS)
def test_convT3x3(Fr, H, W, Cin, Cout, prec, name, tol):
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(H * Cin + Cout + 7)
    x = torch.randn(Fr, H, W, Cin, device="cuda", generator=g)
    w = torch.randn(Cin, Cout, 3, 3, device="cuda", generator=g) / (Cin * 9) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g)
    out = torch.full((Fr, 2 * H, 2 * W, Cout), float("nan"), device="cuda")
    L.check(L.lib().ipk_test_convT3x3(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), Fr, H, W, Cin, Cout, prec, None), "convT3x3")
    ref = F.conv_transpose2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), stride=2, padding=1, output_padding=1).permute(0, 2, 3, 1)
    err = (out.double() - ref).abs().max().item()
    assert err < tol * max(1.0, ref.abs().max().item()), f"{name} convT: max err {err}"
