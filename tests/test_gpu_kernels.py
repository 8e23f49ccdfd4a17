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


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 256, 128), (200, 48, 72), (1000, 144, 200), (4096, 2048, 2048),
                                   (512, 288, 2048), (300, 80, 128), (64, 2048, 2048), (640, 160, 64), (256, 272, 192)])
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


# (B, T, H, W, Cin, Cout, kernel, stride, pad): the encoder's layer shapes plus ragged batches (B not a multiple of the batch box),
# odd T, temporal stride without padding (downsample 1x1x1), and single-time-step volumes where 2/3 of the taps are skipped
CONV3D = [
    (2, 6, 32, 32, 64, 128, (3, 3, 3), (2, 1, 1), (1, 1, 1)),
    (3, 3, 32, 32, 128, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    (2, 3, 32, 32, 128, 256, (3, 3, 3), (2, 2, 2), (1, 1, 1)),
    (3, 2, 16, 16, 256, 256, (3, 3, 3), (2, 2, 2), (1, 1, 1)),
    (5, 1, 8, 8, 256, 256, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    (3, 5, 16, 16, 64, 64, (1, 1, 1), (2, 2, 2), (0, 0, 0)),
    (1, 4, 128, 128, 64, 64, (3, 3, 3), (1, 2, 2), (1, 1, 1)),
]


@pytest.mark.parametrize("shape", CONV3D)
@pytest.mark.parametrize("prec,name,tol", PRECS[1:])
def test_conv3d_tc(shape, prec, name, tol):
    """conv3d_tc_kernel (5-D TMA boxes with element strides, temporal-padding taps skipped, fused statistics) vs torch conv3d in fp64."""
    L = _lib()
    B, T, H, W, Cin, Cout, k, s, p = shape
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T * 100 + Cin + Cout)
    x = torch.randn(B, T, H, W, Cin, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, *k, device="cuda", generator=g) / (Cin * k[0] * k[1] * k[2]) ** 0.5
    ref = F.conv3d(x.permute(0, 4, 1, 2, 3).double(), w.double(), None, stride=s, padding=p).permute(0, 2, 3, 4, 1).contiguous()
    out = torch.full(ref.shape, float("nan"), device="cuda", dtype=torch.float32)
    stats = torch.zeros((B, Cout, 2), device="cuda", dtype=torch.float64)
    dims = (ctypes.c_int32 * 15)(B, T, H, W, Cin, Cout, *k, *s, *p)
    L.check(L.lib().ipk_test_conv3d(x.data_ptr(), w.data_ptr(), out.data_ptr(), stats.data_ptr(), dims, prec, None), "ipk_test_conv3d")
    scale = max(1.0, ref.abs().max().item())
    err = (out.double() - ref).abs().max().item()
    assert err < tol * scale, f"{name} conv3d {shape}: max err {err}"
    # fused GroupNorm statistics = sums of the values the kernel wrote
    s1 = out.double().sum(dim=(1, 2, 3))
    s2 = (out.double() ** 2).sum(dim=(1, 2, 3))
    n = ref[0, ..., 0].numel()
    assert (stats[..., 0] - s1).abs().max().item() < 1e-4 * n ** 0.5 * scale
    assert (stats[..., 1] - s2).abs().max().item() < 1e-4 * n * scale
