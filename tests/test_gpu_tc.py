"""GPU: the tcgen05 / TMA implicit-GEMM kernels against a torch fp32 reference computed on the SAME fp16-rounded operands
(so the comparison isolates the kernel: fp32 accumulation order is the only difference), and against the unrounded
fp32 result with the fp16-operand tolerance."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def s2():
    import sc2bench_b200
    return sc2bench_b200


CASES = [  # (c_in, c_out, k, pad, H, W, batch)
    (24, 512, 2, 1, 55, 55, 2),
    (512, 256, 2, 0, 56, 56, 2),
    (256, 256, 2, 1, 55, 55, 2),
    (64, 64, 1, 0, 9, 13, 3),
    (128, 128, 3, 1, 16, 16, 2),
    (64, 128, 2, 1, 7, 200, 1),
]


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('mode', ['f16', 'f32'])
def test_tc_conv_matches_torch(s2, case, mode):
    cin, cout, k, pad, H, W, B = case
    dev = torch.device('cuda:0')
    torch.manual_seed(sum(case))
    x = torch.randn(B, cin, H, W)
    w = torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5
    xh, wh = x.half().float(), w.half().float()
    ref = F.conv2d(xh.double(), wh.double(), None, 1, pad).float()
    cp = (cin + 63) // 64 * 64
    x_nhwc = s2.ops.nchw_to_nhwc_f16(x.to(dev), cp)
    assert torch.equal(x_nhwc[..., :cin].float().cpu(), xh.permute(0, 2, 3, 1)) and float(x_nhwc[..., cin:].abs().sum()) == 0
    wp = s2.ops.pack_conv_weight_f16(w.to(dev), cp)
    out = s2.ops.tc_conv(x_nhwc, wp, k, k, pad, mode=s2._native.TC_STORE_F32 if mode == 'f32' else s2._native.TC_STORE_F16)
    got = out.float().permute(0, 3, 1, 2).cpu()
    assert got.shape == ref.shape
    tol = 1e-5 if mode == 'f32' else 1.5e-3  # fp16 output rounding dominates in 'f16' mode
    assert rel_err(got, ref) < tol, rel_err(got, ref)


@pytest.mark.parametrize('C,inverse,H,W', [(512, True, 56, 56), (256, True, 55, 55), (64, False, 9, 13), (128, True, 16, 16)])
def test_tc_gdn1_matches_torch(s2, C, inverse, H, W):
    dev = torch.device('cuda:0')
    torch.manual_seed(C + H)
    x = torch.randn(2, C, H, W)
    gamma = 0.1 * torch.eye(C) + 0.02 * torch.rand(C, C) / (C / 64)
    beta = 0.5 + torch.rand(C)
    xh, gh = x.half().float(), gamma.half().float()
    norm = F.conv2d(xh.abs().double(), gh.double().view(C, C, 1, 1), beta.double())
    ref = (xh.double() * norm if inverse else xh.double() / norm).float()
    x_nhwc = s2.ops.nchw_to_nhwc_f16(x.to(dev), C)
    gp = gamma.half().to(dev).view(1, C, C).contiguous()
    out = s2.ops.tc_conv(x_nhwc, gp, 1, 1, 0, mode=s2._native.TC_IGDN1_F16 if inverse else s2._native.TC_GDN1_F16,
                         beta=beta.to(dev), gdn_x=x_nhwc)
    got = out.float().permute(0, 3, 1, 2).cpu()
    assert rel_err(got, ref) < 1.5e-3, rel_err(got, ref)
