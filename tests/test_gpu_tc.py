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


# ---------------------------------------------------------------------------------------------------
# fp32-grade "split fp16" tensor-core kernels (g_a)
# ---------------------------------------------------------------------------------------------------
SPLIT_TOL = 3e-6  # max-abs / max-abs vs an fp64 reference of the SAME fp32 operands


def _planes(s2, x, dev, parity=False):
    """fp32 NCHW -> split NHWC planes on the device; parity=True gives the [B * 4, H/2, W/2, C] parity-plane layout."""
    if parity:
        B, C, H, W = x.shape
        x = torch.stack([x[:, :, py::2, px::2] for py in (0, 1) for px in (0, 1)], dim=1).reshape(B * 4, C, H // 2, W // 2)
    hi, lo = s2.ops.split_f16(x.permute(0, 2, 3, 1).contiguous())
    return hi.to(dev), lo.to(dev)


def _unsplit(hi, lo):
    return (hi.float() + lo.float() / 2048.0).permute(0, 3, 1, 2).cpu()


def test_patchify_matches_unfold(s2):
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    x = torch.randn(2, 3, 36, 44)
    hi, lo = s2.ops.patchify_split(x.to(dev), 5, 5, 2, 2, 80)
    got = hi.float() + lo.float() / 2048.0                      # [B*4, 9, 11, 80]
    cols = F.unfold(x, 5, padding=2, stride=2).view(2, 75, 18, 22)  # K order (c, dy, dx), like the weight
    want = torch.stack([cols[:, :, py::2, px::2] for py in (0, 1) for px in (0, 1)], dim=1).reshape(8, 75, 9, 11).permute(0, 2, 3, 1)
    assert got.shape == (8, 9, 11, 80)
    assert float((got[..., :75].cpu() - want).abs().max()) < 1e-6 and float(got[..., 75:].abs().max()) == 0.0


@pytest.mark.parametrize('cin,cout,k,stride,pad,H,W', [(96, 48, 5, 2, 2, 112, 112), (48, 24, 2, 1, 0, 56, 56), (32, 16, 5, 2, 2, 24, 40),
                                                       (16, 8, 2, 1, 0, 9, 13), (64, 128, 3, 1, 1, 12, 20)])
def test_tc_split_conv_matches_fp64(s2, cin, cout, k, stride, pad, H, W):
    dev = torch.device('cuda:0')
    torch.manual_seed(cin + cout + H)
    x = torch.randn(2, cin, H, W) * 2
    w = torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5
    ref = F.conv2d(x.double(), w.double(), None, stride, pad)
    xh, xl = _planes(s2, x, dev, parity=stride == 2)
    wh, wl = s2.ops.pack_conv_weight_split(w.to(dev), c_in_pad=cin)
    oh, ol = s2.ops.tc_split_conv(xh, xl, wh, wl, cout, k, k, stride, pad, s2._native.TCS_STORE)
    got = _unsplit(oh, ol)[:, :cout]
    assert got.shape == ref.shape
    assert rel_err(got, ref.float()) < SPLIT_TOL, rel_err(got, ref.float())
    med = torch.randn(cout)
    sym = s2.ops.tc_split_conv(xh, xl, wh, wl, cout, k, k, stride, pad, s2._native.TCS_QUANT, medians=med.to(dev))
    want = torch.round(ref.float() - med.view(1, -1, 1, 1)).int()
    assert sym.shape == want.shape
    assert int((sym.cpu() != want).sum()) <= max(1, want.numel() // 100000)  # ties at fp32 resolution only


def test_tc_split_first_layer_via_patches(s2):
    dev = torch.device('cuda:0')
    torch.manual_seed(5)
    x = torch.randn(2, 3, 64, 48)
    w = torch.randn(96, 3, 5, 5) / 75 ** 0.5
    ref = F.conv2d(x.double(), w.double(), None, 2, 2)           # [2, 96, 32, 24]
    ph, pl = s2.ops.patchify_split(x.to(dev), 5, 5, 2, 2, 80)
    wh, wl = s2.ops.pack_conv_weight_split(w.to(dev), as_patches=True)
    assert wh.shape == (1, 96, 80)
    oh, ol = s2.ops.tc_split_conv(ph, pl, wh, wl, 96, 1, 1, 1, 0, s2._native.TCS_STORE)  # parity-plane output [8, 16, 12, 96]
    got = _unsplit(oh, ol).view(2, 2, 2, 96, 16, 12)
    full = torch.zeros(2, 96, 32, 24)
    for py in (0, 1):
        for px in (0, 1):
            full[:, :, py::2, px::2] = got[:, py, px]
    assert rel_err(full, ref.float()) < SPLIT_TOL


@pytest.mark.parametrize('C,H,W', [(96, 20, 28), (48, 56, 56), (32, 7, 9), (16, 5, 5)])
def test_tc_split_gdn1_matches_fp64(s2, C, H, W):
    dev = torch.device('cuda:0')
    torch.manual_seed(C)
    x = torch.randn(2, C, H, W) * 3
    gamma = 0.1 * torch.eye(C) + 0.02 * torch.rand(C, C)
    beta = 0.5 + torch.rand(C)
    norm = F.conv2d(x.abs().double(), gamma.double().view(C, C, 1, 1), beta.double())
    ref = (x.double() / norm).float()
    xh, xl = _planes(s2, x, dev)
    gh, gl = s2.ops.pack_conv_weight_split(gamma.view(C, C, 1, 1).to(dev), c_in_pad=C)
    oh, ol = s2.ops.tc_split_conv(xh, xl, gh, gl, C, 1, 1, 1, 0, s2._native.TCS_GDN1, beta=beta.to(dev), gdn=True)
    assert rel_err(_unsplit(oh, ol)[:, :C], ref) < SPLIT_TOL


@pytest.mark.parametrize('cout,H,W', [(96, 64, 48), (96, 224, 224), (32, 36, 44), (48, 20, 28)])
def test_tc_first_layer_fused_im2col(s2, cout, H, W):
    dev = torch.device('cuda:0')
    torch.manual_seed(cout + H)
    x = torch.randn(2, 3, H, W) * 1.5
    w = torch.randn(cout, 3, 5, 5) / 75 ** 0.5
    ref = F.conv2d(x.double(), w.double(), None, 2, 2)
    wh, wl = s2.ops.pack_conv_weight_split(w.to(dev), as_patches=True)
    oh, ol = s2.ops.tc_first_layer(x.to(dev), wh, wl, cout, 5, 5, 2)
    hp, wp = ref.shape[2] // 2, ref.shape[3] // 2
    got = _unsplit(oh, ol)[:, :cout].view(2, 2, 2, cout, hp, wp)
    full = torch.zeros_like(ref, dtype=torch.float32)
    for py in (0, 1):
        for px in (0, 1):
            full[:, :, py::2, px::2] = got[:, py, px]
    assert rel_err(full, ref.float()) < SPLIT_TOL
    # identical to the unfused route (patchify + 1x1 GEMM)
    ph, pl = s2.ops.patchify_split(x.to(dev), 5, 5, 2, 2, 80)
    uh, ul = s2.ops.tc_split_conv(ph, pl, wh, wl, cout, 1, 1, 1, 0, s2._native.TCS_STORE)
    assert torch.equal(uh, oh) and torch.equal(ul, ol)


# ---------------------------------------------------------------------------------------------------
# round 2: fused conv + GDN1 kernels of g_a (halo tiles, stacked weights, gamma GEMM in the epilogue)
# ---------------------------------------------------------------------------------------------------
def _gdn1_ref64(x64, gamma, beta):
    C = beta.numel()
    norm = F.conv2d(x64.abs(), gamma.double().view(C, C, 1, 1), beta.double())
    return x64 / norm


@pytest.mark.parametrize('cin,cout,H,W,batch', [(96, 48, 112, 112, 2), (96, 48, 112, 112, 24), (96, 48, 40, 72, 3), (32, 16, 24, 40, 2), (64, 80, 36, 20, 1),
                                                 (48, 24, 400, 140, 1), (192, 64, 20, 28, 2)])
def test_ga_halo_conv_gdn_matches_fp64(s2, cin, cout, H, W, batch):
    """sc2_ga_halo_conv_gdn = Conv2d(k5, s2, p2) + GDN1 (layer.py:479-481) vs an fp64 reference of the same fp32 operands, and
    vs the round-1 two-kernel route (which it must match to fp32 rounding)."""
    dev = torch.device('cuda:0')
    torch.manual_seed(cin + cout + H)
    x = torch.randn(batch, cin, H, W) * 2
    w = torch.randn(cout, cin, 5, 5) / (cin * 25) ** 0.5
    gamma = 0.1 * torch.eye(cout) + 0.02 * torch.rand(cout, cout)
    beta = 0.5 + torch.rand(cout)
    ref = _gdn1_ref64(F.conv2d(x.double(), w.double(), None, 2, 2), gamma, beta).float()
    xh, xl = _planes(s2, x, dev, parity=True)
    ws = s2.ops.pack_conv_weight_stacked(w.to(dev))
    n = ws.shape[1] // 2
    gs = s2.ops.pack_conv_weight_stacked(gamma.view(cout, cout, 1, 1).to(dev), n=n, c_in_pad=n)[0]
    oh, ol = s2.ops.ga_halo_conv_gdn(xh, xl, ws, gs, beta.to(dev), cout, 5, 5, 2)
    got = _unsplit(oh, ol)
    assert got.shape[1] == (cout + 7) // 8 * 8 and float(got[:, cout:].abs().max() if got.shape[1] > cout else 0.0) == 0.0
    got = got[:, :cout]
    assert got.shape == ref.shape
    assert rel_err(got, ref) < SPLIT_TOL, rel_err(got, ref)


def _to_parity_full(got, B, cout):
    """split parity planes [B*4, hp, wp, C] (already unsplit to NCHW [B*4, C, hp, wp]) -> full-resolution [B, C, 2hp, 2wp]."""
    hp, wp = got.shape[2], got.shape[3]
    g = got[:, :cout].view(B, 2, 2, cout, hp, wp)
    full = torch.zeros(B, cout, 2 * hp, 2 * wp)
    for py in (0, 1):
        for px in (0, 1):
            full[:, :, py::2, px::2] = g[:, py, px]
    return full


@pytest.mark.parametrize('cout,H,W,batch', [(96, 224, 224, 2), (96, 64, 48, 3), (32, 36, 44, 2), (48, 20, 28, 1), (96, 160, 336, 1), (80, 40, 24, 2)])
def test_ga_first_conv_gdn_matches_fp64(s2, cout, H, W, batch):
    """sc2_ga_first_conv_gdn = Conv2d(3 -> C, k5, s2, p2) + GDN1 (layer.py:476-478) vs an fp64 reference of the same fp32 operands."""
    dev = torch.device('cuda:0')
    torch.manual_seed(cout + H)
    x = torch.randn(batch, 3, H, W) * 1.5
    w = torch.randn(cout, 3, 5, 5) / 75 ** 0.5
    gamma = 0.1 * torch.eye(cout) + 0.02 * torch.rand(cout, cout)
    beta = 0.5 + torch.rand(cout)
    ref = _gdn1_ref64(F.conv2d(x.double(), w.double(), None, 2, 2), gamma, beta).float()
    ws = s2.ops.pack_first_layer_stacked(w.to(dev))
    n = ws.shape[0] // 2
    gs = s2.ops.pack_conv_weight_stacked(gamma.view(cout, cout, 1, 1).to(dev), n=n, c_in_pad=n)[0]
    oh, ol = s2.ops.ga_first_conv_gdn(x.to(dev), ws, gs, beta.to(dev), cout)
    full = _to_parity_full(_unsplit(oh, ol), batch, cout)
    assert full.shape == ref.shape
    assert rel_err(full, ref) < SPLIT_TOL, rel_err(full, ref)


def test_ga_first_uint8_lut_equals_float_input(s2):
    """Device-side ToTensor + Normalize: a uint8 image through the look-up table gives BIT-IDENTICAL planes to
    the fp32 image the data loader would have produced with the same torch ops."""
    dev = torch.device('cuda:0')
    torch.manual_seed(7)
    u8 = torch.randint(0, 256, (2, 3, 224, 224), dtype=torch.uint8)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    xf = u8.float().div(255).sub(torch.tensor(mean).view(1, 3, 1, 1)).div(torch.tensor(std).view(1, 3, 1, 1))
    w = torch.randn(96, 3, 5, 5) / 75 ** 0.5
    gamma = 0.1 * torch.eye(96) + 0.02 * torch.rand(96, 96)
    beta = 0.5 + torch.rand(96)
    ws = s2.ops.pack_first_layer_stacked(w.to(dev))
    gs = s2.ops.pack_conv_weight_stacked(gamma.view(96, 96, 1, 1).to(dev), n=96, c_in_pad=96)[0]
    lut = s2.ops.normalize_lut(mean, std, dev)
    fh, fl = s2.ops.ga_first_conv_gdn(xf.to(dev), ws, gs, beta.to(dev), 96)
    uh, ul = s2.ops.ga_first_conv_gdn(u8.to(dev), ws, gs, beta.to(dev), 96, lut=lut)
    assert torch.equal(fh, uh) and torch.equal(fl, ul)


# ---------------------------------------------------------------------------------------------------
# round 2: synthesis transform of the zoo codecs on the tensor cores (transposed conv as parity sub-convolutions, GDN proper)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('cin,cout,H,W,batch', [(320, 192, 16, 16, 2), (192, 192, 32, 32, 2), (64, 64, 5, 7, 3), (192, 3, 20, 12, 2), (128, 128, 9, 9, 1)])
def test_tc_deconv5_parity_classes_match_torch(s2, cin, cout, H, W, batch):
    """ConvTranspose2d(k5, s2, p2, op1) + bias as four sc2_tc_conv_ex launches vs torch on the same fp16-rounded operands; the x^2
    side output of mode 6; the clamped fp32 NCHW output of the last layer (mode 8)."""
    dev = torch.device('cuda:0')
    torch.manual_seed(cin + cout + H)
    x = torch.randn(batch, cin, H, W)
    w = torch.randn(cin, cout, 5, 5) / (cin * 25 / 4) ** 0.5
    b = torch.randn(cout) * 0.1
    xh, wh = x.half().float(), w.half().float()
    ref = F.conv_transpose2d(xh.double(), wh.double(), b.double(), stride=2, padding=2, output_padding=1).float()
    last = cout <= 32
    x_nhwc = s2.ops.nchw_to_nhwc_f16(x.to(dev), (cin + 63) // 64 * 64)
    packs = s2.ops.pack_deconv5_weight_f16(w.to(dev), rows_pad=32 if last else None)
    T = s2._native
    if last:
        out = torch.full((batch, cout, 2 * H, 2 * W), -7.0, dtype=torch.float32, device=dev)
        bias = torch.zeros(32, device=dev)
        bias[:cout] = b.to(dev)
        sq = None
    else:
        out = torch.full((batch, 2 * H, 2 * W, cout), -7.0, dtype=torch.float16, device=dev)
        sq = torch.empty_like(out)
        bias = b.to(dev)
    for (py, px), pk in packs.items():
        (ky, pad_y), (kx, pad_x) = s2.ops.DECONV5_TAPS[py], s2.ops.DECONV5_TAPS[px]
        s2.ops.tc_conv_ex(x_nhwc, pk, len(ky), len(kx), pad_y, pad_x, T.TC_NCHW_F32_CLAMP if last else T.TC_STORE_SQ_F16, (H, W), out,
                          out_stride=2, out_py=py, out_px=px, vec=bias, out2=sq, c_out=cout, c_in=cin)
    if last:
        assert rel_err(out.cpu(), ref.clamp(0, 1)) < 2e-3
    else:
        got = out.float().permute(0, 3, 1, 2).cpu()
        assert rel_err(got, ref) < 1.5e-3, rel_err(got, ref)
        got_sq = sq.float().permute(0, 3, 1, 2).cpu() * 256.0
        assert rel_err(got_sq, ref * ref) < 3e-3


@pytest.mark.parametrize('C,H,W', [(192, 32, 32), (128, 9, 11), (64, 5, 5)])
def test_tc_inverse_gdn_on_square_pair(s2, C, H, W):
    """mode 7: y = x * sqrt(beta + gamma . x^2) with the gamma GEMM on the x^2 / 256 tensor (compressai.layers.GDN, inverse)."""
    dev = torch.device('cuda:0')
    torch.manual_seed(C + H)
    x = (torch.randn(2, C, H, W) * 3).half().float()
    gamma = (0.1 * torch.eye(C) + 0.02 * torch.rand(C, C)).half().float()
    beta = 0.5 + torch.rand(C)
    ref = (x.double() * torch.sqrt(F.conv2d(x.double() ** 2, gamma.double().view(C, C, 1, 1), beta.double()))).float()
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().half().to(dev)
    sq = (x.permute(0, 2, 3, 1) ** 2 / 256.0).contiguous().half().to(dev)
    out = torch.empty_like(x_nhwc)
    s2.ops.tc_conv_ex(sq, gamma.half().to(dev).view(1, C, C).contiguous(), 1, 1, 0, 0, s2._native.TC_IGDN_SQ_F16, (H, W), out,
                      vec=beta.to(dev), gdn_x=x_nhwc, c_out=C, c_in=C)
    got = out.float().permute(0, 3, 1, 2).cpu()
    assert rel_err(got, ref) < 1.5e-3, rel_err(got, ref)


# ---- sc2_tc_split_conv_ex: N tiles, NHWC stride-2 input, bias / activation, GDN proper (zoo g_a / h_a) ----------------------
@pytest.mark.parametrize('cin,cout,k,stride,pad,H,W,act', [(192, 192, 5, 2, 2, 32, 32, 0), (192, 320, 5, 2, 2, 16, 24, 0), (320, 192, 3, 1, 1, 16, 16, 1),
                                                           (32, 200, 5, 2, 2, 10, 14, 2), (64, 48, 3, 2, 1, 12, 20, 0), (16, 144, 1, 1, 0, 7, 9, 1)])
def test_tc_split_conv_ex_tiles_bias_activation(s2, cin, cout, k, stride, pad, H, W, act):
    dev = torch.device('cuda:0')
    torch.manual_seed(cin + cout + H + act)
    x = torch.randn(2, cin, H, W) * 2
    w = torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5
    b = torch.randn(cout)
    ref = F.conv2d(x.double(), w.double(), b.double(), stride, pad)
    if act == 1:
        ref = torch.relu(ref)
    elif act == 2:
        ref = F.leaky_relu(ref, 0.2)
    xh, xl = _planes(s2, x, dev)  # plain NHWC, also for the stride-2 layers
    tiles = s2.ops.pack_conv_weight_split_tiles(w.to(dev))
    assert sum(t[1] for t in tiles) == cout and all(t[1] <= 128 for t in tiles)
    oh, ol = s2.ops.tc_split_conv_tiled(xh, xl, tiles, k, k, stride, pad, s2._native.TCS_STORE, vec=b.to(dev), act=act, slope=0.2,
                                        in_nhwc=stride == 2)
    assert oh.shape[-1] == (cout + 7) // 8 * 8
    got = _unsplit(oh, ol)[:, :cout]
    assert got.shape == ref.shape
    assert rel_err(got, ref.float()) < SPLIT_TOL, rel_err(got, ref.float())
    if act == 0:
        med = torch.randn(cout)
        sym = s2.ops.tc_split_conv_tiled(xh, xl, tiles, k, k, stride, pad, s2._native.TCS_QUANT, vec=b.to(dev), medians=med.to(dev),
                                         in_nhwc=stride == 2)
        want = torch.round(ref.float() - med.view(1, -1, 1, 1)).int()
        assert sym.shape == want.shape
        assert int((sym.cpu() != want).sum()) <= max(1, want.numel() // 100000)  # ties at fp32 resolution only


@pytest.mark.parametrize('C,H,W,kind', [(192, 32, 32, 'gdn'), (192, 9, 11, 'gdn'), (128, 16, 16, 'gdn'), (320, 8, 8, 'gdn'), (192, 12, 12, 'gdn1'),
                                        (48, 20, 12, 'gdn')])
def test_tc_split_gdn_proper_and_tiled(s2, C, H, W, kind):
    """GDN of the zoo codecs: y = x / sqrt(beta + gamma . x^2), squares formed in shared memory; N-tiled over 192 / 320 channels."""
    dev = torch.device('cuda:0')
    torch.manual_seed(C + H)
    x = torch.randn(2, C, H, W) * 3
    x[0, :, 0, 0] = 1e-4 * torch.randn(C)  # tiny activations: x^2 under fp16's normal range
    gamma = 0.1 * torch.eye(C) + 0.02 * torch.rand(C, C)
    beta = 0.5 + torch.rand(C)
    if kind == 'gdn':
        norm = torch.sqrt(F.conv2d(x.double() ** 2, gamma.double().view(C, C, 1, 1), beta.double()))
    else:
        norm = F.conv2d(x.double().abs(), gamma.double().view(C, C, 1, 1), beta.double())
    ref = (x.double() / norm).float()
    xh, xl = _planes(s2, x, dev)
    tiles = s2.ops.pack_conv_weight_split_tiles(gamma.view(C, C, 1, 1).to(dev))
    mode = s2._native.TCS_GDN if kind == 'gdn' else s2._native.TCS_GDN1
    oh, ol = s2.ops.tc_split_conv_tiled(xh, xl, tiles, 1, 1, 1, 0, mode, vec=beta.to(dev), gdn_x=(xh, xl))
    assert rel_err(_unsplit(oh, ol)[:, :C], ref) < SPLIT_TOL


def test_patchify_nhwc_and_abs_split(s2):
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    x = torch.randn(2, 3, 37, 44)
    hi, lo = s2.ops.patchify_split_nhwc(x.to(dev), 5, 5, 2, 2, 80)
    got = (hi.float() + lo.float() / 2048.0).cpu()
    cols = F.unfold(x, 5, padding=2, stride=2).view(2, 75, 19, 22).permute(0, 2, 3, 1)
    assert got.shape == (2, 19, 22, 80)
    assert float((got[..., :75] - cols).abs().max()) < 1e-6 and float(got[..., 75:].abs().max()) == 0.0
    a = torch.randn(4, 5, 6, 16) * torch.tensor([1.0, 1e-9, 1e-3, 50.0]).view(4, 1, 1, 1)
    h, l = s2.ops.split_f16(a.to(dev))
    ah, al = s2.ops.abs_split(h, l)
    assert torch.equal(ah.float() + al.float() / 2048.0, (h.float() + l.float() / 2048.0).abs())
    assert float(ah.min()) >= 0


@pytest.mark.parametrize('cin,cout,H,W,act', [(192, 192, 4, 4, 1), (192, 192, 8, 12, 1), (64, 200, 5, 7, 2), (16, 24, 9, 3, 0)])
def test_tc_split_deconv5_matches_fp64(s2, cin, cout, H, W, act):
    """ConvTranspose2d(k5, s2, p2, op1) + bias + activation in split precision (h_s of the scale-hyperprior codec): four parity
    sub-convolutions writing interleaved pixels through a 5-D tensor map."""
    dev = torch.device('cuda:0')
    torch.manual_seed(cin + cout + H)
    x = torch.randn(2, cin, H, W) * 2
    w = torch.randn(cin, cout, 5, 5) / (cin * 6.25) ** 0.5
    b = torch.randn(cout)
    ref = F.conv_transpose2d(x.double(), w.double(), b.double(), stride=2, padding=2, output_padding=1)
    if act == 1:
        ref = torch.relu(ref)
    elif act == 2:
        ref = F.leaky_relu(ref, 0.01)
    xh, xl = _planes(s2, x, dev)
    packs = s2.ops.pack_deconv5_weight_split_tiles(w.to(dev))
    pitch = (cout + 7) // 8 * 8
    out = (torch.full((2, 2 * H, 2 * W, pitch), float('nan'), dtype=torch.float16, device=dev),
           torch.full((2, 2 * H, 2 * W, pitch), float('nan'), dtype=torch.float16, device=dev))
    for (py, px), tiles in packs.items():
        (ky, pad_y), (kx, pad_x) = s2.ops.DECONV5_TAPS[py], s2.ops.DECONV5_TAPS[px]
        s2.ops.tc_split_conv_tiled(xh, xl, tiles, len(ky), len(kx), 1, pad_y, s2._native.TCS_STORE, vec=b.to(dev), act=act, slope=0.01,
                                   pad_x=pad_x, out=out, out_parity=(py, px))
    got = _unsplit(*out)
    assert not torch.isnan(got).any()
    assert got[:, :cout].shape == ref.shape
    assert rel_err(got[:, :cout], ref.float()) < SPLIT_TOL, rel_err(got[:, :cout], ref.float())
    assert float(got[:, cout:].abs().max()) == 0.0 if pitch > cout else True


@pytest.mark.parametrize('B,C,H,W,cp', [(3, 24, 55, 55, 64), (2, 64, 9, 13, 64), (1, 3, 5, 5, 64), (2, 24, 8, 8, 32), (2, 320, 4, 4, 320)])
def test_nchw_to_nhwc_f16_layouts(s2, B, C, H, W, cp):
    """fp32 NCHW -> fp16 NHWC with zero-padded channels: the tiled-transpose kernel (c_pad = 64) and the generic one."""
    dev = torch.device('cuda:0')
    torch.manual_seed(B + C + H)
    x = torch.randn(B, C, H, W)
    got = s2.ops.nchw_to_nhwc_f16(x.to(dev), cp).cpu()
    assert got.shape == (B, H, W, cp) and got.dtype == torch.float16
    assert torch.equal(got[..., :C], x.permute(0, 2, 3, 1).half())
    assert float(got[..., C:].abs().max()) == 0.0 if cp > C else True


@pytest.mark.parametrize('key,shape', [('h_s', (16, 7, 7)), ('h_a', (24, 55, 55)), ('h_s_mshp', (16, 5, 9)), ('h_a_mshp', (24, 21, 13))])
def test_split_plan_hyperprior_bottleneck_transforms_match_torch(s2, key, shape):
    """h_a / h_s of the (mean-)scale-hyperprior bottlenecks on the split tensor-core plan: 24 input channels (zero-padded to 32), odd
    sizes in front of stride-2 layers, ConvTranspose2d(k5, s2, p1) with 2H + 1 outputs, LeakyReLU -- against torch fp64 on the CPU."""
    dev = torch.device('cuda:0')
    torch.manual_seed(len(key) + shape[1])
    name = 'MSHPBasedResNetBottleneck' if key.endswith('mshp') else 'SHPBasedResNetBottleneck'
    layer = s2.get_layer(name, num_latent_channels=16, num_bottleneck_channels=24, num_target_channels=256).eval()
    seq = getattr(layer, key[:3])
    x = torch.randn(2, *shape) * 2
    with torch.no_grad():
        ref = seq.double()(x.double()).float()
        seq.float()
    assert s2.models.SplitAnalysisPlan.why_not(seq, shape) is None
    plan = s2.models.SplitAnalysisPlan(seq.to(dev))
    got = plan(x.to(dev)).cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert rel_err(got, ref) < 2 * SPLIT_TOL, rel_err(got, ref)
    if key.startswith('h_a'):
        med = torch.randn(seq[-1].out_channels)
        sym = plan(x.to(dev), medians=med.to(dev), out='symbols').cpu()
        want = torch.round(ref - med.view(1, -1, 1, 1)).int()
        assert int((sym != want).sum()) <= 1
