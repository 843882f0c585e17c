"""GPU operator tests through the C ABI: each stand-alone entry point of include/var_b200.h
against plain fp32 PyTorch on the CPU.  Operands are pre-rounded to tf32 (10-bit mantissa), so
every product is exact in fp32 and the only difference left is the summation order: the
comparison is tight and any layout / indexing / masking mistake is a gross mismatch."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_to_max

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def tf32(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def _pack(lib, w):
    O, I, R, S = w.shape
    kpad = (R * S * I + 31) // 32 * 32
    wd = w.to(DEV).contiguous()
    packed = torch.empty(O, kpad, device=DEV)
    mma = torch.empty(O, kpad, device=DEV)
    assert lib.var_pack_weight(wd.data_ptr(), packed.data_ptr(), mma.data_ptr(), O, I, R, S, kpad, None) == 0
    back = torch.empty_like(wd)
    assert lib.var_unpack_weight(packed.data_ptr(), back.data_ptr(), O, I, R, S, kpad, None) == 0
    assert torch.equal(back, wd)
    return mma, kpad


CONV_CASES = [  # N, H, W, Cin, Cout, R, S, sh, sw, ph, pw
    (3, 24, 24, 32, 32, 3, 3, 2, 2, 1, 1),     # kuka.img.conv2-like
    (5, 12, 12, 64, 64, 3, 3, 2, 2, 1, 1),
    (2, 48, 1, 32, 32, 3, 1, 2, 1, 0, 0),      # kuka.snd.conv2
    (7, 5, 1, 32, 32, 3, 1, 2, 1, 0, 0),       # M = 14 rows only
    (2, 20, 20, 32, 64, 3, 3, 1, 1, 1, 1),     # thor.img.conv3-like
    (1, 12, 12, 64, 128, 3, 3, 1, 1, 1, 1),
    (3, 6, 6, 128, 128, 3, 3, 2, 2, 1, 1),
    (1, 60, 20, 64, 64, 11, 5, 2, 2, 5, 5),    # thor.snd.conv2-like (K = 3520)
    (2, 31, 13, 64, 64, 7, 3, 2, 2, 1, 1),     # thor.snd.conv3-like, odd extents
    (300, 1, 1, 576, 128, 1, 1, 1, 1, 0, 0),   # Linear 576 -> 128
    (41, 1, 1, 1024, 128, 1, 1, 1, 1, 0, 0),
    (130, 1, 1, 128, 64, 1, 1, 1, 1, 0, 0),
    (146, 1, 1, 448, 1536, 1, 1, 1, 1, 0, 0),  # GRU input projection
    (19000, 1, 1, 448, 1536, 1, 1, 1, 1, 0, 0),  # same at training size: persistent tile loop, coalescing epilogue
    (70, 60, 20, 64, 64, 11, 5, 2, 2, 5, 5),   # thor.snd.conv2 with > 296 tiles: persistent im2col path
    (8, 96, 96, 32, 32, 3, 3, 1, 1, 1, 1),     # thor.img.conv2, 576 tiles, odd k-block count (9) in 2-block stages
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_dgrad_wgrad_vs_torch(vb, case):
    lib = vb._lib.lib
    N, H, W, Cin, Cout, R, S, sh, sw, ph, pw = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = tf32(torch.randn(N, Cin, H, W, generator=g)).requires_grad_(True)
    w = tf32(torch.randn(Cout, Cin, R, S, generator=g) / (R * S * Cin) ** 0.5).requires_grad_(True)
    b = torch.randn(Cout, generator=g)
    y_ref = F.relu(F.conv2d(x, w, b, stride=(sh, sw), padding=(ph, pw)))
    dy = tf32(torch.randn(y_ref.shape, generator=g)) * (y_ref > 0)
    y_ref.backward(dy)
    P, Q = y_ref.shape[2], y_ref.shape[3]
    x_nhwc = x.detach().permute(0, 2, 3, 1).contiguous().to(DEV)
    wp, kpad = _pack(lib, w.detach())
    bd = b.to(DEV)
    y = torch.empty(N, P, Q, Cout, device=DEV)
    rc = lib.var_conv2d_fwd(x_nhwc.data_ptr(), 0, None, 1.0, N, H, W, Cin, Cout, R, S, sh, sw, ph, pw,
                            wp.data_ptr(), bd.data_ptr(), y.data_ptr(), 1, 0, None)
    assert rc == 0, vb._lib.last_error()
    assert rel_to_max(y.cpu().permute(0, 3, 1, 2).numpy(), y_ref.detach().numpy()) < 1e-5
    dy_nhwc = dy.permute(0, 2, 3, 1).contiguous().to(DEV)
    # dgrad with the ReLU mask of a (synthetic) previous layer folded in
    mask = (torch.rand(N, H, W, Cin, generator=g) > 0.3).float().to(DEV)
    dx = torch.full((N, H, W, Cin), float("nan"), device=DEV)
    rc = lib.var_conv2d_dgrad(dy_nhwc.data_ptr(), wp.data_ptr(), dx.data_ptr(), mask.data_ptr(), N, H, W, Cin, Cout,
                              R, S, sh, sw, ph, pw, 0, None)
    assert rc == 0, vb._lib.last_error()
    dx_ref = x.grad.permute(0, 2, 3, 1) * mask.cpu()
    assert rel_to_max(dx.cpu().numpy(), dx_ref.numpy()) < 1e-5
    dw = torch.zeros(Cout, kpad, device=DEV)
    db = torch.zeros(Cout, device=DEV)
    rc = lib.var_conv2d_wgrad(x_nhwc.data_ptr(), 0, None, 1.0, dy_nhwc.data_ptr(), dw.data_ptr(), db.data_ptr(), N, H,
                              W, Cin, Cout, R, S, sh, sw, ph, pw, None)
    assert rc == 0, vb._lib.last_error()
    dw_ref = torch.empty(Cout, Cin, R, S, device=DEV)
    assert lib.var_unpack_weight(dw.data_ptr(), dw_ref.data_ptr(), Cout, Cin, R, S, kpad, None) == 0
    assert rel_to_max(dw_ref.cpu().numpy(), w.grad.numpy()) < 2e-5
    assert float(dw[:, R * S * Cin:].abs().max() if kpad > R * S * Cin else 0.0) == 0.0  # K padding stays zero
    assert rel_to_max(db.cpu().numpy(), dy.sum(dim=(0, 2, 3)).numpy()) < 2e-5


FIRST_LAYER_CASES = [  # strided (NCHW / single channel) sources
    ("u8", 3, 96, 96, 3, 32, 3, 3, 2, 2, 1, 1),
    ("u8", 2, 96, 96, 3, 32, 3, 3, 1, 1, 1, 1),
    ("u8", 40, 96, 96, 3, 32, 3, 3, 1, 1, 1, 1),      # 3840 tiles: 13 per resident CTA (cin3_conv.cu)
    ("u8", 3, 64, 64, 3, 32, 3, 3, 1, 1, 1, 1),       # 2 output rows per tile
    ("u8", 5, 66, 64, 3, 32, 3, 3, 2, 2, 1, 1),       # 4 rows per tile, P = 33: ragged last tile
    ("u8", 2, 40, 40, 3, 32, 3, 3, 2, 2, 1, 1),       # narrow frame: direct fp32 kernel
    ("f32", 4, 40, 40, 3, 32, 3, 3, 2, 2, 1, 1),
    ("f32", 5, 100, 40, 1, 32, 5, 40, 2, 1, 0, 0),
    ("f32", 2, 120, 40, 1, 64, 11, 11, 2, 2, 5, 5),   # thor.snd.conv1: persistent kernels (cin1_conv.cu)
    ("f32", 3, 600, 40, 1, 64, 11, 11, 2, 2, 5, 5),   # full-height map, 150 tiles
    ("f32", 2, 50, 40, 1, 64, 11, 11, 2, 2, 5, 5),    # P = 25: last tile of an image has one row
    ("f32", 1, 7, 40, 1, 64, 11, 11, 2, 2, 5, 5),     # shorter than the filter
    ("f32raw", 90, 120, 40, 1, 64, 11, 11, 2, 2, 5, 5),  # 900 tiles: >= 6 per resident CTA; unrounded input
    ("f32", 2, 120, 24, 1, 64, 11, 11, 2, 2, 5, 5),   # other width: generic first-layer path
]


@pytest.mark.parametrize("case", FIRST_LAYER_CASES)
def test_first_layer_conv_vs_torch(vb, case):
    lib = vb._lib.lib
    kind, N, H, W, Cin, Cout, R, S, sh, sw, ph, pw = case
    g = torch.Generator().manual_seed(7)
    # the 3 -> 32 channel 3x3 conv1 runs as a direct fp32 kernel (exact input values); the other
    # first layers go through the tensor path, which rounds the loaded value to tf32
    direct = Cin == 3 and Cout == 32 and R == 3 and S == 3
    rnd = (lambda t: t) if direct else tf32
    if kind == "u8":
        xu = torch.randint(0, 256, (N, Cin, H, W), generator=g, dtype=torch.uint8)
        x = rnd(xu.float() * np.float32(1.0 / 255.0))
        src, src_kind, scale = xu.to(DEV), 2, 1.0 / 255.0
    elif kind == "f32raw":  # the kernel rounds the loaded value itself
        xr = torch.randn(N, Cin, H, W, generator=g)
        x = rnd(xr)
        src, src_kind, scale = xr.to(DEV), 1, 1.0
    else:
        x = rnd(torch.randn(N, Cin, H, W, generator=g))
        src, src_kind, scale = x.to(DEV), 1, 1.0
    w = tf32(torch.randn(Cout, Cin, R, S, generator=g) / (R * S * Cin) ** 0.5).requires_grad_(True)
    b = torch.randn(Cout, generator=g)
    y_ref = F.relu(F.conv2d(x, w, b, stride=(sh, sw), padding=(ph, pw)))
    dy = tf32(torch.randn(y_ref.shape, generator=g))
    y_ref.backward(dy)
    P, Q = y_ref.shape[2], y_ref.shape[3]
    import ctypes as C
    strides = (C.c_int64 * 4)(Cin * H * W, W, 1, H * W)  # N, H, W, C element strides of NCHW
    wp, kpad = _pack(lib, w.detach())
    y = torch.empty(N, P, Q, Cout, device=DEV)
    bd = b.to(DEV)
    rc = lib.var_conv2d_fwd(src.data_ptr(), src_kind, strides, scale, N, H, W, Cin, Cout, R, S, sh, sw, ph, pw,
                            wp.data_ptr(), bd.data_ptr(), y.data_ptr(), 1, 0, None)
    assert rc == 0, vb._lib.last_error()
    assert rel_to_max(y.cpu().permute(0, 3, 1, 2).numpy(), y_ref.detach().numpy()) < 1e-5
    dy_nhwc = (dy * (y_ref > 0)).permute(0, 2, 3, 1).contiguous().to(DEV)
    dw = torch.zeros(Cout, kpad, device=DEV)
    db = torch.zeros(Cout, device=DEV)
    rc = lib.var_conv2d_wgrad(src.data_ptr(), src_kind, strides, scale, dy_nhwc.data_ptr(), dw.data_ptr(),
                              db.data_ptr(), N, H, W, Cin, Cout, R, S, sh, sw, ph, pw, None)
    assert rc == 0, vb._lib.last_error()
    dw_ref = torch.empty(Cout, Cin, R, S, device=DEV)
    assert lib.var_unpack_weight(dw.data_ptr(), dw_ref.data_ptr(), Cout, Cin, R, S, kpad, None) == 0
    assert rel_to_max(dw_ref.cpu().numpy(), w.grad.numpy()) < 2e-5
    assert rel_to_max(db.cpu().numpy(), (dy * (y_ref > 0)).sum(dim=(0, 2, 3)).numpy()) < 2e-5


@pytest.mark.parametrize("shape", [(3, 96, 96, 32), (2, 12, 12, 128), (5, 6, 10, 64)])
def test_maxpool_fwd_bwd_vs_torch(vb, shape):
    lib = vb._lib.lib
    N, H, W, Cc = shape
    g = torch.Generator().manual_seed(3)
    x = F.relu(torch.randn(N, Cc, H, W, generator=g)).requires_grad_(True)  # post-ReLU: many exact zeros
    y_ref = F.max_pool2d(x, 2, 2)
    dy = torch.randn(y_ref.shape, generator=g)
    y_ref.backward(dy)
    xd = x.detach().permute(0, 2, 3, 1).contiguous().to(DEV)
    y = torch.empty(N, H // 2, W // 2, Cc, device=DEV)
    assert lib.var_maxpool2x2_fwd(xd.data_ptr(), y.data_ptr(), N, H, W, Cc, None) == 0
    assert torch.equal(y.cpu().permute(0, 3, 1, 2), y_ref.detach())
    dyd = dy.permute(0, 2, 3, 1).contiguous().to(DEV)
    dx = torch.full((N, H, W, Cc), float("nan"), device=DEV)
    assert lib.var_maxpool2x2_bwd(xd.data_ptr(), dyd.data_ptr(), dx.data_ptr(), N, H, W, Cc, None) == 0
    # reference: pool backward followed by the ReLU backward of the layer that produced x
    ref = (x.grad * (x.detach() > 0)).permute(0, 2, 3, 1)
    assert torch.equal(dx.cpu(), ref)


@pytest.mark.parametrize("B,Ki,Ks", [(1, 128, 128), (37, 128, 64), (1000, 128, 128)])
def test_fused_triplet_kernel_vs_torch(vb, B, Ki, Ks):
    lib = vb._lib.lib
    g = torch.Generator().manual_seed(B)
    D = 3
    mk = lambda *s: torch.randn(*s, generator=g)
    h_img, h_pos, h_neg = F.relu(mk(B, Ki)), F.relu(mk(B, Ks)), F.relu(mk(B, Ks))
    W_i, b_i, W_s, b_s = mk(D, Ki) * 0.1, mk(D) * 0.1, mk(D, Ks) * 0.1, mk(D) * 0.1
    leaves = [t.clone().requires_grad_(True) for t in (h_img, h_pos, h_neg, W_i, b_i, W_s, b_s)]
    hi, hp, hn, Wi, bi, Ws, bs = leaves
    fa = F.normalize(F.linear(hi, Wi, bi), p=2, dim=1)
    fp = F.normalize(F.linear(hp, Ws, bs), p=2, dim=1)
    fn = F.normalize(F.linear(hn, Ws, bs), p=2, dim=1)
    crit = torch.nn.TripletMarginLoss(margin=1.0, p=2, reduction="none")
    rows = crit(fa, fp, fn)
    loss = rows.mean()
    loss.backward()
    d = lambda t: t.to(DEV).contiguous()
    dv = [d(t) for t in (h_img, h_pos, h_neg, W_i, b_i, W_s, b_s)]
    feats = torch.empty(3, B, D, device=DEV)
    out_loss = torch.zeros((), device=DEV)
    loss_rows = torch.empty(B, device=DEV)
    dh = [torch.empty(B, Ki, device=DEV), torch.empty(B, Ks, device=DEV), torch.empty(B, Ks, device=DEV)]
    dW_i, db_i = torch.zeros(D, Ki, device=DEV), torch.zeros(4, device=DEV)
    dW_s, db_s = torch.zeros(D, Ks, device=DEV), torch.zeros(4, device=DEV)
    rc = lib.var_triplet_fwd_bwd(dv[0].data_ptr(), dv[1].data_ptr(), dv[2].data_ptr(), B, D, Ki, Ks, dv[3].data_ptr(),
                                 dv[4].data_ptr(), dv[5].data_ptr(), dv[6].data_ptr(), 1.0, float(B), feats.data_ptr(),
                                 out_loss.data_ptr(), loss_rows.data_ptr(), dh[0].data_ptr(), dh[1].data_ptr(),
                                 dh[2].data_ptr(), dW_i.data_ptr(), db_i.data_ptr(), dW_s.data_ptr(), db_s.data_ptr(),
                                 None)
    assert rc == 0, vb._lib.last_error()
    assert abs(float(out_loss) - float(loss)) < 2e-6 * max(1.0, float(loss))
    assert np.abs(loss_rows.cpu().numpy() - rows.detach().numpy()).max() < 1e-5
    for i, f in enumerate((fa, fp, fn)):
        assert np.abs(feats[i].cpu().numpy() - f.detach().numpy()).max() < 1e-5
    # grads of the head inputs are stored ReLU-masked and tf32-rounded (they feed the next MMAs)
    for got, leaf, hsrc in zip(dh, (hi, hp, hn), (h_img, h_pos, h_neg)):
        ref = leaf.grad * (hsrc > 0)
        assert rel_to_max(got.cpu().numpy(), ref.numpy()) < 1e-3
    assert rel_to_max(dW_i.cpu().numpy(), Wi.grad.numpy()) < 1e-4
    assert rel_to_max(dW_s.cpu().numpy(), Ws.grad.numpy()) < 1e-4
    assert rel_to_max(db_i[:D].cpu().numpy(), bi.grad.numpy()) < 1e-4
    assert rel_to_max(db_s[:D].cpu().numpy(), bs.grad.numpy()) < 1e-4


H16_CASES = [  # N, H, W, Cin, Cout, R, S, sh, sw, ph, pw   (Cin, Cout multiples of 64)
    (1, 60, 20, 64, 64, 11, 5, 2, 2, 5, 5),    # thor.snd.conv2 geometry (K = 3520: free 64-row group -> fused bias grad)
    (70, 60, 20, 64, 64, 11, 5, 2, 2, 5, 5),   # same, many tiles per resident CTA
    (2, 31, 13, 64, 64, 7, 3, 2, 2, 1, 1),     # thor.snd.conv3 geometry, odd extents (K = 1344)
    (5, 12, 12, 64, 64, 3, 3, 2, 2, 1, 1),     # K = 576 = 4.5 k tiles
    (3, 12, 12, 64, 128, 3, 3, 1, 1, 1, 1),    # thor.img.conv5-like, stride 1, N = 128
    (3, 6, 6, 128, 128, 3, 3, 2, 2, 1, 1),     # two 64-channel chunks per tap; K = 1152 = 9 tiles (separate column sum)
]


# launch variants of the stride-2 64 -> 64 convs (csrc/halo_conv.cuh): the defaults (im2col-free forward / data gradient, im2col
# weight gradient), the im2col kernels alone, and the im2col-free weight gradient with other launch shapes
H16_VARIANTS = {
    "default": {},
    "im2col": {"VAR_HALO": "0"},
    "halo_alt": {"VAR_HALO_WGRAD": "1", "VAR_HALO_CPS": "2", "VAR_HALO_DG_CPS": "2"},
}


@pytest.mark.parametrize("variant", list(H16_VARIANTS))
@pytest.mark.parametrize("case", H16_CASES)
def test_conv_h16_fwd_dgrad_wgrad_vs_torch(vb, case, variant, monkeypatch):
    """The 16-bit operand region: f16 activations / weights / scaled f16 gradients through kind::f16 MMAs.
    Operands are pre-rounded to f16 (products exact in fp32), so only the summation order differs."""
    if variant != "default" and not (case[3] == 64 and case[4] == 64 and case[7] == 2):
        pytest.skip("launch variants only exist for the stride-2 64 -> 64 convs")
    for k, v in H16_VARIANTS[variant].items():
        monkeypatch.setenv(k, v)
    lib = vb._lib.lib
    N, H, W, Cin, Cout, R, S, sh, sw, ph, pw = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    h16 = lambda t: t.half().float()
    x = h16(torch.randn(N, Cin, H, W, generator=g)).requires_grad_(True)
    w = h16(torch.randn(Cout, Cin, R, S, generator=g) / (R * S * Cin) ** 0.5).requires_grad_(True)
    b = torch.randn(Cout, generator=g)
    y_ref = F.relu(F.conv2d(x, w, b, stride=(sh, sw), padding=(ph, pw)))
    P, Q = y_ref.shape[2], y_ref.shape[3]
    # incoming gradient: small magnitudes (like 1/B-scaled losses) so the dynamic scale matters
    dy32 = torch.randn(y_ref.shape, generator=g) * 3e-6 * (y_ref > 0)
    x_h = x.detach().permute(0, 2, 3, 1).contiguous().to(DEV).half()
    wp32, kpad = _pack(lib, w.detach())
    w_h = torch.empty(Cout, kpad, dtype=torch.half, device=DEV)
    assert lib.var_cvt_f16(wp32.data_ptr(), w_h.data_ptr(), Cout * kpad, None) == 0
    bd = b.to(DEV)
    # forward, f16 output and fp32 (tf32-rounded) output
    y_h = torch.empty(N, P, Q, Cout, dtype=torch.half, device=DEV)
    rc = lib.var_conv2d_fwd_h16(x_h.data_ptr(), N, H, W, Cin, Cout, R, S, sh, sw, ph, pw, w_h.data_ptr(), bd.data_ptr(),
                                y_h.data_ptr(), 1, 1, 0, None)
    assert rc == 0, vb._lib.last_error()
    y32 = torch.empty(N, P, Q, Cout, device=DEV)
    rc = lib.var_conv2d_fwd_h16(x_h.data_ptr(), N, H, W, Cin, Cout, R, S, sh, sw, ph, pw, w_h.data_ptr(), bd.data_ptr(),
                                y32.data_ptr(), 0, 1, 0, None)
    assert rc == 0, vb._lib.last_error()
    ref_nhwc = y_ref.detach().permute(0, 2, 3, 1)
    assert rel_to_max(y32.cpu().numpy(), ref_nhwc.numpy()) < 1e-5
    assert rel_to_max(y_h.float().cpu().numpy(), ref_nhwc.numpy()) < 1e-3   # f16 storage: 2^-11 relative
    # gradient entering the region: device-chosen power-of-two scale
    dy_nhwc = dy32.permute(0, 2, 3, 1).contiguous().to(DEV)
    dy_h = torch.empty(N, P, Q, Cout, dtype=torch.half, device=DEV)
    scale = torch.zeros(2, device=DEV)
    amax = torch.zeros(1, dtype=torch.int32, device=DEV)
    assert lib.var_grad_to_f16_scaled(dy_nhwc.data_ptr(), dy_h.data_ptr(), dy_nhwc.numel(), scale.data_ptr(),
                                      amax.data_ptr(), None) == 0
    S_, inv = scale.cpu().tolist()
    assert S_ * inv == 1.0 and np.log2(S_) == round(np.log2(S_))
    assert 2048 <= float(dy_nhwc.abs().max()) * S_ <= 4096
    dy_exact = (dy_h.float() / S_).cpu()                       # what the kernels actually multiply
    y_ref.backward(dy_exact.permute(0, 3, 1, 2))
    # dgrad: f16 mask (previous activation), fp32 output unscaled; and f16 output (still scaled)
    mask = (torch.rand(N, H, W, Cin, generator=g) > 0.3).to(DEV)
    mask_h = (mask.half() * 0.37)
    dx = torch.full((N, H, W, Cin), float("nan"), device=DEV)
    rc = lib.var_conv2d_dgrad_h16(dy_h.data_ptr(), w_h.data_ptr(), dx.data_ptr(), 0, mask_h.data_ptr(), 1,
                                  scale[1:].data_ptr(), N, H, W, Cin, Cout, R, S, sh, sw, ph, pw, 0, None)
    assert rc == 0, vb._lib.last_error()
    dx_ref = x.grad.permute(0, 2, 3, 1) * mask.cpu()
    assert rel_to_max(dx.cpu().numpy(), dx_ref.numpy()) < 1e-5
    dx_h = torch.zeros(N, H, W, Cin, dtype=torch.half, device=DEV)
    rc = lib.var_conv2d_dgrad_h16(dy_h.data_ptr(), w_h.data_ptr(), dx_h.data_ptr(), 1, mask_h.data_ptr(), 1, None, N, H,
                                  W, Cin, Cout, R, S, sh, sw, ph, pw, 0, None)
    assert rc == 0, vb._lib.last_error()
    assert rel_to_max((dx_h.float() / S_).cpu().numpy(), dx_ref.numpy()) < 1e-3
    # wgrad (+ bias gradient), unscaled by inv
    dw = torch.zeros(Cout, kpad, device=DEV)
    db = torch.zeros(Cout, device=DEV)
    rc = lib.var_conv2d_wgrad_h16(x_h.data_ptr(), dy_h.data_ptr(), dw.data_ptr(), db.data_ptr(), scale[1:].data_ptr(), N,
                                  H, W, Cin, Cout, R, S, sh, sw, ph, pw, None)
    assert rc == 0, vb._lib.last_error()
    dw_ref = torch.empty(Cout, Cin, R, S, device=DEV)
    assert lib.var_unpack_weight(dw.data_ptr(), dw_ref.data_ptr(), Cout, Cin, R, S, kpad, None) == 0
    assert rel_to_max(dw_ref.cpu().numpy(), w.grad.numpy()) < 2e-5
    assert rel_to_max(db.cpu().numpy(), dy_exact.sum(dim=(0, 1, 2)).numpy()) < 2e-5


LIN16_CASES = [  # M, K, N
    (146, 448, 3072),    # GRU input projection of two sounds: ragged last M tile, 12 column tiles, TMA-store epilogue
    (1000, 3072, 448),   # its dX GEMM: K = 3072 (48 k-blocks), two column tiles of 224, scale + ReLU mask in the epilogue
    (300, 512, 1536),    # W_hh-shaped
]


@pytest.mark.parametrize("case", LIN16_CASES)
def test_linear_h16_fwd_and_wgrad_vs_torch(vb, case):
    """Plain f16 GEMMs of the GRU input projection (forward with bias, backward-data with 1/S and a ReLU mask, weight
    gradient over slabs of 256 output channels).  Operands pre-rounded to f16: only the summation order differs."""
    lib = vb._lib.lib
    M, K, N = case
    g = torch.Generator().manual_seed(M + K + N)
    a = (torch.randn(M, K, generator=g) * 0.5).half()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).half()
    bias = torch.randn(N, generator=g)
    a_d, w_d, b_d = a.to(DEV), w.to(DEV), bias.to(DEV)
    out = torch.full((M, N), float("nan"), device=DEV)
    rc = lib.var_linear_h16(a_d.data_ptr(), K, w_d.data_ptr(), b_d.data_ptr(), out.data_ptr(), N, None, 0, 0, None, M, K, N,
                            0, None)
    assert rc == 0, vb._lib.last_error()
    ref = a.double() @ w.double().t() + bias.double()
    assert rel_to_max(out.cpu().numpy(), ref.float().numpy()) < 1e-5
    # scaled + masked variant (what the dX GEMM uses)
    scale = torch.tensor([0.125], device=DEV)
    mask = (torch.rand(M, N, generator=g) > 0.4).float().to(DEV) * 0.7
    out2 = torch.full((M, N), float("nan"), device=DEV)
    rc = lib.var_linear_h16(a_d.data_ptr(), K, w_d.data_ptr(), None, out2.data_ptr(), N, mask.data_ptr(), 0, N,
                            scale.data_ptr(), M, K, N, 0, None)
    assert rc == 0, vb._lib.last_error()
    ref2 = (a.double() @ w.double().t()) * 0.125 * (mask.cpu().double() > 0)
    assert rel_to_max(out2.cpu().numpy(), ref2.float().numpy()) < 1e-5
    # weight gradient: dw[N][K] += inv * dY^T A with dY = a second f16 matrix [M, N] read at a row pitch of 2 N
    # (output channels in slabs of 256: N must be <= 256 or a multiple of 256)
    if N > 256 and N % 256:
        return
    dy_wide = (torch.randn(M, 2 * N, generator=g) * 0.25).half().to(DEV)
    dw = torch.zeros(N, K, device=DEV)
    rc = lib.var_linear_wgrad_h16(a_d.data_ptr(), K, dy_wide[:, N:].data_ptr(), 2 * N, dw.data_ptr(), K, scale.data_ptr(), M,
                                  K, N, None)
    assert rc == 0, vb._lib.last_error()
    dw_ref = 0.125 * (dy_wide[:, N:].cpu().double().t() @ a.double())
    assert rel_to_max(dw.cpu().numpy(), dw_ref.float().numpy()) < 1e-5
