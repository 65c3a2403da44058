"""Kernel-level parity through the C-ABI: each entry point of include/icf.h against the torch primitive the
reference dispatches to (computed in float64 on the CPU).  Tolerances: fp32 path 1e-3 relative (north_star),
in practice ~1e-6; bf16 path 2e-2; integer work (argmax, masks) bit-exact."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def ops_mod():
    from icf_b200 import ops
    return ops


def nhwc(x, pitch, dtype):
    """(N,C,H,W) cpu -> [N*H*W, pitch] cuda, zero padded channels."""
    n, c, h, w = x.shape
    t = torch.zeros(n * h * w, pitch, dtype=dtype, device=DEV)
    t[:, :c] = x.permute(0, 2, 3, 1).reshape(-1, c).to(DEV).to(dtype)
    return t


def from_nhwc(t, n, h, w, c):
    return t[:, :c].float().cpu().reshape(n, h, w, c).permute(0, 3, 1, 2)


def pack_w(ops, w_kcrs, dtype_code, pitch):
    """reference Conv2d weight [K,C,R,S] -> packed [K][R*S][pitch] through icf_pack."""
    K, C, R, S = w_kcrs.shape
    T = R * S
    src = w_kcrs.contiguous().to(DEV)
    dst = torch.empty(K * T * pitch, dtype=ops.torch_dtype(dtype_code), device=DEV)
    ops.pack(src.data_ptr(), dst.data_ptr(), dtype_code, ops.make_perm(K, T, C, C * T, 1, T, d2_pad=pitch))
    return dst


CONV_CASES = [  # N, C, H, K, k, stride, pad
    (3, 5, 28, 64, 3, 2, 1), (2, 64, 14, 128, 4, 2, 1), (2, 128, 7, 256, 4, 2, 1), (5, 256, 3, 512, 4, 2, 1),
    (7, 512, 1, 512, 1, 2, 0), (2, 5, 28, 32, 5, 1, 0), (2, 32, 24, 64, 4, 2, 0), (3, 7, 32, 64, 5, 2, 1),
    (130, 64, 3, 1, 1, 1, 0), (2, 1024, 1, 1024, 1, 1, 0),
    # several image blocks / columns per persistent CTA of the row-streaming kernel
    (300, 5, 28, 32, 5, 1, 0), (200, 32, 24, 64, 4, 2, 0), (150, 64, 11, 128, 4, 1, 0), (260, 64, 14, 40, 4, 2, 1),
    (140, 24, 12, 16, 3, 1, 1),
]


@pytest.mark.parametrize("code", [0, 1], ids=["fp32", "bf16"])
@pytest.mark.parametrize("case", CONV_CASES, ids=[str(c) for c in CONV_CASES])
def test_conv_gather(case, code):
    ops = ops_mod()
    N, C, H, K, k, s, p = case
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, C, H, H, generator=g)
    w = torch.randn(K, C, k, k, generator=g) / (C * k * k) ** 0.5
    b = torch.randn(K, generator=g)
    dt = ops.torch_dtype(code)
    if code == 1:
        x, w = x.bfloat16().float(), w.bfloat16().float()
    ref = F.leaky_relu(F.conv2d(x.double(), w.double(), b.double(), s, p), 0.2)
    P = ref.shape[-1]
    cp, kp = (C + 7) // 8 * 8, (K + 7) // 8 * 8 if K > 1 else 1
    xt = nhwc(x, cp, dt)
    wt = pack_w(ops, w, code, cp)
    y = torch.empty(N * P * P, kp, dtype=dt, device=DEV)
    bias = b.to(DEV)
    ops.conv_forward(code, ops.GATHER, N, H, H, C, cp, P, P, K, kp, k, k, s, p, xt.data_ptr(), wt.data_ptr(), K, cp,
                     y.data_ptr(), bias=bias.data_ptr(), act="lrelu", slope=0.2)
    torch.cuda.synchronize()
    err = rel_err(from_nhwc(y, N, P, P, K), ref)
    assert err < (1e-5 if code == 0 else 8e-3), err


CONVT_CASES = [  # N, C, H, K, k, stride, pad, opad
    (3, 771, 1, 512, 3, 1, 0, 0), (2, 512, 3, 256, 3, 2, 0, 0), (2, 256, 7, 128, 3, 2, 1, 0),
    (2, 128, 13, 64, 3, 2, 1, 0), (3, 64, 25, 1, 4, 1, 0, 0), (2, 64, 4, 32, 5, 2, 2, 1), (2, 64, 16, 1, 5, 2, 2, 1),
    (260, 64, 25, 1, 4, 1, 0, 0), (200, 32, 24, 5, 5, 1, 0, 0), (150, 128, 13, 64, 3, 2, 1, 0), (130, 64, 11, 32, 4, 2, 0, 0),
]


@pytest.mark.parametrize("code", [0, 1], ids=["fp32", "bf16"])
@pytest.mark.parametrize("case", CONVT_CASES, ids=[str(c) for c in CONVT_CASES])
def test_conv_transposed(case, code):
    ops = ops_mod()
    N, C, H, K, k, s, p, op = case
    g = torch.Generator().manual_seed(2)
    x = torch.randn(N, C, H, H, generator=g)
    w = torch.randn(C, K, k, k, generator=g) / (C * k * k / s / s) ** 0.5      # ConvTranspose2d layout
    b = torch.randn(K, generator=g)
    dt = ops.torch_dtype(code)
    if code == 1:
        x, w = x.bfloat16().float(), w.bfloat16().float()
    ref = torch.tanh(F.conv_transpose2d(x.double(), w.double(), b.double(), s, p, op))
    P = ref.shape[-1]
    cp, kp = (C + 7) // 8 * 8, (K + 7) // 8 * 8 if K > 1 else 1
    T = k * k
    wt = torch.empty(K * T * cp, dtype=dt, device=DEV)
    wsrc = w.contiguous().to(DEV)
    ops.pack(wsrc.data_ptr(), wt.data_ptr(), code, ops.make_perm(K, T, C, T, 1, K * T, d2_pad=cp))
    xt = nhwc(x, cp, dt)
    y = torch.empty(N * P * P, kp, dtype=dt, device=DEV)
    bias = b.to(DEV)
    ops.conv_forward(code, ops.TRANSPOSED, N, H, H, C, cp, P, P, K, kp, k, k, s, p, xt.data_ptr(), wt.data_ptr(), K,
                     cp, y.data_ptr(), bias=bias.data_ptr(), act="tanh")
    torch.cuda.synchronize()
    err = rel_err(from_nhwc(y, N, P, P, K), ref)
    assert err < (1e-5 if code == 0 else 8e-3), err


WGRAD_CASES = [(3, 5, 28, 64, 3, 2, 1), (2, 64, 14, 128, 4, 2, 1), (9, 256, 3, 512, 4, 2, 1), (4, 32, 24, 64, 4, 2, 0),
               (33, 512, 1, 512, 1, 1, 0), (2, 7, 32, 64, 5, 2, 1),
               # single-channel gradient operand (Cout = 1 generator tail): streaming-reduction kernel
               (6, 1, 28, 64, 4, 1, 0), (3, 1, 20, 40, 5, 2, 2), (2, 1, 12, 130, 3, 1, 1)]


@pytest.mark.parametrize("code", [0, 1], ids=["fp32", "bf16"])
@pytest.mark.parametrize("case", WGRAD_CASES, ids=[str(c) for c in WGRAD_CASES])
def test_conv_wgrad(case, code):
    ops = ops_mod()
    N, C, H, K, k, s, p = case
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, C, H, H, generator=g)
    P = (H + 2 * p - k) // s + 1
    dy = torch.randn(N, K, P, P, generator=g)
    dt = ops.torch_dtype(code)
    if code == 1:
        x, dy = x.bfloat16().float(), dy.bfloat16().float()
    w = torch.zeros(K, C, k, k, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), w, None, s, p).backward(dy.double())
    cp, kp = (C + 7) // 8 * 8, (K + 7) // 8 * 8
    xt, dyt = nhwc(x, cp, dt), nhwc(dy, kp, dt)
    T = k * k
    dwp = torch.zeros(K * T * C, dtype=torch.float32, device=DEV)
    ops.conv_wgrad(code, N, P, P, K, kp, H, H, C, cp, k, k, s, p, dyt.data_ptr(), xt.data_ptr(), dwp.data_ptr())
    dw = torch.empty(K, C, k, k, dtype=torch.float32, device=DEV)
    ops.unpack(dwp.data_ptr(), dw.data_ptr(), ops.make_perm(K, T, C, C * T, 1, T))
    torch.cuda.synchronize()
    err = rel_err(dw, w.grad)
    assert err < (1e-5 if code == 0 else 1e-4), err      # inputs pre-rounded to bf16, fp32 accumulation


@pytest.mark.parametrize("case", [(6, 5, 28, 32, 5), (150, 5, 28, 32, 5), (3, 7, 20, 64, 3), (2, 2, 33, 16, 4)],
                         ids=lambda c: str(c))
def test_conv_wgrad_first_layer_folded(case):
    """Unit-stride first-layer weight gradient on the 8-channel-pitch feature tensor in the folded ("win") packing
    dw[k][r][s*8 + c] (the form engine.LayerExec uses for D.dx.1): overlapping-window descriptor kernel."""
    ops = ops_mod()
    N, C, H, K, k = case
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, C, H, H, generator=g).bfloat16().float()
    P = H - k + 1
    dy = torch.randn(N, K, P, P, generator=g).bfloat16().float()
    w = torch.zeros(K, C, k, k, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), w, None, 1, 0).backward(dy.double())
    kp = (K + 7) // 8 * 8
    xt, dyt = nhwc(x, 8, torch.bfloat16), nhwc(dy, kp, torch.bfloat16)
    xt = torch.cat([xt, torch.zeros(H, 8, dtype=torch.bfloat16, device=DEV)])      # slack rows the folded reads run into
    dwp = torch.zeros(K * k * k * 8, dtype=torch.float32, device=DEV)
    ops.conv_wgrad(1, N, P, P, K, kp, H, H, k * 8, 8, k, 1, 1, 0, dyt.data_ptr(), xt.data_ptr(), dwp.data_ptr(), win=k)
    torch.cuda.synchronize()
    got = dwp.view(K, k, k, 8)[..., :C].permute(0, 3, 1, 2).cpu().double()        # [K][r][s][c] -> [K][c][r][s]
    err = rel_err(got, w.grad)
    assert err < 1e-4, err


@pytest.mark.parametrize("case", [(6, 5, 28, 32, 5), (300, 5, 28, 32, 5), (1200, 5, 28, 32, 5), (37, 7, 20, 64, 3), (33, 3, 24, 32, 4),
                                  (10, 5, 30, 128, 5), (300, 5, 30, 64, 3, 2), (45, 3, 22, 64, 5, 2), (9, 7, 130, 64, 5, 2), (3, 2, 258, 64, 5, 2),
                                  (5, 4, 150, 32, 3, 1), (20, 3, 16, 16, 2, 1)], ids=lambda c: str(c))
def test_conv_forward_first_layer_folded(case):
    """Unit-stride first-layer forward on the 8-channel-pitch feature tensor in the folded ("win") packing (the form
    engine.LayerExec uses for D.dx.1) with bias, LeakyReLU, Dropout2d mask and BatchNorm statistics in the epilogue:
    channel-major accumulator kernel (icf_conv_cm.cu; ragged image groups and row blocks included; stride 2 = E.layers.0)."""
    ops = ops_mod()
    N, C, H, K, k = case[:5]
    st = case[5] if len(case) > 5 else 1
    g = torch.Generator().manual_seed(6)
    x = torch.randn(N, C, H, H, generator=g).bfloat16().float()
    w = (torch.randn(K, C, k, k, generator=g) / (C * k * k) ** 0.5).bfloat16().float()
    b = torch.randn(K, generator=g)
    mask = (torch.rand(N, K, generator=g) > 0.2).float() / 0.8
    P = (H - k) // st + 1
    ref = F.leaky_relu(F.conv2d(x.double(), w.double(), b.double(), st, 0), 0.1) * mask.double().reshape(N, K, 1, 1)
    xt = nhwc(x, 8, torch.bfloat16)
    wf = torch.zeros(K * k * 64, dtype=torch.bfloat16, device=DEV)
    wsrc = w.contiguous().to(DEV)
    ops.pack4(wsrc.data_ptr(), wf.data_ptr(), 1, ops.make_perm4(K, k, k, C, C * k * k, k, 1, k * k, 8, 64))
    kp = (K + 7) // 8 * 8
    y = torch.full((N * P * P, kp), 7.0, dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros(2, K, dtype=torch.float32, device=DEV)
    md, bias = mask.to(DEV), b.to(DEV)
    ops.conv_forward(1, ops.GATHER, N, H, H, k * 8, 8, P, P, K, kp, k, 1, st, 0, xt.data_ptr(), wf.data_ptr(), K, 64,
                     y.data_ptr(), bias=bias.data_ptr(), act="lrelu", slope=0.1, mask=md.data_ptr(), mask_pitch=K,
                     stats=stats.data_ptr(), win=k)
    torch.cuda.synchronize()
    from icf_b200 import lib
    assert lib.load().icf_last_conv_path() == 4                 # ICF_PATH_CM
    got = from_nhwc(y, N, P, P, K)
    assert rel_err(got, ref) < 8e-3
    gd = got.double()
    assert rel_err(stats[0], gd.sum(dim=(0, 2, 3))) < 1e-4 and rel_err(stats[1], gd.square().sum(dim=(0, 2, 3))) < 1e-4


@pytest.mark.parametrize("case", [(32, 1024, 3, 512, 5, 2, 1, "g"), (32, 512, 7, 1024, 5, 2, 1, "g"), (8, 1024, 1, 1024, 1, 1, 0, "g"),
                                  (8, 1024, 3, 512, 5, 2, 2, "t"), (5, 1024, 3, 200, 3, 1, 0, "g")], ids=lambda c: str(c))
def test_conv_forward_splitk(case):
    """icf_conv_forward_splitk: layers whose tile grid is a handful of CTAs (a few hundred output pixels, a long tap x channel
    reduction: the 512..1024-channel layers of the spectrogram families at batch 32) with bias / activation / mask applied by the
    finishing pass; gather and transposed form, ragged channel count."""
    ops = ops_mod()
    N, C, H, K, k, s, p, form = case
    g = torch.Generator().manual_seed(9)
    x = torch.randn(N, C, H, H, generator=g).bfloat16().float()
    b = torch.randn(K, generator=g)
    mask = (torch.rand(N, K, generator=g) > 0.3).float() / 0.7
    cp, kp = (C + 7) // 8 * 8, (K + 7) // 8 * 8
    T = k * k
    if form == "g":
        w = (torch.randn(K, C, k, k, generator=g) / (C * k * k) ** 0.5).bfloat16().float()
        ref = F.leaky_relu(F.conv2d(x.double(), w.double(), b.double(), s, p), 0.2)
        wt = pack_w(ops, w, 1, cp)
        fcode = ops.GATHER
    else:
        w = (torch.randn(C, K, k, k, generator=g) / (C * k * k / s / s) ** 0.5).bfloat16().float()
        ref = F.leaky_relu(F.conv_transpose2d(x.double(), w.double(), b.double(), s, p, 1), 0.2)
        wt = torch.empty(K * T * cp, dtype=torch.bfloat16, device=DEV)
        wsrc = w.contiguous().to(DEV)
        ops.pack(wsrc.data_ptr(), wt.data_ptr(), 1, ops.make_perm(K, T, C, T, 1, K * T, d2_pad=cp))
        fcode = ops.TRANSPOSED
    ref = ref * mask.double().reshape(N, K, 1, 1)
    P = ref.shape[-1]
    xt = nhwc(x, cp, torch.bfloat16)
    y = torch.full((N * P * P, kp), 7.0, dtype=torch.bfloat16, device=DEV)
    part = torch.zeros((N * P * P, (K + 255) // 256 * 256), dtype=torch.float32, device=DEV)
    md, bias = mask.to(DEV), b.to(DEV)
    ops.conv_forward(1, fcode, N, H, H, C, cp, P, P, K, kp, k, k, s, p, xt.data_ptr(), wt.data_ptr(), K, cp, y.data_ptr(),
                     bias=bias.data_ptr(), act="lrelu", slope=0.2, mask=md.data_ptr(), mask_pitch=K, partial=part)
    torch.cuda.synchronize()
    assert float(part.abs().max()) > 0                           # the split path ran (the scratch is left dirty)
    assert rel_err(from_nhwc(y, N, P, P, K), ref) < 8e-3


def test_argmax_first_max_wins():
    ops = ops_mod()
    x = torch.tensor([[0, 0, 0], [0, 1, 1], [2, 2, 1], [0.5, 0.2, 0.9]], device=DEV)
    for t in (x, x.int(), x.long(), x.bfloat16()):
        assert ops.argmax_rows(t).tolist() == t.cpu().float().argmax(1).tolist()      # CPU argmax: first max wins
    big = torch.nn.functional.one_hot(torch.randint(0, 15, (1000,)), 15).to(DEV)
    assert torch.equal(ops.argmax_rows(big).long().cpu(), big.cpu().argmax(1))


@pytest.mark.parametrize("code", [0, 1], ids=["fp32", "bf16"])
def test_bn_and_act_backward(code):
    """conv-epilogue statistics -> icf_bn_finalize -> icf_scale_shift_mask and the fused backward, against
    F.batch_norm + leaky_relu + dropout-mask autograd."""
    ops = ops_mod()
    N, C, H = 6, 40, 5
    g = torch.Generator().manual_seed(4)
    pre = torch.randn(N, C, H, H, generator=g, dtype=torch.float64, requires_grad=True)
    m1 = (torch.rand(N, C, generator=g) > 0.2).double() / 0.8
    m2 = (torch.rand(N, C, generator=g) > 0.5).double() / 0.5
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).double().requires_grad_()
    beta = (0.1 * torch.randn(C, generator=g)).double().requires_grad_()
    y = F.leaky_relu(pre, 0.1) * m1[:, :, None, None]
    rm, rv = torch.zeros(C, dtype=torch.float64), torch.ones(C, dtype=torch.float64)
    u = F.batch_norm(y, rm, rv, gamma, beta, True, 0.1, 1e-5) * m2[:, :, None, None]
    dU = torch.randn(N, C, H, H, generator=g, dtype=torch.float64)
    u.backward(dU)
    dt = ops.torch_dtype(code)
    cp = 40
    # forward pieces: statistics come from a 1x1 identity "conv" so that the epilogue path is exercised
    eye = torch.eye(C).reshape(C, C, 1, 1)
    wt = pack_w(ops, eye, code, cp)
    pre_q = pre.detach().float()
    if code == 1:
        pre_q = pre_q.bfloat16().float()
    xt = nhwc(pre_q, cp, dt)
    yt = torch.empty(N * H * H, cp, dtype=dt, device=DEV)
    stats = torch.zeros(2 * C, device=DEV)
    m1d, m2d = m1.float().to(DEV).contiguous(), m2.float().to(DEV).contiguous()
    ops.conv_forward(code, ops.GATHER, N, H, H, C, cp, H, H, C, cp, 1, 1, 1, 0, xt.data_ptr(), wt.data_ptr(), C, cp,
                     yt.data_ptr(), act="lrelu", slope=0.1, mask=m1d.data_ptr(), mask_pitch=C, stats=stats.data_ptr())
    ss = torch.empty(4 * C, device=DEV)
    gm, bt = gamma.detach().float().to(DEV), beta.detach().float().to(DEV)
    rmd, rvd = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    nbt = torch.zeros((), dtype=torch.long, device=DEV)
    pix = N * H * H
    ops.bn_finalize(stats.data_ptr(), C, pix, gm.data_ptr(), bt.data_ptr(), 1e-5, 0.1, rmd.data_ptr(), rvd.data_ptr(),
                    nbt.data_ptr(), ops.ptr(ss), ops.ptr(ss, C), ops.ptr(ss, 2 * C), ops.ptr(ss, 3 * C))
    ut = torch.empty(pix, cp, dtype=dt, device=DEV)
    ops.scale_shift_mask(yt.data_ptr(), code, cp, ut.data_ptr(), code, cp, pix, H * H, C, scale=ops.ptr(ss),
                         shift=ops.ptr(ss, C), mask=m2d.data_ptr(), mask_pitch=C)
    torch.cuda.synchronize()
    tol = 1e-5 if code == 0 else 1.5e-2
    assert rel_err(from_nhwc(ut, N, H, H, C), u) < tol
    assert int(nbt) == 1 and float(stats.abs().max()) == 0.0
    assert rel_err(rmd, rm) < max(tol, 1e-4) and rel_err(rvd, rv) < max(tol, 1e-4)
    # backward
    dUt = nhwc(dU.float(), cp, dt)
    sums = torch.zeros(2 * C, device=DEV)
    ops.bn_bwd_reduce(dUt.data_ptr(), code, cp, yt.data_ptr(), code, cp, pix, H * H, C, m2d.data_ptr(), C,
                      ops.ptr(ss, 2 * C), ops.ptr(ss, 3 * C), sums.data_ptr())
    dpre = torch.empty(pix, cp, dtype=dt, device=DEV)
    dbias, dgam, dbet = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    ops.act_backward(dUt.data_ptr(), code, cp, yt.data_ptr(), code, cp, dpre.data_ptr(), code, cp, pix, H * H, C,
                     "lrelu", 0.1, out_mask=m1d.data_ptr(), mask_pitch=C, dbias=dbias.data_ptr(),
                     bn_sums=sums.data_ptr(), bn_mask=m2d.data_ptr(), bn_mask_pitch=C, bn_gamma=gm.data_ptr(),
                     bn_mean=ops.ptr(ss, 2 * C), bn_invstd=ops.ptr(ss, 3 * C), bn_dgamma=dgam.data_ptr(),
                     bn_dbeta=dbet.data_ptr())
    torch.cuda.synchronize()
    tol = 1e-4 if code == 0 else 3e-2
    assert rel_err(from_nhwc(dpre, N, H, H, C), pre.grad) < tol
    assert rel_err(dgam, gamma.grad) < tol and rel_err(dbet, beta.grad) < tol
    assert rel_err(dbias, pre.grad.sum((0, 2, 3))) < tol


def test_bce_adam_cast_fill():
    ops = ops_mod()
    g = torch.Generator().manual_seed(5)
    n = 777
    l = (3 * torch.randn(n, 1, generator=g)).requires_grad_()
    loss = F.binary_cross_entropy_with_logits(l, torch.ones(n, 1)) * 0.5
    loss.backward()
    ld = l.detach().to(DEV)
    out, dl = torch.zeros(2, device=DEV), torch.empty(n, 1, device=DEV)
    ops.bce_logits(ld.data_ptr(), 0, 1, n, 1.0, 0.5, out.data_ptr(), dl.data_ptr(), 0, 1)
    ops.sigmoid_mean(ld.data_ptr(), 0, 1, n, ops.ptr(out, 1))
    torch.cuda.synchronize()
    assert abs(float(out[0]) - float(loss)) < 1e-5 * float(loss)
    assert rel_err(dl, l.grad) < 1e-5
    assert abs(float(out[1]) - float(torch.sigmoid(l).mean())) < 1e-5
    # Adam, 3 steps against torch.optim.Adam
    p = torch.randn(1003, generator=g)
    pt = p.clone().requires_grad_()
    opt = torch.optim.Adam([pt], lr=1e-3, betas=(0.5, 0.999))
    n4 = 1004
    pd = torch.zeros(n4, device=DEV)
    pd[:1003] = p.to(DEV)
    m, v, gd = torch.zeros(n4, device=DEV), torch.zeros(n4, device=DEV), torch.zeros(n4, device=DEV)
    state = torch.tensor([0, 1e-3, 0.5, 0.999, 1e-8, 1.0, 0, 0], dtype=torch.float32, device=DEV)
    for _ in range(3):
        gr = torch.randn(1003, generator=g)
        pt.grad = gr.clone()
        opt.step()
        gd[:1003] = gr.to(DEV)
        ops.adam_step(pd.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n4, state.data_ptr())
    torch.cuda.synchronize()
    assert rel_err(pd[:1003], pt) < 1e-6 and float(state[0]) == 3.0
    a = torch.randn(1000, device=DEV)
    b = torch.empty(1000, dtype=torch.bfloat16, device=DEV)
    ops.cast(a.data_ptr(), 0, b.data_ptr(), 1, 1000)
    torch.cuda.synchronize()
    assert torch.equal(b, a.bfloat16())
    ops.fill_f32(a.data_ptr(), 2.5, 1000)
    torch.cuda.synchronize()
    assert float(a.min()) == 2.5 == float(a.max())


@pytest.mark.parametrize("case", [(3, 16, 9, 1, 5, 2, 2, 1), (2, 24, 7, 2, 5, 2, 1, 0), (2, 8, 6, 1, 3, 2, 1, 0), (130, 64, 5, 1, 5, 2, 2, 1)],
                         ids=str)
def test_taps_as_channels_transposed_conv(case):
    """icf_conv_forward (plain GEMM over the input pixels, taps as output columns) + icf_col2im_taps == conv_transpose2d for
    stride-2 layers with 1-2 output channels (audio_mnist.py:242 ConvTranspose2d(64, 1, 5, 2, 2, 1)), and icf_im2col_taps +
    GEMM == its data gradient (conv2d of the one-channel gradient)."""
    ops = ops_mod()
    n, C, H, K, k, s, p, op = case
    g = torch.Generator().manual_seed(11)
    x = torch.randn(n, C, H, H, generator=g)
    w = 0.2 * torch.randn(C, K, k, k, generator=g)                     # ConvTranspose2d layout [Cin][Cout][R][S]
    b = torch.randn(K, generator=g)
    P = (H - 1) * s - 2 * p + k + op
    T_, TP = k * k, (k * k + 7) // 8 * 8 if K == 1 else k * k
    cols = K * TP if K == 1 else (K * T_ + 7) // 8 * 8
    xb = x.bfloat16().float()
    wb = w.bfloat16().float()
    # operand rows (k, tap) x channels:  rows[kk*TP + t][c] = w[c][kk][t]
    wrows = torch.zeros(cols, C)
    for kk in range(K):
        wrows[kk * TP:kk * TP + T_] = wb[:, kk].reshape(C, T_).t()
    wd = wrows.bfloat16().to(DEV).contiguous()
    xd = nhwc(xb, C, torch.bfloat16)
    Tm = torch.empty(n * H * H, cols, dtype=torch.bfloat16, device=DEV)
    ops.conv_forward(1, ops.GATHER, n, H, H, C, C, H, H, K * TP if K == 1 else K * T_, cols, 1, 1, 1, 0, xd.data_ptr(), wd.data_ptr(),
                     cols, C, Tm.data_ptr())
    out = torch.empty(n * P * P, 8, dtype=torch.float32, device=DEV)
    b_d = b.to(DEV)
    ops.col2im_taps(Tm.data_ptr(), cols, TP, n, H, H, P, P, K, k, k, s, p, b_d.data_ptr(), "tanh", 0.0, out.data_ptr(), 0, 8)
    ref = torch.tanh(F.conv_transpose2d(xb.double(), wb.double(), b.double(), stride=s, padding=p, output_padding=op))
    got = from_nhwc(out, n, P, P, K)
    assert rel_err(got, ref) < 2e-2
    if K == 1:
        # data gradient of the one-output-channel layer: dX[c] = conv2d(dOut, w[c]) = im2col(dOut) [pixels, taps] x w[taps, c]
        dy = torch.randn(n, 1, P, P, generator=g)
        dyd = nhwc(dy.bfloat16().float(), 8, torch.bfloat16)
        A = torch.empty(n * H * H, TP, dtype=torch.bfloat16, device=DEV)
        ops.im2col_taps(dyd.data_ptr(), 1, 8, n, P, P, H, H, k, k, s, p, A.data_ptr(), TP)
        wt = torch.zeros(C, TP)
        wt[:, :T_] = wb[:, 0].reshape(C, T_)
        dx = torch.empty(n * H * H, C, dtype=torch.bfloat16, device=DEV)
        wt_d = wt.bfloat16().to(DEV)
        ops.conv_forward(1, ops.GATHER, n, H, H, T_, TP, H, H, C, C, 1, 1, 1, 0, A.data_ptr(), wt_d.data_ptr(), C, TP, dx.data_ptr())
        ref_dx = F.conv2d(dy.bfloat16().double(), wb.double()[:, :1], stride=s, padding=p)
        # conv2d with weight [C][1][k][k] over the one-channel gradient (cross-correlation, as in conv_transpose2d's adjoint)
        assert rel_err(from_nhwc(dx, n, H, H, C), ref_dx[:, :, :H, :H]) < 2e-2


def test_bn_fold_kernels():
    """icf_bn_fold_weights / icf_bn_fold_wgrad: conv(scale*y + shift; w, b) == conv(y; w', b') and the weight-gradient fix-up."""
    ops = ops_mod()
    g = torch.Generator().manual_seed(12)
    K, T_, C, Cp = 24, 16, 20, 24
    w = torch.randn(K, T_, Cp, generator=g)
    w[:, :, C:] = 0
    scale, shift, bias = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g), torch.randn(K, generator=g)
    wd, wo, bo = w.to(DEV).contiguous(), torch.empty(K * T_ * Cp, device=DEV), torch.empty(K, device=DEV)
    sc_d, sh_d, b_d = scale.to(DEV), shift.to(DEV), bias.to(DEV)          # kept alive: raw addresses go to the library
    ops.bn_fold_weights(wd.data_ptr(), 0, K, T_, Cp, C, sc_d.data_ptr(), sh_d.data_ptr(), b_d.data_ptr(), wo.data_ptr(), bo.data_ptr())
    ref_w = w.clone()
    ref_w[:, :, :C] *= scale
    assert rel_err(wo.reshape(K, T_, Cp), ref_w) < 1e-6
    assert rel_err(bo, bias + (w[:, :, :C] * shift).sum(dim=(1, 2))) < 1e-5
    G = torch.randn(K, T_, C, generator=g)
    db = torch.randn(K, generator=g)
    Gd, db_d = G.to(DEV).contiguous(), db.to(DEV)
    ops.bn_fold_wgrad(Gd.data_ptr(), K, T_, C, sc_d.data_ptr(), sh_d.data_ptr(), db_d.data_ptr())
    assert rel_err(Gd, G * scale + shift * db[:, None, None]) < 1e-6
