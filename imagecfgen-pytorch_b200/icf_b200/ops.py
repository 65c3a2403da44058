"""Thin functional layer over the C-ABI (include/icf.h): torch tensors in, raw device pointers out.

Every function launches on torch's *current* CUDA stream and never synchronises, so the whole hot path can
be captured into a CUDA graph.  torch is used for device memory and streams only; no torch arithmetic
stands in for a kernel here, and a missing libicf_b200.so raises (icf_b200.lib.load).
"""
import ctypes as C
import functools

import torch

from . import lib as _l

F32, BF16 = _l.F32, _l.BF16
ACT = {"none": _l.ACT_NONE, "lrelu": _l.ACT_LRELU, "tanh": _l.ACT_TANH}
GATHER, TRANSPOSED = _l.FORM_GATHER, _l.FORM_TRANSPOSED
_TORCH = {F32: torch.float32, BF16: torch.bfloat16}


def torch_dtype(code):
    return _TORCH[code]


def code_of(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"libicf_b200 handles float32 / bfloat16 tensors, got {t.dtype}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "icf_b200: the BiGAN hot path runs only in the sm_100a CUDA extension (libicf_b200.so); "
                f"got a tensor on {t.device}. There is no CPU / eager fallback.")


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t, offset_elems: int = 0):
    """Device address of ``t`` (+ an element offset), or None."""
    if t is None:
        return None
    return t.data_ptr() + offset_elems * t.element_size()


LAUNCHES = 0          # C-ABI launch calls made so far (bench.py reports the per-step count)
PROFILE = None        # set to a list to record (name, start_event, end_event, flops, bytes) per launch


def _launch(name, fn, *args, flops=0.0, nbytes=0.0, detail=""):
    global LAUNCHES
    LAUNCHES += 1
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = fn(*args, stream())
        e1.record()
        PROFILE.append((name, e0, e1, flops, nbytes, detail))
    else:
        st = fn(*args, stream())
    _l.check(st, name)


@functools.lru_cache(maxsize=None)
def valid_taps(form, H, P, R, stride, pad):
    """Sum over output positions of the kernel taps that land inside the un-padded input."""
    if form == GATHER:
        return sum(1 for p in range(P) for r in range(R) if 0 <= p * stride - pad + r < H)
    return sum(1 for p in range(P) for r in range(R)
               if (p + pad - r) >= 0 and (p + pad - r) % stride == 0 and (p + pad - r) // stride < H)


def conv_forward(dtype, form, N, H, W, Cc, in_pitch, P, Q, K, out_pitch, R, S, stride, pad,
                 src, w, w_rows, w_pitch, dst, bias=None, act="none", slope=0.0, out_f32=False,
                 mask=None, mask_pitch=0, stats=None, accumulate=False, win=0, alg_flops=None, alg_bytes=None, partial=None):
    """src/w/dst/bias/mask/stats are raw addresses (ints) or None; ``partial``: zeroed fp32 scratch tensor that allows split-K."""
    a = _l.ConvArgs(dtype, form, N, H, W, Cc, in_pitch, P, Q, K, out_pitch, R, S, stride, pad,
                    w_rows, w_pitch, ACT[act], slope, 1 if out_f32 else 0, mask_pitch,
                    1 if accumulate else 0, win, src, w, bias, dst, mask, stats)
    fl = nb = 0.0
    det = ""
    if PROFILE is not None:
        es = 4 if dtype == F32 else 2
        fl = 2.0 * N * Cc * K * valid_taps(form, H, P, R, stride, pad) * valid_taps(form, W, Q, S, stride, pad)
        nb = float(es) * (N * (H * W * Cc + P * Q * K) + K * Cc * R * S)
        if alg_flops is not None:
            fl = alg_flops
        if alg_bytes is not None:
            nb = alg_bytes
        det = f"{'gather' if form == GATHER else 'transp'} {Cc}x{H}x{W}->{K}x{P}x{Q} k{R}s{stride}p{pad}" + (f" win{win}" if win > 1 else "")
    if partial is not None:
        _launch("icf_conv_forward", _l.load().icf_conv_forward_splitk, C.byref(a), partial.data_ptr(), partial.numel(), flops=fl,
                nbytes=nb, detail=det)
        return
    _launch("icf_conv_forward", _l.load().icf_conv_forward, C.byref(a), flops=fl, nbytes=nb, detail=det)


def conv_wgrad(dtype, N, P, Q, A, a_pitch, H, W, B, b_pitch, R, S, stride, pad, small, big, dw, win=0, alg_flops=None,
               alg_bytes=None):
    a = _l.WgradArgs(dtype, N, P, Q, A, a_pitch, H, W, B, b_pitch, R, S, stride, pad, small, big, dw, win)
    fl = nb = 0.0
    if PROFILE is not None:
        es = 4 if dtype == F32 else 2
        fl = 2.0 * N * A * B * valid_taps(GATHER, H, P, R, stride, pad) * valid_taps(GATHER, W, Q, S, stride, pad)
        nb = float(es) * N * (H * W * B + P * Q * A) + 4.0 * A * B * R * S
    if alg_flops is not None:
        fl = alg_flops
    if alg_bytes is not None:
        nb = alg_bytes
    det = (f"wgrad A{A}x{P}x{Q} B{B}x{H}x{W} k{R}s{stride}p{pad}" + (f" win{win}" if win > 1 else "")) if PROFILE is not None else ""
    _launch("icf_conv_wgrad", _l.load().icf_conv_wgrad, C.byref(a), flops=fl, nbytes=nb, detail=det)


def make_perm(d0, d1, d2, s0, s1, s2, d2_pad=None, d0_pad=None):
    return _l.Perm(d0, d1, d2, s0, s1, s2, d2 if d2_pad is None else d2_pad, d0 if d0_pad is None else d0_pad)


def pack(src, dst, dst_dtype, perm):
    _launch("icf_pack", _l.load().icf_pack, src, dst, dst_dtype, C.byref(perm))


class PackTable:
    """Device-resident table of re-packing jobs (icf_pack_job[]) launched as one kernel by ``run``."""

    def __init__(self, jobs, device, unpack=False):
        """jobs: list of (src_ptr, dst_ptr, dst_dtype, perm) with perm an icf_perm or icf_perm4; ``unpack``: the table
        is for icf_unpack_multi (packed fp32 accumulators -> checkpoint-layout gradients)."""
        arr = (_l.PackJob * len(jobs))()
        self.max_elems = 0
        self.unpack = unpack
        for j, (src, dst, dt, perm) in zip(arr, jobs):
            j.src, j.dst, j.dst_dtype = src, dst, dt
            if isinstance(perm, _l.Perm4):
                j.kind, j.p4 = 1, perm
                n = perm.d0 * perm.d1 * (perm.d2 * perm.d3 if unpack else perm.row_pitch)
            else:
                j.kind, j.p = 0, perm
                n = perm.d0 * perm.d1 * perm.d2 if unpack else perm.d0_pad * perm.d1 * perm.d2_pad
            self.max_elems = max(self.max_elems, n)
        self.n = len(jobs)
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8) if self.n else torch.zeros(0, dtype=torch.uint8)
        self.dev = raw.to(device)

    def run(self):
        if self.n:
            fn = _l.load().icf_unpack_multi if self.unpack else _l.load().icf_pack_multi
            _launch("icf_unpack_multi" if self.unpack else "icf_pack_multi", fn, self.dev.data_ptr(), self.n, self.max_elems)


def make_perm4(d0, d1, d2, d3, s0, s1, s2, s3, d3_pad, row_pitch):
    return _l.Perm4(d0, d1, d2, d3, s0, s1, s2, s3, d3_pad, row_pitch)


def pack4(src, dst, dst_dtype, perm):
    _launch("icf_pack4", _l.load().icf_pack4, src, dst, dst_dtype, C.byref(perm))


def unpack4(src_packed, dst, perm):
    _launch("icf_unpack4", _l.load().icf_unpack4, src_packed, dst, C.byref(perm))


def unpack(src_packed, dst, perm, atomic_add=False):
    _launch("icf_unpack", _l.load().icf_unpack, src_packed, dst, C.byref(perm), 1 if atomic_add else 0)


_ARGMAX_DT = {torch.float32: 0, torch.bfloat16: 1, torch.int32: 2, torch.int64: 3}


def argmax_rows(x: torch.Tensor) -> torch.Tensor:
    """First-max argmax over dim 1 of an (N,K) tensor -> int32 (N,)."""
    require_cuda(x)
    if x.dtype not in _ARGMAX_DT:
        x = x.float()
    x = x.contiguous()
    n, k = x.shape
    out = torch.empty(n, dtype=torch.int32, device=x.device)
    _launch("icf_argmax_rows", _l.load().icf_argmax_rows, x.data_ptr(), _ARGMAX_DT[x.dtype], n, k, out.data_ptr())
    return out


def _vp_array(ptrs):
    arr = (C.c_void_p * _l.MAX_PLANES)()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


def image_features(dtype, N, H, W, feat_pitch, x, x_dtype, x_pitch, tables, indices, conts, mask, mask_pitch,
                   feat, dfeat=None, dtables=None, backward=False, pad=0):
    a = _l.ImgFeatArgs()
    a.dtype, a.N, a.H, a.W, a.feat_pitch = dtype, N, H, W, feat_pitch
    a.x_dtype, a.x_pitch, a.n_emb, a.n_cont, a.mask_pitch = x_dtype, x_pitch, len(tables), len(conts), mask_pitch
    a.pad = pad
    a.x = x
    a.emb_table = _vp_array(tables)
    a.emb_index = _vp_array(indices)
    a.cont = _vp_array(conts)
    a.mask, a.feat, a.dfeat = mask, feat, dfeat
    a.demb_table = _vp_array(dtables or [])
    if backward:
        _launch("icf_image_features_bwd", _l.load().icf_image_features_bwd, C.byref(a))
    else:
        _launch("icf_image_features_fwd", _l.load().icf_image_features_fwd, C.byref(a))


def latent_features(dtype, N, latent, feat_pitch, z, z_dtype, z_pitch, emb_k, tables, onehots, conts, feat,
                    dfeat=None, dz=None, dtables=None, donehots=None, dconts=None, backward=False):
    a = _l.LatFeatArgs()
    a.dtype, a.N, a.latent, a.feat_pitch = dtype, N, latent, feat_pitch
    a.z_dtype, a.z_pitch, a.n_emb, a.n_cont = z_dtype, z_pitch, len(tables), len(conts)
    for i, k in enumerate(emb_k):
        a.emb_k[i] = k
    a.z = z
    a.emb_table = _vp_array(tables)
    a.onehot = _vp_array(onehots)
    a.cont = _vp_array(conts)
    a.feat, a.dfeat, a.dz = feat, dfeat, dz
    a.demb_table = _vp_array(dtables or [])
    a.donehot = _vp_array(donehots or [])
    a.dcont = _vp_array(dconts or [])
    if backward:
        _launch("icf_latent_features_bwd", _l.load().icf_latent_features_bwd, C.byref(a))
    else:
        _launch("icf_latent_features_fwd", _l.load().icf_latent_features_fwd, C.byref(a))


def bn_finalize(stats, Cc, count, gamma, beta, eps, momentum, rmean, rvar, nbt, scale, shift, save_mean,
                save_invstd):
    _launch("icf_bn_finalize", _l.load().icf_bn_finalize, stats, Cc, float(count), gamma, beta, eps, momentum, rmean, rvar, nbt, scale,
                               shift, save_mean, save_invstd)


def scale_shift_mask(y, y_dtype, y_pitch, u, u_dtype, u_pitch, pixels, pixels_per_sample, Cc, scale=None,
                     shift=None, mask=None, mask_pitch=0):
    es = lambda d: 4 if d == F32 else 2
    _launch("icf_scale_shift_mask", _l.load().icf_scale_shift_mask, y, y_dtype, y_pitch, u, u_dtype, u_pitch, pixels, pixels_per_sample, Cc,
                                    scale, shift, mask, mask_pitch, nbytes=float(pixels) * Cc * (es(y_dtype) + es(u_dtype)),
            detail=f"scale_shift_mask C{Cc} pix{pixels}" if PROFILE is not None else "")


def bn_bwd_reduce(dU, d_dtype, d_pitch, y, y_dtype, y_pitch, pixels, pps, Cc, mask, mask_pitch, mean, invstd,
                  sums):
    es = lambda d: 4 if d == F32 else 2
    _launch("icf_bn_bwd_reduce", _l.load().icf_bn_bwd_reduce, dU, d_dtype, d_pitch, y, y_dtype, y_pitch, pixels, pps, Cc, mask, mask_pitch,
                                 mean, invstd, sums, nbytes=float(pixels) * Cc * (es(d_dtype) + es(y_dtype)),
            detail=f"bn_bwd_reduce C{Cc} pix{pixels}" if PROFILE is not None else "")


def act_backward(dOut, d_dtype, d_pitch, y, y_dtype, y_pitch, dPre, p_dtype, p_pitch, pixels, pps, Cc, act,
                 slope, out_mask=None, mask_pitch=0, dbias=None, bn_sums=None, bn_mask=None, bn_mask_pitch=0,
                 bn_gamma=None, bn_mean=None, bn_invstd=None, bn_dgamma=None, bn_dbeta=None, bn_inv_world=0.0):
    a = _l.ActBwdArgs(d_dtype, d_pitch, y_dtype, y_pitch, p_dtype, p_pitch, pixels, pps, Cc, ACT[act], slope,
                      mask_pitch, bn_mask_pitch, dOut, y, dPre, out_mask, dbias, 0, bn_sums, bn_mask, bn_gamma,
                      bn_mean, bn_invstd, bn_dgamma, bn_dbeta, bn_inv_world)
    es = lambda d: 4 if d == F32 else 2
    _launch("icf_act_backward", _l.load().icf_act_backward, C.byref(a),
            nbytes=float(pixels) * Cc * (es(d_dtype) + es(y_dtype) + es(p_dtype)),
            detail=(f"act_backward C{Cc} pix{pixels}" + (" bn" if bn_sums else "")) if PROFILE is not None else "")


def bce_logits(logits, l_dtype, l_pitch, n, target, weight, loss_out, dlogits, d_dtype, d_pitch):
    _launch("icf_bce_logits", _l.load().icf_bce_logits, logits, l_dtype, l_pitch, n, target, weight, loss_out, dlogits, d_dtype, d_pitch)


def sigmoid_mean(logits, l_dtype, l_pitch, n, score_out):
    _launch("icf_sigmoid_mean", _l.load().icf_sigmoid_mean, logits, l_dtype, l_pitch, n, score_out)


def adam_step(param, grad, exp_avg, exp_avg_sq, n, state):
    _launch("icf_adam_step", _l.load().icf_adam_step, param, grad, exp_avg, exp_avg_sq, n, state)


def cast(src, src_dtype, dst, dst_dtype, n):
    _launch("icf_cast", _l.load().icf_cast, src, src_dtype, dst, dst_dtype, n)


def fill_f32(dst, value, n):
    _launch("icf_fill_f32", _l.load().icf_fill_f32, dst, value, n)


def mse_loss(x, target_stride, xr, xr_dtype, xr_pitch, n_img, pixels, weight, extra, loss_out, dxr, d_dtype, d_pitch):
    _launch("icf_mse_loss", _l.load().icf_mse_loss, x, target_stride, xr, xr_dtype, xr_pitch, n_img, pixels, weight, extra,
            loss_out, dxr, d_dtype, d_pitch)


def col_mean(x, n, p, xbar, var_out):
    _launch("icf_col_mean", _l.load().icf_col_mean, x, n, p, xbar, var_out)


def latent_l2(z, z_dtype, z_pitch, n, latent, weight, loss_out, dz, accumulate):
    _launch("icf_latent_l2", _l.load().icf_latent_l2, z, z_dtype, z_pitch, n, latent, weight, loss_out, dz,
            1 if accumulate else 0)


def explain_groups(specs):
    """specs: [(mode, width, offset, dout_ptr | None)] -> ctypes array for icf_explain_transform / icf_explain_backward."""
    arr = (_l.ExplainGroup * len(specs))()
    for g, (mode, width, offset, dout) in zip(arr, specs):
        g.mode, g.width, g.offset, g.dout = mode, width, offset, dout
    return arr


def explain_transform(raw, out, groups, rows):
    _launch("icf_explain_transform", _l.load().icf_explain_transform, raw, out, groups, len(groups), rows)


def explain_backward(out, draw, groups, rows):
    _launch("icf_explain_backward", _l.load().icf_explain_backward, out, draw, groups, len(groups), rows)


def scm_affine_cf(args):
    _launch("icf_scm_affine_cf", _l.load().icf_scm_affine_cf, C.byref(args))


def onehot_swap(idx: torch.Tensor, mask, rows: torch.Tensor):
    """rows (N,K) fp32 <- one_hot(idx) where mask (bool / uint8, or None = everywhere)."""
    require_cuda(idx, rows, mask)
    assert rows.dtype == torch.float32 and rows.is_contiguous() and idx.dtype in (torch.int32, torch.int64)
    m = None
    if mask is not None:
        m = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8).contiguous()
    n, k = rows.shape
    _launch("icf_onehot_swap", _l.load().icf_onehot_swap, idx.contiguous().data_ptr(), 1 if idx.dtype == torch.int64 else 0,
            ptr(m), n, k, rows.data_ptr())
    return rows


def bn_fold_weights(w, dtype, K, T, Cp, Cc, scale, shift, bias, w_out, bias_out):
    _launch("icf_bn_fold_weights", _l.load().icf_bn_fold_weights, w, dtype, K, T, Cp, Cc, scale, shift, bias, w_out, bias_out)


def bn_fold_wgrad(dw, K, T, Cc, scale, shift, dbias):
    _launch("icf_bn_fold_wgrad", _l.load().icf_bn_fold_wgrad, dw, K, T, Cc, scale, shift, dbias)


def col2im_taps(T, t_pitch, TP, N, H, W, P, Q, K, R, S, stride, pad, bias, act, slope, out, out_dtype, out_pitch):
    nb = 2.0 * N * H * W * t_pitch + (4.0 if out_dtype == F32 else 2.0) * N * P * Q * K
    _launch("icf_col2im_taps", _l.load().icf_col2im_taps, T, t_pitch, TP, N, H, W, P, Q, K, R, S, stride, pad, bias, ACT[act],
            slope, out, out_dtype, out_pitch, nbytes=nb, detail=f"col2im_taps {H}x{W}->{K}x{P}x{Q}" if PROFILE is not None else "")


def im2col_taps(src, src_dtype, src_pitch, N, P, Q, H, W, R, S, stride, pad, A, a_pitch):
    nb = 2.0 * N * H * W * a_pitch + (4.0 if src_dtype == F32 else 2.0) * N * P * Q
    _launch("icf_im2col_taps", _l.load().icf_im2col_taps, src, src_dtype, src_pitch, N, P, Q, H, W, R, S, stride, pad, A, a_pitch,
            nbytes=nb, detail=f"im2col_taps {P}x{Q}->{H}x{W}x{a_pitch}" if PROFILE is not None else "")
