"""Thin functional layer over the C-ABI (include/icf.h): torch tensors in, raw device pointers out.

Every function launches on torch's *current* CUDA stream and never synchronises, so the whole hot path can
be captured into a CUDA graph.  torch is used for device memory and streams only; no torch arithmetic
stands in for a kernel here, and a missing libicf_b200.so raises (icf_b200.lib.load).
"""
import ctypes as C

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


def conv_forward(dtype, form, N, H, W, Cc, in_pitch, P, Q, K, out_pitch, R, S, stride, pad,
                 src, w, w_rows, w_pitch, dst, bias=None, act="none", slope=0.0, out_f32=False,
                 mask=None, mask_pitch=0, stats=None, accumulate=False):
    """src/w/dst/bias/mask/stats are raw addresses (ints) or None."""
    a = _l.ConvArgs(dtype, form, N, H, W, Cc, in_pitch, P, Q, K, out_pitch, R, S, stride, pad,
                    w_rows, w_pitch, ACT[act], slope, 1 if out_f32 else 0, mask_pitch,
                    1 if accumulate else 0, src, w, bias, dst, mask, stats)
    L = _l.load()
    _l.check(L.icf_conv_forward(C.byref(a), stream()), "icf_conv_forward")


def conv_wgrad(dtype, N, P, Q, A, a_pitch, H, W, B, b_pitch, R, S, stride, pad, small, big, dw):
    a = _l.WgradArgs(dtype, N, P, Q, A, a_pitch, H, W, B, b_pitch, R, S, stride, pad, small, big, dw)
    L = _l.load()
    _l.check(L.icf_conv_wgrad(C.byref(a), stream()), "icf_conv_wgrad")


def make_perm(d0, d1, d2, s0, s1, s2, d2_pad=None, d0_pad=None):
    return _l.Perm(d0, d1, d2, s0, s1, s2, d2 if d2_pad is None else d2_pad, d0 if d0_pad is None else d0_pad)


def pack(src, dst, dst_dtype, perm):
    L = _l.load()
    _l.check(L.icf_pack(src, dst, dst_dtype, C.byref(perm), stream()), "icf_pack")


def unpack(src_packed, dst, perm, atomic_add=False):
    L = _l.load()
    _l.check(L.icf_unpack(src_packed, dst, C.byref(perm), 1 if atomic_add else 0, stream()), "icf_unpack")


_ARGMAX_DT = {torch.float32: 0, torch.bfloat16: 1, torch.int32: 2, torch.int64: 3}


def argmax_rows(x: torch.Tensor) -> torch.Tensor:
    """First-max argmax over dim 1 of an (N,K) tensor -> int32 (N,)."""
    require_cuda(x)
    if x.dtype not in _ARGMAX_DT:
        x = x.float()
    x = x.contiguous()
    n, k = x.shape
    out = torch.empty(n, dtype=torch.int32, device=x.device)
    L = _l.load()
    _l.check(L.icf_argmax_rows(x.data_ptr(), _ARGMAX_DT[x.dtype], n, k, out.data_ptr(), stream()),
             "icf_argmax_rows")
    return out


def _vp_array(ptrs):
    arr = (C.c_void_p * _l.MAX_PLANES)()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


def image_features(dtype, N, H, W, feat_pitch, x, x_dtype, x_pitch, tables, indices, conts, mask, mask_pitch,
                   feat, dfeat=None, dtables=None, backward=False):
    a = _l.ImgFeatArgs()
    a.dtype, a.N, a.H, a.W, a.feat_pitch = dtype, N, H, W, feat_pitch
    a.x_dtype, a.x_pitch, a.n_emb, a.n_cont, a.mask_pitch = x_dtype, x_pitch, len(tables), len(conts), mask_pitch
    a.x = x
    a.emb_table = _vp_array(tables)
    a.emb_index = _vp_array(indices)
    a.cont = _vp_array(conts)
    a.mask, a.feat, a.dfeat = mask, feat, dfeat
    a.demb_table = _vp_array(dtables or [])
    L = _l.load()
    if backward:
        _l.check(L.icf_image_features_bwd(C.byref(a), stream()), "icf_image_features_bwd")
    else:
        _l.check(L.icf_image_features_fwd(C.byref(a), stream()), "icf_image_features_fwd")


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
    L = _l.load()
    if backward:
        _l.check(L.icf_latent_features_bwd(C.byref(a), stream()), "icf_latent_features_bwd")
    else:
        _l.check(L.icf_latent_features_fwd(C.byref(a), stream()), "icf_latent_features_fwd")


def bn_finalize(stats, Cc, count, gamma, beta, eps, momentum, rmean, rvar, nbt, scale, shift, save_mean,
                save_invstd):
    L = _l.load()
    _l.check(L.icf_bn_finalize(stats, Cc, float(count), gamma, beta, eps, momentum, rmean, rvar, nbt, scale,
                               shift, save_mean, save_invstd, stream()), "icf_bn_finalize")


def scale_shift_mask(y, y_dtype, y_pitch, u, u_dtype, u_pitch, pixels, pixels_per_sample, Cc, scale=None,
                     shift=None, mask=None, mask_pitch=0):
    L = _l.load()
    _l.check(L.icf_scale_shift_mask(y, y_dtype, y_pitch, u, u_dtype, u_pitch, pixels, pixels_per_sample, Cc,
                                    scale, shift, mask, mask_pitch, stream()), "icf_scale_shift_mask")


def bn_bwd_reduce(dU, d_dtype, d_pitch, y, y_dtype, y_pitch, pixels, pps, Cc, mask, mask_pitch, mean, invstd,
                  sums):
    L = _l.load()
    _l.check(L.icf_bn_bwd_reduce(dU, d_dtype, d_pitch, y, y_dtype, y_pitch, pixels, pps, Cc, mask, mask_pitch,
                                 mean, invstd, sums, stream()), "icf_bn_bwd_reduce")


def act_backward(dOut, d_dtype, d_pitch, y, y_dtype, y_pitch, dPre, p_dtype, p_pitch, pixels, pps, Cc, act,
                 slope, out_mask=None, mask_pitch=0, dbias=None, bn_sums=None, bn_mask=None, bn_mask_pitch=0,
                 bn_gamma=None, bn_mean=None, bn_invstd=None, bn_dgamma=None, bn_dbeta=None):
    a = _l.ActBwdArgs(d_dtype, d_pitch, y_dtype, y_pitch, p_dtype, p_pitch, pixels, pps, Cc, ACT[act], slope,
                      mask_pitch, bn_mask_pitch, dOut, y, dPre, out_mask, dbias, 0, bn_sums, bn_mask, bn_gamma,
                      bn_mean, bn_invstd, bn_dgamma, bn_dbeta)
    L = _l.load()
    _l.check(L.icf_act_backward(C.byref(a), stream()), "icf_act_backward")


def bce_logits(logits, l_dtype, l_pitch, n, target, weight, loss_out, dlogits, d_dtype, d_pitch):
    L = _l.load()
    _l.check(L.icf_bce_logits(logits, l_dtype, l_pitch, n, target, weight, loss_out, dlogits, d_dtype, d_pitch,
                              stream()), "icf_bce_logits")


def sigmoid_mean(logits, l_dtype, l_pitch, n, score_out):
    L = _l.load()
    _l.check(L.icf_sigmoid_mean(logits, l_dtype, l_pitch, n, score_out, stream()), "icf_sigmoid_mean")


def adam_step(param, grad, exp_avg, exp_avg_sq, n, state):
    L = _l.load()
    _l.check(L.icf_adam_step(param, grad, exp_avg, exp_avg_sq, n, state, stream()), "icf_adam_step")


def cast(src, src_dtype, dst, dst_dtype, n):
    L = _l.load()
    _l.check(L.icf_cast(src, src_dtype, dst, dst_dtype, n, stream()), "icf_cast")


def fill_f32(dst, value, n):
    L = _l.load()
    _l.check(L.icf_fill_f32(dst, value, n, stream()), "icf_fill_f32")
