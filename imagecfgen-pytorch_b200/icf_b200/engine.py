"""Execution engine of the conditional-BiGAN hot path on raw NHWC buffers.

One ``NetExec`` runs one network (Encoder / Generator / Discriminator) of one family (icf_b200.arch) through
the C-ABI kernels: attribute/latent feature assembly -> towers of fused conv layers (bias + activation +
Dropout2d mask + BatchNorm statistics in the conv epilogue, BatchNorm apply + Dropout2d in one pass) and the
matching backward (fused activation/dropout/BatchNorm backward, dgrad, wgrad).  It replaces the forward /
autograd of the reference modules image_scms/mnist.py:21-154, audio_mnist.py:173-318, whalecalls.py:230-387,
esrf_acoustic.py:134-260.

Layout: activations are 2-D tensors [N*H*W, pitch] (NHWC, pitch = channels rounded up to 8) in the compute
dtype (fp32 or bf16); parameters stay fp32 in the checkpoint layout on the owning nn.Module and are re-packed
into K-major operand copies ([rows][tap][channels]) by ``repack()``.
"""
from typing import Dict, List, Optional

import torch

from . import ops
from .arch import Family, L, conv_out, convT_out, pad8

F32, BF16 = ops.F32, ops.BF16


def dtype_code(name) -> int:
    if name in (F32, "fp32", "f32", "float32", torch.float32):
        return F32
    if name in (BF16, "bf16", "bfloat16", torch.bfloat16):
        return BF16
    raise ValueError(f"compute dtype must be 'fp32' or 'bf16', got {name!r}")


class Act:
    """An NHWC activation: 2-D tensor [N*H*W, pitch] plus a channel offset / count inside the pitch."""
    __slots__ = ("t", "off", "C")

    def __init__(self, t, C, off=0):
        self.t, self.C, self.off = t, C, off

    @property
    def pitch(self):
        return self.t.shape[1]

    @property
    def ptr(self):
        return ops.ptr(self.t, self.off)

    @property
    def code(self):
        return ops.code_of(self.t)


class LayerExec:
    """One conv-shaped contraction with its packed operands (Conv2d / ConvTranspose2d / Linear+Unflatten)."""

    def __init__(self, spec: L, hin: int, win: int, code: int, device, fold: bool = False):
        self.spec, self.code, self.device = spec, code, device
        self.Hin, self.Win = hin, win
        self.fold = 0
        if spec.kind == "conv":
            self.form, self.taps_r = ops.GATHER, spec.k
            self.P, self.Q = conv_out(hin, spec.k, spec.stride, spec.pad), conv_out(win, spec.k, spec.stride, spec.pad)
            self.Cin, self.Kout = spec.cin, spec.cout
            self.Hout, self.Wout, self.Cout = self.P, self.Q, spec.cout
        elif spec.kind == "convT":
            self.form, self.taps_r = ops.TRANSPOSED, spec.k
            self.P = convT_out(hin, spec.k, spec.stride, spec.pad, spec.opad)
            self.Q = convT_out(win, spec.k, spec.stride, spec.pad, spec.opad)
            self.Cin, self.Kout = spec.cin, spec.cout
            self.Hout, self.Wout, self.Cout = self.P, self.Q, spec.cout
        elif spec.kind == "linear":
            # nn.Linear + nn.Unflatten(1,(Cu,Hu,Wu)) as one GEMM whose output columns are permuted to
            # NHWC order: column (h*Wu+w)*Cu + c  <-  reference row c*Hu*Wu + h*Wu + w
            assert hin == 1 and win == 1 and spec.unflatten is not None
            cu, hu, wu = spec.unflatten
            assert cu * hu * wu == spec.cout
            self.form, self.taps_r = ops.GATHER, 1
            self.P = self.Q = 1
            self.Cin, self.Kout = spec.cin, spec.cout
            self.Hout, self.Wout, self.Cout = hu, wu, cu
        else:
            raise ValueError(spec.kind)
        self.R = self.S = self.taps_r
        self.taps = self.R * self.S
        self.stride = spec.stride if spec.kind != "linear" else 1
        self.pad = spec.pad if spec.kind != "linear" else 0
        self.in_pitch_min = pad8(self.Cin)
        self.wf_pitch = pad8(self.Cin)
        self.wb_pitch = pad8(self.Kout)
        dt = ops.torch_dtype(code)
        self.w_fwd = torch.zeros(self.Kout * self.taps * self.wf_pitch, dtype=dt, device=device)
        self.w_bwd = torch.zeros(self.Cin * self.taps * self.wb_pitch, dtype=dt, device=device)
        self.bias = torch.zeros(self.Kout, dtype=torch.float32, device=device)
        C_, K_, T = self.Cin, self.Kout, self.taps
        if spec.kind == "conv":          # reference weight [K][C][R][S]
            self.perm_f = ops.make_perm(K_, T, C_, C_ * T, 1, T, d2_pad=self.wf_pitch)
            self.perm_b = ops.make_perm(C_, T, K_, T, 1, C_ * T, d2_pad=self.wb_pitch)
            self.perm_g = ops.make_perm(K_, T, C_, C_ * T, 1, T)                 # dw packed [K][T][C]
            self.perm_bias = None
        elif spec.kind == "convT":       # reference weight [C][K][R][S]
            self.perm_f = ops.make_perm(K_, T, C_, T, 1, K_ * T, d2_pad=self.wf_pitch)
            self.perm_b = ops.make_perm(C_, T, K_, K_ * T, 1, T, d2_pad=self.wb_pitch)
            self.perm_g = ops.make_perm(C_, T, K_, K_ * T, 1, T)                 # dw packed [C][T][K]
            self.perm_bias = None
        else:                            # reference weight [Cu*HW][C]
            cu, hu, wu = spec.unflatten
            hw = hu * wu
            self.perm_f = ops.make_perm(hw, cu, C_, C_, hw * C_, 1, d2_pad=self.wf_pitch)
            self.perm_b = ops.make_perm(C_, hw, cu, 1, C_, hw * C_)
            self.wb_pitch = K_
            self.w_bwd = torch.zeros(self.Cin * K_, dtype=dt, device=device)
            self.perm_g = ops.make_perm(hw, cu, C_, C_, hw * C_, 1)              # dw packed [hw*cu][C]
            self.perm_bias = ops.make_perm(1, hw, cu, 0, 1, hw)
        self.wgrad_elems = self.Kout * self.taps * self.Cin
        # ConvTranspose2d on a 1x1 input (stride 1, no padding): output pixel (p, q) sees exactly tap (p, q), so the
        # forward is a plain GEMM [N, Cin] x [Cin, taps*Cout] whose output row is already the NHWC [k, k, Cout] image;
        # the transposed-conv form would push all taps through every output pixel (taps-1 of them multiplying zeros)
        self.gemm_fwd = (spec.kind == "convT" and hin == 1 and win == 1 and spec.pad == 0 and spec.opad == 0
                         and self.P == spec.k and self.Q == spec.k)
        if self.gemm_fwd:
            self.w_gemm = torch.zeros(self.taps * K_ * self.wf_pitch, dtype=dt, device=device)
            self.perm_gemm = ops.make_perm(self.taps, K_, C_, 1, self.taps, K_ * self.taps, d2_pad=self.wf_pitch)
            self.bias_gemm = torch.zeros(self.taps * K_, dtype=torch.float32, device=device)
            self.perm_bias_gemm = ops.make_perm(1, self.taps, K_, 0, 0, 1)
        # folded form of a small-channel first conv (bf16 tensor-core path): the S filter columns become part of
        # the channel dimension of a pre-padded 8-channel input, cutting the K loop from R*S taps to R rows
        if fold and spec.kind == "conv" and code == BF16 and pad8(self.Cin) == 8 and self.S * 8 <= 64 and self.S > 1:
            self.fold = self.S
            self.fold_pitch = 64
            self.w_fold = torch.zeros(K_ * self.R * self.fold_pitch, dtype=dt, device=device)
            self.perm_fold = ops.make_perm4(K_, self.R, self.S, C_, C_ * T, self.S, 1, T, 8, self.fold_pitch)
            self.perm_fold_g = ops.make_perm4(K_, self.R, self.S, C_, C_ * T, self.S, 1, T, 8, self.S * 8)
            self.wgrad_elems = K_ * self.R * self.S * 8
        # single-output-channel ConvTranspose2d tail (stride 1, no padding, bf16): its weight gradient is the
        # overlapping-window form dw[c][r][s*8 + 0] over the 8-channel-pitch gradient tensor (icf_wgrad_px8.cu)
        self.px8 = (spec.kind == "convT" and code == BF16 and self.Kout == 1 and self.stride == 1 and self.pad == 0
                    and self.Cin % 8 == 0 and self.Cin <= 128 and self.S >= 2)
        if self.px8:
            self.perm_px8_g = ops.make_perm4(C_, self.R, self.S, 1, self.R * self.S, self.S, 1, self.R * self.S, 8, self.S * 8)
            self.wgrad_elems = max(self.wgrad_elems, C_ * self.R * self.S * 8)
        # "Taps as channels" (icf.h, icf_col2im_taps / icf_im2col_taps): stride-2 layers with one channel on one side run as
        # plain GEMMs over the pixels with the filter taps as the other dimension.
        #  tail: ConvTranspose2d(C, 1, k, 2, ...) — forward = GEMM [pix, C] x [C, taps] + col2im (+bias, tanh); data gradient =
        #        im2col of the one-channel gradient + GEMM [pix, taps] x [taps, C]; weight gradient = GEMM-wgrad X^T x im2col
        #  head: the data gradient of a first Conv2d(<= 8 channels, K, k, 2, ...) = GEMM [pix, K] x [K, C*taps] + col2im
        import os as _os
        _off = bool(_os.environ.get("ICF_NO_TAPS"))
        self.tail_taps = (not _off and spec.kind == "convT" and code == BF16 and self.Kout == 1 and self.stride == 2
                          and self.Cin % 8 == 0 and self.taps <= 32 and not self.px8)
        self.head_ok = (not _off and spec.kind == "conv" and code == BF16 and self.stride == 2 and self.Cin <= 8
                        and self.Kout % 8 == 0 and self.taps <= 32)
        self.head_taps = False                    # enabled by NetExec.enable_head_taps for the attribute-plane channels
        self.head_active = False
        self._im2col = None                       # (key of the gradient tensor, its im2col matrix): shared by wgrad and dgrad
        if self.tail_taps:
            self.TP = pad8(self.taps)
            self.w_tail_f = torch.zeros(self.TP * self.wf_pitch, dtype=dt, device=device)       # [taps][1][C]
            self.perm_tail_f = ops.make_perm(T, 1, C_, 1, 0, T, d2_pad=self.wf_pitch, d0_pad=self.TP)
            self.w_tail_b = torch.zeros(C_ * self.TP, dtype=dt, device=device)                  # [C][1][taps]
            self.perm_tail_b = ops.make_perm(C_, 1, T, T, 0, 1, d2_pad=self.TP)
            self.perm_tail_g = ops.make_perm(C_, 1, T, T, 0, 1)                                 # dw packed [C][1][taps] = checkpoint order
        # BatchNorm folding (set by Tower): operand copy / bias of this layer with the preceding BatchNorm's scale / shift
        # folded in, rebuilt per forward from the batch statistics
        self.w_bnfold = None
        self.bias_bnfold = None
        # algorithmic bytes per image (each tensor touched once, real channels — not the folded / padded operand widths)
        self.alg_bytes_img = 2.0 * (hin * win * self.Cin + self.Hout * self.Wout * self.Cout)
        self.alg_flops_img = 2.0 * self.Cin * self.Kout * ops.valid_taps(self.form, hin, self.P, self.R, self.stride, self.pad) \
            * ops.valid_taps(self.form, win, self.Q, self.S, self.stride, self.pad)

    # ---- operands -------------------------------------------------------------------------------
    def enable_head_taps(self, c_lo: int, c_hi: int):
        """The data gradient of this first conv is only ever needed for input channels [c_lo, c_hi) (the embedded attribute
        planes: the image and the constant planes are data).  With one or two such channels it runs as
        GEMM [out pixel, K] x [K, (c, tap)] + col2im instead of the gather form's N = 16 MMA chains per parity class and tap."""
        if not self.head_ok or c_hi - c_lo < 1 or c_hi - c_lo > 2:
            return
        T = self.taps
        self.head_lo, self.head_n = c_lo, c_hi - c_lo
        self.head_rows = self.head_n * T                                                    # T columns (c, tap)
        self.head_cols = pad8(self.head_rows)
        dt = ops.torch_dtype(self.code)
        self.w_head = torch.zeros(self.head_cols * self.wb_pitch, dtype=dt, device=self.device)   # [(c,tap)][1][K]
        self.perm_head = ops.make_perm(self.head_rows, 1, self.Kout, 1, 0, self.Cin * T, d2_pad=self.wb_pitch,
                                       d0_pad=self.head_cols)
        self.head_taps = True

    def pack_jobs(self, weight: torch.Tensor, bias: torch.Tensor):
        """[(src ptr, dst ptr, dst dtype, perm)] that refresh this layer's operand copies from its parameters."""
        wp, bp = weight.data_ptr(), bias.data_ptr()
        jobs = [(wp, self.w_fwd.data_ptr(), self.code, self.perm_f), (wp, self.w_bwd.data_ptr(), self.code, self.perm_b)]
        if self.fold:
            jobs.append((wp, self.w_fold.data_ptr(), self.code, self.perm_fold))
        if self.perm_bias is None:
            jobs.append((bp, self.bias.data_ptr(), F32, ops.make_perm(1, 1, self.Kout, 0, 0, 1)))
        else:
            jobs.append((bp, self.bias.data_ptr(), F32, self.perm_bias))
        if self.gemm_fwd:
            jobs.append((wp, self.w_gemm.data_ptr(), self.code, self.perm_gemm))
            jobs.append((bp, self.bias_gemm.data_ptr(), F32, self.perm_bias_gemm))
        if self.tail_taps:
            jobs.append((wp, self.w_tail_f.data_ptr(), self.code, self.perm_tail_f))
            jobs.append((wp, self.w_tail_b.data_ptr(), self.code, self.perm_tail_b))
        if self.head_taps:      # rows (c, tap) of channels head_lo .. : the source starts head_lo*taps floats into each k-slab
            jobs.append((wp + 4 * self.head_lo * self.taps, self.w_head.data_ptr(), self.code, self.perm_head))
        return jobs

    def repack(self, weight: torch.Tensor, bias: torch.Tensor):
        """Stand-alone refresh of this layer (tools / tests); NetExec.repack batches all layers into one launch."""
        from .lib import Perm4
        for src, dst, dt, perm in self.pack_jobs(weight, bias):
            (ops.pack4 if isinstance(perm, Perm4) else ops.pack)(src, dst, dt, perm)

    # ---- launches -------------------------------------------------------------------------------
    def forward(self, N, x: Act, y: Act, mask=None, mask_pitch=0, stats=None, out_f32=False, w_override=None,
                bias_override=None):
        """``w_override`` / ``bias_override``: device addresses of a BatchNorm-folded operand copy / bias (see NetExec._tower_fwd)."""
        sp = self.spec
        if self.fold:      # x is the pre-padded [N][Hin+2p][Win+2p][8] feature tensor
            ops.conv_forward(self.code, self.form, N, self.Hin + 2 * self.pad, self.Win + 2 * self.pad,
                             self.fold * x.pitch, x.pitch, self.P, self.Q, self.Kout, y.pitch, self.R, 1, self.stride, 0,
                             x.ptr, self.w_fold.data_ptr(), self.Kout, self.fold_pitch, y.ptr, bias=self.bias.data_ptr(),
                             act=sp.act, slope=sp.slope, out_f32=out_f32, mask=mask, mask_pitch=mask_pitch, stats=stats,
                             win=self.fold, alg_flops=self.alg_flops_img * N,
                             alg_bytes=self.alg_bytes_img * N + 2.0 * self.wgrad_elems)
            return
        if self.tail_taps and mask is None and stats is None and w_override is None:
            # T[pixel][tap] = x[pixel][:] . w[:, tap]  (plain GEMM over the input pixels), then col2im + bias + activation
            Tm = torch.empty((N * self.Hin * self.Win, self.TP), dtype=ops.torch_dtype(self.code), device=self.device)
            ops.conv_forward(self.code, ops.GATHER, N, self.Hin, self.Win, self.Cin, x.pitch, self.Hin, self.Win, self.taps,
                             self.TP, 1, 1, 1, 0, x.ptr, self.w_tail_f.data_ptr(), self.TP, self.wf_pitch, Tm.data_ptr(),
                             alg_flops=self.alg_flops_img * N, alg_bytes=self.alg_bytes_img * N)
            ops.col2im_taps(Tm.data_ptr(), self.TP, self.TP, N, self.Hin, self.Win, self.P, self.Q, 1, self.R, self.S,
                            self.stride, self.pad, self.bias.data_ptr(), sp.act, sp.slope, y.ptr, y.code, y.pitch)
            return
        if self.gemm_fwd and mask is None and stats is None and y.pitch == self.Kout and y.off == 0:
            ops.conv_forward(self.code, ops.GATHER, N, 1, 1, self.Cin, x.pitch, 1, 1, self.taps * self.Kout,
                             y.pitch * self.taps, 1, 1, 1, 0, x.ptr, self.w_gemm.data_ptr(), self.taps * self.Kout,
                             self.wf_pitch, y.ptr, bias=self.bias_gemm.data_ptr(), act=sp.act, slope=sp.slope,
                             out_f32=out_f32, alg_flops=self.alg_flops_img * N)
            return
        ops.conv_forward(self.code, self.form, N, self.Hin, self.Win, self.Cin, x.pitch,
                         self.P, self.Q, self.Kout, y.pitch if sp.kind != "linear" else y.pitch * self.Hout * self.Wout,
                         self.R, self.S, self.stride, self.pad, x.ptr,
                         self.w_fwd.data_ptr() if w_override is None else w_override, self.Kout,
                         self.wf_pitch, y.ptr, bias=self.bias.data_ptr() if bias_override is None else bias_override,
                         act=sp.act, slope=sp.slope, out_f32=out_f32, mask=mask, mask_pitch=mask_pitch, stats=stats,
                         partial=self._splitk_scratch(N * self.P * self.Q, self.Kout, self.taps * self.Cin) if stats is None else None)

    def _splitk_scratch(self, pixels, K, reduction):
        """Zeroed fp32 scratch that lets icf_conv_forward_splitk deal the K loop of a small-grid layer to many CTAs: a few thousand
        output pixels against a long reduction (the 512..4096-channel layers of the spectrogram families at batch 32-128)."""
        # (MorphoMNIST never qualifies — its longest reduction is 4608 — so its forward stays bit-reproducible: split-K sums in
        # float-atomic order)
        if self.code != BF16 or pixels > 2048 or reduction <= 4608:
            return None
        return torch.zeros((pixels, (K + 255) // 256 * 256), dtype=torch.float32, device=self.device)

    def dgrad(self, N, dpre: Act, dx: Act):
        """dx[n,h,w,c] = sum_{k,taps} dpre[...]*w  — the other conv form with the transposed operand."""
        if self.tail_taps:
            A = self._im2col_of(N, dpre, consume=True)
            ops.conv_forward(self.code, ops.GATHER, N, self.Hin, self.Win, self.taps, self.TP, self.Hin, self.Win, self.Cin,
                             dx.pitch, 1, 1, 1, 0, A.data_ptr(), self.w_tail_b.data_ptr(), self.Cin, self.TP, dx.ptr,
                             alg_flops=self.alg_flops_img * N, alg_bytes=self.alg_bytes_img * N)
            return
        if self.head_taps and self.head_active:
            # dX towards the attribute-plane channels only: T[out pixel][(c, tap)] = dpre[out pixel][:] . w[:, c, tap], then col2im
            Tm = torch.empty((N * self.P * self.Q, self.head_cols), dtype=ops.torch_dtype(self.code), device=self.device)
            ops.conv_forward(self.code, ops.GATHER, N, self.P, self.Q, self.Kout, dpre.pitch, self.P, self.Q, self.head_rows,
                             self.head_cols, 1, 1, 1, 0, dpre.ptr, self.w_head.data_ptr(), self.head_cols, self.wb_pitch,
                             Tm.data_ptr(), alg_flops=self.alg_flops_img * N * self.head_n / self.Cin,
                             alg_bytes=2.0 * N * (self.P * self.Q * self.Kout + self.Hin * self.Win * self.head_n))
            ops.col2im_taps(Tm.data_ptr(), self.head_cols, self.taps, N, self.P, self.Q, self.Hin, self.Win, self.head_n, self.R,
                            self.S, self.stride, self.pad, None, "none", 0.0, ops.ptr(dx.t, dx.off + self.head_lo), dx.code,
                            dx.pitch)
            return
        if self.px8 and dx.pitch == self.Cin and dpre.pitch == 8 and self.wb_pitch == 8:
            # data gradient of the one-channel stride-1 ConvTranspose2d tail = a unit-stride conv over the 16-byte pixels of
            # dpre: w_bwd [Cin][R*S][8] IS the folded (win = S) packing [Cin][R][S*8], so the channel-major first-layer
            # kernel serves it (icf_conv_cm.cu)
            ops.conv_forward(self.code, ops.GATHER, N, self.P, self.Q, self.S * 8, 8, self.Hin, self.Win,
                             self.Cin, dx.pitch, self.R, 1, 1, 0, dpre.ptr, self.w_bwd.data_ptr(), self.Cin,
                             self.S * 8, dx.ptr, win=self.S, alg_flops=self.alg_flops_img * N,
                             alg_bytes=self.alg_bytes_img * N)
            return
        form = ops.TRANSPOSED if self.form == ops.GATHER else ops.GATHER
        lin = self.spec.kind == "linear"
        dp_pitch = dpre.pitch * self.Hout * self.Wout if lin else dpre.pitch
        ops.conv_forward(self.code, form, N, self.P, self.Q, self.Kout, dp_pitch,
                         self.Hin, self.Win, self.Cin, dx.pitch, self.R, self.S, self.stride, self.pad,
                         dpre.ptr, self.w_bwd.data_ptr(), self.Cin, self.wb_pitch, dx.ptr,
                         partial=None if lin else self._splitk_scratch(N * self.Hin * self.Win, self.Cin, self.taps * self.Kout))

    def _im2col_of(self, N, dpre: Act, consume: bool):
        """[N*Hin*Win, TP] tap matrix of the one-channel gradient ``dpre``.  A backward pass calls wgrad, then dgrad, on the same
        gradient tensor: wgrad always builds the matrix and leaves it for the dgrad that follows (``consume``), which drops it —
        a matrix is never reused across backward passes (the allocator hands the same address to other contents)."""
        key = (dpre.ptr, N)
        if consume and self._im2col is not None and self._im2col[0] == key:
            A = self._im2col[1]
            self._im2col = None
            return A
        A = torch.empty((N * self.Hin * self.Win, self.TP), dtype=ops.torch_dtype(self.code), device=self.device)
        ops.im2col_taps(dpre.ptr, dpre.code, dpre.pitch, N, self.P, self.Q, self.Hin, self.Win, self.R, self.S, self.stride,
                        self.pad, A.data_ptr(), self.TP)
        self._im2col = None if consume else (key, A)
        return A

    def unpack_perm(self):
        """Permutation that takes this layer's packed fp32 weight-gradient accumulator to the checkpoint layout."""
        if self.tail_taps:
            return self.perm_tail_g
        if self.fold:
            return self.perm_fold_g
        if self.px8:
            return self.perm_px8_g
        return self.perm_g

    def wgrad(self, N, dpre: Act, x: Act, gw: Optional[torch.Tensor], scratch: torch.Tensor):
        """gw (checkpoint layout, fp32) = weight gradient; ``scratch`` holds the packed fp32 accumulator.  With
        gw = None the caller has zeroed ``scratch`` and unpacks it later (NetExec batches both over all layers)."""
        n = self.wgrad_elems
        if gw is None:
            self._wgrad_accumulate(N, dpre, x, scratch)
            return
        ops.fill_f32(scratch.data_ptr(), 0.0, n)
        if self.tail_taps:
            self._wgrad_accumulate(N, dpre, x, scratch)
            ops.unpack(scratch.data_ptr(), gw.data_ptr(), self.perm_tail_g)
            return
        if self.fold:
            ops.conv_wgrad(self.code, N, self.P, self.Q, self.Kout, dpre.pitch, self.Hin + 2 * self.pad,
                           self.Win + 2 * self.pad, self.fold * x.pitch, x.pitch, self.R, 1, self.stride, 0, dpre.ptr,
                           x.ptr, scratch.data_ptr(), win=self.fold, alg_flops=self.alg_flops_img * N,
                           alg_bytes=self.alg_bytes_img * N + 4.0 * self.wgrad_elems)
            ops.unpack4(scratch.data_ptr(), gw.data_ptr(), self.perm_fold_g)
            return
        if self.px8 and dpre.pitch == 8:
            ops.conv_wgrad(self.code, N, self.Hin, self.Win, self.Cin, x.pitch, self.P, self.Q, self.S * 8, 8,
                           self.R, 1, 1, 0, x.ptr, dpre.ptr, scratch.data_ptr(), win=self.S,
                           alg_flops=self.alg_flops_img * N, alg_bytes=self.alg_bytes_img * N + 4.0 * self.wgrad_elems)
            ops.unpack4(scratch.data_ptr(), gw.data_ptr(), self.perm_px8_g)
            return
        lin = self.spec.kind == "linear"
        dp_pitch = dpre.pitch * self.Hout * self.Wout if lin else dpre.pitch
        if self.spec.kind == "convT":     # small = X (A = Cin), big = dY (B = Cout)
            ops.conv_wgrad(self.code, N, self.Hin, self.Win, self.Cin, x.pitch, self.P, self.Q, self.Kout,
                           dp_pitch, self.R, self.S, self.stride, self.pad, x.ptr, dpre.ptr, scratch.data_ptr())
        else:                             # small = dY (A = Cout), big = X (B = Cin)
            ops.conv_wgrad(self.code, N, self.P, self.Q, self.Kout, dp_pitch, self.Hin, self.Win, self.Cin,
                           x.pitch, self.R, self.S, self.stride, self.pad, dpre.ptr, x.ptr, scratch.data_ptr())
        ops.unpack(scratch.data_ptr(), gw.data_ptr(), self.perm_g)

    def _wgrad_accumulate(self, N, dpre: Act, x: Act, scratch: torch.Tensor):
        """The wgrad launch alone: accumulates into the (already zeroed) packed fp32 buffer ``scratch``."""
        if self.tail_taps:
            A = self._im2col_of(N, dpre, consume=False)           # dw[c][tap] = sum_pixels x[pixel][c] * A[pixel][tap]
            ops.conv_wgrad(self.code, N, self.Hin, self.Win, self.Cin, x.pitch, self.Hin, self.Win, self.taps, self.TP, 1, 1, 1,
                           0, x.ptr, A.data_ptr(), scratch.data_ptr(), alg_flops=self.alg_flops_img * N,
                           alg_bytes=self.alg_bytes_img * N + 4.0 * self.wgrad_elems)
            return
        if self.fold:
            ops.conv_wgrad(self.code, N, self.P, self.Q, self.Kout, dpre.pitch, self.Hin + 2 * self.pad,
                           self.Win + 2 * self.pad, self.fold * x.pitch, x.pitch, self.R, 1, self.stride, 0, dpre.ptr,
                           x.ptr, scratch.data_ptr(), win=self.fold, alg_flops=self.alg_flops_img * N,
                           alg_bytes=self.alg_bytes_img * N + 4.0 * self.wgrad_elems)
            return
        if self.px8 and dpre.pitch == 8:
            ops.conv_wgrad(self.code, N, self.Hin, self.Win, self.Cin, x.pitch, self.P, self.Q, self.S * 8, 8,
                           self.R, 1, 1, 0, x.ptr, dpre.ptr, scratch.data_ptr(), win=self.S,
                           alg_flops=self.alg_flops_img * N, alg_bytes=self.alg_bytes_img * N + 4.0 * self.wgrad_elems)
            return
        lin = self.spec.kind == "linear"
        dp_pitch = dpre.pitch * self.Hout * self.Wout if lin else dpre.pitch
        if self.spec.kind == "convT":     # small = X (A = Cin), big = dY (B = Cout)
            ops.conv_wgrad(self.code, N, self.Hin, self.Win, self.Cin, x.pitch, self.P, self.Q, self.Kout,
                           dp_pitch, self.R, self.S, self.stride, self.pad, x.ptr, dpre.ptr, scratch.data_ptr())
        else:                             # small = dY (A = Cout), big = X (B = Cin)
            ops.conv_wgrad(self.code, N, self.P, self.Q, self.Kout, dp_pitch, self.Hin, self.Win, self.Cin,
                           x.pitch, self.R, self.S, self.stride, self.pad, dpre.ptr, x.ptr, scratch.data_ptr())


def mask_sites(fam: Family):
    """[(tower, layer index, 'in'|'out'|'bn', p, channels)] of the Discriminator's Dropout2d sites in the order
    torch consumes the RNG: dx, then dz, then dxz (mnist.py:151-154), Sequential order inside each."""
    sites = []
    for tname in ("Dx", "Dz", "Dxz"):
        for i, l in enumerate(getattr(fam, tname)):
            if l.in_drop:
                sites.append((tname, i, "in", l.in_drop, l.cin))
            if l.out_drop:
                sites.append((tname, i, "out", l.out_drop, l.cout))
            if l.bn_drop:
                sites.append((tname, i, "bn", l.bn_drop, l.cout))
    return sites


def draw_masks(fam: Family, n: int, device) -> List[torch.Tensor]:
    """Dropout2d masks of ONE Discriminator forward, drawn with the aten calls nn.Dropout2d makes on a 4-D
    input (feature_dropout: ``empty(N,C,1,1).bernoulli_(1-p).div_(1-p)``) in the reference's order, so that
    the same seed yields the same masks as the reference on the same device."""
    return [torch.empty(n, c, 1, 1, device=device).bernoulli_(1 - p).div_(1 - p)
            for (_, _, _, p, c) in mask_sites(fam)]


class Tower:
    def __init__(self, specs, hin, win, code, device, fold_first=False):
        self.layers: List[LayerExec] = []
        h, w = hin, win
        for i, s in enumerate(specs):
            le = LayerExec(s, h, w, code, device, fold=(fold_first and i == 0))
            self.layers.append(le)
            h, w = le.Hout, le.Wout
        self.Hout, self.Wout = h, w
        self.code, self.device = code, device
        # BatchNorm -> Conv2d with nothing in between (no Dropout2d after the BatchNorm) and no padding in the consumer:
        # the affine transform folds into the consumer's weights and bias, the normalised tensor is never written
        # (mnist.py:111-112, dx.4 -> dx.5).  ICF_NO_BN_FOLD=1 keeps the separate pass (tuning / bisecting aid).
        self.bn_fold = {}
        import os
        if not os.environ.get("ICF_NO_BN_FOLD"):
            for i in range(len(self.layers) - 1):
                a, b = self.layers[i], self.layers[i + 1]
                if (a.spec.bn and not a.spec.bn_drop and b.spec.kind == "conv" and b.pad == 0 and not b.fold
                        and not b.gemm_fwd and not b.spec.in_drop):
                    dt = ops.torch_dtype(code)
                    b.w_bnfold = torch.zeros_like(b.w_fwd)
                    b.bias_bnfold = torch.zeros(b.Kout, dtype=torch.float32, device=device)
                    self.bn_fold[i] = i + 1
        # BatchNorm statistic accumulators, zero between uses (icf_bn_finalize clears them)
        self.stats = {i: torch.zeros(2 * le.Kout, dtype=torch.float32, device=device)
                      for i, le in enumerate(self.layers) if le.spec.bn}


class NetExec:
    """Runs one network of a family.  ``module`` owns the fp32 parameters / buffers (checkpoint layout)."""

    def __init__(self, fam: Family, role: str, module, code: int, device):
        assert role in ("E", "G", "D")
        self.fam, self.role, self.module, self.code, self.device = fam, role, module, code, device
        self.dt = ops.torch_dtype(code)
        H, W = fam.image
        self.H, self.W = H, W
        self.n_emb, self.n_cont = len(fam.cat_attrs), len(fam.cont_attrs)
        self.feat_ch = 1 + self.n_emb + self.n_cont
        if role == "E":
            self.towers = {"E": Tower(fam.E, H, W, code, device, fold_first=True)}
        elif role == "G":
            self.towers = {"G": Tower(fam.G, 1, 1, code, device)}
            self.lat_dim = fam.latent + 256 * self.n_emb + self.n_cont
            assert self.lat_dim == fam.G[0].cin, (self.lat_dim, fam.G[0].cin)
        else:
            self.towers = {"Dx": Tower(fam.Dx, H, W, code, device, fold_first=True), "Dz": Tower(fam.Dz, 1, 1, code, device),
                           "Dxz": Tower(fam.Dxz, 1, 1, code, device)}
            self.sites = mask_sites(fam)
        if role in ("E", "D") and self.n_emb > 0:
            self.towers["E" if role == "E" else "Dx"].layers[0].enable_head_taps(1, 1 + self.n_emb)
        self._versions = None
        self._pack_table = None
        self._pack_ptrs = None
        self._wg_flat = None
        self._wg_off = None
        self._wg_table = None
        self._wg_key = None
        self._scratch = None
        self.bn_sync = None                         # icf_b200.dp.Group: BatchNorm statistics span all ranks (SyncBN)
        self.idx_cache = None                       # per-step cache of attribute argmax indices (set by the trainers)
        self.keep_state = False                     # tests: keep the saved activations of the last forward
        self.last_state = None
        self.emb_key = 2 if role != "G" else 3      # index into fam.cat_attrs tuples

    # ---- parameters -----------------------------------------------------------------------------
    def tensors(self) -> Dict[str, torch.Tensor]:
        d = dict(self.module.named_parameters())
        d.update(dict(self.module.named_buffers()))
        return d

    def all_layers(self):
        for t in self.towers.values():
            for le in t.layers:
                yield le

    def repack(self, force=False):
        """Refresh the packed operand copies if any parameter changed (tracked by tensor version)."""
        ts = self.tensors()
        vers = tuple((k, v._version, v.data_ptr()) for k, v in ts.items() if k.endswith(("weight", "bias")))
        if not force and vers == self._versions:
            return
        ptrs = tuple(v.data_ptr() for k, v in ts.items() if k.endswith(("weight", "bias")))
        if self._pack_table is None or self._pack_ptrs != ptrs:        # one device-side job table per parameter placement
            jobs = []
            for le in self.all_layers():
                jobs += le.pack_jobs(ts[le.spec.key + ".weight"], ts[le.spec.key + ".bias"])
            self._pack_table, self._pack_ptrs = ops.PackTable(jobs, self.device), ptrs
        self._pack_table.run()                                         # every operand copy of the network, one launch
        self._versions = vers

    # ---- batched weight-gradient plumbing: one zero-fill before, one unpack launch after a whole backward ----------
    def _wg_layout(self):
        if self._wg_flat is None:
            off, self._wg_off = 0, {}
            for le in self.all_layers():
                self._wg_off[id(le)] = off
                off += (le.wgrad_elems + 3) // 4 * 4
            self._wg_flat = torch.empty(off, dtype=torch.float32, device=self.device)
        return self._wg_flat

    def wgrad_scratch(self, le):
        flat = self._wg_layout()
        o = self._wg_off[id(le)]
        return flat[o:o + le.wgrad_elems]

    def wgrad_begin(self):
        flat = self._wg_layout()
        ops.fill_f32(flat.data_ptr(), 0.0, flat.numel())

    def wgrad_end(self, grads):
        key = tuple(g.data_ptr() for g in grads.values())
        if self._wg_table is None or self._wg_key != key:
            jobs = [(self.wgrad_scratch(le).data_ptr(), grads[le.spec.key + ".weight"].data_ptr(), F32, le.unpack_perm())
                    for le in self.all_layers()]
            self._wg_table, self._wg_key = ops.PackTable(jobs, self.device, unpack=True), key
        self._wg_table.run()

    def scratch(self):
        if self._scratch is None:
            n = max(le.wgrad_elems for le in self.all_layers())
            self._scratch = torch.empty(n, dtype=torch.float32, device=self.device)
        return self._scratch

    def new_grads(self) -> Dict[str, torch.Tensor]:
        return {k: torch.zeros_like(v) for k, v in self.module.named_parameters()}

    # ---- tower forward / backward ---------------------------------------------------------------
    def _tower_fwd(self, tname, N, x: Act, masks: Dict, training: bool, save: bool, final: Optional[Act] = None,
                   final_mask=None, final_f32=False):
        """masks: {(layer index, 'out'|'bn'): tensor [N,C]}.  Returns (output Act, saved list)."""
        tw = self.towers[tname]
        ts = self.tensors()
        saved = []
        n_layers = len(tw.layers)
        folded_ss = None
        for i, le in enumerate(tw.layers):
            sp = le.spec
            last = i == n_layers - 1
            pix = N * le.Hout * le.Wout
            if last and final is not None:
                y = final
            else:
                ydt = torch.float32 if (last and final_f32) else self.dt
                y = Act(torch.empty((pix, pad8(le.Cout) if le.Cout > 1 else 1), dtype=ydt, device=self.device), le.Cout)
            om = masks.get((i, "out"))
            om_ptr, om_pitch = (ops.ptr(om), om.shape[1]) if om is not None else (None, 0)
            if last and final_mask is not None:
                assert om is None
                fm, fm_off = final_mask
                om, om_ptr, om_pitch = fm, ops.ptr(fm, fm_off), fm.shape[1]
            use_stats = sp.bn is not None and training
            stats = tw.stats[i].data_ptr() if use_stats else None
            wo = bo = None
            if folded_ss is not None:            # the BatchNorm in front of this layer lives in its operand copy and bias
                K_prev = le.Cin
                ops.bn_fold_weights(le.w_fwd.data_ptr(), self.code, le.Kout, le.taps, le.wf_pitch, le.Cin, ops.ptr(folded_ss),
                                    ops.ptr(folded_ss, K_prev), le.bias.data_ptr(), le.w_bnfold.data_ptr(),
                                    le.bias_bnfold.data_ptr())
                wo, bo = le.w_bnfold.data_ptr(), le.bias_bnfold.data_ptr()
            le.forward(N, x, y, mask=om_ptr, mask_pitch=om_pitch, stats=stats,
                       out_f32=(y.t.dtype == torch.float32 and self.code == BF16), w_override=wo, bias_override=bo)
            rec = {"x": x, "y": y, "out_mask": (om, om_ptr, om_pitch) if om is not None else None, "bn": None,
                   "in_fold": folded_ss}
            folded_ss = None
            nxt = y
            if sp.bn is not None:
                K = le.Kout
                gamma, beta = ts[sp.bn + ".weight"], ts[sp.bn + ".bias"]
                rm, rv = ts[sp.bn + ".running_mean"], ts[sp.bn + ".running_var"]
                ss = torch.empty(4 * K, dtype=torch.float32, device=self.device)   # scale|shift|mean|invstd
                if training:
                    nbt = ts.get(sp.bn + ".num_batches_tracked")
                    count = pix
                    if self.bn_sync is not None:         # SyncBN: sum / sum of squares over the whole (global) batch
                        self.bn_sync.all_reduce(tw.stats[i])
                        count = pix * self.bn_sync.world
                    ops.bn_finalize(stats, K, count, gamma.data_ptr(), beta.data_ptr(), 1e-5, 0.1, rm.data_ptr(),
                                    rv.data_ptr(), ops.ptr(nbt), ops.ptr(ss), ops.ptr(ss, K), ops.ptr(ss, 2 * K),
                                    ops.ptr(ss, 3 * K))
                else:
                    # eval mode: affine transform from the running statistics (tiny per-channel vectors)
                    inv = torch.rsqrt(rv + 1e-5)
                    ss[:K] = gamma * inv
                    ss[K:2 * K] = beta - rm * gamma * inv
                    ss[2 * K:3 * K] = rm
                    ss[3 * K:] = inv
                bm = masks.get((i, "bn"))
                if i in tw.bn_fold and bm is None:
                    folded_ss = ss                     # consumed by the next layer's operand copy: no pass over y
                    nxt = y
                else:
                    u = Act(torch.empty((pix, y.pitch), dtype=self.dt, device=self.device), le.Cout)
                    ops.scale_shift_mask(y.ptr, y.code, y.pitch, u.ptr, u.code, u.pitch, pix, le.Hout * le.Wout, K,
                                         scale=ops.ptr(ss), shift=ops.ptr(ss, K), mask=ops.ptr(bm),
                                         mask_pitch=bm.shape[1] if bm is not None else 0)
                    nxt = u
                rec["bn"] = {"ss": ss, "mask": bm, "training": training}
            if save:
                saved.append(rec)
            x = nxt
        return x, saved

    def _tower_bwd(self, tname, N, saved, dout: Act, grads: Optional[Dict], need_dx: bool, inplace_ok=True):
        """dout: gradient w.r.t. the tower's (consumer visible) output.  Returns dX Act of the tower input or None."""
        tw = self.towers[tname]
        ts = self.tensors()
        g = dout
        for i in range(len(tw.layers) - 1, -1, -1):
            le, rec = tw.layers[i], saved[i]
            sp = le.spec
            y, x = rec["y"], rec["x"]
            pix = N * le.Hout * le.Wout
            pps = le.Hout * le.Wout
            K = le.Kout if sp.kind != "linear" else le.Cout
            Kfull = le.Kout
            # for Linear+Unflatten the activation tensor is [N*hw, Cu]; treat it as [N, hw*Cu] rows
            if sp.kind == "linear":
                a_pix, a_pps, a_C = N, 1, Kfull
                ypitch, gpitch = y.pitch * pps, g.pitch * pps
            else:
                a_pix, a_pps, a_C = pix, pps, K
                ypitch, gpitch = y.pitch, g.pitch
            bn = rec["bn"]
            kw = {}
            if bn is not None:
                if not bn["training"]:
                    raise RuntimeError("backward through an eval-mode BatchNorm2d is not part of the hot path")
                ss, bm = bn["ss"], bn["mask"]
                sums = torch.zeros(2 * K, dtype=torch.float32, device=self.device)
                bm_ptr, bm_pitch = (ops.ptr(bm), bm.shape[1]) if bm is not None else (None, 0)
                ops.bn_bwd_reduce(g.ptr, g.code, g.pitch, y.ptr, y.code, y.pitch, pix, pps, K, bm_ptr, bm_pitch,
                                  ops.ptr(ss, 2 * K), ops.ptr(ss, 3 * K), sums.data_ptr())
                inv_world = 0.0
                if self.bn_sync is not None:
                    self.bn_sync.all_reduce(sums)
                    inv_world = 1.0 / self.bn_sync.world
                kw = dict(bn_sums=sums.data_ptr(), bn_mask=bm_ptr, bn_mask_pitch=bm_pitch, bn_inv_world=inv_world,
                          bn_gamma=ts[sp.bn + ".weight"].data_ptr(), bn_mean=ops.ptr(ss, 2 * K),
                          bn_invstd=ops.ptr(ss, 3 * K),
                          bn_dgamma=ops.ptr(grads[sp.bn + ".weight"]) if grads is not None else None,
                          bn_dbeta=ops.ptr(grads[sp.bn + ".bias"]) if grads is not None else None)
            om = rec["out_mask"]
            om_ptr, om_pitch = (om[1], om[2]) if om is not None else (None, 0)
            # dPre in the compute dtype; reuse g's storage when it is ours, same dtype and same pitch
            # (pitch is always a multiple of 8 so that single-channel gradients stay TMA-addressable)
            if inplace_ok and g.t.dtype == self.dt and g.off == 0 and g.pitch == pad8(le.Cout):
                dpre = Act(g.t, le.Cout)
            else:
                dpre = Act(torch.empty((pix, pad8(le.Cout)), dtype=self.dt, device=self.device), le.Cout)
            dpitch = dpre.pitch * pps if sp.kind == "linear" else dpre.pitch
            if grads is not None:
                if le.perm_bias is None:
                    dbias_t = grads[sp.key + ".bias"]
                else:
                    dbias_t = torch.zeros(Kfull, dtype=torch.float32, device=self.device)
                dbias = dbias_t.data_ptr()
            else:
                dbias_t, dbias = None, None
            ops.act_backward(g.ptr, g.code, gpitch, y.ptr, y.code, ypitch, dpre.ptr, dpre.code, dpitch, a_pix, a_pps,
                             a_C, sp.act, sp.slope, out_mask=om_ptr, mask_pitch=om_pitch, dbias=dbias, **kw)
            if grads is not None:
                if le.perm_bias is not None:
                    ops.unpack(dbias_t.data_ptr(), grads[sp.key + ".bias"].data_ptr(), le.perm_bias)
                le.wgrad(N, dpre, x, None, self.wgrad_scratch(le))
                fs = rec.get("in_fold")
                if fs is not None:                 # gradient was taken against the un-normalised input: dW = G*scale + shift*dbias
                    ops.bn_fold_wgrad(self.wgrad_scratch(le).data_ptr(), le.Kout, le.taps, le.Cin, ops.ptr(fs),
                                      ops.ptr(fs, le.Cin), dbias)
            if i > 0 or need_dx:
                dx = Act(torch.empty((N * le.Hin * le.Win, x.pitch), dtype=self.dt, device=self.device), le.Cin)
                le.dgrad(N, dpre, dx)
                g = dx
            else:
                g = None
            inplace_ok = True
        return g

    # ---- attribute plumbing ---------------------------------------------------------------------
    def _attr_inputs(self, c: Dict[str, torch.Tensor], N):
        """-> (onehot float tensors per categorical attr, continuous float tensors (N,))."""
        cats, conts = [], []
        for (name, K, _, _) in self.fam.cat_attrs:
            if name not in c:
                raise KeyError(f"attribute {name!r} missing from the attribute dict")
            t = c[name]
            ops.require_cuda(t)
            if t.shape != (N, K):
                raise ValueError(f"attribute {name!r}: expected shape {(N, K)}, got {tuple(t.shape)}")
            cats.append(t)
        for name in self.cont_names(c):
            t = c[name]
            ops.require_cuda(t)
            if t.numel() != N:
                raise ValueError(f"attribute {name!r}: expected {N} values, got shape {tuple(t.shape)}")
            conts.append(t.detach().reshape(N).float().contiguous())
        return cats, conts

    def cont_names(self, c):
        """Continuous attribute order.  MorphoMNIST treats every key != 'digit' as continuous, in sorted order
        (mnist.py:47-55); the other families name theirs explicitly."""
        if self.fam.name == "mnist":
            names = sorted(k for k in c if k != "digit")
            if len(names) != self.n_cont:
                raise ValueError(f"expected {self.n_cont} continuous attributes, got {names}")
            return names
        return list(self.fam.cont_attrs)

    def _argmax(self, t):
        """argmax(1) of a one-hot attribute; inside one train step the same attribute tensors feed two E and six D forwards,
        so the trainer lends a per-step cache (never kept across steps: a replayed CUDA graph must recompute it)."""
        cache = self.idx_cache
        if cache is None:
            return ops.argmax_rows(t.detach())
        key = (t.data_ptr(), tuple(t.shape), t.dtype)
        if key not in cache:
            cache[key] = ops.argmax_rows(t.detach())
        return cache[key]

    def _image_feats(self, N, x_ptr, x_code, x_pitch, c, mask):
        ts = self.tensors()
        cats, conts = self._attr_inputs(c, N)
        idx = [self._argmax(t) for t in cats]
        tables = [ts[a[self.emb_key]] for a in self.fam.cat_attrs]
        first = self.towers["E" if self.role == "E" else "Dx"].layers[0]
        fpad = first.pad if first.fold else 0            # folded first conv reads a zero-bordered tensor
        Hp, Wp = self.H + 2 * fpad, self.W + 2 * fpad
        rows = N * Hp * Wp
        feat = Act(torch.empty((rows + (Wp if first.fold else 0), pad8(self.feat_ch)), dtype=self.dt,
                               device=self.device), self.feat_ch)
        if first.fold:
            feat.t[rows:].zero_()                          # slack the folded rows of the last pixels run into
        ops.image_features(self.code, N, self.H, self.W, feat.pitch, x_ptr, x_code, x_pitch,
                           [t.data_ptr() for t in tables], [t.data_ptr() for t in idx],
                           [t.data_ptr() for t in conts], ops.ptr(mask), mask.shape[1] if mask is not None else 0,
                           feat.ptr, pad=fpad)
        return feat, {"idx": idx, "conts": conts, "mask": mask}

    def _image_feats_bwd(self, N, fstate, dfeat: Act, grads):
        ts = self.tensors()
        tables = [ts[a[self.emb_key]] for a in self.fam.cat_attrs]
        dtables = [grads[a[self.emb_key]] for a in self.fam.cat_attrs]
        mask = fstate["mask"]
        ops.image_features(self.code, N, self.H, self.W, dfeat.pitch, None, F32, 1,
                           [t.data_ptr() for t in tables], [t.data_ptr() for t in fstate["idx"]],
                           [t.data_ptr() for t in fstate["conts"]], ops.ptr(mask),
                           mask.shape[1] if mask is not None else 0, None, dfeat=dfeat.ptr,
                           dtables=[t.data_ptr() for t in dtables], backward=True)

    def _dX_from_dfeat(self, N, dfeat: Act, mask):
        """Gradient w.r.t. the image channel: dfeat[..., 0] * mask[n, 0] -> fp32 [N*H*W, 1]."""
        pix = N * self.H * self.W
        dX = torch.empty((pix, 1), dtype=torch.float32, device=self.device)
        ops.scale_shift_mask(dfeat.ptr, dfeat.code, dfeat.pitch, dX.data_ptr(), F32, 1, pix, self.H * self.W, 1,
                             mask=ops.ptr(mask), mask_pitch=mask.shape[1] if mask is not None else 0)
        return dX

    # ---- Encoder --------------------------------------------------------------------------------
    def encoder_forward(self, N, x_ptr, x_code, x_pitch, c, save=True):
        self.repack()
        feat, fstate = self._image_feats(N, x_ptr, x_code, x_pitch, c, None)
        out, saved = self._tower_fwd("E", N, feat, {}, True, save)
        st = {"N": N, "feat": fstate, "tower": saved}
        if self.keep_state:
            self.last_state = st
        return out, st

    def encoder_backward(self, st, dout: Act, grads, need_dX=False):
        if grads is not None:
            self.wgrad_begin()
        r = self._encoder_backward(st, dout, grads, need_dX)
        if grads is not None:
            self.wgrad_end(grads)
        return r

    def _encoder_backward(self, st, dout: Act, grads, need_dX=False):
        N = st["N"]
        need_feat = need_dX or (grads is not None and self.n_emb > 0)
        self.towers["E"].layers[0].head_active = not need_dX          # only the attribute-plane channels of dfeat are read
        dfeat = self._tower_bwd("E", N, st["tower"], dout, grads, need_feat, inplace_ok=False)
        if dfeat is None:
            return None
        if grads is not None and self.n_emb > 0:
            self._image_feats_bwd(N, st["feat"], dfeat, grads)
        return self._dX_from_dfeat(N, dfeat, None) if need_dX else None

    # ---- Generator ------------------------------------------------------------------------------
    def generator_forward(self, N, z_ptr, z_code, z_pitch, c, save=True):
        self.repack()
        ts = self.tensors()
        cats, conts = self._attr_inputs(c, N)
        onehots = [t.detach().float().contiguous() for t in cats]
        tables = [ts[a[self.emb_key]] for a in self.fam.cat_attrs]
        lat = Act(torch.empty((N, pad8(self.lat_dim)), dtype=self.dt, device=self.device), self.lat_dim)
        ops.latent_features(self.code, N, self.fam.latent, lat.pitch, z_ptr, z_code, z_pitch,
                            [a[1] for a in self.fam.cat_attrs], [t.data_ptr() for t in tables],
                            [t.data_ptr() for t in onehots], [t.data_ptr() for t in conts], lat.ptr)
        out, saved = self._tower_fwd("G", N, lat, {}, True, save)
        st = {"N": N, "onehots": onehots, "conts": conts, "tower": saved}
        if self.keep_state:
            self.last_state = st
        return out, st

    def generator_backward(self, st, dout: Act, grads, need_dz=False, need_dattr=False):
        """-> (dz fp32 [N,latent] | None, [d_onehot per cat attr] | None, [d_cont per cont attr] | None)."""
        if grads is not None:
            self.wgrad_begin()
        r = self._generator_backward(st, dout, grads, need_dz, need_dattr)
        if grads is not None:
            self.wgrad_end(grads)
        return r

    def _generator_backward(self, st, dout: Act, grads, need_dz=False, need_dattr=False):
        N = st["N"]
        need_lat = need_dz or need_dattr or (grads is not None and self.n_emb > 0)
        dlat = self._tower_bwd("G", N, st["tower"], dout, grads, need_lat, inplace_ok=False)
        if dlat is None:
            return None, None, None
        ts = self.tensors()
        tables = [ts[a[self.emb_key]] for a in self.fam.cat_attrs]
        dz = torch.empty((N, self.fam.latent), dtype=torch.float32, device=self.device) if need_dz else None
        doh = [torch.empty_like(t) for t in st["onehots"]] if need_dattr else None
        dco = [torch.empty_like(t) for t in st["conts"]] if need_dattr else None
        dtab = [grads[a[self.emb_key]].data_ptr() for a in self.fam.cat_attrs] if grads is not None else []
        ops.latent_features(self.code, N, self.fam.latent, dlat.pitch, None, F32, 0,
                            [a[1] for a in self.fam.cat_attrs], [t.data_ptr() for t in tables],
                            [t.data_ptr() for t in st["onehots"]], [t.data_ptr() for t in st["conts"]], None,
                            dfeat=dlat.ptr, dz=ops.ptr(dz), dtables=dtab,
                            donehots=[t.data_ptr() for t in doh] if doh else [],
                            dconts=[t.data_ptr() for t in dco] if dco else [], backward=True)
        return dz, doh, dco

    # ---- Discriminator --------------------------------------------------------------------------
    def discriminator_forward(self, N, x_ptr, x_code, x_pitch, z_ptr, z_code, z_pitch, c, masks=None,
                              training=True, save=True):
        """masks: list of (N,C,1,1)/(N,C) fp32 tensors in RNG order (draw_masks) or None to draw them here."""
        self.repack()
        fam = self.fam
        per = {"Dx": {}, "Dz": {}, "Dxz": {}}
        in_masks = {}
        if training and self.sites:
            if masks is None:
                masks = draw_masks(fam, N, self.device)
            if len(masks) != len(self.sites):
                raise ValueError(f"expected {len(self.sites)} dropout masks, got {len(masks)}")
            for (tname, i, where, _, ch), m in zip(self.sites, masks):
                m = m.reshape(N, ch)
                if m.dtype != torch.float32 or not m.is_contiguous():
                    m = m.float().contiguous()
                if where == "in":
                    in_masks[tname] = m
                else:
                    per[tname][(i, where)] = m
        feat, fstate = self._image_feats(N, x_ptr, x_code, x_pitch, c, in_masks.get("Dx"))
        kx, kz = fam.Dx[-1].cout, fam.Dz[-1].cout
        cat = torch.empty((N, kx + kz), dtype=self.dt, device=self.device)
        cm = in_masks.get("Dxz")
        _, sx = self._tower_fwd("Dx", N, feat, per["Dx"], training, save, final=Act(cat, kx, 0),
                                final_mask=(cm, 0) if cm is not None else None)
        zm = in_masks.get("Dz")
        zin = Act(torch.empty((N, fam.latent), dtype=self.dt, device=self.device), fam.latent)
        ops.scale_shift_mask(z_ptr, z_code, z_pitch, zin.ptr, zin.code, zin.pitch, N, 1, fam.latent,
                             mask=ops.ptr(zm), mask_pitch=zm.shape[1] if zm is not None else 0)
        _, sz = self._tower_fwd("Dz", N, zin, per["Dz"], training, save, final=Act(cat, kz, kx),
                                final_mask=(cm, kx) if cm is not None else None)
        logits, sxz = self._tower_fwd("Dxz", N, Act(cat, kx + kz), per["Dxz"], training, save, final_f32=True)
        st = {"N": N, "feat": fstate, "Dx": sx, "Dz": sz, "Dxz": sxz, "zmask": zm, "cat": cat}
        if self.keep_state:
            self.last_state = st
        return logits, st

    def discriminator_backward(self, st, dlogits: Act, grads, need_dX=False, need_dz=False):
        """-> (dX fp32 [N*H*W,1] | None, dz fp32 [N,latent] | None)."""
        if grads is not None:
            self.wgrad_begin()
        r = self._discriminator_backward(st, dlogits, grads, need_dX, need_dz)
        if grads is not None:
            self.wgrad_end(grads)
        return r

    def _discriminator_backward(self, st, dlogits: Act, grads, need_dX=False, need_dz=False):
        N = st["N"]
        fam = self.fam
        kx, kz = fam.Dx[-1].cout, fam.Dz[-1].cout
        dcat = self._tower_bwd("Dxz", N, st["Dxz"], dlogits, grads, True, inplace_ok=False)
        dz = None
        if need_dz or grads is not None:
            dzin = self._tower_bwd("Dz", N, st["Dz"], Act(dcat.t, kz, kx), grads, need_dz, inplace_ok=False)
            if need_dz:
                zm = st["zmask"]
                dz = torch.empty((N, fam.latent), dtype=torch.float32, device=self.device)
                ops.scale_shift_mask(dzin.ptr, dzin.code, dzin.pitch, dz.data_ptr(), F32, fam.latent, N, 1,
                                     fam.latent, mask=ops.ptr(zm), mask_pitch=zm.shape[1] if zm is not None else 0)
        dX = None
        if need_dX or grads is not None:
            need_feat = need_dX or (grads is not None and self.n_emb > 0)
            self.towers["Dx"].layers[0].head_active = not need_dX
            dfeat = self._tower_bwd("Dx", N, st["Dx"], Act(dcat.t, kx, 0), grads, need_feat, inplace_ok=False)
            if dfeat is not None:
                if grads is not None and self.n_emb > 0:
                    self._image_feats_bwd(N, st["feat"], dfeat, grads)
                if need_dX:
                    dX = self._dX_from_dfeat(N, dfeat, st["feat"]["mask"])
        return dX, dz
