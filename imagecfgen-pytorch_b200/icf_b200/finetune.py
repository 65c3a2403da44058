"""Fused encoder fine-tune step: the loop body of finetune_mnist_bigan.py:68-86,
finetune_audio_mnist_bigan.py:79-92 and finetune_whale_bigan.py:58-73.

    codes = E(x, a);  xr = G(codes, a);  loss = rec(x, xr) + mean(codes^2);  loss.backward();  Adam(E).step()

The reference back-propagates through G with autograd, which also computes (and accumulates forever, only
``opt.zero_grad()`` over E is called) every weight gradient of G.  Here G runs its data-gradient kernels only; the
reconstruction loss and its gradient are one pass over the reconstruction, the latent penalty is folded into the latent
gradient, and Adam over E's flat parameter buffer is one kernel.  3 F_E + 2 F_G of tensor work instead of 3 F_E + 3 F_G.
"""
from typing import Dict, Optional

import torch

from . import ops
from .dp import Group
from .engine import Act
from .trainer import _FlatGroup

F32 = ops.F32


class EncoderFineTuner:
    def __init__(self, E, G, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, dtype=None, metric="mse", all_pairs=False,
                 process_group=None):
        """``metric``: 'mse' (fused) or 'ssim' (1 - SSIM through torch on the reconstruction, then the fused backward).
        ``all_pairs``: mirror finetune_whale_bigan.py:59-65, where x (N,H,W) minus xr (N,1,H,W) broadcasts to all
        (image, reconstruction) pairs — the loss is then the MSE against the batch-mean image plus the mean pixel variance."""
        if metric not in ("mse", "ssim"):
            raise ValueError(f"Invalid metric {metric}")
        self.E, self.G, self.metric, self.all_pairs = E, G, metric, all_pairs
        if dtype is not None:
            E.set_compute_dtype(dtype)
            G.set_compute_dtype(dtype)
        self.exE, self.exG = E.engine(), G.engine()
        self.fam, self.device = self.exE.fam, self.exE.device
        self.gE = _FlatGroup([("E." + n, p) for n, p in E.named_parameters()], self.device)
        self.gradsE = {n[2:]: v for n, v in self.gE.grad_views.items()}
        self.group = Group(process_group)
        self.world = self.group.world
        if self.world > 1:
            self.group.broadcast_state([self.gE.flat])
        self.state = torch.tensor([0, lr, betas[0], betas[1], eps, 1.0 / self.world, 0, 0], dtype=torch.float32,
                                  device=self.device)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.exE.repack(force=True)
        self.exG.repack(force=True)

    def step(self, x: torch.Tensor, c: Dict[str, torch.Tensor], out: Optional[torch.Tensor] = None):
        """x: images already scaled to [-1,1], any of (N,H,W) / (N,1,H,W) / (N,H*W); returns a device tensor
        [rec_loss, latent_loss] (accumulated into ``out``)."""
        exE, exG, fam = self.exE, self.exG, self.fam
        ops.require_cuda(x)
        xx = x.contiguous().float()
        P = exE.H * exE.W
        N = xx.numel() // P
        if out is None:
            out = torch.zeros(2, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            ops.fill_f32(self.gE.grad.data_ptr(), 0.0, self.gE.n)
            zE, stE = exE.encoder_forward(N, xx.data_ptr(), F32, 1, c, save=True)
            xr, stG = exG.generator_forward(N, zE.ptr, zE.code, zE.pitch, c, save=True)
            dxr = torch.empty((N * P, 1), dtype=torch.float32, device=self.device)
            if self.metric == "mse":
                if self.all_pairs:
                    xbar = torch.empty(P, dtype=torch.float32, device=self.device)
                    var = torch.zeros(1, dtype=torch.float32, device=self.device)
                    ops.col_mean(xx.data_ptr(), N, P, xbar.data_ptr(), var.data_ptr())
                    ops.mse_loss(xbar.data_ptr(), 0, xr.ptr, xr.code, xr.pitch, N, P, 1.0, var.data_ptr(), ops.ptr(out, 0),
                                 dxr.data_ptr(), F32, 1)
                else:
                    ops.mse_loss(xx.data_ptr(), P, xr.ptr, xr.code, xr.pitch, N, P, 1.0, None, ops.ptr(out, 0),
                                 dxr.data_ptr(), F32, 1)
            else:
                from image_scms.training_utils import ssim
                rec = torch.empty((N, 1, exE.H, exE.W), dtype=torch.float32, device=self.device)
                ops.cast(xr.ptr, xr.code, rec.data_ptr(), F32, rec.numel())
                rec.requires_grad_(True)
                loss = 1 - ssim(xx.reshape(N, 1, exE.H, exE.W), rec, data_range=1.0).mean()
                (g,) = torch.autograd.grad(loss, rec)
                dxr.copy_(g.reshape(N * P, 1))
                out[0] += loss.detach()
            dz, _, _ = exG.generator_backward(stG, Act(dxr, 1), None, need_dz=True)      # data gradients only: no wgrad of G
            ops.latent_l2(zE.ptr, zE.code, zE.pitch, N, fam.latent, 1.0, ops.ptr(out, 1), dz.data_ptr(), True)
            exE.encoder_backward(stE, Act(dz, fam.latent), self.gradsE)
            self.group.all_reduce(self.gE.grad)
            ops.adam_step(self.gE.flat.data_ptr(), self.gE.grad.data_ptr(), self.gE.exp_avg.data_ptr(),
                          self.gE.exp_avg_sq.data_ptr(), self.gE.n, self.state.data_ptr())
            exE.repack(force=True)
        return out

    def export_optimizer(self):
        """torch.optim.Adam over E.parameters() carrying the fused optimiser's state (the scripts save / reuse `opt`)."""
        opt = torch.optim.Adam(list(self.E.parameters()), lr=self.lr, betas=self.betas, eps=self.eps)
        steps = float(self.state[0].item())
        for p, off in zip(self.gE.params, self.gE.offsets):
            n = p.numel()
            opt.state[p] = {"step": torch.tensor(steps), "exp_avg": self.gE.exp_avg[off:off + n].view_as(p).clone(),
                            "exp_avg_sq": self.gE.exp_avg_sq[off:off + n].view_as(p).clone()}
        return opt
