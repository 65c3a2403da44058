"""Device-side attribute intervention — the "intervene" step between E and G of a counterfactual.

The reference runs it on the host with pyro (attribute_scms/graph.py:144-184 ``sample_cf``: abduct the exogenous noise of
the observed attributes, regenerate the descendants of the intervened ones; one-hots rebuilt with ``torch.eye(K)[idx]`` CPU
tensors, graph.py:154) and then rescales the attributes (mnist_gan_counterfactuals.py:57-68).  For MorphoMNIST the graph
has one edge, thickness -> intensity, whose mechanism is a conditional affine flow followed by a sigmoid and an affine map
(attribute_scms/mnist.py:28-33,48); thickness, slant and digit are roots (their abducted noise regenerates them unchanged).
``AffineSigmoidMechanism`` holds that mechanism's parameters and runs abduction + regeneration + rescaling as ONE kernel
(icf_scm_affine_cf), so a counterfactual batch never leaves the device:

    pipe = CounterfactualPipeline(E, G, stats, mechanism)
    x_cf = pipe(x_scaled, a_raw, delta_thickness=2.0)

Categorical interventions (mnist_bigan_score.py:83-91, audiomnist_cf_eval.py:82-83) use ``onehot_swap``.
"""
import math
from typing import Dict, Optional

import torch

from . import lib as _l
from . import ops


class AffineSigmoidMechanism:
    """child = lo + span * sigmoid(loc(parent) + exp(ls(parent)) * noise).

    ``hyper``: None for a closed form loc = a0 + a1*parent, ls = a2 (``closed``), or a dict with the weights of pyro's
    ConditionalAutoRegressiveNN(1, 1, [H]) as trained by attribute_scms.mnist.train: w1 (H,), b1 (H,) — the context column
    and bias of the first masked layer — w2 (2,H), b2 (2,) — rows (loc, log scale) of the second."""

    def __init__(self, lo: float, span: float, closed=(0.0, 0.0, 0.0), hyper: Optional[Dict[str, torch.Tensor]] = None,
                 clip=(-5.0, 3.0), device="cuda"):
        self.lo, self.span, self.closed, self.clip = float(lo), float(span), tuple(float(v) for v in closed), clip
        self.hyper = None
        if hyper is not None:
            self.hyper = {k: hyper[k].detach().to(device=device, dtype=torch.float32).contiguous() for k in ("w1", "b1", "w2", "b2")}
            H = self.hyper["w1"].numel()
            assert self.hyper["b1"].numel() == H and self.hyper["w2"].shape == (2, H) and self.hyper["b2"].numel() == 2

    @staticmethod
    def morphomnist_ground_truth(device="cuda"):
        """The data-generating mechanism (create_train_dataset.py:42-46): intensity = 191*sigmoid(0.5*eps + 2t - 5) + 64."""
        return AffineSigmoidMechanism(64.0, 191.0, closed=(-5.0, 2.0, math.log(0.5)), device=device)

    def counterfactual(self, value: torch.Tensor, parent: torch.Tensor, parent_cf: Optional[torch.Tensor] = None,
                       parent_shift: float = 0.0, value_stats=None, parent_stats=None, want_noise=False):
        """-> dict(value_cf, parent_cf [, value_cf_scaled, parent_cf_scaled, noise]); all (N,1) fp32 on the device."""
        ops.require_cuda(value, parent, parent_cf)
        n = value.numel()
        v = value.detach().reshape(n).float().contiguous()
        p = parent.detach().reshape(n).float().contiguous()
        pc = parent_cf.detach().reshape(n).float().contiguous() if parent_cf is not None else None
        dev = v.device
        new = lambda: torch.empty((n, 1), dtype=torch.float32, device=dev)
        out = {"value_cf": new(), "parent_cf": new()}
        a = _l.ScmAffineArgs()
        a.n, a.hidden = n, 0 if self.hyper is None else self.hyper["w1"].numel()
        for i, cv in enumerate(self.closed):
            a.closed[i] = cv
        a.clip_lo, a.clip_hi, a.lo, a.span = self.clip[0], self.clip[1], self.lo, self.span
        fi = torch.finfo(torch.float32)
        a.u_min, a.u_max, a.parent_shift = fi.tiny, 1.0 - fi.eps, parent_shift      # torch SigmoidTransform._inverse clamp
        if self.hyper is not None:
            a.w1, a.b1, a.w2, a.b2 = (self.hyper[k].data_ptr() for k in ("w1", "b1", "w2", "b2"))
        a.value, a.parent, a.parent_cf = v.data_ptr(), p.data_ptr(), ops.ptr(pc)
        a.value_cf, a.parent_cf_out = out["value_cf"].data_ptr(), out["parent_cf"].data_ptr()
        if want_noise:
            out["noise"] = new()
            a.noise_out = out["noise"].data_ptr()
        if value_stats is not None and parent_stats is not None:
            a.v_min, a.v_max = float(value_stats[0]), float(value_stats[1])
            a.p_min, a.p_max = float(parent_stats[0]), float(parent_stats[1])
            out["value_cf_scaled"], out["parent_cf_scaled"] = new(), new()
            a.value_cf_scaled, a.parent_cf_scaled = out["value_cf_scaled"].data_ptr(), out["parent_cf_scaled"].data_ptr()
        with torch.cuda.device(dev):
            ops.scm_affine_cf(a)
        return out


def onehot_swap(onehot: torch.Tensor, new_index: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Copy of ``onehot`` (N,K) whose rows selected by ``mask`` (all rows if None) become one_hot(new_index): the device form of
    ``torch.eye(K)[idx]`` + masked assignment (mnist_bigan_score.py:83-91, audiomnist_cf_eval.py:82-83)."""
    rows = onehot.detach().float().contiguous().clone()
    with torch.cuda.device(rows.device):
        return ops.onehot_swap(new_index, mask, rows)


class CounterfactualPipeline:
    """encode -> intervene on the attributes -> decode, device-resident (mnist_gan_counterfactuals.py:57-73): the attribute
    SCM step, the rescale, E and G run back to back on one stream with the latent code kept in the engine's buffer;
    ``capture(n)`` records the whole pipeline for a fixed batch size into one CUDA graph (a single launch per batch)."""

    def __init__(self, E, G, stats: Dict[str, tuple], mechanism: Optional[AffineSigmoidMechanism] = None,
                 parent="thickness", child="intensity"):
        self.E, self.G, self.parent, self.child = E, G, parent, child
        self.stats = {k: (float(v[0]), float(v[1])) for k, v in stats.items()}
        self.mech = mechanism if mechanism is not None else AffineSigmoidMechanism.morphomnist_ground_truth(E.device)
        self.graph, self.static = None, None

    def attributes(self, a_raw: Dict[str, torch.Tensor], delta: float = 0.0, parent_cf: Optional[torch.Tensor] = None):
        """-> (c, c_cf): min-max scaled factual and counterfactual attribute dicts (categorical entries pass through)."""
        st = self.stats
        c = {k: (2 * (a_raw[k].float() - st[k][0]) / (st[k][1] - st[k][0]) - 1) if k in st else a_raw[k] for k in a_raw}
        r = self.mech.counterfactual(a_raw[self.child], a_raw[self.parent], parent_cf=parent_cf, parent_shift=delta,
                                     value_stats=st[self.child], parent_stats=st[self.parent])
        c_cf = dict(c)
        c_cf[self.parent], c_cf[self.child] = r["parent_cf_scaled"], r["value_cf_scaled"]
        return c, c_cf

    def __call__(self, x: torch.Tensor, a_raw: Dict[str, torch.Tensor], delta: float = 0.0, parent_cf=None,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        from .trainer import counterfactual
        c, c_cf = self.attributes(a_raw, delta, parent_cf)
        return counterfactual(self.E, self.G, x, c, c_cf, out=out)

    def capture(self, x: torch.Tensor, a_raw: Dict[str, torch.Tensor], delta: float):
        self.static = {"x": x.clone(), "a": {k: v.clone() for k, v in a_raw.items()}, "out": None}
        s = torch.cuda.Stream(device=x.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.static["out"] = self(self.static["x"], self.static["a"], delta)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self(self.static["x"], self.static["a"], delta, out=self.static["out"])
        self.graph = g
        return g

    def replay(self, x=None, a_raw=None):
        if x is not None:
            self.static["x"].copy_(x, non_blocking=True)
        if a_raw is not None:
            for k, v in a_raw.items():
                self.static["a"][k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.static["out"]
