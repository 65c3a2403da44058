"""Development aid: run each tensor-core conv case in its own process (a faulting kernel kills the CUDA context)."""
import subprocess
import sys

CASES = [  # kind, N, C, H, K, k, stride, pad, opad
    ("conv", 128, 64, 1, 64, 1, 1, 0, 0), ("conv", 256, 128, 1, 256, 1, 1, 0, 0), ("conv", 300, 1024, 1, 1024, 1, 1, 0, 0),
    ("conv", 2, 64, 14, 128, 4, 2, 1, 0), ("conv", 4, 64, 8, 64, 3, 1, 1, 0), ("conv", 3, 128, 7, 256, 4, 2, 1, 0),
    ("conv", 20, 256, 3, 512, 4, 2, 1, 0), ("conv", 2, 32, 24, 64, 4, 2, 0, 0), ("conv", 2, 64, 31, 128, 5, 2, 1, 0),
    ("convT", 3, 771, 1, 512, 3, 1, 0, 0), ("convT", 2, 512, 3, 256, 3, 2, 0, 0), ("convT", 2, 256, 7, 128, 3, 2, 1, 0),
    ("convT", 2, 128, 13, 64, 3, 2, 1, 0), ("convT", 2, 64, 4, 32, 5, 2, 2, 1), ("convT", 2, 1024, 4, 512, 5, 2, 2, 1),
    ("conv", 3, 5, 28, 32, 5, 1, 0, 0), ("conv", 3, 5, 28, 64, 3, 2, 1, 0), ("conv", 2, 7, 64, 64, 5, 2, 1, 0),
    ("convT", 3, 64, 25, 1, 4, 1, 0, 0), ("convT", 2, 64, 16, 1, 5, 2, 2, 1), ("convT", 3, 32, 24, 5, 5, 1, 0, 0),
    ("conv", 3, 1, 28, 64, 4, 1, 0, 0), ("conv", 200, 1024, 1, 1, 1, 1, 0, 0),
]


WG_CASES = [  # kind, N, C(in), H, K(out), k, stride, pad, opad
    ("conv", 64, 64, 1, 128, 1, 1, 0, 0), ("conv", 130, 512, 1, 512, 1, 1, 0, 0), ("conv", 4, 64, 14, 128, 4, 2, 1, 0),
    ("conv", 5, 128, 7, 256, 4, 2, 1, 0), ("conv", 9, 256, 3, 512, 4, 2, 1, 0), ("conv", 4, 32, 24, 64, 4, 2, 0, 0),
    ("conv", 3, 64, 11, 128, 4, 1, 0, 0), ("conv", 2, 64, 31, 128, 5, 2, 1, 0),
    ("convT", 5, 771, 1, 512, 3, 1, 0, 0), ("convT", 3, 512, 3, 256, 3, 2, 0, 0), ("convT", 2, 256, 7, 128, 3, 2, 1, 0),
    ("convT", 2, 128, 13, 64, 3, 2, 1, 0), ("convT", 2, 1024, 4, 512, 5, 2, 2, 1),
    ("conv", 3, 5, 28, 32, 5, 1, 0, 0), ("conv", 3, 5, 28, 64, 3, 2, 1, 0), ("convT", 3, 64, 25, 1, 4, 1, 0, 0),
    ("conv", 70, 1024, 1, 1, 1, 1, 0, 0), ("conv", 2, 7, 64, 64, 5, 2, 1, 0),
]


def run_wg(i):
    import torch
    import torch.nn.functional as F
    sys.path.insert(0, "imagecfgen-pytorch_b200")
    from icf_b200 import ops
    kind, N, C, H, K, k, s, p, op = WG_CASES[i]
    g = torch.Generator().manual_seed(100 + i)
    x = torch.randn(N, C, H, H, generator=g).bfloat16().float()
    T = k * k
    dev = "cuda"
    cp, kp = (C + 7) // 8 * 8, (K + 7) // 8 * 8
    if kind == "conv":
        P = (H + 2 * p - k) // s + 1
        dy = torch.randn(N, K, P, P, generator=g).bfloat16().float()
        w = torch.zeros(K, C, k, k, dtype=torch.float64, requires_grad=True)
        F.conv2d(x.double(), w, None, s, p).backward(dy.double())
    else:
        P = (H - 1) * s - 2 * p + k + op
        dy = torch.randn(N, K, P, P, generator=g).bfloat16().float()
        w = torch.zeros(C, K, k, k, dtype=torch.float64, requires_grad=True)
        F.conv_transpose2d(x.double(), w, None, s, p, op).backward(dy.double())

    def nhwc(t, pitch):
        n, c, h, _ = t.shape
        o = torch.zeros(n * h * h, pitch, dtype=torch.bfloat16, device=dev)
        o[:, :c] = t.permute(0, 2, 3, 1).reshape(-1, c).to(dev).bfloat16()
        return o
    xt, dyt = nhwc(x, cp), nhwc(dy, kp)
    if kind == "conv":      # small = dY (A=K, grid P), big = X (B=C, grid H)
        dwp = torch.zeros(K * T * C, dtype=torch.float32, device=dev)
        ops.conv_wgrad(1, N, P, P, K, kp, H, H, C, cp, k, k, s, p, dyt.data_ptr(), xt.data_ptr(), dwp.data_ptr())
        dw = torch.empty(K, C, k, k, dtype=torch.float32, device=dev)
        ops.unpack(dwp.data_ptr(), dw.data_ptr(), ops.make_perm(K, T, C, C * T, 1, T))
    else:                   # small = X (A=C, grid H), big = dY (B=K, grid P)
        dwp = torch.zeros(C * T * K, dtype=torch.float32, device=dev)
        ops.conv_wgrad(1, N, H, H, C, cp, P, P, K, kp, k, k, s, p, xt.data_ptr(), dyt.data_ptr(), dwp.data_ptr())
        dw = torch.empty(C, K, k, k, dtype=torch.float32, device=dev)
        ops.unpack(dwp.data_ptr(), dw.data_ptr(), ops.make_perm(C, T, K, K * T, 1, T))
    torch.cuda.synchronize()
    got, ref = dw.double().cpu(), w.grad
    err = float((got - ref).norm() / ref.norm())
    bad = (got - ref).abs() > 0.02 * ref.abs().max()
    print(f"wgrad case {i} {WG_CASES[i]}: rel err {err:.3e}; bad elems {int(bad.sum())}/{bad.numel()}", flush=True)
    if bad.any():
        print("   first bad idx:", bad.nonzero()[:6].tolist(), "got", got[bad][:4].tolist(), "ref", ref[bad][:4].tolist())


def run_case(i):
    import torch
    import torch.nn.functional as F
    sys.path.insert(0, "imagecfgen-pytorch_b200")
    from icf_b200 import ops
    kind, N, C, H, K, k, s, p, op = CASES[i]
    g = torch.Generator().manual_seed(i)
    x = torch.randn(N, C, H, H, generator=g).bfloat16().float()
    b = torch.randn(K, generator=g)
    cp, kp = (C + 7) // 8 * 8, (K + 7) // 8 * 8
    T = k * k
    if kind == "conv":
        w = (torch.randn(K, C, k, k, generator=g) / (C * T) ** 0.5).bfloat16().float()
        ref = F.conv2d(x.double(), w.double(), b.double(), s, p)
        perm = ops.make_perm(K, T, C, C * T, 1, T, d2_pad=cp)
        form = ops.GATHER
    else:
        w = (torch.randn(C, K, k, k, generator=g) / (C * T / s / s) ** 0.5).bfloat16().float()
        ref = F.conv_transpose2d(x.double(), w.double(), b.double(), s, p, op)
        perm = ops.make_perm(K, T, C, T, 1, K * T, d2_pad=cp)
        form = ops.TRANSPOSED
    P = ref.shape[-1]
    dev = "cuda"
    xt = torch.zeros(N * H * H, cp, dtype=torch.bfloat16, device=dev)
    xt[:, :C] = x.permute(0, 2, 3, 1).reshape(-1, C).to(dev).bfloat16()
    wsrc = w.contiguous().to(dev)
    wt = torch.empty(K * T * cp, dtype=torch.bfloat16, device=dev)
    ops.pack(wsrc.data_ptr(), wt.data_ptr(), 1, perm)
    y = torch.zeros(N * P * P, kp, dtype=torch.bfloat16, device=dev)
    bias = b.to(dev)
    ops.conv_forward(1, form, N, H, H, C, cp, P, P, K, kp, k, k, s, p, xt.data_ptr(), wt.data_ptr(), K, cp, y.data_ptr(),
                     bias=bias.data_ptr())
    torch.cuda.synchronize()
    got = y[:, :K].float().cpu().reshape(N, P, P, K).permute(0, 3, 1, 2).double()
    err = float((got - ref).norm() / ref.norm())
    bad = (got - ref).abs() > 0.05 * ref.abs().max()
    print(f"case {i} {CASES[i]}: rel err {err:.3e}; bad elems {int(bad.sum())}/{bad.numel()}", flush=True)
    if bad.any():
        idx = bad.nonzero()
        print("   first bad (n,k,p,q):", idx[:6].tolist(), "got", got[bad][:4].tolist(), "ref", ref[bad][:4].tolist(),
              "bad per n", bad.sum((1, 2, 3)).tolist()[:8], "bad per k (first 16)", bad.sum((0, 2, 3)).tolist()[:16])


if __name__ == "__main__":
    if len(sys.argv) > 2:
        (run_wg if sys.argv[1] == "wg" else run_case)(int(sys.argv[2]))
    else:
        which = sys.argv[1] if len(sys.argv) > 1 else "wg"
        for i in range(len(WG_CASES if which == "wg" else CASES)):
            r = subprocess.run([sys.executable, __file__, which, str(i)], capture_output=True, text=True, timeout=300)
            out = (r.stdout + r.stderr).strip().splitlines()
            print("\n".join(l for l in out if l.startswith(("case", "wgrad", "   ", "icf")) or "rror" in l)[:1500] or f"case {i}: no output rc={r.returncode}", flush=True)
