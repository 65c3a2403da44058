"""TEST INFRASTRUCTURE ONLY — numpy restatement of the torch primitives the reference's hot path uses.

The reference's arithmetic lives in third-party ``torch`` (``/root/reference/requirements.txt:1``, unpinned;
torch 2.11.0 in this image).  ``oracle/bigan_ref.py`` executes those primitives with torch's CPU kernels;
this file restates their published definitions independently (float64 numpy, loops over kernel taps only) so
the oracle does not rest on torch alone.  ``tests/test_oracle_cpu.py`` checks each function against torch.

Call sites restated: nn.Conv2d (image_scms/mnist.py:31-39,100-135), nn.ConvTranspose2d (mnist.py:64-72),
nn.BatchNorm2d train mode (mnist.py:111-122), nn.LeakyReLU, nn.BCEWithLogitsLoss (mnist.py:181),
torch.optim.Adam (mnist.py:176-179), nn.Upsample(nearest) index map (mnist.py:27).
"""
import numpy as np


def conv2d(x, w, b, stride, pad):
    """x (N,C,H,W), w (K,C,R,S) -> (N,K,P,Q); out[n,k,p,q] = b[k] + sum x[n,c,p*s-pad+r,q*s-pad+t] w[k,c,r,t]."""
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    N, C, H, W = x.shape
    K, _, R, S = w.shape
    P = (H + 2 * pad - R) // stride + 1
    Q = (W + 2 * pad - S) // stride + 1
    xp = np.zeros((N, C, H + 2 * pad, W + 2 * pad))
    xp[:, :, pad:pad + H, pad:pad + W] = x
    out = np.zeros((N, K, P, Q))
    for r in range(R):
        for t in range(S):
            win = xp[:, :, r:r + stride * (P - 1) + 1:stride, t:t + stride * (Q - 1) + 1:stride]
            out += np.einsum("ncpq,kc->nkpq", win, w[:, :, r, t])
    if b is not None:
        out += np.asarray(b, np.float64).reshape(1, K, 1, 1)
    return out


def conv_transpose2d(x, w, b, stride, pad, out_pad=0):
    """x (N,C,H,W), w (C,K,R,S) -> (N,K,Ho,Wo), Ho=(H-1)s-2pad+R+out_pad; scatter form:
    out[n,k,i*s-pad+r,j*s-pad+t] += x[n,c,i,j] w[c,k,r,t]."""
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    N, C, H, W = x.shape
    _, K, R, S = w.shape
    Ho = (H - 1) * stride - 2 * pad + R + out_pad
    Wo = (W - 1) * stride - 2 * pad + S + out_pad
    full = np.zeros((N, K, (H - 1) * stride + R + out_pad, (W - 1) * stride + S + out_pad))
    for r in range(R):
        for t in range(S):
            contrib = np.einsum("nchw,ck->nkhw", x, w[:, :, r, t])
            full[:, :, r:r + stride * (H - 1) + 1:stride, t:t + stride * (W - 1) + 1:stride] += contrib
    out = full[:, :, pad:pad + Ho, pad:pad + Wo].copy()
    if b is not None:
        out += np.asarray(b, np.float64).reshape(1, K, 1, 1)
    return out


def leaky_relu(x, slope):
    x = np.asarray(x, np.float64)
    return np.where(x > 0, x, slope * x)


def batch_norm_train(x, gamma, beta, running_mean, running_var, momentum=0.1, eps=1e-5):
    """Train-mode BatchNorm2d: normalise with batch mean / biased var; running_var gets the unbiased one."""
    x = np.asarray(x, np.float64)
    m = x.shape[0] * x.shape[2] * x.shape[3]
    mean = x.mean(axis=(0, 2, 3))
    var = x.var(axis=(0, 2, 3))
    y = (x - mean.reshape(1, -1, 1, 1)) / np.sqrt(var.reshape(1, -1, 1, 1) + eps)
    y = y * np.asarray(gamma, np.float64).reshape(1, -1, 1, 1) + np.asarray(beta, np.float64).reshape(1, -1, 1, 1)
    new_rm = (1 - momentum) * np.asarray(running_mean, np.float64) + momentum * mean
    new_rv = (1 - momentum) * np.asarray(running_var, np.float64) + momentum * var * m / max(m - 1, 1)
    return y, new_rm, new_rv


def bce_with_logits_mean(l, t):
    l = np.asarray(l, np.float64)
    t = np.asarray(t, np.float64)
    return float(np.mean(np.maximum(l, 0) - l * t + np.log1p(np.exp(-np.abs(l)))))


def adam_update(p, g, m, v, step, lr, b1, b2, eps=1e-8):
    """One Adam step (no weight decay, no amsgrad); ``step`` is the 1-based step index. Returns (p, m, v)."""
    p, g, m, v = (np.asarray(a, np.float64) for a in (p, g, m, v))
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    denom = np.sqrt(v) / np.sqrt(1 - b2 ** step) + eps
    p = p - (lr / (1 - b1 ** step)) * m / denom
    return p, m, v


def nearest_index(dst_size, src_size):
    """torch 'nearest' upsample source index: floor(dst * src / dst_size)."""
    return np.minimum((np.arange(dst_size) * src_size) // dst_size, src_size - 1)


def embedding_plane(table, idx, out_size):
    """Embedding(K,256) -> (1,16,16) -> nearest upsample -> tanh."""
    e = np.asarray(table, np.float64)[np.asarray(idx)].reshape(-1, 16, 16)
    iy = nearest_index(out_size[0], 16)
    ix = nearest_index(out_size[1], 16)
    return np.tanh(e[:, iy][:, :, ix])[:, None]
