"""The quantisation-time flow of one Linear on the CPU (oracle; test infrastructure only, see oracle/__init__.py):
Hessian accumulation -> diagonal rescale -> random orthogonal projection -> GPTQ damping -> grid parameters ->
LDLQ rounding.  SURVEY section 8(f) ranks 1-2: what a GPU quantiser will have to reproduce; restated here so that it has
a checker.  Reference: method.py:98-123 (add_batch / post_batch), :139-156 (rescale), :157-180 (projection),
:182-192 (damping), bal.py:15-45 (Balance.fasterquant), vector_balance.py:500-530 (quantize_weight_vecbal).

torch on the CPU is used for the arithmetic so that float32 / float64 reductions follow the reference's own library
calls; integer results (codes) are compared bit-exactly, the floating-point preprocessing with a tolerance.
"""
import torch

from . import ldlq


def accumulate_hessian(batches):
    """sum over batches of X^T X in float64; `nsamples` counts BATCHES, not tokens (method.py:105,118: tmp = inp.shape[0]
    of the (batch, tokens, features) input)."""
    H, n = None, 0
    for X in batches:
        X3 = X if X.dim() == 3 else X.unsqueeze(0)
        n += X3.shape[0]
        flat = X3.reshape(-1, X3.shape[-1]).to(torch.float64)
        G = flat.T @ flat
        H = G if H is None else H + G
    return H, n


def finalize_hessian(H, n):
    return (H / n).to(torch.float32)                        # method.py:122-123


def rescale(w, H):
    """method.py:139-156: scaleWH = (diag(H/max|H|) / diag(W^T W))^(1/4), clamps at 1e-8; W <- W diag(s), H <- D^-1 H D^-1."""
    w = w.to(torch.float32).clone()
    H = H.to(torch.float32).clone()
    H = H / H.abs().max()
    dH = torch.clamp(torch.diag(H), min=1e-8)
    dW = torch.clamp(torch.diag(w.T @ w), min=1e-8)
    s = (dH / dW).sqrt().sqrt().to(torch.float32).clamp(min=1e-8)
    w = w * s[None, :]
    H = H / s[None, :]
    H = H / s[:, None]
    return w, H, s


def project(w, H, U, V):
    """method.py:157-180: H <- H n/tr(H) + 1e-2 I, then W <- U W V^T, H <- V H V^T (U, V dense float32 here)."""
    w = w.to(torch.float32)
    H = H.to(torch.float32)
    n = H.shape[0]
    H = H * (n / (torch.trace(H) + 1e-8)) + 1e-2 * torch.eye(n)
    return U @ w @ V.T, V @ H @ V.T


def gptq_damp(w, H, percdamp=0.01):
    """method.py:182-192: dead columns (H_ii == 0) are zeroed in W and set to 1 in H; diag(H) += percdamp mean(diag H)."""
    w, H = w.clone(), H.clone()
    dead = torch.diag(H) == 0
    H[dead, dead] = 1
    w[:, dead] = 0
    H[torch.arange(H.shape[0]), torch.arange(H.shape[0])] += percdamp * torch.mean(torch.diag(H))
    return w, H


def find_params(w, bits):
    """Quantizer.find_params for qfn 'a', per output channel, asymmetric (quant.py:57-127) -- torch version of
    oracle.qmath.find_params_qfna."""
    maxq = float(2 ** bits - 1)
    x = w.flatten(1)
    zero_ = torch.zeros(x.shape[0])
    xmin = torch.minimum(x.min(1)[0].float(), zero_)
    xmax = torch.maximum(x.max(1)[0].float(), zero_)
    dead = (xmin == 0) & (xmax == 0)
    xmin[dead], xmax[dead] = -1, 1
    scale = (xmax - xmin) / maxq
    zero = torch.round(-xmin / scale)
    return scale.reshape(-1, 1), zero.reshape(-1, 1)


def preprocess(W0, batches, rescale_on, U=None, V=None, gptqH=True, percdamp=0.01):
    """W0 (N, K) in the layer's dtype, calibration batches -> (w, H, scaleWH) as Balance.fasterquant sees them: after every
    preprocessing step the weight is stored back in the layer's dtype (method.py:155,178,191)."""
    H, n = accumulate_hessian(batches)
    H = finalize_hessian(H, n)
    w, s = W0.clone(), None
    if rescale_on:
        w32, H, s = rescale(w, H)
        w = w32.to(W0.dtype)
    if U is not None:
        w32, H = project(w, H, U, V)
        w = w32.to(W0.dtype)
    if gptqH:
        w, H = gptq_damp(w, H, percdamp)
    return w, H, s


def quantize(w, H, bits, method, greedy_passes, qfn):
    """Balance.fasterquant's call (bal.py:26-43): returns (grid fp16, codes uint8, scale, zero)."""
    maxq = float(2 ** bits - 1)
    scale, zero = (find_params(w, bits) if qfn == 'a' else (None, None))
    grid, codes = ldlq.quantize_weight(w, H, bits, greedy_passes, scale, zero, maxq, qfn=qfn, method=method)
    return grid, codes, scale, zero
