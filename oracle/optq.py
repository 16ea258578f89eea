"""TEST INFRASTRUCTURE -- OPTQ (GPTQ) rounding restated on the CPU, for the OPTQ == LDLQ equivalence check.

Follows the reference's GPTQ.fasterquant (gptq.py:25-100) with blocksize = all columns, groupsize = -1 and the
`debug_equiv` float64 path, i.e. the configuration of the reference's only self-check, optq_ldlq_equiv.py:16-66: columns
first to last, each rounded to the per-row grid (qfn 'c', quant.py:17-20: round(x/scale) + zero, clamped) and its error
spread over the later columns through the upper Cholesky factor of H^-1.  The reference version needs CUDA
(gptq.py:98 `torch.cuda.synchronize()`); nothing else in it is device-specific.
"""
import torch


def fake_layer(m, d, seed=0):
    """optq_ldlq_equiv.py:9-13: uniform random float64 weight, H = X^T X + 0.01 I."""
    g = torch.Generator().manual_seed(seed)
    w = torch.rand(m, d, dtype=torch.float64, generator=g)
    x = torch.rand(d, d, dtype=torch.float64, generator=g)
    return w, x.T @ x + 0.01 * torch.eye(d, dtype=torch.float64)


def row_grid(w, bits):
    """Quantizer.find_params(weight=True), perchannel, asymmetric (quant.py:57-127) -> scale, zero (m, 1)."""
    maxq = 2 ** bits - 1
    zero_ = torch.zeros(w.shape[0], dtype=w.dtype)
    lo, hi = torch.minimum(w.min(1)[0], zero_), torch.maximum(w.max(1)[0], zero_)
    flat = (lo == 0) & (hi == 0)
    lo[flat], hi[flat] = -1, 1
    scale = ((hi - lo) / maxq).unsqueeze(1)
    return scale, torch.round(-lo.unsqueeze(1) / scale)


def optq_codes(w, H, bits):
    """Integer codes (m, d) of OPTQ on weight w against the proxy Hessian H."""
    maxq = 2 ** bits - 1
    w = w.clone().double()
    scale, zero = row_grid(w, bits)
    hinv = torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(H.double())), upper=True)   # gptq.py:51-54
    codes = torch.zeros_like(w)
    for i in range(w.shape[1]):                                         # gptq.py:66-85
        col = w[:, i]
        c = torch.clamp(torch.round(col / scale[:, 0]) + zero[:, 0], 0, maxq)
        codes[:, i] = c
        err = (col - scale[:, 0] * (c - zero[:, 0])) / hinv[i, i]
        w[:, i:] -= err.unsqueeze(1) * hinv[i, i:].unsqueeze(0)
    return codes, scale, zero
