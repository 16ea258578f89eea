"""Synthetic quantized layers: random codes and random orthogonal butterflies of the real shapes.

There is no network (no checkpoints, no calibration data), and running the reference's LDLQ over a
7B model takes hours of CPU; benchmarks and the full-size property tests therefore use *synthetic*
quantized layers with the exact structure the reference produces with `--incoh_processing`:
uniform random integer codes, the qfn 'b' grid (scale = 2.4*rms, quant.py:150), Haar-random
orthogonal blocks with the reference's factorisation rule (method.py:16-43) and random permutations,
and a log-normal scaleWH.
"""
import torch

from .capture import Butterfly, LayerParts, butterfly_factors


def _haar(m, p, gen, device):
    a = torch.randn(m, p, p, generator=gen, device=device)
    q, r = torch.linalg.qr(a)
    d = torch.diagonal(r, dim1=-2, dim2=-1)
    return q * torch.sign(d).unsqueeze(-2)


def synth_butterfly(n, mode='blocked', seed=0, device='cpu'):
    gen = torch.Generator(device=device).manual_seed(seed)
    p1, p2 = butterfly_factors(n)
    kron = mode == 'kron'
    B0 = _haar(1 if kron else n // p1, p1, gen, device)
    B1 = _haar(1 if kron else n // p2, p2, gen, device)
    if mode == 'noperm':
        p_in = p_out = torch.arange(n, device=device)
    else:
        p_in = torch.randperm(n, generator=gen, device=device)
        p_out = torch.randperm(n, generator=gen, device=device)
    return Butterfly(n, B0.float().cpu(), B1.float().cpu(), p_in.cpu(), p_out.cpu())


def synth_layer_parts(K, N, bits=2, incoh='blocked', rescale=True, bias=False, seed=0, w_std=0.02, device='cpu',
                      qfn='b'):
    gen = torch.Generator(device=device).manual_seed(seed)
    maxq = 2 ** bits - 1
    codes = torch.randint(0, maxq + 1, (N, K), generator=gen, device=device, dtype=torch.uint8).cpu()
    if qfn == 'b':
        s = torch.tensor(2.4 * w_std).half().float()
        scales = (2 * s / maxq).reshape(1, 1).repeat(N, 1)
        zeros = s.reshape(1, 1).repeat(N, 1)
    else:
        scales = (w_std * (0.5 + torch.rand(N, 1, generator=gen, device=device))).cpu() * 4 / maxq
        zeros = scales * torch.randint(0, maxq + 1, (N, 1), generator=gen, device=device).float().cpu()
    b = (0.1 * torch.randn(N, generator=gen, device=device)).half().cpu() if bias else None
    sWH = torch.exp(0.3 * torch.randn(K, generator=gen, device=device)).float().cpu() if rescale else None
    U = V = None
    if incoh:
        U = synth_butterfly(N, incoh, seed * 2 + 1, device)
        V = synth_butterfly(K, incoh, seed * 2 + 2, device)
    return LayerParts(bits=bits, qfn=qfn, codes=codes, scales=scales, zeros=zeros, bias=b, scaleWH=sWH, U=U, V=V)
