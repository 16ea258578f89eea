"""Capture adapter: record what the reference's quantizer throws away.

The reference is a fake-quant pipeline: after rounding it immediately
de-quantizes and un-projects back to a dense fp16 weight (bal.py:44-45 ->
method.py:195-214) and `free()` drops U, V and scaleWH (method.py:216-226).  A
packed QuantLinear needs exactly those: the integer codes, the dequant
parameters, scaleWH and the *structured* butterfly factors.

`Capture(method_mod, bal_mod)` wraps -- without editing any reference file --
  * `method.rand_ortho_butterfly{,_noblock,_nopermute}` (method.py:71-78): the
    wrapper draws the factors with the reference's own generator (same RNG
    consumption, method.py:34-43), records them, and returns the same dense matrix;
  * `QuantMethod.preproc` (method.py:125-193): binds the two recorded butterflies
    to the layer (U is drawn first, then V: method.py:162-169);
  * `bal.quantize_weight_vecbal` (bal.py:10,32; vector_balance.py:499-532):
    records scale / zero / maxq / qfn and the fp16 grid matrix it returns;
  * `Balance.fasterquant` (bal.py:21-48): after the original returns, collects the
    record for that layer (scaleWH is still alive until the driver calls free()).

Use:
    import method, bal                      # the reference's modules
    with Capture(method, bal) as cap:
        quantizers, errors = opt.opt_sequential(model, loader, dev, args)
    parts = cap.parts_for(layer_module)     # -> LayerParts
"""
import math
from dataclasses import dataclass, field
from typing import Optional

import torch


@dataclass
class Butterfly:
    """([B0, B1], p_in, p_out) of reference method.py:34-43, for a side of size n."""
    n: int
    B0: torch.Tensor          # (n/p1, p1, p1) or (1, p1, p1), float32
    B1: torch.Tensor          # (n/p2, p2, p2) or (1, p2, p2), float32
    p_in: torch.Tensor        # (n,) int64
    p_out: torch.Tensor       # (n,) int64

    @property
    def p1(self):
        return self.B0.shape[-1]

    @property
    def p2(self):
        return self.B1.shape[-1]

    @staticmethod
    def from_bpp(bpp, n):
        (B, p_in, p_out) = bpp
        p1, p2 = B[0].shape[-1], B[1].shape[-1]
        assert p1 * p2 == n
        return Butterfly(n, B[0].reshape(-1, p1, p1).float().clone(),
                         B[1].reshape(-1, p2, p2).float().clone(),
                         p_in.long().clone(), p_out.long().clone())

    def as_bpp(self):
        return ([self.B0, self.B1], self.p_in, self.p_out)


@dataclass
class LayerParts:
    """Everything a packed QuantLinear is built from (one reference Linear)."""
    bits: int
    qfn: str
    codes: torch.Tensor                   # (N, K) uint8, in the incoherent basis
    scales: torch.Tensor                  # (N, 1) float32:  Q = scales*codes - zeros
    zeros: torch.Tensor                   # (N, 1) float32   (pre-multiplied, quant.py:186)
    bias: Optional[torch.Tensor] = None   # (N,)
    scaleWH: Optional[torch.Tensor] = None    # (K,) float32, method.py:147-154
    U: Optional[Butterfly] = None         # size N
    V: Optional[Butterfly] = None         # size K
    W_ref: Optional[torch.Tensor] = None  # dense fp16 the reference leaves in layer.weight
    grid: Optional[torch.Tensor] = None   # fp16 grid matrix returned by quantize_weight_vecbal
    raw: dict = field(default_factory=dict)   # raw quantizer scale / zero as the reference held them


def codes_from_grid(grid, qfn, scale, zero, maxq):
    """Inverse of the value maps at vector_balance.py:514-530 (exact: the grid has
    at most maxq+1 distinct values per row and the map is evaluated in float64)."""
    g = grid.double()
    if qfn == 'b':
        c = torch.round((g / float(scale) + 1.0) / 2.0 * float(maxq))
    else:
        c = torch.round(g / scale.double() + zero.double())
    if c.min() < 0 or c.max() > maxq:
        raise ValueError('grid value outside the code range')
    return c.to(torch.uint8)


def affine_from_quantizer(qfn, scale, zero, maxq, n_rows):
    """(scales, zeros) such that grid = scales*code - zeros in real arithmetic."""
    if qfn == 'b':
        s = torch.tensor(float(scale), dtype=torch.float32)
        scales = (2.0 * s / float(maxq)).reshape(1, 1).repeat(n_rows, 1)
        zeros = s.reshape(1, 1).repeat(n_rows, 1)
    else:
        scales = scale.float().reshape(-1, 1).clone()
        zeros = (zero.float().reshape(-1, 1) * scales)
        if scales.shape[0] == 1:
            scales, zeros = scales.repeat(n_rows, 1), zeros.repeat(n_rows, 1)
    return scales.contiguous(), zeros.contiguous()


class Capture:
    def __init__(self, method_mod, bal_mod=None):
        self.method = method_mod
        self.bal = bal_mod
        self._saved = []
        self._pending = []
        self._last_q = None
        self._uv = {}
        self.records = {}

    # -- patch helpers ------------------------------------------------------
    def _patch(self, obj, name, new):
        self._saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, new)

    def __enter__(self):
        m = self.method
        cap = self

        def wrap_gen(gen_name):
            gen = getattr(m, gen_name)

            def dense_from_recorded(n):
                bpp = gen(n)
                cap._pending.append(Butterfly.from_bpp(bpp, n))
                return m.mul_ortho_butterfly(bpp, torch.eye(n))
            return dense_from_recorded

        self._patch(m, 'rand_ortho_butterfly', wrap_gen('gen_rand_ortho_butterfly'))
        self._patch(m, 'rand_ortho_butterfly_noblock', wrap_gen('gen_rand_ortho_butterfly_noblock'))
        self._patch(m, 'rand_ortho_butterfly_nopermute', wrap_gen('gen_rand_ortho_butterfly_nopermute'))

        orig_preproc = m.QuantMethod.preproc

        def preproc(qm, *a, **kw):
            cap._pending = []
            out = orig_preproc(qm, *a, **kw)
            if len(cap._pending) == 2:
                cap._uv[id(qm)] = (cap._pending[0], cap._pending[1])
            elif cap._pending:
                raise RuntimeError('expected U then V from preproc (method.py:162-169)')
            cap._pending = []
            return out
        self._patch(m.QuantMethod, 'preproc', preproc)

        if self.bal is not None:
            b = self.bal
            orig_q = b.quantize_weight_vecbal

            def quantize_weight_vecbal(*a, **kw):
                out = orig_q(*a, **kw)
                w = kw['w'] if 'w' in kw else a[0]
                qfn = kw.get('qfn', 'a')
                scale, zero = kw.get('scale'), kw.get('zero')
                if qfn == 'b':
                    scale = 2.4 * w.square().mean().sqrt() + 1e-16     # vector_balance.py:522
                cap._last_q = dict(qfn=qfn, scale=scale, zero=zero, maxq=int(kw['maxq']),
                                   bits=int(kw['nbits']), grid=out.clone())
                return out
            self._patch(b, 'quantize_weight_vecbal', quantize_weight_vecbal)

            orig_fq = b.Balance.fasterquant

            def fasterquant(qm, *a, **kw):
                cap._last_q = None
                out = orig_fq(qm, *a, **kw)
                cap._collect(qm)
                return out
            self._patch(b.Balance, 'fasterquant', fasterquant)
        return self

    def __exit__(self, *exc):
        for obj, name, old in reversed(self._saved):
            setattr(obj, name, old)
        self._saved = []
        return False

    # -- record assembly ----------------------------------------------------
    def _collect(self, qm):
        q = self._last_q
        if q is None:
            raise RuntimeError('fasterquant did not go through quantize_weight_vecbal')
        layer = qm.layer
        N = qm.rows
        codes = codes_from_grid(q['grid'], q['qfn'], q['scale'], q['zero'], q['maxq'])
        scales, zeros = affine_from_quantizer(q['qfn'], q['scale'], q['zero'], q['maxq'], N)
        U = V = None
        if getattr(qm, 'preproc_proj', False):
            U, V = self._uv.pop(id(qm))
        s = qm.scaleWH.float().clone() if getattr(qm, 'preproc_rescale', False) else None
        bias = layer.bias.data.clone() if getattr(layer, 'bias', None) is not None else None
        raw = dict(scale=None if q['scale'] is None else torch.as_tensor(q['scale']).clone(),
                   zero=None if q['zero'] is None else torch.as_tensor(q['zero']).clone())
        self.records[id(layer)] = LayerParts(
            bits=q['bits'], qfn=q['qfn'], codes=codes, scales=scales, zeros=zeros, bias=bias,
            scaleWH=s, U=U, V=V, W_ref=layer.weight.data.clone(), grid=q['grid'], raw=raw)
        self._last_q = None

    def parts_for(self, layer):
        return self.records[id(layer)]


def butterfly_factors(n):
    """Same rule as method.py:16-18 (sorted prime factors, alternating products)."""
    pf, d, m = [], 2, int(n)
    while d * d <= m:
        while m % d == 0:
            pf.append(d)
            m //= d
        d += 1 if d == 2 else 2
    if m > 1:
        pf.append(m)
    return math.prod(pf[0::2]), math.prod(pf[1::2])
