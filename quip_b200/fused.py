"""Llama decoder layers of the eval loop with the glue between the packed linears fused (csrc/glue.cu).

The reference's eval loop calls the HF decoder layer (llama.py:227); around the seven linears that layer issues ~25
elementwise / reduction launches (LlamaRMSNorm 8, apply_rotary_pos_emb 9 for q and k, residual adds, SiLU and the gate
product) which cost 27 % of a 2048-token step once the linears are packed (profiles/launches_r01.json).  `llama_stack`
runs the same layers -- the layer's own modules for every linear and its own weights for the norms -- with that glue as
four kernel launches per layer:

    x        = rmsnorm(h)                                   first layer only
    q, k, v  = q_proj(x), k_proj(x), v_proj(x);  rope_(q, k)           in place, token-major layout
    o        = o_proj(SDPA(q, k, v))                        torch SDPA (cuDNN / flash), strided head views, no copies
    h, x     = add_rmsnorm(h, o)                            residual add + post-attention norm
    d        = down_proj(silu_mul(gate_proj(x), up_proj(x)))
    h, x     = add_rmsnorm(h, d)                            residual add + the NEXT layer's input norm

Every fp16 rounding point of the HF modules is kept (see glue.cu), so the hidden states differ from the HF layer's only
through the summation order of the norms' fp32 mean.

`ops` is the provider of the three glue kernels: `CudaGlue` (the product: the C ABI, CUDA only, raises without the
extension) or, in the CPU tests, the torch restatement under oracle/ -- which this module never imports.
"""
import ctypes as C
import os

import torch
import torch.nn.functional as F

from . import _lib


class CudaGlue:
    """quip_rmsnorm / quip_rope / quip_silu_mul of the C ABI on torch tensors (fp16, CUDA, contiguous)."""

    @staticmethod
    def _check(*ts):
        for t in ts:
            if t is None:
                continue
            if not t.is_cuda:
                raise RuntimeError('the fused glue kernels run on a CUDA device only (there is no CPU fallback)')
            if t.dtype != torch.float16 or not t.is_contiguous():
                raise ValueError('fused glue kernels take contiguous fp16 tensors')

    @staticmethod
    def _stream(t):
        return torch.cuda.current_stream(t.device).cuda_stream

    def rmsnorm(self, x, weight, eps, residual=None):
        """y = weight * norm(x [+ residual]).  Returns y, or (x + residual, y) when a residual is given."""
        self._check(x, weight, residual)
        d = x.shape[-1]
        rows = x.numel() // d
        y = torch.empty_like(x)
        s = torch.empty_like(x) if residual is not None else None
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().quip_rmsnorm(x.data_ptr(), residual.data_ptr() if residual is not None else None,
                                                weight.data_ptr(), s.data_ptr() if s is not None else None, y.data_ptr(),
                                                rows, d, C.c_float(eps), self._stream(x)))
        return y if residual is None else (s, y)

    def rope_(self, q, k, cos, sin, head_dim):
        """In place: q (..., nq*hd), k (..., nkv*hd) token-major; cos, sin (rows, hd)."""
        self._check(q, k, cos, sin)
        rows = q.numel() // q.shape[-1]
        if cos.numel() != rows * head_dim or sin.numel() != rows * head_dim:
            raise ValueError(f'cos/sin must hold one row of {head_dim} per token ({rows} tokens), got {tuple(cos.shape)}')
        with torch.cuda.device(q.device):
            _lib.check(_lib.load().quip_rope(q.data_ptr(), k.data_ptr(), cos.data_ptr(), sin.data_ptr(), rows,
                                             q.shape[-1] // head_dim, k.shape[-1] // head_dim, head_dim, self._stream(q)))

    def silu_mul(self, gate, up):
        self._check(gate, up)
        out = torch.empty_like(gate)
        with torch.cuda.device(gate.device):
            _lib.check(_lib.load().quip_silu_mul(gate.data_ptr(), up.data_ptr(), out.data_ptr(), gate.numel(),
                                                 self._stream(gate)))
        return out


    def silu_mul_gather(self, gate, up, idx):
        """silu(gate[..., ig]) * up[..., iu] with idx = ig | iu << 16 per output feature (int32 tensor): quip_silu_mul_gather."""
        self._check(gate, up)
        n = gate.shape[-1]
        out = torch.empty_like(gate)
        with torch.cuda.device(gate.device):
            _lib.check(_lib.load().quip_silu_mul_gather(gate.data_ptr(), up.data_ptr(), idx.data_ptr(), out.data_ptr(),
                                                        gate.numel() // n, n, self._stream(gate)))
        return out


def mlp_layout_plan(mlp):
    """The combined index of the three permutations around SiLU(gate) * up of a packed Llama MLP -- the output gathers of
    gate_proj / up_proj (y[j] = layout[u_idx[j]]) and the input gather of down_proj (layout[l] = x[v_idx[l]]) -- so that one
    kernel (quip_silu_mul_gather) replaces three gather launches and silu_mul: for an 11008-wide side the gather is a kernel of
    its own (29 us at 2048 tokens each, profiles/launches_r01.json).  None when the layers are not packed, have a bias /
    unfolded 1/s at those gathers, or are too wide for 16-bit positions.  Cached on the module."""
    from .quant import QuantLinear
    gate, up, down = mlp.gate_proj, mlp.up_proj, mlp.down_proj
    if not all(isinstance(m, QuantLinear) for m in (gate, up, down)):
        return None
    dev = gate.qweight.device
    cached = getattr(mlp, '_quip_layout_plan', None)
    if cached is not None and cached[0] == dev:
        return cached[1]
    plan = None
    n = gate.outfeatures
    if (dev.type == 'cuda' and n == up.outfeatures == down.infeatures and n % 8 == 0 and n < 65536 and
            gate.layout_variant_ok(skip_out=True) and up.layout_variant_ok(skip_out=True) and down.layout_variant_ok(skip_in=True)):
        ar = torch.arange(n, device=dev)
        ig, iu, idn = (i if i is not None else ar for i in (gate.gather_index('u'), up.gather_index('u'), down.gather_index('v')))
        comb = ig[idn] | (iu[idn] << 16)                                     # int64, < 2^32
        plan = torch.where(comb >= 2 ** 31, comb - 2 ** 32, comb).to(torch.int32).contiguous()
    mlp._quip_layout_plan = (dev, plan)
    return plan


def enabled():
    """The fused stack is opt-in (QUIP_FUSED_LAYER=1): the library default is the HF modules' own glue, the reference's; bench.py
    and GraphDecoder switch it on after checking it against the HF layers in the run (bit-exact ops, one-ulp norms)."""
    return os.environ.get('QUIP_FUSED_LAYER') == '1'


def supports(model, h, kwargs):
    """Llama-family model, one fp16 sample, SiLU MLP, HF default rotary application, SDPA attention."""
    cfg = getattr(model, 'config', None)
    if cfg is None or getattr(cfg, 'model_type', None) != 'llama':
        return False
    if getattr(cfg, 'hidden_act', 'silu') != 'silu' or getattr(cfg, '_attn_implementation', 'sdpa') not in ('sdpa', None):
        return False
    if h.dim() != 3 or h.shape[0] != 1 or h.dtype != torch.float16:
        return False
    pe = kwargs.get('position_embeddings')
    if pe is None or pe[0].shape[0] != 1 or pe[0].dtype != torch.float16:
        return False
    hd = getattr(cfg, 'head_dim', None) or cfg.hidden_size // cfg.num_attention_heads
    return hd % 16 == 0 and cfg.hidden_size % 8 == 0 and cfg.intermediate_size % 8 == 0


def llama_stack(layers, h, kwargs, ops=None, trace=None, fold_gathers=None):
    """`for layer in layers: h = layer(h, **kwargs)` for LlamaDecoderLayers on one sample h (1, S, hidden).  `trace`: an optional
    list that receives (layer index, stage name, tensor) for tools/glue_bisect.py."""
    ops = ops or CudaGlue()
    if fold_gathers is None:
        fold_gathers = trace is None and os.environ.get('QUIP_FOLD_GATHERS', '1') == '1'
    cos, sin = kwargs['position_embeddings']
    cos, sin = cos[0].contiguous(), sin[0].contiguous()                     # (S, head_dim)
    mask = kwargs.get('attention_mask')
    S = h.shape[1]
    pend = None                        # previous layer's MLP output, not yet added to h
    x = None
    h = h.contiguous()

    def rec(li, name, t):
        if trace is not None:
            trace.append((li, name, t.clone()))

    for li, layer in enumerate(layers):
        a, mlp = layer.self_attn, layer.mlp
        n1, n2 = layer.input_layernorm, layer.post_attention_layernorm
        hd = a.head_dim
        if pend is None:
            x = ops.rmsnorm(h, n1.weight, n1.variance_epsilon)
        else:
            h, x = ops.rmsnorm(h, n1.weight, n1.variance_epsilon, residual=pend)
        rec(li, 'h_in', h)
        rec(li, 'x_attn', x)
        q, k, v = a.q_proj(x), a.k_proj(x), a.v_proj(x)                     # sibling groups launch these concurrently
        for nm, t in (('q_lin', q), ('k_lin', k), ('v_lin', v)):
            rec(li, nm, t)
        ops.rope_(q, k, cos, sin, hd)
        rec(li, 'q_rope', q)
        rec(li, 'k_rope', k)
        nq, nkv = q.shape[-1] // hd, k.shape[-1] // hd
        qh = q.view(1, S, nq, hd).transpose(1, 2)
        kh = k.view(1, S, nkv, hd).transpose(1, 2)
        vh = v.view(1, S, nkv, hd).transpose(1, 2)
        extra = {}
        # the same call transformers' sdpa_attention_forward makes (enable_gqa whenever there is no mask, also for
        # nkv == nq), so that torch picks the same SDPA backend for both glues
        if mask is None:
            extra['enable_gqa'] = True
        elif nkv != nq:
            kh = kh.repeat_interleave(nq // nkv, dim=1)
            vh = vh.repeat_interleave(nq // nkv, dim=1)
        o = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=mask, dropout_p=0.0, scale=a.scaling,
                                           is_causal=(mask is None and S > 1), **extra)
        o = o.transpose(1, 2).reshape(1, S, nq * hd).contiguous()
        rec(li, 'attn_out', o)
        ao = a.o_proj(o)
        rec(li, 'o_lin', ao)
        h, x = ops.rmsnorm(h, n2.weight, n2.variance_epsilon, residual=ao)
        rec(li, 'h_mid', h)
        rec(li, 'x_mlp', x)
        plan = mlp_layout_plan(mlp) if (fold_gathers and hasattr(ops, 'silu_mul_gather')) else None
        if plan is not None:
            # gate / up stay in their N-side layout order, the product lands in down_proj's K-side layout order: one kernel
            # instead of three gathers + silu_mul (same arithmetic per element, only the data movement differs)
            g, u = mlp.gate_proj.forward_layout(x, skip_out=True), mlp.up_proj.forward_layout(x, skip_out=True)
            act = ops.silu_mul_gather(g, u, plan)
            pend = mlp.down_proj.forward_layout(act, skip_in=True)
        else:
            g, u = mlp.gate_proj(x), mlp.up_proj(x)
            rec(li, 'gate', g)
            rec(li, 'up', u)
            act = ops.silu_mul(g, u)
            rec(li, 'act', act)
            pend = mlp.down_proj(act)
        rec(li, 'down', pend)
    return h if pend is None else h + pend
