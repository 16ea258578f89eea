"""Quantise one Linear with QuIP (incoherence processing + LDLQ) into the parts a packed QuantLinear is built from.

This is the producer side of the hot path: what the reference does in `Balance.fasterquant` after `QuantMethod.preproc`
(bal.py:15-48, method.py:125-214, vector_balance.py:160-212 and :500-530), but keeping the integer codes, the structured
butterfly factors and 1/s instead of collapsing them back into a dense fp16 weight.  The arithmetic is plain torch (runs on
CPU or CUDA): device-agnostic set-up code, not a kernel -- the rounding recursion is blocked so that the feedback of all
already-rounded columns is one GEMM per block of 128 (the "lazy batch" structure of vector_balance.py:218-257) and only the
128 columns inside a block are visited one by one.  `oracle/ldlq.py` and `oracle/quantflow.py` are the checkers.

    parts = quantize_linear(layer.weight.data, H, bits=2, method='ldlq', qfn='b', incoh='blocked')
    qlinear.pack_parts(parts)            # quip_b200.quant.QuantLinear

`HessianAccumulator` is the calibration half (method.py:98-123): X^T X in float64 over batches.
"""
import math

import torch

from .capture import Butterfly, LayerParts, affine_from_quantizer, butterfly_factors


# ----------------------------------------------------------------------------------------------------------------
# calibration
# ----------------------------------------------------------------------------------------------------------------
class HessianAccumulator:
    """H = sum_b X_b^T X_b / #batches, float64 accumulation (method.py:98-123; the count is of batches, not tokens).

    On a CUDA device fp16 activations go through quip_hessian_accumulate (csrc/hessian.cu): tensor-core products -- exact in
    float32 for fp16 inputs -- summed in float32 over 256 tokens and carried in the float64 H, instead of the reference's
    float64 GEMM per batch.  Other dtypes / the CPU use the float64 matmul."""

    def __init__(self, features, device='cpu'):
        self.H = torch.zeros((features, features), dtype=torch.float64, device=device)
        self.batches = 0
        self._upper_only = False                            # the kernel fills the upper block triangle only

    @torch.no_grad()
    def add_batch(self, x):
        x = x if x.dim() == 3 else x.unsqueeze(0)
        self.batches += x.shape[0]
        flat = x.reshape(-1, x.shape[-1])
        K = flat.shape[1]
        if self.H.is_cuda and flat.dtype == torch.float16 and K % 8 == 0:
            import ctypes as C
            from . import _lib
            flat = flat.to(self.H.device).contiguous()
            with torch.cuda.device(self.H.device):
                _lib.check(_lib.load().quip_hessian_accumulate(_lib.ptr(flat), _lib.ptr(self.H), flat.shape[0], K,
                                                               C.c_void_p(torch.cuda.current_stream(self.H.device).cuda_stream)))
            self._upper_only = True
            return
        if self._upper_only:
            self._mirror()
        flat = flat.to(self.H.device, torch.float64)
        self.H.addmm_(flat.T, flat)

    def _mirror(self):
        up = torch.triu(self.H)
        self.H = up + torch.triu(self.H, 1).T
        self._upper_only = False

    def result(self):
        if self._upper_only:
            self._mirror()
        return (self.H / max(self.batches, 1)).to(torch.float32)


# ----------------------------------------------------------------------------------------------------------------
# random orthogonal butterflies (method.py:16-67)
# ----------------------------------------------------------------------------------------------------------------
def _haar(m, p, generator, device):
    """m Haar-distributed p x p orthogonal matrices: QR of a Gaussian with the sign of diag(R) folded into Q."""
    g = torch.randn((m, p, p), generator=generator, dtype=torch.float64)
    q, r = torch.linalg.qr(g)
    q = q * torch.sign(torch.diagonal(r, dim1=-2, dim2=-1)).unsqueeze(-2)
    return q.to(torch.float32).to(device)


def random_butterfly(n, mode='blocked', generator=None, device='cpu'):
    """P_out (B-diag_2) Pi^T (B-diag_1) Pi P_in with distinct blocks ('blocked', the reference's as-run default), one shared
    block per stage ('kron': P_out (B0 x B1) P_in) or no permutations ('noperm')."""
    p1, p2 = butterfly_factors(n)
    shared = mode == 'kron'
    B0 = _haar(1 if shared else n // p1, p1, generator, device)
    B1 = _haar(1 if shared else n // p2, p2, generator, device)
    if mode == 'noperm':
        p_in = p_out = torch.arange(n, device=device)
    else:
        p_in = torch.randperm(n, generator=generator).to(device)
        p_out = torch.randperm(n, generator=generator).to(device)
    return Butterfly(n, B0, B1, p_in, p_out)


def butterfly_apply(bf, x):
    """M @ x for x (n, q): gather p_in, stage 0 down the columns of the (p1, p2) view, stage 1 along its rows, gather
    p_out (method.py:46-67)."""
    n, p1, p2 = bf.n, bf.p1, bf.p2
    q = x.shape[1]
    t = x[bf.p_in].reshape(p1, p2, q)
    B0 = bf.B0.to(x.dtype).expand(p2, p1, p1)
    B1 = bf.B1.to(x.dtype).expand(p1, p2, p2)
    t = torch.einsum('bij,jbq->ibq', B0, t)                 # T[:, b] <- B0[b] T[:, b]
    t = torch.einsum('aij,ajq->aiq', B1, t)                 # T[a, :] <- B1[a] T[a, :]
    return t.reshape(n, q)[bf.p_out]


def butterfly_apply_t(bf, x):
    """M^T @ x (the inverse: M is orthogonal)."""
    n, p1, p2 = bf.n, bf.p1, bf.p2
    q = x.shape[1]
    t = torch.empty_like(x)
    t[bf.p_out] = x
    t = t.reshape(p1, p2, q)
    B0 = bf.B0.to(x.dtype).expand(p2, p1, p1)
    B1 = bf.B1.to(x.dtype).expand(p1, p2, p2)
    t = torch.einsum('aji,ajq->aiq', B1, t)
    t = torch.einsum('bji,jbq->ibq', B0, t)
    out = torch.empty_like(x)
    out[bf.p_in] = t.reshape(n, q)
    return out


# ----------------------------------------------------------------------------------------------------------------
# rounding
# ----------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def ldlq_round_cuda(w, H, nbits, greedy_passes=0, block=128):
    """ldlq_round on a CUDA device with the column-sequential part in hand-written kernels (csrc/ldlq.cu through the C ABI:
    quip_ldlq_block / quip_greedy_block).  Per block of 128 columns: one fp32 GEMM for the feedback of the finished blocks
    (torch / cuBLAS, TF32 off) and ONE kernel launch that walks the block's columns for all rows -- instead of the
    reference's d GEMV launches per pass (vector_balance.py:179-183, :186-196)."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    assert w.is_cuda and H.is_cuda and 1 <= block <= 128
    m, d = w.shape
    prev_tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        H = H.float()
        wT = w.float().T.contiguous()                            # (d, m): rows contiguous
        Cf = torch.linalg.cholesky(H)
        Lf = (Cf / torch.diagonal(Cf).unsqueeze(0) - torch.eye(d, dtype=H.dtype, device=H.device)).contiguous()
        qT = torch.empty_like(wT)
        errT = torch.zeros_like(wT)
        st = C.c_void_p(torch.cuda.current_stream(w.device).cuda_stream)
        with torch.cuda.device(w.device):
            for i2 in range(d, 0, -block):
                i1 = max(0, i2 - block)
                baseT = wT[i1:i2] if i2 == d else torch.addmm(wT[i1:i2], Lf[i2:, i1:i2].T, errT[i2:])
                Lb = Lf[i1:i2, i1:i2].contiguous()
                _lib.check(lib.quip_ldlq_block(_lib.ptr(baseT.contiguous()), _lib.ptr(wT[i1:i2]), _lib.ptr(Lb), _lib.ptr(qT[i1:i2]),
                                               _lib.ptr(errT[i1:i2]), m, m, i2 - i1, nbits, st))
            if greedy_passes:
                top = float(2 ** nbits - 1)
                Hn = (H / torch.diagonal(H).max()).contiguous()
                wrT = qT.clone()
                sT = qT - wT                                      # s = w_hat - w
                for _ in range(greedy_passes):
                    for i2 in range(d, 0, -block):
                        i1 = max(0, i2 - block)
                        # s of every column outside the block (finished blocks of this sweep and untouched ones) times H
                        preT = Hn[i1:i2, :i1] @ sT[:i1] + Hn[i1:i2, i2:] @ sT[i2:]
                        Hb = Hn[i1:i2, i1:i2].contiguous()
                        _lib.check(lib.quip_greedy_block(_lib.ptr(preT.contiguous()), _lib.ptr(Hb), _lib.ptr(wrT[i1:i2]),
                                                         _lib.ptr(sT[i1:i2]), m, m, i2 - i1, st))
                    wrT.clamp_(0, top)                            # s is NOT updated by the clamp (reference behaviour)
                    if bool((qT == wrT).all()):
                        break
                    qT.copy_(wrT)
                qT = wrT
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev_tf32
    return qT.T.contiguous()


@torch.no_grad()
def ldlq_round(w, H, nbits, greedy_passes=0, block=128, kernels=None):
    """LDLQ adaptive rounding of w (m, d), already in grid units, against the proxy Hessian H (d, d): columns last to first,
    column i rounded to nearest after adding the feedback of the rounding errors of columns > i through the strictly lower
    LDL factor (vector_balance.py:174-183).  Blocked: the errors of all finished blocks enter through one matmul.  On a
    CUDA device the in-block column loop runs in the kernels of csrc/ldlq.cu (`kernels=False` keeps this torch loop)."""
    if (w.is_cuda if kernels is None else kernels) and block <= 128:
        return ldlq_round_cuda(w, H, nbits, greedy_passes, block)
    w = w.float()
    H = H.float()
    m, d = w.shape
    top = float(2 ** nbits - 1)
    C = torch.linalg.cholesky(H)
    Lf = C / torch.diagonal(C).unsqueeze(0) - torch.eye(d, dtype=H.dtype, device=H.device)
    q = torch.empty_like(w)
    err = torch.zeros_like(w)                               # w - q of the finished columns
    for i2 in range(d, 0, -block):
        i1 = max(0, i2 - block)
        base = w[:, i1:i2] + err[:, i2:] @ Lf[i2:, i1:i2]
        Lb = Lf[i1:i2, i1:i2]
        eb = torch.zeros((m, i2 - i1), dtype=w.dtype, device=w.device)
        for j in range(i2 - i1 - 1, -1, -1):
            v = base[:, j] + eb[:, j + 1:] @ Lb[j + 1:, j]
            qj = torch.clamp(torch.floor(v + 0.5), min=0, max=top)
            q[:, i1 + j] = qj
            eb[:, j] = w[:, i1 + j] - qj
        err[:, i1:i2] = eb
    if greedy_passes:
        out = q.clone()
        s = q - w
        Hn = H / torch.diagonal(H).max()
        for _ in range(greedy_passes):                      # coordinate descent on the proxy loss (:185-202)
            for i in range(d - 1, -1, -1):
                move = out[:, i] - torch.round(out[:, i] - (s @ Hn[:, i]) / Hn[i, i])
                out[:, i] -= move
                s[:, i] -= move
            out = torch.clamp(out, min=0, max=top)
            if bool((q == out).all()):
                break
            q.copy_(out)
        q = out
    return q


@torch.no_grad()
def ldlq_rg_round(w, H, nbits, greedy_passes=0, block=128):
    """LDLQ-RG: the same on columns sorted by ascending diag(H) (round_sorted_ldlqRG, vector_balance.py:139-152)."""
    p = torch.argsort(torch.diagonal(H))
    out = torch.empty_like(w, dtype=torch.float32)
    out[:, p] = ldlq_round(w[:, p], H[p][:, p], nbits, greedy_passes, block)
    return out


# ----------------------------------------------------------------------------------------------------------------
# one Linear
# ----------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def quantize_linear(weight, H, bits=2, method='ldlq', greedy_passes=0, qfn='b', rescale=True, incoh='blocked',
                    percdamp=0.01, bias=None, generator=None, return_dense=False, damp=True):
    """weight (N, K) fp16/fp32, H (K, K) float32 from HessianAccumulator -> LayerParts (codes, affine map, 1/s, U, V).

    method: 'ldlq' | 'ldlq_rg' | 'nearest' (near.py: plain rounding, H unused).  incoh: 'blocked' | 'kron' | 'noperm' |
    None (no projection).  damp: the --pre_gptqH step (dead columns + percdamp on the diagonal).  With return_dense also the fake-quantised dense fp16 weight
    the reference would leave in the layer (method.py:195-214), e.g. to propagate activations to the next layer."""
    dev, wdtype = weight.device, weight.dtype
    N, K = weight.shape
    w = weight.clone()
    H = H.to(dev, torch.float32).clone()
    s = None
    if rescale:                                             # method.py:139-156
        w32 = w.float()
        H = H / H.abs().max()
        dH = torch.clamp(torch.diagonal(H), min=1e-8)
        dW = torch.clamp((w32 * w32).sum(0), min=1e-8)
        s = (dH / dW).sqrt().sqrt().clamp(min=1e-8)
        w = (w32 * s[None, :]).to(wdtype)
        H = H / s[None, :] / s[:, None]
    U = V = None
    if incoh is not None:                                   # method.py:157-180, U drawn before V
        U = random_butterfly(N, incoh, generator, dev)
        V = random_butterfly(K, incoh, generator, dev)
        H = H * (K / (torch.trace(H) + 1e-8)) + 1e-2 * torch.eye(K, device=dev)
        w32 = butterfly_apply(U, w.float())                 # U W
        w = butterfly_apply(V, w32.T.contiguous()).T.contiguous().to(wdtype)      # (U W) V^T
        H = butterfly_apply(V, butterfly_apply(V, H).T.contiguous()).T.contiguous()
    if damp:                                                # --pre_gptqH, method.py:182-192
        dead = torch.diagonal(H) == 0
        H[dead, dead] = 1
        w[:, dead] = 0
        H = H + percdamp * torch.mean(torch.diagonal(H)) * torch.eye(K, device=dev)

    maxq = float(2 ** bits - 1)
    method = {'ldlqRG': 'ldlq_rg'}.get(method, method)      # the reference's spelling
    if method not in ('ldlq', 'ldlq_rg', 'nearest'):
        raise NotImplementedError(f'rounding method {method!r} (ldlq, ldlq_rg and nearest are implemented)')
    rnd = {'ldlq': ldlq_round, 'ldlq_rg': ldlq_rg_round,
           'nearest': lambda t, H_, nbits, passes: torch.clamp(torch.round(t), 0, 2 ** nbits - 1)}[method]
    if qfn == 'a':                                          # per-row asymmetric grid (quant.py:57-127)
        x = w.flatten(1)
        zero_ = torch.zeros(N, device=dev)
        xmin = torch.minimum(x.min(1)[0].float(), zero_)
        xmax = torch.maximum(x.max(1)[0].float(), zero_)
        flat = (xmin == 0) & (xmax == 0)
        xmin[flat], xmax[flat] = -1, 1
        scale = ((xmax - xmin) / maxq).reshape(-1, 1)
        zero = torch.round(-xmin.reshape(-1, 1) / scale)
        codes = rnd(torch.clamp(w / scale + zero, 0, maxq).float(), H, bits, greedy_passes)
        grid = (scale * (codes - zero)).half()
    else:                                                   # symmetric grid of 2.4 rms (vector_balance.py:521-530)
        scale = 2.4 * w.square().mean().sqrt() + 1e-16
        zero = None
        codes = rnd(torch.clamp(((w / scale) + 1) / 2 * maxq, 0, maxq).float(), H, bits, greedy_passes)
        grid = (((codes / maxq) * 2 - 1) * scale).half()
    scales, zeros = affine_from_quantizer(qfn, scale, zero, maxq, N)
    parts = LayerParts(bits=bits, qfn=qfn, codes=codes.to(torch.uint8).cpu(), scales=scales.cpu(), zeros=zeros.cpu(),
                       bias=None if bias is None else bias.detach().clone().cpu(),
                       scaleWH=None if s is None else s.float().cpu(),
                       U=None if U is None else Butterfly(U.n, U.B0.cpu(), U.B1.cpu(), U.p_in.cpu(), U.p_out.cpu()),
                       V=None if V is None else Butterfly(V.n, V.B0.cpu(), V.B1.cpu(), V.p_in.cpu(), V.p_out.cpu()),
                       grid=grid.cpu(), raw=dict(scale=torch.as_tensor(scale).detach().cpu()))
    if not return_dense:
        return parts
    wq = grid.float()                                       # postproc (method.py:195-214)
    if U is not None:
        wq = butterfly_apply_t(U, wq)                       # U^T Q
        wq = butterfly_apply_t(V, wq.T.contiguous()).T.contiguous()               # (U^T Q) V
    wq = wq.to(wdtype)
    if s is not None:
        wq = (wq.float() / s[None, :]).to(wdtype)          # fp16 / fp32 -> fp32 -> layer dtype (method.py:210-213)
    parts.W_ref = wq.cpu()
    return parts, wq


def proxy_loss(w_hat, w, H):
    """tr((W_hat - W) H (W_hat - W)^T): what LDLQ minimises (normalised by tr(W H W^T))."""
    d = (w_hat - w).double()
    return float(torch.trace(d @ H.double() @ d.T) / torch.trace(w.double() @ H.double() @ w.double().T))


# ----------------------------------------------------------------------------------------------------------------
# a whole decoder stack, layer by layer
# ----------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def quantize_model(model, arch, calib_batches, dev=None, bits=2, method='ldlq', greedy_passes=0, qfn='b', rescale=True,
                   incoh='blocked', percdamp=0.01, generator=None, pack=True, verbose=False, damp=True):
    """Quantise every Linear inside the decoder layers of an OPT / Llama model, one decoder layer at a time, with the
    calibration activations produced by the already-quantised layers before it -- the flow of the reference's
    opt_sequential / llama_sequential (opt.py:29-190): per layer, (1) run the calibration samples through the layer with
    hooks accumulating X^T X for each Linear, (2) quantise all of its Linears from those Hessians, (3) run the samples again
    through the now fake-quantised layer to get the next layer's inputs.  Embeddings, final norm and lm_head stay fp16
    (opt.py:367-372).  Returns {dotted module name: LayerParts}; with pack=True the Linears are replaced by packed
    QuantLinear modules (quip_b200.opt.opt_pack) so the model is ready for `opt_eval` / `llama_eval`.

    calib_batches: list of (1, S) LongTensors."""
    from . import evalloop
    from .modelutils import find_layers
    dev = torch.device(dev) if dev is not None else next(iter(model.parameters())).device
    use_cache = model.config.use_cache
    model.config.use_cache = False
    layers = arch.layers(model)
    prefix = _layers_prefix(model, layers)
    inps, kw = [], {}
    for b in calib_batches:
        h, kw = evalloop.layer_inputs(model, arch, b.to(dev))
        inps.append(h)
    parts_by_name = {}
    for li, layer in enumerate(layers):
        linears = find_layers(layer)
        accs = {n: HessianAccumulator(m.in_features, dev) for n, m in linears.items()}
        hooks = [m.register_forward_hook(lambda mod, inp, out, n=n: accs[n].add_batch(inp[0].detach())) for n, m in linears.items()]
        for h in inps:
            evalloop._call_layer(layer, h, kw)
        for hk in hooks:
            hk.remove()
        for n, m in linears.items():
            parts, w_hat = quantize_linear(m.weight.data, accs[n].result(), bits=bits, method=method,
                                           greedy_passes=greedy_passes, qfn=qfn, rescale=rescale, incoh=incoh,
                                           percdamp=percdamp, bias=None if m.bias is None else m.bias.data,
                                           generator=generator, return_dense=True, damp=damp)
            if verbose:
                print(f'{prefix}.{li}.{n}: proxy loss {proxy_loss(w_hat.float(), m.weight.data.float(), accs[n].result()):.4f}')
            m.weight.data = w_hat.to(m.weight.data.dtype)
            parts_by_name[f'{prefix}.{li}.{n}'] = parts
        inps = [evalloop._call_layer(layer, h, kw) for h in inps]
    model.config.use_cache = use_cache
    if pack:
        from .opt import opt_pack
        opt_pack(model, parts_by_name)
    return parts_by_name


def _layers_prefix(model, layers):
    """Dotted name of the decoder ModuleList inside `model` (e.g. 'model.decoder.layers', 'model.layers')."""
    for name, mod in model.named_modules():
        if mod is layers:
            return name
    raise ValueError('decoder layers are not a submodule of the model')


_METHODS = {'ldlq': 'ldlq', 'ldlqRG': 'ldlq_rg', 'nearest': 'nearest'}
_PROJ = {0: 'blocked', 1: 'kron', 2: 'noperm'}


def options_from_args(args):
    """The reference's command-line namespace (opt.py:484-600) -> quantize_model keywords.  `--incoh_processing` expands as
    the reference does at opt.py:591-597 -- including its slip of setting `proj_extra` (unused) instead of
    `pre_proj_extra`, so the projection stays the blocked one."""
    g = lambda k, d=None: getattr(args, k, d)            # noqa: E731
    quant = g('quant', 'nearest')
    if quant not in _METHODS:
        raise NotImplementedError(f'--quant {quant}: ldlq, ldlqRG and nearest are implemented')
    if g('unbiased', False) or g('groupsize', -1) != -1 or g('lazy_batch', False):
        raise NotImplementedError('--unbiased / --groupsize / --lazy_batch are not implemented')
    pre_gptqH, pre_rescale, pre_proj, qfn = g('pre_gptqH', False), g('pre_rescale', False), g('pre_proj', False), g('qfn', 'a')
    if g('incoh_processing', False):
        pre_gptqH = pre_rescale = pre_proj = True
        qfn = 'b'
    # preproc runs for every method, nearest included (opt.py:154-157): dead-column zeroing (method.py:185-187) applies
    # there too; the damping of H is harmless because nearest ignores H
    return dict(bits=g('wbits'), method=_METHODS[quant], greedy_passes=g('npasses', 0), qfn=qfn, rescale=bool(pre_rescale),
                incoh=_PROJ[g('pre_proj_extra', 0)] if pre_proj else None, percdamp=g('percdamp', 0.01), damp=bool(pre_gptqH))


def sequential(model, arch, dataloader, dev, args, pack=False, generator=None, verbose=False):
    """`opt_sequential(model, dataloader, dev)` / `llama_sequential` with the reference's flags (opt.py:29-190): dataloader is
    the list of (input_ids, target) pairs datautils.get_loaders returns; `args` the namespace (the reference reads a module
    global).  Returns {module name: LayerParts} -- the role of the reference's `quantizers` dict."""
    batches = [b[0] if isinstance(b, (tuple, list)) else b for b in dataloader]
    return quantize_model(model, arch, batches, dev=dev, pack=pack, generator=generator, verbose=verbose,
                          **options_from_args(args))
