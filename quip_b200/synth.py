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
    """m Haar-random p x p orthogonal matrices (QR of a Gaussian, sign-fixed) -- what scipy's
    special_ortho_group gives the reference (method.py:22)."""
    a = torch.randn(m, p, p, generator=gen, device=device)
    q, r = torch.linalg.qr(a)
    d = torch.diagonal(r, dim1=-2, dim2=-1)
    return q * torch.sign(d).unsqueeze(-2)


def _fast_ortho(m, p, gen, device):
    """m dense random orthogonal p x p matrices as a product of two Householder reflections: exactly
    orthogonal, fully dense, and ~100x cheaper to generate than a batched QR (used to build whole synthetic
    models; the kernels' cost does not depend on which orthogonal matrix they multiply by)."""
    eye = torch.eye(p, device=device).expand(m, p, p)
    out = None
    for _ in range(2):
        v = torch.randn(m, p, 1, generator=gen, device=device)
        v = v / v.norm(dim=1, keepdim=True)
        h = eye - 2.0 * v @ v.transpose(1, 2)
        out = h if out is None else out @ h
    return out


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


# ------------------------------------------------------------------------------------------------
# whole synthetic models, built directly on the target device (nothing dense is ever materialised)
# ------------------------------------------------------------------------------------------------
def init_synthetic_(ql, seed=0, w_std=0.02):
    """Fill a QuantLinear's buffers in place, on its own device, with a synthetic quantized layer:
    uniform random packed words, the symmetric qfn 'b' grid, Haar-random butterfly factors, random
    gather indices; 1/scaleWH is treated as already folded into the first V pass (as pack_parts does for
    the as-run blocked butterfly)."""
    dev = ql.qweight.device
    gen = torch.Generator(device=dev).manual_seed(seed)
    maxq = 2 ** ql.bits - 1
    n_words = ql.qweight.numel()
    ql.qweight.copy_(torch.randint(-2 ** 31, 2 ** 31 - 1, (n_words,), generator=gen, device=dev, dtype=torch.int64)
                     .to(torch.int32))
    s = float(torch.tensor(2.4 * w_std).half())
    ql.scales.fill_(2 * s / maxq)
    ql.zeros.fill_(s)
    ql.meta[1] = 1
    if ql.bias is not None:
        ql.bias.copy_((0.02 * torch.randn(ql.outfeatures, generator=gen, device=dev)).half())
    if ql.incoh:
        for side, n in (('v', ql.infeatures), ('u', ql.outfeatures)):
            idx = torch.arange(n, device=dev) if ql.incoh == 'noperm' else torch.randperm(n, generator=gen, device=dev)
            getattr(ql, f'{side}_idx').copy_(idx.to(torch.int32))
            ql.meta[3 if side == 'v' else 4] = int(ql.incoh == 'noperm')
            for i in range(2):
                buf = getattr(ql, f'{side}_f{i}')
                nb, p, _ = buf.shape
                for lo in range(0, nb, 256):
                    hi = min(nb, lo + 256)
                    buf[lo:hi].copy_(_fast_ortho(hi - lo, p, gen, dev).half())
    if ql.rescale:
        ql.inv_scale.fill_(1.0)
        ql.meta[2] = 1 if ql.incoh in ('blocked', 'noperm') else 0
    ql._desc = None
    return ql


def build_synthetic_model(config, device, bits=2, incoh='blocked', rescale=True, seed=0, seqlen=2048, dtype=torch.float16,
                          layer_range=None, head=True):
    """A random-init HF OPT / Llama model whose decoder Linears are synthetic packed QuantLinears, created on
    `device` without ever allocating the dense decoder weights (meta-device construction).

    layer_range=(lo, hi): a pipeline stage -- only decoder layers lo..hi-1 are materialised (the others stay on the meta
    device and must not be called); head=False leaves final norm / lm_head on the meta device too.  The embedding and the
    rotary / position modules are always real: every stage derives the layer kwargs from them once."""
    import torch.nn as nn
    from .modelutils import find_layers
    from .quant import QuantLinear
    from transformers import LlamaConfig
    is_llama = isinstance(config, LlamaConfig)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        with torch.device('meta'):
            if is_llama:
                from transformers import LlamaForCausalLM
                model = LlamaForCausalLM(config)
            else:
                from transformers import OPTForCausalLM
                model = OPTForCausalLM(config)
    finally:
        torch.set_default_dtype(prev)
    layers = model.model.layers if is_llama else model.model.decoder.layers
    lo, hi = layer_range or (0, len(layers))
    k = 0
    for li, layer in enumerate(layers):
        names = find_layers(layer)
        if not (lo <= li < hi):
            k += len(names)                                       # same seeds per layer whatever the stage split
            continue
        for name, lin in names.items():
            ql = QuantLinear(bits, lin.in_features, lin.out_features, bias=lin.bias is not None, incoh=incoh,
                             rescale=rescale).to(device)
            init_synthetic_(ql, seed=seed * 100003 + k)
            k += 1
            parent = layer
            *path, leaf = name.split('.')
            for part in path:
                parent = getattr(parent, part)
            setattr(parent, leaf, ql)
    gen = torch.Generator(device=device).manual_seed(seed + 7)
    skip = set()
    for li, layer in enumerate(layers):
        if not (lo <= li < hi):
            skip.update(id(m) for m in layer.modules())
    if not head:
        dec = model.model if is_llama else model.model.decoder
        for mod in (getattr(dec, 'norm', None), getattr(dec, 'final_layer_norm', None), getattr(dec, 'project_out', None),
                    model.lm_head):
            if mod is not None:
                skip.update(id(m) for m in mod.modules())
    for mod in model.modules():
        if id(mod) in skip:
            continue
        for pname, p in list(mod._parameters.items()):
            if p is None or not p.is_meta:
                continue
            t = torch.empty(p.shape, dtype=dtype, device=device)
            if p.dim() == 1:
                t.fill_(0.0 if pname == 'bias' else 1.0)          # norm weights 1, biases 0
            else:
                t.normal_(0.0, 0.02, generator=gen)
            mod._parameters[pname] = nn.Parameter(t, requires_grad=False)
        for bname, b in list(mod._buffers.items()):
            if b is not None and b.is_meta:
                mod._buffers[bname] = torch.zeros(b.shape, dtype=b.dtype, device=device)
    if is_llama and hasattr(model.model, 'rotary_emb'):
        model.model.rotary_emb = type(model.model.rotary_emb)(config=config, device=device)
    model.seqlen = seqlen if is_llama else config.max_position_embeddings
    return model.eval()


LLAMA2_7B = dict(vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                 num_attention_heads=32, num_key_value_heads=32, max_position_embeddings=4096, rms_norm_eps=1e-5)
OPT_1_3B = dict(vocab_size=50272, hidden_size=2048, ffn_dim=8192, num_hidden_layers=24, num_attention_heads=32,
                max_position_embeddings=2048, word_embed_proj_dim=2048, do_layer_norm_before=True)
OPT_30B = dict(vocab_size=50272, hidden_size=7168, ffn_dim=28672, num_hidden_layers=48, num_attention_heads=56,
               max_position_embeddings=2048, word_embed_proj_dim=7168, do_layer_norm_before=True)
LLAMA2_70B = dict(vocab_size=32000, hidden_size=8192, intermediate_size=28672, num_hidden_layers=80,
                  num_attention_heads=64, num_key_value_heads=8, max_position_embeddings=4096, rms_norm_eps=1e-5)
OPT_125M = dict(vocab_size=50272, hidden_size=768, ffn_dim=3072, num_hidden_layers=12, num_attention_heads=12,
                max_position_embeddings=2048, word_embed_proj_dim=768, do_layer_norm_before=True)


MODELS = {'llama7b': ('llama', LLAMA2_7B), 'llama70b': ('llama', LLAMA2_70B), 'opt125m': ('opt', OPT_125M),
          'opt1.3b': ('opt', OPT_1_3B), 'opt30b': ('opt', OPT_30B)}


def model_config(name, **overrides):
    """HF config of one of the BASELINE.json model shapes (SURVEY section 8 header)."""
    from transformers import LlamaConfig, OPTConfig
    family, kw = MODELS[name]
    return (LlamaConfig if family == 'llama' else OPTConfig)(**{**kw, **overrides})
