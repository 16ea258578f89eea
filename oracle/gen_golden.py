"""Generate tests/golden/*.npz from the LIVE reference (build container only).

Run:  python -m oracle.gen_golden            (needs /root/reference; CPU only)

The reference ships no golden vectors, so the oracle is pinned
against the reference's own outputs produced here.  Every array written below
comes out of unmodified reference code imported from /root/reference
(quant.py, method.py, bal.py, vector_balance.py, zeroShot/models/quant.py,
opt.py), with `oracle/_shims/primefac.py` standing in for the one missing
import.  quip_b200.capture.Capture only *records* the intermediates the
reference discards.  The files are committed; nothing at test or bench time
reads /root/reference.
"""
import argparse
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = '/root/reference'
OUT = os.path.join(ROOT, 'tests', 'golden')


def import_reference():
    sys.path.insert(0, os.path.join(HERE, '_shims'))
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    with contextlib.redirect_stdout(io.StringIO()):       # "CUDA extension not installed."
        import quant, method, bal, vector_balance           # noqa
    return quant, method, bal, vector_balance


def bpp_np(b):
    if b is None:
        return {}
    return dict(B0=b.B0.numpy(), B1=b.B1.numpy(), p_in=b.p_in.numpy(), p_out=b.p_out.numpy())


def save_parts(path, parts, extra):
    d = dict(bits=np.int32(parts.bits), qfn=np.array(parts.qfn),
             codes=parts.codes.numpy(), scales=parts.scales.numpy(), zeros=parts.zeros.numpy(),
             W_ref=parts.W_ref.numpy(), grid=parts.grid.numpy())
    if parts.bias is not None:
        d['bias'] = parts.bias.numpy()
    if parts.scaleWH is not None:
        d['scaleWH'] = parts.scaleWH.numpy()
    for side in 'UV':
        for k, v in bpp_np(getattr(parts, side)).items():
            d[f'{side}_{k}'] = v
    for k, v in parts.raw.items():
        if v is not None:
            d[f'raw_{k}'] = v.numpy()
    d.update(extra)
    np.savez_compressed(path, **d)
    print('wrote', os.path.relpath(path, ROOT), f'{os.path.getsize(path)/1024:.0f} KiB')


def synth_inputs(K, n_tok, seed):
    """X ~ N(0,1) * (1 + 3 U(0,1)) per-feature scale."""
    g = torch.Generator().manual_seed(seed)
    feat = 1.0 + 3.0 * torch.rand(K, generator=g)
    return (torch.randn(n_tok, K, generator=g) * feat).half()


def quantize_layer(mods, N, K, bits, qmethod, seed, bias, npasses=0, qfn='a',
                   rescale=False, proj=False, extra=0, gptqH=False):
    quant, method, bal, _ = mods
    from quip_b200.capture import Capture
    torch.manual_seed(seed)
    np.random.seed(seed)
    layer = nn.Linear(K, N, bias=bias).half()
    if bias:
        layer.bias.data = (torch.randn(N) * 0.1).half()
    X = synth_inputs(K, 512, seed + 1)
    with Capture(method, bal) as cap:
        qm = bal.Balance(layer)
        qm.configure(qmethod, bits, npasses, False)                 # bal.py:15
        qm.quantizer = quant.Quantizer()
        qm.quantizer.configure(bits, perchannel=True, sym=False, qfn=qfn, mse=False)   # opt.py:124-128
        qm.add_batch(X.unsqueeze(0), None)                          # method.py:98-120
        qm.post_batch()
        qm.preproc(preproc_gptqH=gptqH, percdamp=0.01, preproc_rescale=rescale,
                   preproc_proj=proj, preproc_proj_extra=extra)     # opt.py:154-157
        with contextlib.redirect_stderr(io.StringIO()):
            qm.fasterquant(lazy_batch=False)                        # opt.py:161
        parts = cap.parts_for(layer)
    x = synth_inputs(K, 24, seed + 2)
    with torch.no_grad():
        y_ref = nn.functional.linear(x, layer.weight.data, layer.bias)   # the reference's effective forward
    return parts, dict(x=x.numpy(), y_ref=y_ref.numpy())


LAYER_CASES = [
    # name              N    K   bits method   seed bias  kwargs
    ('l2b_incoh',       256, 384, 2, 'ldlq',   11, True,  dict(qfn='b', rescale=True, proj=True, gptqH=True)),
    ('l2b_incoh_rg',    384, 256, 2, 'ldlqRG', 12, False, dict(qfn='b', rescale=True, proj=True, gptqH=True, npasses=2)),
    ('l3b_incoh',       128, 640, 3, 'ldlq',   13, True,  dict(qfn='b', rescale=True, proj=True, gptqH=True)),
    ('l4b_plain',       256, 128, 4, 'ldlq',   14, True,  dict(qfn='a', gptqH=True)),
    ('l2b_kron',        256, 256, 2, 'ldlq',   15, False, dict(qfn='b', rescale=True, proj=True, gptqH=True, extra=1)),
    ('l4b_noperm',      128, 384, 4, 'ldlq',   16, True,  dict(qfn='b', rescale=True, proj=True, gptqH=True, extra=2)),
    ('l3b_rescale',     128, 256, 3, 'ldlq',   17, False, dict(qfn='a', rescale=True, gptqH=True)),
    ('l2b_qfna_proj',   128, 128, 2, 'ldlq',   18, True,  dict(qfn='a', proj=True, gptqH=True)),
]


def gen_layers(mods):
    for name, N, K, bits, qmethod, seed, bias, kw in LAYER_CASES:
        parts, extra = quantize_layer(mods, N, K, bits, qmethod, seed, bias, **kw)
        save_parts(os.path.join(OUT, f'layer_{name}.npz'), parts, extra)


def gen_quantizer(mods):
    quant = mods[0]
    out = {}
    for bits in (2, 3, 4):
        torch.manual_seed(100 + bits)
        w = (torch.randn(48, 160) * 0.05).half()
        w[5] = 0                                   # all-zero row -> [-1, 1] branch, quant.py:86-88
        w[7] = w[7].abs()                          # min clamps to 0, quant.py:77
        qz = quant.Quantizer()
        qz.configure(bits, perchannel=True, sym=False, qfn='a', mse=False)
        qz.find_params(w, weight=True)
        out[f'w{bits}'] = w.numpy()
        out[f'scale{bits}'] = qz.scale.numpy()
        out[f'zero{bits}'] = qz.zero.numpy()
        out[f'qa{bits}'] = quant.quantize_qfna(w, qz.scale, qz.zero, qz.maxq).numpy()
        s = 2.4 * w.square().mean().sqrt() + 1e-16
        out[f'sb{bits}'] = s.numpy()
        out[f'qb{bits}'] = quant.quantize_qfnb(w, s, qz.maxq).numpy()
    p = os.path.join(OUT, 'quantizer.npz')
    np.savez_compressed(p, **out)
    print('wrote', os.path.relpath(p, ROOT))


def gen_packing(mods):
    quant = mods[0]
    out = {}
    # 3-bit: Quant3Linear.pack (quant.py:185-220); K must be a multiple of 1024 there.
    torch.manual_seed(7)
    N, K = 16, 1024
    codes = torch.randint(0, 8, (N, K))
    scales = torch.rand(N, 1) * 0.02 + 0.01
    zeros = torch.randint(0, 8, (N, 1)).float()
    lin = nn.Linear(K, N)
    lin.weight.data = scales * (codes.float() - zeros)
    q3 = quant.Quant3Linear(K, N)
    q3.pack(lin, scales, zeros)
    out.update(codes3=codes.numpy().astype(np.uint8), scales3=scales.numpy(), zeros3=zeros.numpy(),
               qweight3=q3.qweight.numpy(), stored_zeros3=q3.zeros.numpy())
    # 4-bit: Quant4Linear (zeroShot/models/quant.py:185-199); the module imports quant_cuda at top level.
    sys.modules.setdefault('quant_cuda', types.ModuleType('quant_cuda'))
    import importlib.util
    spec = importlib.util.spec_from_file_location('zs_quant', os.path.join(REF, 'zeroShot/models/quant.py'))
    zs = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(zs)
    N, K = 16, 256
    codes = torch.randint(0, 16, (N, K))
    scales = torch.rand(N, 1) * 0.02 + 0.01
    zeros = torch.randint(0, 16, (N, 1)).float()
    lin = nn.Linear(K, N)
    lin.weight.data = scales * (codes.float() - zeros)
    q4 = zs.Quant4Linear(lin, scales, zeros)
    out.update(codes4=codes.numpy().astype(np.uint8), scales4=scales.numpy(), zeros4=zeros.numpy(),
               qweight4=q4.qweight.numpy(), stored_zeros4=q4.zeros.numpy())
    p = os.path.join(OUT, 'packing_ref.npz')
    np.savez_compressed(p, **out)
    print('wrote', os.path.relpath(p, ROOT))


def gen_butterfly(mods):
    method = mods[1]
    out = {}
    cases = [('blk', 'gen_rand_ortho_butterfly'), ('kron', 'gen_rand_ortho_butterfly_noblock'),
             ('noperm', 'gen_rand_ortho_butterfly_nopermute')]
    for n in (96, 128, 384, 640):
        out[f'factors_{n}'] = np.array(method.butterfly_factors(n))
        for tag, fn in cases:
            torch.manual_seed(n)
            np.random.seed(n)
            bpp = getattr(method, fn)(n)
            probe = torch.randn(n, 5)
            res = method.mul_ortho_butterfly(bpp, probe)
            p1, p2 = method.butterfly_factors(n)
            k = f'{tag}_{n}'
            out[f'{k}_B0'] = bpp[0][0].reshape(-1, p1, p1).numpy()
            out[f'{k}_B1'] = bpp[0][1].reshape(-1, p2, p2).numpy()
            out[f'{k}_pin'] = bpp[1].numpy()
            out[f'{k}_pout'] = bpp[2].numpy()
            out[f'{k}_probe'] = probe.numpy()
            out[f'{k}_out'] = res.numpy()
    for n in (768, 1024, 2048, 3072, 4096, 7168, 8192, 11008, 28672):
        out[f'factors_{n}'] = np.array(method.butterfly_factors(n))
    p = os.path.join(OUT, 'butterfly.npz')
    np.savez_compressed(p, **out)
    print('wrote', os.path.relpath(p, ROOT))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default='')
    a = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    mods = import_reference()
    todo = a.only.split(',') if a.only else ['quantizer', 'packing', 'butterfly', 'layers']
    if 'quantizer' in todo:
        gen_quantizer(mods)
    if 'packing' in todo:
        gen_packing(mods)
    if 'butterfly' in todo:
        gen_butterfly(mods)
    if 'layers' in todo:
        gen_layers(mods)
    if 'models' in todo:
        from oracle import gen_golden_models
        gen_golden_models.main(mods)
