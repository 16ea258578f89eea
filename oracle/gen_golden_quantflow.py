"""tests/golden/quantflow_*.npz from the LIVE reference (build container only; CPU):  python -m oracle.gen_golden_quantflow

One Linear through the reference's whole quantisation flow (add_batch, post_batch, preproc, Balance.fasterquant) with the
exact arguments of vector_balance.quantize_weight_vecbal recorded at the call (w and H after preprocessing, scale, zero) next
to the inputs (initial weight, calibration activations, dense U / V) and the outputs (grid, codes, scaleWH)."""
import contextlib
import io
import os

import numpy as np
import torch
import torch.nn as nn

from oracle.gen_golden import OUT, import_reference, synth_inputs

CASES = [
    # name        N    K   bits method    seed npasses qfn  rescale proj
    ('incoh_rg', 128, 192, 2, 'ldlqRG', 31, 2,      'b', True,   True),
    ('plain_a',   96, 128, 4, 'ldlq',   32, 0,      'a', False,  False),
    ('rescale_a', 64, 128, 3, 'ldlq',   33, 1,      'a', True,   False),
]


def main():
    quant, method, bal, vb = import_reference()
    from quip_b200.capture import Capture
    for name, N, K, bits, qmethod, seed, npasses, qfn, rescale, proj in CASES:
        torch.manual_seed(seed)
        np.random.seed(seed)
        layer = nn.Linear(K, N, bias=False).half()
        W0 = layer.weight.data.clone()
        X = synth_inputs(K, 256, seed + 1)
        seen = {}
        with Capture(method, bal) as cap:
            inner = bal.quantize_weight_vecbal              # Capture's wrapper
            bal.quantize_weight_vecbal = lambda *a, **kw: spy_through(inner, seen, *a, **kw)
            qm = bal.Balance(layer)
            qm.configure(qmethod, bits, npasses, False)
            qm.quantizer = quant.Quantizer()
            qm.quantizer.configure(bits, perchannel=True, sym=False, qfn=qfn, mse=False)
            qm.add_batch(X.unsqueeze(0), None)
            qm.post_batch()
            qm.preproc(preproc_gptqH=True, percdamp=0.01, preproc_rescale=rescale, preproc_proj=proj, preproc_proj_extra=0)
            projU = qm.projU.clone() if proj else None
            projV = qm.projV.clone() if proj else None
            scaleWH = qm.scaleWH.clone() if rescale else None
            with contextlib.redirect_stderr(io.StringIO()):
                qm.fasterquant(lazy_batch=False)
            parts = cap.parts_for(layer)
            bal.quantize_weight_vecbal = inner
        d = dict(bits=np.int32(bits), method=np.array(qmethod), qfn=np.array(qfn), npasses=np.int32(npasses),
                 rescale=np.int32(rescale), proj=np.int32(proj), W0=W0.numpy(), X=X.numpy(),
                 w_pre=seen['w'].numpy(), H_pre=seen['H'].numpy(), grid=seen['grid'].numpy(), codes=parts.codes.numpy())
        if seen['scale'] is not None and qfn == 'a':
            d['scale'], d['zero'] = seen['scale'].numpy(), seen['zero'].numpy()
        if proj:
            d['U'], d['V'] = projU.numpy(), projV.numpy()
        if rescale:
            d['scaleWH'] = scaleWH.numpy()
        path = os.path.join(OUT, f'quantflow_{name}.npz')
        np.savez_compressed(path, **d)
        print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def spy_through(inner, seen, *a, **kw):
    seen.update(w=kw['w'].clone(), H=kw['H'].clone(),
                scale=None if kw.get('scale') is None else torch.as_tensor(kw['scale']).clone(),
                zero=None if kw.get('zero') is None else torch.as_tensor(kw['zero']).clone())
    out = inner(*a, **kw)
    seen['grid'] = out.clone()
    return out


if __name__ == '__main__':
    main()
