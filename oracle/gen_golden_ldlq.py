"""tests/golden/ldlq.npz from the LIVE reference (build container only; CPU):  python -m oracle.gen_golden_ldlq

Inputs (w in grid units, H) and the reference's rounded outputs for LDLQ and LDLQ-RG, with and without greedy
passes, straight from vector_balance.round_vecbal_Hsort (vector_balance.py:426-466).  Committed; nothing at test
time reads /root/reference."""
import os

import numpy as np
import torch

from oracle.gen_golden import OUT, import_reference


def main():
    quant, method, bal, vb = import_reference()
    out = {}
    cases = []
    for ci, (m, d, nbits, npasses) in enumerate([(24, 96, 2, 0), (24, 96, 2, 2), (16, 128, 3, 0), (16, 128, 4, 3), (8, 64, 2, 9)]):
        g = torch.Generator().manual_seed(100 + ci)
        X = torch.randn(4 * d, d, generator=g) * (1 + 3 * torch.rand(d, generator=g))
        H = (X.T @ X) / X.shape[0]
        H = H + 0.01 * torch.mean(torch.diag(H)) * torch.eye(d)
        w = torch.rand(m, d, generator=g) * (2 ** nbits - 1)
        for meth in ('ldlq', 'ldlqRG'):
            torch.manual_seed(0)
            ref = vb.round_vecbal_Hsort(w.clone(), H.clone(), nbits, npasses, unbiased=False, qmethod=meth, lazy_batch=False)
            key = f'c{ci}_{meth}'
            out[key + '_out'] = ref.numpy().astype(np.float32)
            cases.append(f'{key}:{nbits}:{npasses}')
        out[f'c{ci}_w'] = w.numpy()
        out[f'c{ci}_H'] = H.numpy()
    out['cases'] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, 'ldlq.npz'), **out)
    print('wrote', os.path.join(OUT, 'ldlq.npz'), len(cases), 'cases')


if __name__ == '__main__':
    main()
