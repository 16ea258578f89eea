"""One-token forwards of the three Llama-2-7B QuantLinear shapes for an ncu capture of the decode kernels (B200 only)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from quip_b200 import quant as Q
    from quip_b200.synth import synth_layer_parts
    g = torch.Generator(device='cuda').manual_seed(0)
    for (K, N) in [(4096, 4096), (4096, 11008), (11008, 4096)]:
        tp = synth_layer_parts(K=K, N=N, bits=2, incoh='blocked', rescale=True, bias=False, seed=K + N)
        ql = Q.QuantLinear(infeatures=K, outfeatures=N, **Q.spec_from_parts(tp))
        ql.pack_parts(tp)
        ql = ql.cuda()
        for M in (1, 8):
            xt = torch.randn(M, K, device='cuda', generator=g).half()
            for _ in range(3):
                ql(xt)
    torch.cuda.synchronize()


if __name__ == '__main__':
    main()
