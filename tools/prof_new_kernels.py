"""Drive the round-2 kernels once each so that ncu can capture them (B200 only):
    ncu --set full --clock-control none -k regex:'hessian|ldlq_block|greedy_block|side_fewtok|vecquant' -o gpurun_out/prof_r2new python tools/prof_new_kernels.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from quip_b200 import quant as Q
    from quip_b200 import quantize as qz
    from quip_b200.synth import synth_layer_parts
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(0)
    # Hessian of a 4096-wide Linear from one 2048-token sample
    x = (torch.randn(1, 2048, 4096, device=dev, generator=g) * (1 + 3 * torch.rand(4096, device=dev, generator=g))).half()
    acc = qz.HessianAccumulator(4096, device=dev)
    for _ in range(2):
        acc.add_batch(x)
    H = acc.result()
    H = H + 0.01 * torch.diagonal(H).mean() * torch.eye(4096, device=dev)
    # LDLQ + two greedy passes of a 4096 x 4096 weight in grid units
    w = torch.rand(4096, 4096, device=dev, generator=g) * 3
    qz.ldlq_round(w, H, 2, 2)
    # decode-time sides and the reference-layout GEMV
    for (K, N) in [(4096, 4096), (4096, 11008), (11008, 4096)]:
        tp = synth_layer_parts(K=K, N=N, bits=2, incoh='blocked', rescale=True, bias=False, seed=K + N)
        ql = Q.QuantLinear(infeatures=K, outfeatures=N, **Q.spec_from_parts(tp))
        ql.pack_parts(tp)
        ql = ql.cuda()
        xt = torch.randn(1, K, device=dev, generator=g).half()
        for _ in range(3):
            ql(xt)
    m = Q.Quant3Linear(4096, 4096).cuda()
    m.qweight.random_(-2 ** 31, 2 ** 31 - 1)
    m.scales.fill_(0.01)
    m.zeros.fill_(0.03)
    for _ in range(3):
        m(torch.randn(1, 4096, device=dev, generator=g))
    torch.cuda.synchronize()


if __name__ == '__main__':
    main()
