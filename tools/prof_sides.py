"""One 2048-token QuantLinear forward per Llama-2-7B layer shape, for an ncu --set full capture of every kernel of the
many-token route (one-kernel sides, gather / dense 688 pass / 16-wide pass of the 11008 sides, the tcgen05 GEMM):
    ncu --set full --clock-control none --import-source on -k regex:quip -s <warm-up launches> -o gpurun_out/prof_sides_r02 python tools/prof_sides.py
The script prints the number of library launches of the warm-up pass so that -s can skip exactly those."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from quip_b200 import _lib, quant as Q
    from quip_b200.synth import synth_layer_parts
    lib = _lib.load()
    mods = []
    g = torch.Generator(device='cuda').manual_seed(0)
    for (K, N) in [(4096, 4096), (4096, 11008), (11008, 4096)]:
        tp = synth_layer_parts(K=K, N=N, bits=2, incoh='blocked', rescale=True, bias=False, seed=K + N)
        ql = Q.QuantLinear(infeatures=K, outfeatures=N, **Q.spec_from_parts(tp))
        ql.pack_parts(tp)
        mods.append((ql.cuda(), torch.randn(2048, K, device='cuda', generator=g).half()))
    n0 = lib.quip_launch_count()
    for ql, x in mods:                     # warm-up: descriptors, workspaces, module loads
        ql(x)
    torch.cuda.synchronize()
    print('warm-up launches', lib.quip_launch_count() - n0, flush=True)
    for ql, x in mods:
        ql(x)
    torch.cuda.synchronize()


if __name__ == '__main__':
    main()
