"""Kernel-level timings at Llama-2-7B layer shapes (CUDA events, warm-up, rotating weight copies so the
packed loads come from HBM).  Writes gpurun_out/microbench.json.  Not the bench contract (bench.py is)."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quip_b200 import _lib, quant as Q          # noqa: E402
from quip_b200.synth import synth_layer_parts   # noqa: E402

DEV = 'cuda:0'


ITERS = None


GRAPH = False


def timeit(fn, iters=50, warm=5):
    if ITERS:
        iters, warm = ITERS, 2
    for w in range(max(warm, 8)):
        fn(w)
    torch.cuda.synchronize()
    if GRAPH:
        # replay `iters` back-to-back launches from one CUDA graph: GPU time without the Python/ctypes launch cost
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for i in range(iters):
                    fn(i)
        torch.cuda.current_stream().wait_stream(side)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (3 * iters) * 1e3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3      # microseconds


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def bench_qgemm(N, K, M, bits, path, copies, peaks):
    lib = _lib.load()
    words = Q.packed_words(N, K, bits)
    qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (copies, words), dtype=torch.int32, device=DEV)
    sc = torch.full((N,), 0.01, device=DEV)
    ze = sc * ((2 ** bits - 1) / 2)
    x = torch.randn(M, K, device=DEV).half()
    z = torch.empty(M, N, dtype=torch.float16, device=DEV)
    descs = []
    for c in range(copies):
        d = _lib.QuipLinearDesc()
        d.K, d.N, d.bits, d.flags = K, N, bits, _lib.QUIP_FLAG_SYMMETRIC
        d.qweight, d.scales, d.zeros = qw[c].data_ptr(), sc.data_ptr(), ze.data_ptr()
        descs.append(d)
    need = C.c_size_t()
    _lib.check(lib.quip_qlinear_workspace_bytes(C.byref(descs[0]), M, C.byref(need)))
    ws = torch.zeros(need.value, dtype=torch.uint8, device=DEV)

    def fn(i):
        _lib.check(lib.quip_qgemm(C.byref(descs[i % copies]), _lib.ptr(x), None, None, _lib.ptr(z), M, path,
                                  _lib.ptr(ws), ws.numel(), stream()))
    us = timeit(fn)
    code_bytes = N * K * bits / 8
    alg_bytes = code_bytes + 2 * M * K + 2 * M * N
    flops = 2.0 * M * N * K
    return dict(kind='qgemm', N=N, K=K, M=M, bits=bits, path=path, us=us, GBps_codes_plus_act=alg_bytes / us / 1e3,
                hbm_frac=alg_bytes / us / 1e3 / peaks['hbm_gbs'], TFLOPs=flops / us / 1e6,
                tensor_frac=flops / us / 1e6 / peaks['bf16_tflops'])


def bench_dense(N, K, M, peaks):
    w = torch.randn(N, K, device=DEV).half()
    x = torch.randn(M, K, device=DEV).half()
    us = timeit(lambda i: torch.nn.functional.linear(x, w))
    flops = 2.0 * M * N * K
    return dict(kind='cublas_fp16_dense', N=N, K=K, M=M, us=us, TFLOPs=flops / us / 1e6,
                GBps=(2 * N * K + 2 * M * K + 2 * M * N) / us / 1e3)


def bench_layer(N, K, M, bits, incoh, peaks, copies=1):
    mods = []
    for c in range(copies):
        tp = synth_layer_parts(K=K, N=N, bits=bits, incoh=incoh, rescale=incoh is not None, seed=c, device=DEV)
        ql = Q.QuantLinear(infeatures=K, outfeatures=N, **Q.spec_from_parts(tp)).to(DEV)
        ql.pack_parts(tp)
        mods.append(ql)
    x = torch.randn(M, K, device=DEV).half()
    us = timeit(lambda i: mods[i % copies](x), iters=30)
    flops = 2.0 * M * N * K
    return dict(kind='qlinear_forward', N=N, K=K, M=M, bits=bits, incoh=incoh, us=us, TFLOPs_main=flops / us / 1e6)


def bench_pass(n, p, nblk, strided, M):
    lib = _lib.load()
    f = (torch.randn(nblk, p, p, device=DEV) / p ** 0.5).half()
    x = torch.randn(M, n, device=DEV).half()
    out = torch.empty_like(x)
    ps = _lib.QuipPass(p=p, nblk=nblk, strided=int(strided), shared=0, factors=f.data_ptr())

    def fn(i):
        _lib.check(lib.quip_rot_pass(C.byref(ps), _lib.ptr(x), _lib.ptr(out), M, n, 0, stream()))
    us = timeit(fn, iters=30)
    return dict(kind='rot_pass', n=n, p=p, nblk=nblk, strided=strided, M=M, us=us,
                GBps=(4.0 * M * n + 2.0 * nblk * p * p) / us / 1e3, TFLOPs=2.0 * M * n * p / us / 1e6)


def bench_gather(n, M):
    lib = _lib.load()
    x = torch.randn(M, n, device=DEV).half()
    out = torch.empty_like(x)
    idx = torch.randperm(n, device=DEV).int()
    us = timeit(lambda i: _lib.check(lib.quip_gather(_lib.ptr(x), _lib.ptr(out), M, n, _lib.ptr(idx), None, None, stream())), iters=30)
    return dict(kind='gather', n=n, M=M, us=us, GBps=4.0 * M * n / us / 1e3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--what', default='qgemm,dense,pass,layer')
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'microbench.json'))
    ap.add_argument('--iters', type=int, default=0)
    ap.add_argument('--graph', action='store_true', help='time CUDA-graph replays (GPU time, no host launch cost)')
    a = ap.parse_args()
    global ITERS, GRAPH
    ITERS = a.iters or None
    GRAPH = a.graph
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(
        os.path.join(ROOT, 'MEASURED_PEAKS.json')) else dict(hbm_gbs=6650.0, bf16_tflops=1590.0)
    res = []
    what = a.what.split(',')
    shapes = [(4096, 4096), (11008, 4096), (4096, 11008)]
    if 'qgemm' in what:
        for (N, K) in shapes:
            copies = max(2, int(300e6 // (N * K // 4)))
            for M in (1, 8, 16, 32):
                res.append(bench_qgemm(N, K, M, 2, 1, copies, peaks)); print(res[-1], flush=True)
            for M in (64, 128, 256, 2048):
                res.append(bench_qgemm(N, K, M, 2, 2, 2, peaks)); print(res[-1], flush=True)
        for bits in (3, 4):
            res.append(bench_qgemm(4096, 4096, 1, bits, 1, 20, peaks)); print(res[-1], flush=True)
            res.append(bench_qgemm(4096, 4096, 2048, bits, 2, 2, peaks)); print(res[-1], flush=True)
        # a stacked matrix (32 x 11008 rows): the skinny kernel's steady-state bandwidth, launch ramp amortised
        res.append(bench_qgemm(32 * 11008, 4096, 1, 2, 1, 2, peaks)); print(res[-1], flush=True)
        res.append(bench_qgemm(32 * 11008, 4096, 16, 2, 1, 2, peaks)); print(res[-1], flush=True)
    if 'gemv' in what:
        lib = _lib.load()
        keys = ('gemv', 'gv_int', 'gv_rbc', 'gv_persist', 'gv_tma', 'gv_cw', 'gv_stream')
        defaults = dict(gemv=1, gv_int=1, gv_rbc=0, gv_persist=1, gv_tma=1, gv_cw=16, gv_stream=32)

        def run(N, K, M, bits, copies, **cfg):
            for k in keys:
                lib.quip_config(k.encode(), cfg.get(k, defaults[k]))
            r = bench_qgemm(N, K, M, bits, 1, copies, peaks); r.update(cfg); res.append(r)
            print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()
                   if k not in ('TFLOPs', 'tensor_frac', 'kind', 'path')}, flush=True)
        if 'route' in what:
            for (N, K) in [(4096, 4096), (11008, 4096), (4096, 11008), (12288, 4096), (22016, 4096)]:
                copies = max(2, int(300e6 // (N * K // 4)))
                for M in (1, 2):
                    run(N, K, M, 2, copies, gv_tma=1, gv_stream=0)
                    run(N, K, M, 2, copies, gv_tma=0, gv_stream=0, gv_rbc=1)
                    run(N, K, M, 2, copies, gv_tma=0, gv_stream=0, gv_rbc=2)
                    run(N, K, M, 2, copies, gv_stream=1)
        if 'probe' in what:
            for bits in (2, 4):
                for st in (0, 8):
                    for M in (1, 2, 4):
                        run(32 * 11008, 4096, M, bits, 2, gv_stream=st)
            run(32 * 4096, 11008, 1, 2, 2, gv_stream=32)
        if 'gemv3' in what:
            for (N, K) in [(32 * 11008, 4096), (32 * 4096, 11008), (3 * 4096, 4096), (2 * 11008, 4096)]:
                copies = max(2, int(300e6 // (N * K // 4)))
                for M in (1, 2, 4):
                    for st in (0, 1, 8):
                        run(N, K, M, 2, copies, gv_stream=st)
            run(32 * 11008, 4096, 1, 4, 2, gv_stream=32)
        if 'gemv2' in what:
            for (N, K) in shapes + [(32 * 11008, 4096)]:
                copies = max(2, int(300e6 // (N * K // 4)))
                for M in (1, 4):
                    for cw in (8, 16):
                        for rbc in (1, 2):
                            run(N, K, M, 2, copies, gv_cw=cw, gv_rbc=rbc)
        for (N, K) in ([] if ('gemv2' in what or 'gemv3' in what or 'probe' in what or 'route' in what) else shapes + [(32 * 11008, 4096)]):
            copies = max(2, int(300e6 // (N * K // 4)))
            for M in (1, 2, 4, 8):
                for rbc in (1, 2):
                    if M <= 5:
                        run(N, K, M, 2, copies, gv_tma=1, gv_rbc=rbc)
                        run(N, K, M, 2, copies, gv_tma=0, gv_rbc=rbc)
                    run(N, K, M, 2, copies, gv_int=0, gv_rbc=rbc)
        for bits in (3, 4):
            for g in (() if ('gemv3' in what or 'probe' in what or 'route' in what) else (1,) if 'gemv2' in what else (0, 1)):
                run(11008, 4096, 1, bits, 8, gemv=g)
        for k in keys:
            lib.quip_config(k.encode(), defaults[k])
    if 'prof_tc' in what:
        res.append(bench_qgemm(4096, 4096, 2048, 2, 2, 2, peaks)); print(res[-1], flush=True)
        res.append(bench_qgemm(11008, 4096, 2048, 2, 2, 2, peaks)); print(res[-1], flush=True)
    if 'tune' in what:
        lib = _lib.load()
        for mt in (1, 2, 4):
            lib.quip_config(b'pass_min_tiles', mt)
            for (n, p, nblk, st) in [(4096, 64, 64, False), (4096, 64, 64, True), (11008, 16, 688, True)]:
                r = bench_pass(n, p, nblk, st, 2048); r['pass_min_tiles'] = mt; res.append(r); print(r, flush=True)
        lib.quip_config(b'pass_min_tiles', 4)
        for R in (1, 2, 4, 8):
            lib.quip_config(b'gather_rows', R)
            for n in (4096, 11008):
                r = bench_gather(n, 2048); r['gather_rows'] = R; res.append(r); print(r, flush=True)
        lib.quip_config(b'gather_rows', 0)
    if 'prof_skinny' in what:
        res.append(bench_qgemm(11008, 4096, 1, 2, 1, 8, peaks)); print(res[-1], flush=True)
        res.append(bench_qgemm(32 * 11008, 4096, 1, 2, 1, 2, peaks)); print(res[-1], flush=True)
    if 'dense' in what:
        for (N, K) in shapes:
            for M in (1, 2048):
                res.append(bench_dense(N, K, M, peaks)); print(res[-1], flush=True)
    if 'pass2048' in what:
        for (n, p, nblk, st) in [(4096, 64, 64, False), (4096, 64, 64, True), (11008, 688, 16, False), (11008, 16, 688, True)]:
            res.append(bench_pass(n, p, nblk, st, 2048)); print(res[-1], flush=True)
        res.append(bench_gather(4096, 2048)); print(res[-1], flush=True)
    if 'pass' in what:
        for M in (1, 2048):
            for (n, p, nblk, st) in [(4096, 64, 64, False), (4096, 64, 64, True), (11008, 688, 16, False), (11008, 16, 688, True)]:
                res.append(bench_pass(n, p, nblk, st, M)); print(res[-1], flush=True)
            res.append(bench_gather(4096, M)); print(res[-1], flush=True)
            res.append(bench_gather(11008, M)); print(res[-1], flush=True)
    if 'skinny' in what:
        # 9..32 tokens (batched decode): the split-K contraction alone (by number of K splits) and whole QuantLinears
        lib = _lib.load()
        for (N, K) in shapes:
            copies = max(2, int(300e6 // (N * K // 4)))
            for M in (12, 16, 24, 32):
                for ks in (0, 1, 2, 4):
                    lib.quip_config(b'sk_ksplit', ks)
                    try:
                        r = bench_qgemm(N, K, M, 2, 1, copies, peaks); r['sk_ksplit'] = ks; res.append(r)
                    except Exception as e:                      # a K slice that does not fit shared memory
                        print(dict(N=N, K=K, M=M, sk_ksplit=ks, error=str(e)[-60:]), flush=True)
                        continue
                    print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items() if k in ('N', 'K', 'M', 'us', 'hbm_frac', 'sk_ksplit')}, flush=True)
        lib.quip_config(b'sk_ksplit', 0)
        for (N, K) in shapes:
            for M in (8, 16, 32):
                r = bench_layer(N, K, M, 2, 'blocked', peaks, copies=4); res.append(r); print(r, flush=True)
    if 'structure' in what:
        # blocked (one factor block per position: n(p1+p2) factor elements per side) against Kronecker (one block per stage,
        # p1^2+p2^2: every CTA of a pass reads the same block) butterflies on the three Llama-2-7B QuantLinear shapes
        for (N, K) in shapes:
            for M in (1, 2048):
                for incoh in ('blocked', 'kron'):
                    r = bench_layer(N, K, M, 2, incoh, peaks, copies=4 if M == 1 else 2); res.append(r); print(r, flush=True)
    if 'decode' in what:
        # one token through the three Llama-2-7B QuantLinear shapes, blocked butterflies + rescale (the as-run default)
        lib = _lib.load()
        for (sf, pdl) in ((1, 1), (0, 1), (1, 0)):        # one-launch sides (rot_side_fewtok.cu) / two passes per side / no PDL
            lib.quip_config(b'pdl', pdl)
            lib.quip_config(b'side_fewtok', sf)
            for (N, K) in shapes:
                for M in (1, 4, 8):
                    r = bench_layer(N, K, M, 2, 'blocked', peaks, copies=4); r['pdl'] = pdl; r['side_fewtok'] = sf
                    res.append(r); print(r, flush=True)
        lib.quip_config(b'pdl', 1)
        lib.quip_config(b'side_fewtok', 0)
    if 'side' in what:
        lib = _lib.load()
        for sf in (0, 1):
            lib.quip_config(b'side_fused', sf)
            for (N, K) in [(4096, 4096), (11008, 4096), (4096, 11008)]:
                r = bench_layer(N, K, 2048, 2, 'blocked', peaks); r['side_fused'] = sf; res.append(r); print(r, flush=True)
        lib.quip_config(b'side_fused', 1)
    if 'batchdecode' in what:
        # 9..32 tokens through one QuantLinear: default route (16-token-tile side kernel) vs the few-token passes
        lib = _lib.load()
        for lim in (8, 32):
            lib.quip_config(b'fewtok_max_m', lim)
            for (N, K) in shapes:
                for M in (8, 16, 32):
                    r = bench_layer(N, K, M, 2, 'blocked', peaks, copies=4); r['fewtok_max_m'] = lim
                    res.append(r); print(r, flush=True)
        lib.quip_config(b'fewtok_max_m', 32)
    if 'glue' in what:
        # glue kernels of csrc/glue.cu against the torch launches they replace, Llama-2-7B sizes at 2048 tokens
        from transformers.models.llama import modeling_llama as ML
        from quip_b200.fused import CudaGlue

        class HFGlue:                                 # the HF modules themselves: what the kernels replace
            def rmsnorm(self, x, weight, eps, residual=None):
                norm = ML.LlamaRMSNorm(x.shape[-1], eps=eps).to(x.device, x.dtype)
                norm.weight.data = weight
                if residual is None:
                    return norm(x)
                s = residual + x
                return s, norm(s)

            def rope_(self, q, k, cos, sin, head_dim):
                S_ = q.shape[1]
                qh = q.view(1, S_, -1, head_dim).transpose(1, 2)
                kh = k.view(1, S_, -1, head_dim).transpose(1, 2)
                return ML.apply_rotary_pos_emb(qh, kh, cos[None], sin[None])

            def silu_mul(self, gate, up):
                return torch.nn.functional.silu(gate) * up

        TorchGlue = HFGlue
        S, d, inter, nh, hd = 2048, 4096, 11008, 32, 128
        cg, tg = CudaGlue(), TorchGlue()
        cp = 24                                       # rotate over > L2 of distinct buffers
        xs = [torch.randn(1, S, d, device='cuda').half() for _ in range(cp)]
        w = torch.randn(d, device='cuda').half()
        gs = [torch.randn(1, S, inter, device='cuda').half() for _ in range(6)]
        cos, sin = torch.randn(S, hd, device='cuda').half(), torch.randn(S, hd, device='cuda').half()
        hbm = peaks['hbm_gbs']
        cases = [
            ('rmsnorm', 2 * S * d * 2, lambda o: (lambda i: o.rmsnorm(xs[i % cp], w, 1e-5))),
            ('add_rmsnorm', 4 * S * d * 2, lambda o: (lambda i: o.rmsnorm(xs[i % cp], w, 1e-5, residual=xs[(i + 1) % cp]))),
            ('rope', 4 * S * d * 2, lambda o: (lambda i: o.rope_(xs[i % cp], xs[(i + 7) % cp], cos, sin, hd))),
            ('silu_mul', 3 * S * inter * 2, lambda o: (lambda i: o.silu_mul(gs[i % 6], gs[(i + 1) % 6]))),
        ]
        for name, nbytes, mk in cases:
            t_c, t_t = timeit(mk(cg)), timeit(mk(tg))
            r = dict(what='glue', op=name, us_kernel=t_c, us_torch_ops=t_t, bytes=nbytes, gbs=nbytes / t_c * 1e-3,
                     hbm_frac=nbytes / t_c * 1e-3 / hbm)
            res.append(r); print(r, flush=True)
    if 'stack' in what:
        # one 2048-token sample through 4 packed Llama-2-7B decoder layers: HF layers vs the fused stack, graph replay
        from transformers import LlamaConfig
        from quip_b200 import evalloop
        from quip_b200.quant import group_siblings
        from quip_b200.synth import LLAMA2_7B, build_synthetic_model
        cfg = LlamaConfig(**{**LLAMA2_7B, 'num_hidden_layers': 4})
        model = build_synthetic_model(cfg, torch.device('cuda:0'), bits=2, incoh='blocked', rescale=True, seed=0, seqlen=2048)
        group_siblings(model)
        ids = torch.randint(0, cfg.vocab_size, (1, 2048), device='cuda')
        for flag in ('0', '1'):
            os.environ['QUIP_FUSED_LAYER'] = flag
            with torch.no_grad():
                for _ in range(2):
                    nll = evalloop.sample_nll(model, evalloop.LLAMA, ids)
                step = evalloop.GraphedSampleNLL(model, evalloop.LLAMA, ids)
                t = timeit(lambda i: step.graph.replay(), iters=10)
            r = dict(what='stack', fused=int(flag), layers=4, us_per_step=t, us_per_layer=t / 4, nll=float(nll))
            res.append(r); print(r, flush=True)
        os.environ.pop('QUIP_FUSED_LAYER', None)
    if 'layer' in what:
        for (N, K) in shapes:
            for M in (1, 2048):
                res.append(bench_layer(N, K, M, 2, 'blocked', peaks)); print(res[-1], flush=True)
        res.append(bench_layer(4096, 4096, 2048, 2, None, peaks)); print(res[-1], flush=True)
        res.append(bench_layer(4096, 4096, 2048, 2, 'kron', peaks)); print(res[-1], flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
