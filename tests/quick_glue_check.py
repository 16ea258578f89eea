"""(A script, not a pytest module.)  Smallest possible device check of the staged glue kernels (csrc/glue.cu): a few seconds after `import torch`.
Appends one JSON line per finished check to gpurun_out/quick_glue.jsonl (flushed as it goes, so a cut-off run still
leaves what it finished).  No transformers import."""
import json
import os
import sys
import time

t_start = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
OUT = open(os.path.join(ROOT, 'gpurun_out', 'quick_glue.jsonl'), 'a')


def emit(**kw):
    kw['t'] = round(time.time() - t_start, 2)
    OUT.write(json.dumps(kw) + '\n')
    OUT.flush()
    os.fsync(OUT.fileno())
    print(kw, flush=True)


emit(stage='start')
import torch  # noqa: E402

emit(stage='torch imported', cuda=torch.cuda.is_available())
from oracle.glue import TorchGlue  # noqa: E402
from quip_b200.fused import CudaGlue  # noqa: E402

cg, tg = CudaGlue(), TorchGlue()
dev = 'cuda'
g = torch.Generator(device=dev).manual_seed(0)


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev, generator=g) * scale).half()


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


S, d, inter, nh, hd = 2048, 4096, 11008, 32, 128
# 1. rotary: bit-exact
q, k = rnd(1, S, nh * hd), rnd(1, S, nh * hd)
ang = torch.rand(S, hd // 2, device=dev, generator=g) * 100
cos, sin = torch.cat((ang.cos(), ang.cos()), -1).half(), torch.cat((ang.sin(), ang.sin()), -1).half()
q0, k0 = q.clone(), k.clone()
tg.rope_(q0, k0, cos, sin, hd)
q1, k1 = q.clone(), k.clone()
cg.rope_(q1, k1, cos, sin, hd)
torch.cuda.synchronize()
emit(check='rope', bit_exact=bool(torch.equal(q0, q1) and torch.equal(k0, k1)),
     mismatches=int((q0 != q1).sum() + (k0 != k1).sum()))
# 2. silu * up
a, b = rnd(1, S, inter, scale=3.0), rnd(1, S, inter)
w_, g_ = tg.silu_mul(a, b), cg.silu_mul(a, b)
torch.cuda.synchronize()
emit(check='silu_mul', bit_exact=bool(torch.equal(w_, g_)), mismatches=int((w_ != g_).sum()),
     max_abs=float((w_.float() - g_.float()).abs().max()))
# 3. rmsnorm with and without the residual
x, r, w = rnd(1, S, d, scale=2.0), rnd(1, S, d), rnd(d)
y0, y1 = tg.rmsnorm(x, w, 1e-5), cg.rmsnorm(x, w, 1e-5)
s0, z0 = tg.rmsnorm(x, w, 1e-5, residual=r)
s1, z1 = cg.rmsnorm(x, w, 1e-5, residual=r)
torch.cuda.synchronize()
emit(check='rmsnorm', mismatch_frac=float((y0 != y1).float().mean()), max_abs=float((y0.float() - y1.float()).abs().max()),
     sum_bit_exact=bool(torch.equal(s0, s1)), res_mismatch_frac=float((z0 != z1).float().mean()),
     res_max_abs=float((z0.float() - z1.float()).abs().max()), ref_abs_max=float(y0.float().abs().max()))
# 4. timings (hot in L2: small working sets; the microbench rotates buffers)
emit(check='time_us', rope_kernel=timeit(lambda: cg.rope_(q1, k1, cos, sin, hd)), rope_torch=timeit(lambda: tg.rope_(q0, k0, cos, sin, hd)),
     silu_kernel=timeit(lambda: cg.silu_mul(a, b)), silu_torch=timeit(lambda: tg.silu_mul(a, b)),
     rms_kernel=timeit(lambda: cg.rmsnorm(x, w, 1e-5)), rms_torch=timeit(lambda: tg.rmsnorm(x, w, 1e-5)),
     addrms_kernel=timeit(lambda: cg.rmsnorm(x, w, 1e-5, residual=r)), addrms_torch=timeit(lambda: tg.rmsnorm(x, w, 1e-5, residual=r)))
emit(stage='done')
