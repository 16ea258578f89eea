"""Thin helpers that call the C ABI directly with torch CUDA tensors (used by the -m gpu tests)."""
import ctypes as C

import numpy as np
import torch

from quip_b200 import _lib
from quip_b200 import quant as Q

DEV = 'cuda:0'


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def t16(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV).half()


def run_pass(X, F, p, nblk, strided, impl=0):
    """One block-diagonal pass on X (M, n) fp16 numpy -> numpy fp16."""
    lib = _lib.load()
    M, n = X.shape
    x = t16(X)
    f = t16(F).contiguous()
    out = torch.empty_like(x)
    ps = _lib.QuipPass(p=p, nblk=nblk, strided=int(strided), shared=int(F.shape[0] == 1 and nblk > 1),
                       factors=f.data_ptr())
    _lib.check(lib.quip_rot_pass(C.byref(ps), _lib.ptr(x), _lib.ptr(out), M, n, impl, stream()))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def run_gather(X, idx=None, scale=None, bias=None):
    lib = _lib.load()
    M, n = X.shape
    x = t16(X)
    out = torch.empty_like(x)
    ti = None if idx is None else torch.from_numpy(np.asarray(idx, np.int32)).to(DEV)
    ts = None if scale is None else torch.from_numpy(np.asarray(scale, np.float32)).to(DEV)
    tb = None if bias is None else t16(bias)
    _lib.check(lib.quip_gather(_lib.ptr(x), _lib.ptr(out), M, n, _lib.ptr(ti), _lib.ptr(ts), _lib.ptr(tb), stream()))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def run_qgemm(codes, scales, zeros, bits, X, path, bias=None, symmetric=False):
    """z = x . Q^T (+bias) through quip_qgemm; codes (N,K) uint8, X (M,K) fp16 numpy."""
    lib = _lib.load()
    N, K = codes.shape
    M = X.shape[0]
    qw = Q.pack_codes(torch.from_numpy(codes).to(DEV), bits)
    sc = torch.from_numpy(np.asarray(scales, np.float32).reshape(-1)).to(DEV)
    ze = torch.from_numpy(np.asarray(zeros, np.float32).reshape(-1)).to(DEV)
    x = t16(X)
    tb = None if bias is None else t16(bias)
    d = _lib.QuipLinearDesc()
    d.K, d.N, d.bits, d.flags = K, N, bits, (_lib.QUIP_FLAG_SYMMETRIC if symmetric else 0)
    d.qweight, d.scales, d.zeros = qw.data_ptr(), sc.data_ptr(), ze.data_ptr()
    need = C.c_size_t()
    _lib.check(lib.quip_qlinear_workspace_bytes(C.byref(d), M, C.byref(need)))
    ws = torch.zeros(need.value, dtype=torch.uint8, device=DEV)
    xsum = torch.empty(M, dtype=torch.float32, device=DEV)
    _lib.check(lib.quip_rowsum(_lib.ptr(x), _lib.ptr(xsum), M, K, stream()))
    z = torch.full((M, N), float('nan'), dtype=torch.float16, device=DEV)
    _lib.check(lib.quip_qgemm(C.byref(d), _lib.ptr(x), _lib.ptr(xsum), _lib.ptr(tb), _lib.ptr(z), M, path,
                              _lib.ptr(ws), ws.numel(), stream()))
    torch.cuda.synchronize()
    assert int(ws[:_lib.WS_HEADER_BYTES].max()) == 0, 'split-K counters must be left zeroed'
    return z.cpu().numpy(), xsum.cpu().numpy()


def torch_reference_forward(ql, x):
    """fp32 restatement of QuantLinear.forward in plain torch (quip_b200/selfcheck.py, shared with bench.py's self-check)."""
    from quip_b200.selfcheck import restated_forward
    return restated_forward(ql, x)


def run_qgemm_dev(codes, scales, zeros, bits, x, path, bias=None, symmetric=False):
    """quip_qgemm on DEVICE tensors (the Llama-sized cases: nothing goes through numpy).  codes (N, K) uint8, scales / zeros
    (N) fp32, x (M, K) fp16, bias (N) fp16 or None -> z (M, N) fp16 on the device."""
    lib = _lib.load()
    N, K = codes.shape
    M = x.shape[0]
    qw = Q.pack_codes(codes, bits)
    d = _lib.QuipLinearDesc()
    d.K, d.N, d.bits, d.flags = K, N, bits, (_lib.QUIP_FLAG_SYMMETRIC if symmetric else 0)
    d.qweight, d.scales, d.zeros = qw.data_ptr(), scales.data_ptr(), zeros.data_ptr()
    need = C.c_size_t()
    _lib.check(lib.quip_qlinear_workspace_bytes(C.byref(d), M, C.byref(need)))
    ws = torch.zeros(max(need.value, 1 << 20), dtype=torch.uint8, device=x.device)
    xsum = torch.empty(M, dtype=torch.float32, device=x.device)
    _lib.check(lib.quip_rowsum(_lib.ptr(x), _lib.ptr(xsum), M, K, stream()))
    z = torch.full((M, N), float('nan'), dtype=torch.float16, device=x.device)
    _lib.check(lib.quip_qgemm(C.byref(d), _lib.ptr(x), _lib.ptr(xsum), _lib.ptr(bias), _lib.ptr(z), M, path,
                              _lib.ptr(ws), ws.numel(), stream()))
    torch.cuda.synchronize()
    return z
