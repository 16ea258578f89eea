"""Layer forward restatements (CPU oracle, numpy).

`parts` is the captured state of one quantized Linear (see quip_b200/capture.py):
  bits, codes (N,K) uint8, scales (N,1) f32, zeros (N,1) f32  [W2 = scales*c - zeros
  is the reference's grid matrix Q in the incoherent basis], bias (N,) or None,
  scaleWH (K,) f32 or None, U / V butterfly triples ([B0,B1], p_in, p_out) or None.

  w_ref()            reference postproc, method.py:195-214:
                         W_ref = fp16( fp16(U^T Q V) / scaleWH )
  dense_forward()    what the reference actually runs at inference: F.linear on
                     W_ref (bal.py:44-45 then HF nn.Linear)
  factored_forward() y = ((x / s) V^T) Q^T U + b in float64, no intermediate rounding
  kernel_forward()   the same through the kernel-side plan (layouts, folded
                     permutations, (c-cbar)/2^b contraction, P/R epilogue), optionally
                     rounding to fp16 where the CUDA pipeline does

Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np

from . import butterfly as bf
from . import qmath

f16, f32, f64 = np.float16, np.float32, np.float64


def grid_matrix(parts, dtype=f32):
    """Q = scales*c - zeros, rounded to fp16 like quantize_weight_vecbal's `.half()`."""
    c = parts['codes'].astype(f32)
    q = (parts['scales'].astype(f32) * c - parts['zeros'].astype(f32)).astype(f16)
    return q.astype(dtype)


def w_ref(parts, Q=None):
    """method.py:195-214.  `Q` may pass the reference's own fp16 grid matrix."""
    w = (grid_matrix(parts) if Q is None else np.asarray(Q)).astype(f32)
    if parts.get('U') is not None:
        N, K = w.shape
        U = bf.dense(parts['U'], N)
        V = bf.dense(parts['V'], K)
        w = (U.T @ w @ V).astype(f32)                       # method.py:202
    w = w.astype(f16)                                       # :204
    if parts.get('scaleWH') is not None:
        w = (w.astype(f32) / parts['scaleWH'].astype(f32)[None, :]).astype(f16)   # :210-213
    return w


def dense_forward(x, W, bias=None):
    y = np.asarray(x, f32) @ np.asarray(W, f32).T
    if bias is not None:
        y = y + np.asarray(bias, f32)
    return y.astype(f16)


def factored_forward(x, parts, dtype=f64):
    x = np.asarray(x).astype(dtype)
    Q = (parts['scales'].astype(dtype) * parts['codes'].astype(dtype) - parts['zeros'].astype(dtype))
    if parts.get('scaleWH') is not None:
        x = x / parts['scaleWH'].astype(dtype)[None, :]
    if parts.get('V') is not None:
        x = bf.rows_times_Vt(x, _cast(parts['V'], dtype))
    z = x @ Q.T
    if parts.get('U') is not None:
        z = bf.rows_times_U(z, _cast(parts['U'], dtype))
    if parts.get('bias') is not None:
        z = z + parts['bias'].astype(dtype)
    return z


def _cast(Bpp, dtype):
    (B, p_in, p_out) = Bpp
    return ([np.asarray(b).astype(dtype) for b in B], p_in, p_out)


def _r(a, on):
    return a.astype(f16).astype(f32) if on else a


def kernel_plan(parts, fold=True):
    """Everything the packed module stores, derived from captured parts.  `fold`: multiply 1/scaleWH
    into the columns of the first V pass when every block has its own factor (what
    quip_b200.incoherence.fold_inv_scale does), instead of scaling x in the gather."""
    bits = parts['bits']
    codes = parts['codes']
    N, K = codes.shape
    scales = parts['scales'].reshape(-1).astype(f32)
    zeros = parts['zeros'].reshape(-1).astype(f32)
    plan = dict(bits=bits, N=N, K=K, V=None, U=None)
    if parts.get('V') is not None:
        plan['V'] = bf.side_plan(parts['V'], K, 'V')
        codes = codes[:, plan['V']['order']]
    if parts.get('U') is not None:
        plan['U'] = bf.side_plan(parts['U'], N, 'U')
        o = plan['U']['order']
        codes, scales, zeros = codes[o], scales[o], zeros[o]
    plan['codes'] = np.ascontiguousarray(codes)
    plan['P'], plan['R'] = qmath.kernel_affine(scales, zeros, bits)
    plan['inv_scale'] = None if parts.get('scaleWH') is None else (f32(1) / parts['scaleWH'].astype(f32))
    plan['bias'] = parts.get('bias')
    if fold and plan['inv_scale'] is not None and plan['V'] is not None:
        v = plan['V']
        F, p, nblk, strided = v['passes'][0]
        s = plan['inv_scale']
        if F.shape[0] == nblk and s.max() / s.min() <= 1e3:
            per_pos = s[v['io_idx']]
            cols = per_pos.reshape(p, nblk).T if strided else per_pos.reshape(nblk, p)
            v['passes'][0] = ((F * cols[:, None, :]).astype(f32), p, nblk, strided)
            plan['inv_scale'] = None
    return plan


def kernel_forward(x, parts, fp16_points=True, plan=None):
    """Emulates the CUDA pipeline stage by stage (fp32 math, fp16 rounding points)."""
    plan = plan or kernel_plan(parts)
    bits, N, K = plan['bits'], plan['N'], plan['K']
    x = np.asarray(x).astype(f32)
    # --- K side: [scale] -> gather -> passes ---
    if plan['inv_scale'] is not None:
        x = x * plan['inv_scale'][None, :]
    if plan['V'] is not None:
        v = plan['V']
        x = _r(x[:, v['io_idx']], fp16_points)
        for F, p, nblk, strided in v['passes']:
            x = _r(bf.apply_pass(x, _r(F, fp16_points), p, nblk, strided), fp16_points)
    else:
        x = _r(x, fp16_points)
    # --- contraction against d = (c - cbar)/2^b, then the per-row affine epilogue ---
    cbar = f32(qmath.maxq_of(bits)) / f32(2)
    d = (plan['codes'].astype(f32) - cbar) / f32(1 << bits)
    S = x @ d.T
    z = plan['P'][None, :] * S + plan['R'][None, :] * x.sum(axis=1, keepdims=True)
    # --- N side: passes -> scatter -> bias ---
    if plan['U'] is not None:
        u = plan['U']
        z = _r(z, fp16_points)
        for i, (F, p, nblk, strided) in enumerate(u['passes']):
            z = bf.apply_pass(z, _r(F, fp16_points), p, nblk, strided)
            if i + 1 < len(u['passes']):
                z = _r(z, fp16_points)
        y = np.empty_like(z)
        y[:, u['io_idx']] = z
    else:
        y = z
    if plan['bias'] is not None:
        y = y + plan['bias'].astype(f32)[None, :]
    return y.astype(f16) if fp16_points else y


def rel_err(y, y_ref):
    y, y_ref = np.asarray(y, f64), np.asarray(y_ref, f64)
    return float(np.linalg.norm(y - y_ref) / max(np.linalg.norm(y_ref), 1e-30))
