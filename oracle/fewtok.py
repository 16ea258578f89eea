"""Numpy restatements of the two few-token contraction schemes of quip_b200/csrc/qgemv.cu (CPU oracle; test
infrastructure only, see oracle/__init__.py).  They pin the *arithmetic* of the kernels -- what is exact, what is
rounded and how large the error can get -- independently of the CUDA code; the GPU parity tests compare the kernels
with the float64 contraction `x @ (scales * codes - zeros).T + bias` (reference semantics: Quant3Linear, quant.py:186-191
and :222-233).

1. int8 tensor-core path (1-5 tokens).  A token is scaled per row by 2^22 / amax and split into three balanced signed
   bytes (hi, mid, lo) with x_q = 65536 hi + 256 mid + lo; the sums of code x byte products are exact integers;
       y = scales * s * (65536 I_hi + 256 I_mid + I_lo) - zeros * S + bias,    s = amax / 2^22, S = sum_k x_k.
2. offset-free fp16 path (6-8 tokens).  A code is read in place inside an fp16 mantissa as 1 + c / 4^(e+1) with e = 0 or 2
   by bit position; the token operand of that k is pre-scaled by 4^(e-2) (exact unless it underflows); products are exact
   in fp32 and sum to T + sum(c x) / 64 with T = sum_k 4^(e(k)-2) x_k.
"""
import numpy as np

QMAX = float(1 << 22)


def limb_split(x):
    """x (M, K) fp16/float -> (hi, mid, lo int8 arrays, s (M,) float32).  Balanced base-256 digits of
    round(x * 2^22 / amax): lo, mid in [-128, 127], hi in [-64, 64]."""
    x = np.asarray(x, np.float32)
    amax = np.abs(x).max(axis=1)
    inv = np.where(amax > 0, np.float32(QMAX) / amax, 0).astype(np.float32)
    q = np.rint(x * inv[:, None]).astype(np.int64)
    lo = ((q + 128) % 256) - 128
    q1 = (q - lo) >> 8
    mid = ((q1 + 128) % 256) - 128
    hi = (q1 - mid) >> 8
    assert np.abs(hi).max(initial=0) <= 64
    return hi.astype(np.int8), mid.astype(np.int8), lo.astype(np.int8), (amax / np.float32(QMAX)).astype(np.float32)


def int8_gemv(codes, scales, zeros, x, bias=None):
    """The IMMA scheme: integer dot products of codes (N, K) with the three limbs, fp32 epilogue, fp16 output."""
    c = np.asarray(codes).astype(np.int64)
    hi, mid, lo, s = limb_split(x)
    I = [c @ l.astype(np.int64).T for l in (hi, mid, lo)]                     # (N, M) exact
    dot = (np.float32(65536.0) * I[0].astype(np.float32) + np.float32(256.0) * I[1].astype(np.float32)
           + I[2].astype(np.float32)) * s[None, :]
    S = np.asarray(x, np.float32).sum(axis=1)
    y = np.asarray(scales, np.float32).reshape(-1, 1) * dot - np.asarray(zeros, np.float32).reshape(-1, 1) * S[None, :]
    if bias is not None:
        y = y + np.asarray(bias, np.float32).reshape(-1, 1)
    return y.T.astype(np.float16)


def k_exponent(K, bits):
    """e(k) of the offset-free expansion: by position inside a lane's run of 8 k (pairs u = 0..3)."""
    u = (np.arange(K) % 8) // 2
    if bits == 2:
        return np.where(u < 2, 2, 0)
    if bits == 4:
        return np.where(u % 2 == 0, 2, 0)
    raise ValueError('the offset-free expansion covers 2 and 4 bits')


def offset_free_gemv(codes, scales, zeros, x, bits, bias=None):
    """A = 1 + c / 2^bits / 4^e (exact in fp16), B = fp16(x * 4^(e-2)), fp32 accumulation, then the T / S correction."""
    e = k_exponent(np.asarray(codes).shape[1], bits)
    A = (1.0 + np.asarray(codes, np.float64) / (2 ** bits) / (4.0 ** e)[None, :]).astype(np.float16)
    assert np.array_equal(A.astype(np.float64), 1.0 + np.asarray(codes, np.float64) / (2 ** bits) / (4.0 ** e)[None, :])
    B = (np.asarray(x, np.float16).astype(np.float32) * (4.0 ** (e - 2)).astype(np.float32)[None, :]).astype(np.float16)
    acc = (A.astype(np.float32) @ B.astype(np.float32).T)                    # fp32 accumulate (order differs on the GPU)
    T = B.astype(np.float32).sum(axis=1)
    S = np.asarray(x, np.float16).astype(np.float32).sum(axis=1)
    A1 = np.float32(16.0 * 2 ** bits)
    y = np.asarray(scales, np.float32).reshape(-1, 1) * (A1 * (acc - T[None, :])) - np.asarray(zeros, np.float32).reshape(-1, 1) * S[None, :]
    if bias is not None:
        y = y + np.asarray(bias, np.float32).reshape(-1, 1)
    return y.T.astype(np.float16)


def reference_gemv(codes, scales, zeros, x, bias=None):
    Q = np.asarray(scales, np.float64).reshape(-1, 1) * np.asarray(codes, np.float64) - np.asarray(zeros, np.float64).reshape(-1, 1)
    y = np.asarray(x, np.float64) @ Q.T
    if bias is not None:
        y = y + np.asarray(bias, np.float64)[None, :]
    return y
