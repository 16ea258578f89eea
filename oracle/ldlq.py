"""LDLQ / LDLQ-RG adaptive rounding on the CPU (oracle; test infrastructure only, see oracle/__init__.py).

SURVEY section 8(f) rank 1: the quantisation-time hot loop the reference runs as d sequential column updates in
Python (vector_balance.py:155-199 `round_ldl`, :218-291 the blocked variant, :139-152 / :202-215 the sorted
"RG" wrappers, :500-530 the dispatch in quantize_weight_vecbal).  This file restates the algorithm; it is the checker for the integer codes of the GPU
implementation (quip_b200/csrc/ldlq.cu through quip_b200/quantize.py, tests/test_gpu_quantize.py):

  * L = strictly lower part of the unit-diagonal Cholesky factor of H (vector_balance.py:174-176);
  * columns are rounded last to first, column i seeing the feedback (w - w_hat)[:, i:] @ L[i:, i]
    (vector_balance.py:182-183), nearest rounding = floor(. + 1/2), clamped to [0, 2^b - 1];
  * optional greedy coordinate passes on H / max(diag H) (vector_balance.py:185-202): a column moves to
    round(wr_i - (s @ H[:, i]) / H_ii) with s = wr - w kept incrementally; the clamp at the end of a pass does
    NOT update s (reference behaviour, kept);
  * "RG" = the same on columns sorted by ascending diag(H) (vector_balance.py:139-152).

torch (CPU, float32) is used rather than numpy so that every rounding decision reproduces the reference's own
float32 arithmetic; the golden codes under tests/golden/ldlq.npz come from the live reference (oracle/gen_golden_ldlq.py).
"""
import torch


def ldl_feedback_matrix(H):
    """L - I with L the unit-lower-triangular LDL factor of H."""
    C = torch.linalg.cholesky(H)
    L = C @ torch.diag(1.0 / torch.diag(C))
    return L - torch.eye(H.shape[0], dtype=H.dtype)


def ldlq_round(w, H, nbits, greedy_passes=0):
    """w (m, d) float32 in grid units (already w/scale + zero), H (d, d) float32 -> integer-valued (m, d) float32."""
    w = w.float()
    H = H.float()
    d = H.shape[0]
    top = float(2 ** nbits - 1)
    Lf = ldl_feedback_matrix(H)
    q = w.clone()
    for i in range(d - 1, -1, -1):
        fb = (w[:, i:] - q[:, i:]) @ Lf[i:, i]
        q[:, i] = torch.clamp(torch.floor(w[:, i] + fb + 0.5), min=0, max=top)
    out = q.clone()
    s = q - w
    Hn = H / H.diag().max()
    for _ in range(greedy_passes):
        for i in range(d - 1, -1, -1):
            move = out[:, i] - torch.round(out[:, i] - (s @ Hn[:, i]) / Hn[i, i])
            out[:, i] -= move
            s[:, i] -= move
        out = torch.clamp(out, min=0, max=top)
        if bool((q == out).all()):
            break
        q.copy_(out)
    return out


def ldlq_rg_round(w, H, nbits, greedy_passes=0):
    """LDLQ-RG: columns visited in ascending order of diag(H)."""
    p = torch.argsort(torch.diag(H))
    out = torch.zeros_like(w, dtype=torch.float32)
    out[:, p] = ldlq_round(w[:, p], H[p, :][:, p], nbits, greedy_passes)
    return out


def quantize_weight(w, H, nbits, greedy_passes, scale, zero, maxq, qfn='a', method='ldlq'):
    """The reference's quantize_weight_vecbal for the two LDL methods (vector_balance.py:500-530): returns
    (grid values fp16, integer codes uint8)."""
    rnd = ldlq_round if method == 'ldlq' else ldlq_rg_round
    if qfn == 'a':
        g = torch.clamp((w / scale) + zero, 0, maxq)
        codes = rnd(g.float(), H, nbits, greedy_passes)
        return (scale * (codes - zero)).half(), codes.to(torch.uint8)
    s = 2.4 * w.square().mean().sqrt() + 1e-16
    g = torch.clamp(((w / s) + 1) / 2 * maxq, 0, maxq)
    codes = rnd(g.float(), H, nbits, greedy_passes)
    return (((codes / maxq) * 2 - 1) * s).half(), codes.to(torch.uint8)
