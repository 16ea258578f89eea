"""Host-side planning of the incoherence un-projection for the packed kernels.

The reference multiplies by dense orthogonal U (N x N) and V (K x K) at quantization time and folds
them back into a dense fp16 weight (method.py:157-180, 195-214).  Here the *factors* of those
butterflies (method.py:34-67) are kept and applied to activations at run time:

    x @ V^T :  u = x[p_in];  T = u.view(p1,p2);  T[:,b] = B0[b] T[:,b];  T[a,:] = B1[a] T[a,:];  T.flat[p_out]
    z @ U   :  w[p_out] = z; T = w.view(q1,q2);  T[a,:] = B1[a]^T T[a,:];  T[:,b] = B0[b]^T T[:,b];  y[p_in] = T.flat

`plan_side` turns one butterfly into what the C ABI's QuipSide needs:
  * a *layout* for the side's work buffers so that the larger block is contiguous
    ('A': index a*p2+b when p2 >= p1, else 'B': index b*p1+a);
  * `idx`, the gather index (V: layout[l] = x[idx[l]];  U: y[j] = layout[idx[j]]);
  * `order`, the permutation folded into the packed matrix (p_out disappears from the run time);
  * the two block-diagonal passes in execution order, as fp16 factors with out_i = sum_j f[i][j] in_j.
The numpy twin of this function, oracle/butterfly.py:side_plan, is its checker.
"""
from dataclasses import dataclass
from typing import List, Optional

import torch

from .capture import Butterfly


@dataclass
class PassPlan:
    factors: torch.Tensor     # (nb, p, p) float32, nb in {nblk, 1}
    p: int
    nblk: int
    strided: bool

    @property
    def shared(self):
        return self.factors.shape[0] == 1 and self.nblk > 1


@dataclass
class SidePlan:
    n: int
    layout: str
    idx: Optional[torch.Tensor]       # (n,) int64 gather index or None (identity)
    order: torch.Tensor               # (n,) int64: layout position l holds original column/row order[l]
    passes: List[PassPlan]


def plan_side(bf: Butterfly, side: str) -> SidePlan:
    n, p1, p2 = bf.n, bf.p1, bf.p2
    i = torch.arange(n)
    a, b = i // p2, i % p2
    layout = 'A' if p2 >= p1 else 'B'
    l_of_i = i if layout == 'A' else b * p1 + a
    inv_pout = torch.argsort(bf.p_out)
    io = torch.empty(n, dtype=torch.long)
    io[l_of_i] = bf.p_in
    order = torch.empty(n, dtype=torch.long)
    order[l_of_i] = inv_pout
    col = dict(p=p1, nblk=p2, strided=(layout == 'A'))
    row = dict(p=p2, nblk=p1, strided=(layout == 'B'))
    B0, B1 = bf.B0.float(), bf.B1.float()
    if side == 'V':
        passes = [PassPlan(B0.contiguous(), **col), PassPlan(B1.contiguous(), **row)]
        idx = io
    else:
        passes = [PassPlan(B1.transpose(1, 2).contiguous(), **row), PassPlan(B0.transpose(1, 2).contiguous(), **col)]
        idx = torch.argsort(io)          # scatter y[io[l]] = T[l]  ==  gather y[j] = T[argsort(io)[j]]
    if torch.equal(idx, torch.arange(n)):
        idx = None
    return SidePlan(n=n, layout=layout, idx=idx, order=order, passes=passes)


def fold_inv_scale(plan: SidePlan, inv_scale: torch.Tensor, ratio_limit: float = 1e3) -> bool:
    """Fold 1/scaleWH into the columns of the first V pass (only possible when every block has its
    own factor, i.e. the as-run blocked butterfly).  Saves one fp16 rounding of the
    activations.  Returns False (and leaves the plan alone) when factors are shared or the scale's
    dynamic range would cost fp16 precision in the factors."""
    ps = plan.passes[0]
    if ps.shared or ps.factors.shape[0] != ps.nblk:
        return False
    s = inv_scale.float()
    if not torch.isfinite(s).all() or s.min() <= 0 or (s.max() / s.min()) > ratio_limit:
        return False
    src = plan.idx if plan.idx is not None else torch.arange(plan.n)
    per_pos = s[src]                                  # scale of the feature sitting at layout position l
    if ps.strided:                                    # pos = j*nblk + blk
        cols = per_pos.view(ps.p, ps.nblk).t()        # (blk, j)
    else:                                             # pos = blk*p + j
        cols = per_pos.view(ps.nblk, ps.p)
    ps.factors = ps.factors * cols[:, None, :]
    return True
