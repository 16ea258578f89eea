"""NCCL sanity of the data-parallel and pipeline eval drivers on the packed tiny-OPT fixture (run under torchrun)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_tiny_opt          # noqa: E402
from quip_b200 import evalloop, pipeline    # noqa: E402
from quip_b200.opt import opt_eval, opt_pack  # noqa: E402

rank, world = pipeline.init_distributed()
dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
model, parts, ids, ref = load_tiny_opt()
opt_pack(model, parts)
single = opt_eval(model, ids, dev, verbose=False)
dp = pipeline.dp_eval(model, evalloop.OPT, ids, dev)
model2, parts, ids, ref = load_tiny_opt()
opt_pack(model2, parts)
pp = pipeline.pp_eval(model2, evalloop.OPT, ids, dev)
if rank == 0:
    print(dict(world=world, reference_cpu_ppl=ref, single_gpu=single, dp=dp, pp=pp,
               dp_rel=abs(dp - single) / single, pp_rel=abs(pp - single) / single))
assert abs(dp - single) / single < 2e-3 and abs(pp - single) / single < 2e-3
torch.distributed.destroy_process_group()
