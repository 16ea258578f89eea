"""world_size-2 gloo tests (CPU) of the data-parallel and pipeline eval drivers."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from conftest import load_tiny_opt
    from quip_b200 import evalloop, pipeline
    pipeline.init_distributed(backend='gloo')
    model, parts, ids, ref_ppl = load_tiny_opt()
    dev = torch.device('cpu')
    if mode == 'dp':
        ppl = pipeline.dp_eval(model, evalloop.OPT, ids, dev)
    else:
        ppl = pipeline.pp_eval(model, evalloop.OPT, ids, dev)
    out[rank] = (ppl, ref_ppl)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize('mode,world', [('dp', 2), ('pp', 2), ('pp', 3)])
def test_multi_rank_eval_matches_single_process(mode, world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), mode, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        ppl, ref = out[rank]
        assert abs(ppl - ref) / ref < 1e-5, (mode, rank, ppl, ref)
    for rank in range(1, world):                      # a middle stage (world 3) both receives and sends
        assert abs(out[0][0] - out[rank][0]) < 1e-9 * out[0][0] + 1e-12


def test_stage_ranges_follow_the_reference_rule():
    from quip_b200.pipeline import stage_ranges
    assert stage_ranges(48, 8) == [(i * 6, i * 6 + 6) for i in range(8)]        # ceil(L/G), opt.py:424-426
    assert stage_ranges(32, 3) == [(0, 11), (11, 22), (22, 32)]
    assert stage_ranges(80, 3, '20:30:30') == [(0, 20), (20, 50), (50, 80)]    # --layers-dist, llama.py:400-413
    with pytest.raises(AssertionError):
        stage_ranges(80, 2, '20:30')
