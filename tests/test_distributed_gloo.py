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


def _decode_worker(rank, world, port, family, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from quip_b200 import evalloop, pipeline
    pipeline.init_distributed(backend='gloo')
    torch.manual_seed(0)
    if family == 'opt':
        from transformers import OPTConfig, OPTForCausalLM
        model = OPTForCausalLM(OPTConfig(hidden_size=64, ffn_dim=128, num_hidden_layers=5, num_attention_heads=4, vocab_size=199,
                                         max_position_embeddings=32, word_embed_proj_dim=64)).float().eval()
        arch = evalloop.OPT
    else:
        from transformers import LlamaConfig, LlamaForCausalLM
        model = LlamaForCausalLM(LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=5, num_attention_heads=4,
                                             num_key_value_heads=2, vocab_size=199, max_position_embeddings=32)).float().eval()
        arch = evalloop.LLAMA
    ids = torch.randint(0, 199, (1, 10), generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        ref_sec, ref_ppl = evalloop.decode_benchmark(model, ids, check=True)      # the whole model in one process (benchmark())
        res = pipeline.pp_decode_benchmark(model, arch, ids, 'cpu', max_len=16, check=True)
        # logits of the last stage, token by token, against the HF forward with a KV cache
        pd = pipeline.PipelinedDecoder(model, arch, 'cpu', max_len=16, batch=1)
        past, worst = None, 0.0
        for i in range(ids.shape[1]):
            got = pd.step(ids[0, i:i + 1])
            o = model(ids[:, i:i + 1], past_key_values=past, use_cache=True)
            past = o.past_key_values
            if pd.last:
                worst = max(worst, float((got - o.logits[:, -1]).abs().max()))
            else:
                assert got is None
    out[rank] = (res, ref_ppl, worst)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize('family,world', [('opt', 2), ('llama', 3)])
def test_pipelined_decode_matches_the_single_process_benchmark(family, world):
    """benchmark() over a layer pipeline (opt_multigpu + benchmark, opt.py:384-482): every rank one GraphDecoder stage with its own
    KV cache, hidden states over the per-link groups; 5 layers on 2 / 3 ranks (uneven stages, a middle stage)."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_decode_worker, args=(world, _free_port(), family, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        res, ref_ppl, worst = out[rank]
        assert res['stages'] == world and res['tokens'] == 10 and res['latency_s'] > 0 and res['pipelined_s'] > 0
        assert abs(res['ppl'] - ref_ppl) / ref_ppl < 1e-4, (rank, res['ppl'], ref_ppl)
        assert worst < 2e-4, (rank, worst)
