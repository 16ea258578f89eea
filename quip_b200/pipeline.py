"""Multi-GPU drivers for the per-layer eval loop: one process per GPU, torch.distributed (NCCL on GPUs,
gloo for the CPU tests).

The reference has no distributed backend: cross-GPU movement is `tensor.to(device)` between layer
groups inside one process (opt.py:413-426, llama.py:370-413).  The path shards
two ways:

  * data parallel over samples (`dp_eval`): the eval loop's samples are independent (opt.py:262-264),
    the 2-bit model fits one B200 many times over, so every rank holds the packed weights and
    evaluates samples rank, rank+G, ...; the only collective is one all-reduce of (sum NLL, tokens).
  * layer pipeline (`pp_eval`), for models that do not fit one GPU: contiguous layer ranges per rank
    exactly as opt.py:424-426 (or --layers-dist, llama.py:400-413), micro-batch = one sample, hidden
    states (1,S,H) fp16 sent to the next stage with isend/irecv (ncclSend/ncclRecv over NVLink),
    receives posted one sample ahead so the transfer of sample j+1 overlaps the compute of sample j.
"""
import os

import torch
import torch.distributed as dist

from . import evalloop


def init_distributed(backend=None, device=None):
    """Initialise from the torchrun environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT)."""
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if world == 1:
        return 0, 1
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29500')
    kw = {}
    if backend == 'nccl':
        local = int(os.environ.get('LOCAL_RANK', rank))
        torch.cuda.set_device(local)
        kw['device_id'] = torch.device('cuda', local)
    dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world


def _allreduce_pair(nll, count):
    t = torch.stack([nll.float(), count.float()])
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t[0], t[1]


@torch.no_grad()
def dp_eval(model, arch, testenc, dev, verbose=False):
    """Data-parallel perplexity: samples round-robin over ranks, one all-reduce at the end."""
    ids = testenc.input_ids if hasattr(testenc, 'input_ids') else testenc
    nsamples = ids.numel() // model.seqlen
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    mine = list(range(rank, nsamples, world))
    return evalloop.eval_ppl(model, arch, ids, dev, sample_ids=mine, verbose=verbose and rank == 0,
                             reduce_fn=_allreduce_pair if world > 1 else None)


def stage_ranges(nlayers, world, layers_dist=None):
    if layers_dist:
        from .llama import parse_layers_dist
        return parse_layers_dist(layers_dist, nlayers)
    from .opt import layer_placement
    return layer_placement(nlayers, world)


def place_stage(model, arch, lo, hi, dev, first, last):
    """Move only what this pipeline stage needs to `dev`."""
    layers = arch.layers(model)
    for i in range(lo, hi):
        layers[i].to(dev)
    for mod in arch.pre(model):        # every stage derives the layer kwargs (positions / rotary) locally
        mod.to(dev)
    if last:
        for mod in arch.post(model):
            mod.to(dev)
        arch.head(model).to(dev)


@torch.no_grad()
def pp_eval(model, arch, testenc, dev, layers_dist=None, verbose=False):
    """Layer-pipelined perplexity over all ranks of the default process group."""
    ids = testenc.input_ids if hasattr(testenc, 'input_ids') else testenc
    seqlen = model.seqlen
    nsamples = ids.numel() // seqlen
    rank, world = dist.get_rank(), dist.get_world_size()
    layers = arch.layers(model)
    lo, hi = stage_ranges(len(layers), world, layers_dist)[rank]
    first, last = rank == 0, rank == world - 1
    place_stage(model, arch, lo, hi, dev, first, last)
    use_cache = model.config.use_cache
    model.config.use_cache = False
    dtype = next(iter(layers[lo].parameters())).dtype if hi > lo else torch.float16
    shape = (1, seqlen, model.config.hidden_size)
    bufs = [torch.empty(shape, dtype=dtype, device=dev) for _ in range(2)]
    recv_req = [None, None]
    send_req = []
    nll = torch.zeros((), dtype=torch.float32, device=dev)
    if not first and nsamples:
        recv_req[0] = dist.irecv(bufs[0], src=rank - 1)
    for j in range(nsamples):
        batch = ids[:, j * seqlen:(j + 1) * seqlen].to(dev)
        h0, kw = evalloop.layer_inputs(model, arch, batch)
        if first:
            h = h0
        else:
            if j + 1 < nsamples:                      # post the next receive before computing this sample
                recv_req[(j + 1) & 1] = dist.irecv(bufs[(j + 1) & 1], src=rank - 1)
            recv_req[j & 1].wait()
            h = bufs[j & 1].clone()
        for i in range(lo, hi):
            h = evalloop._call_layer(layers[i], h, kw)
        if last:
            nll += evalloop.sample_logits_nll(model, arch, h, batch, seqlen)
        else:
            send_req.append((dist.isend(h.contiguous(), dst=rank + 1), h))
            if len(send_req) > 2:
                send_req.pop(0)[0].wait()
    for r, _ in send_req:
        r.wait()
    count = torch.tensor(float(nsamples * seqlen), device=dev)
    t = torch.stack([nll, count]) if last else torch.zeros(2, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)         # only the last stage contributes
    ppl = torch.exp(t[0] / t[1]).item()
    if verbose and rank == 0:
        print(ppl)
    model.config.use_cache = use_cache
    return ppl
