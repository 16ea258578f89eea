"""CPU port of the reference's per-layer evaluation loop (oracle; test infrastructure only).

Follows reference opt.py:193-299 (opt_eval) and llama.py:174-253 (llama_eval) step by step:
  1. Catcher on decoder layer 0 records each sample's hidden states and the layer kwargs
     (opt.py:222-241);
  2. layer-major loop: for every decoder layer, for every sample (batch 1) `outs[j] = layer(inps[j])`
     (opt.py:258-268) -- on dense fp16/fp32 weights this is the reference's effective
     QuantLinear.forward (F.linear on W_ref);
  3. final norm / project_out / lm_head, shifted CrossEntropy on the logits' own dtype,
     nll = loss.float() * seqlen, ppl = exp(sum / (nsamples * seqlen)) (opt.py:270-297).
The only change: Llama decoder layers receive `position_embeddings`, which transformers >= 4.48
requires and the reference does not pass (its llama_eval crashes here).
Returns the perplexity (the reference prints it).

Pinned by tests/test_evalloop.py against the perplexity the live reference printed for the committed
tiny-OPT fixture (tests/golden/tiny_opt_2bit_incoh.npz).
"""
import time

import torch
import torch.nn as nn


class _Abort(Exception):
    pass


def _layers(model):
    return model.model.decoder.layers if hasattr(model.model, 'decoder') else model.model.layers


def _final(model, h):
    if hasattr(model.model, 'decoder'):
        d = model.model.decoder
        if d.final_layer_norm is not None:
            h = d.final_layer_norm(h)
        if getattr(d, 'project_out', None) is not None:
            h = d.project_out(h)
    elif model.model.norm is not None:
        h = model.model.norm(h)
    return model.lm_head(h)


@torch.no_grad()
def reference_eval(model, input_ids, nlayers=None, timing=None):
    """input_ids: (1, nsamples*seqlen).  `nlayers` bounds the layer loop (CPU baseline sampling);
    `timing`, if a dict, receives the seconds spent in the layer loop."""
    seqlen = model.seqlen
    nsamples = input_ids.numel() // seqlen
    use_cache = model.config.use_cache
    model.config.use_cache = False
    layers = _layers(model)
    dtype = next(iter(model.parameters())).dtype
    inps = torch.zeros((nsamples, seqlen, model.config.hidden_size), dtype=dtype)
    cache = {'i': 0, 'kw': None}

    class Catcher(nn.Module):
        def __init__(self, module):
            super().__init__()
            self.module = module

        def forward(self, inp, **kwargs):
            inps[cache['i']] = inp
            cache['i'] += 1
            cache['kw'] = {k: kwargs[k] for k in ('attention_mask', 'position_ids', 'position_embeddings') if k in kwargs}
            raise _Abort

    layers[0] = Catcher(layers[0])
    for i in range(nsamples):
        try:
            model(input_ids[:, i * seqlen:(i + 1) * seqlen])
        except _Abort:
            pass
    layers[0] = layers[0].module
    outs = torch.zeros_like(inps)
    kw = cache['kw']
    t0 = time.perf_counter()
    for li in range(len(layers) if nlayers is None else nlayers):
        layer = layers[li]
        for j in range(nsamples):
            out = layer(inps[j].unsqueeze(0), **kw)
            outs[j] = (out[0] if isinstance(out, (tuple, list)) else out)[0]
        inps, outs = outs, inps
    if timing is not None:
        timing['layer_loop_s'] = time.perf_counter() - t0
    nlls = []
    for i in range(nsamples):
        lm_logits = _final(model, inps[i].unsqueeze(0))
        shift_logits = lm_logits[:, :-1, :].contiguous()
        shift_labels = input_ids[:, i * seqlen:(i + 1) * seqlen][:, 1:]
        loss = nn.CrossEntropyLoss()(shift_logits.view(-1, shift_logits.size(-1)), shift_labels.reshape(-1))
        nlls.append(loss.float() * seqlen)
    ppl = torch.exp(torch.stack(nlls).sum() / (nsamples * seqlen))
    model.config.use_cache = use_cache
    return ppl.item()
